"""ctypes binding of libsqcircuit_b200.so (the C ABI declared in include/sqcircuit_b200.h).

There is NO fallback: if the CUDA library is missing or does not export the ABI the import of
any compute path raises.  Build it with ``python __graft_entry__.py`` (nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libsqcircuit_b200.so')
MAX_MODES = 8
MAX_TERMS = 16
EINVAL, ELIMIT = -1, -2

c_i32, c_i64, c_dbl, c_vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p


class Block(C.Structure):
    """sqc_block_t"""
    _fields_ = [('ptr', c_vp), ('stride_b', c_i64), ('off', c_vp), ('cnt', c_vp),
                ('cnt_fixed', c_i32)]


class Space(C.Structure):
    """sqc_space_t"""
    _fields_ = [('n_modes', c_i32), ('dims', c_i32 * MAX_MODES), ('harmonic', c_i32 * MAX_MODES)]


class Potential(C.Structure):
    """sqc_potential_t"""
    _fields_ = [('n_terms', c_i32), ('shift', (c_i32 * MAX_MODES) * MAX_TERMS), ('lam', c_vp),
                ('etab', c_vp), ('etab_stride_b', c_i64), ('g', c_vp), ('a', c_vp),
                ('diag', c_vp), ('diag_stride_b', c_i64)]


class DavidsonState(C.Structure):
    """sqc_davidson_state_t"""
    _fields_ = [('p', c_i32), ('k', c_i32), ('mmax', c_i32), ('ldg', c_i32),
                ('tol_rel', c_dbl), ('tol_abs_floor', c_dbl),
                ('msz', c_vp), ('nact', c_vp), ('act', c_vp), ('restart', c_vp), ('done', c_vp),
                ('n_not_done', c_vp), ('theta', c_vp), ('rn2', c_vp), ('resmax', c_vp),
                ('den_floor', c_vp), ('G', c_vp), ('GB', c_vp), ('mtot', c_vp), ('nexp', c_i32),
                ('xoff', c_vp), ('xlast', c_vp), ('xrow', c_i32), ('real_basis', c_i32)]


# name -> (restype, argtypes); must list every symbol of include/sqcircuit_b200.h
SIGNATURES = {
    'sqc_abi_version': (c_i32, []),
    'sqc_error_string': (C.c_char_p, [c_i32]),
    'sqc_launch_count': (c_i64, []),
    'sqc_source_hash': (C.c_char_p, []),
    'sqc_xbasis': (c_i32, [c_i32, c_vp, c_vp, c_vp]),
    'sqc_phase_tables': (c_i32, [C.POINTER(Space), c_i32, c_vp, c_vp, c_i64, c_vp, c_i64, c_i32, c_vp]),
    'sqc_lc_diagonal': (c_i32, [C.POINTER(Space), c_vp, c_vp, c_vp, c_i64, c_vp, c_i32, c_vp]),
    'sqc_diag_mul': (c_i32, [Block, Block, c_vp, c_i64, c_i64, c_i32, c_i32, c_vp]),
    'sqc_mode_transform': (c_i32, [C.POINTER(Space), c_i32, c_vp, c_i64, c_i32, Block, Block, c_vp,
                                   c_i64, Block, c_i32, c_i32, c_vp]),
    'sqc_potential_apply': (c_i32, [C.POINTER(Space), C.POINTER(Potential), Block, Block, c_i32, c_vp]),
    'sqc_potential_elements_work_bytes': (c_i64, [c_i64, c_i32, c_i32]),
    'sqc_potential_elements': (c_i32, [C.POINTER(Space), C.POINTER(Potential), Block, c_i32, c_vp, c_vp, c_i32, c_vp]),
    'sqc_ladder_apply': (c_i32, [C.POINTER(Space), c_i32, c_vp, c_vp, Block, Block, c_i32, c_i32, c_vp]),
    'sqc_gram_work_bytes': (c_i64, [c_i32, c_i32, c_i32, c_i64]),
    'sqc_gram': (c_i32, [Block, Block, c_i64, c_i32, c_vp, c_i32, c_i32, c_vp, c_i32, c_vp]),
    'sqc_gram_pair': (c_i32, [Block, Block, Block, c_i64, c_i32, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp]),
    'sqc_gram_pair_real': (c_i32, [Block, Block, Block, c_i64, c_i32, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp]),
    'sqc_gen_rr': (c_i32, [c_vp, c_vp, c_vp, c_i32, c_dbl, c_vp, c_vp, c_i32, c_vp]),
    'sqc_gen_rr_real': (c_i32, [c_vp, c_vp, c_vp, c_i32, c_dbl, c_vp, c_vp, c_i32, c_i32, c_vp]),
    'sqc_davidson_insert2': (c_i32, [C.POINTER(DavidsonState), c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]),
    'sqc_ritz_update_work_bytes': (c_i64, [c_i32, c_i64]),
    'sqc_ritz_update': (c_i32, [Block, Block, c_vp, c_i32, c_i32, Block, Block, c_vp, c_i32, c_vp, c_vp, c_i64, c_i32,
                                c_vp]),
    'sqc_block_update': (c_i32, [Block, c_vp, c_i32, c_i32, c_i32, Block, c_i64, c_i32, c_i32, c_i32, c_vp]),
    'sqc_hx_fused': (c_i32, [C.POINTER(Space), c_vp, c_vp, C.POINTER(Potential), c_vp, c_i64, Block, Block, c_i32, c_vp]),
    'sqc_hx_generic': (c_i32, [C.POINTER(Space), c_vp, C.POINTER(Potential), c_vp, c_i64, Block, Block, c_i32, c_vp]),
    'sqc_block_rightmul': (c_i32, [Block, c_vp, c_i32, c_i32, c_vp, c_i64, c_i32, c_vp]),
    'sqc_heev': (c_i32, [c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp]),
    'sqc_residual_norms': (c_i32, [Block, Block, c_vp, c_i32, c_i64, c_i32, c_vp, c_vp]),
    'sqc_davidson_decide': (c_i32, [C.POINTER(DavidsonState), c_i32, c_vp]),
    'sqc_restart_copy': (c_i32, [Block, Block, Block, Block, c_vp, c_i64, c_i32, c_vp]),
    'sqc_precond_residual': (c_i32, [Block, Block, Block, c_vp, c_i32, c_vp, c_i32, c_vp, c_i64, c_vp,
                                     c_i64, c_i32, c_vp]),
    'sqc_lowblock_assemble': (c_i32, [C.POINTER(Space), c_i32, c_vp, c_i64, c_vp, C.POINTER(Potential), c_vp, c_i64,
                                      c_vp, c_dbl, c_vp, c_i32, c_vp]),
    'sqc_lowblock_assemble_generic': (c_i32, [C.POINTER(Space), c_i32, c_vp, c_i64, c_vp, c_i64, C.POINTER(Potential), c_vp,
                                              c_i64, c_vp, c_dbl, c_vp, c_i32, c_vp]),
    'sqc_hamiltonian_dense': (c_i32, [C.POINTER(Space), c_vp, c_i64, C.POINTER(Potential), c_vp, c_i64, c_vp, c_i32, c_i32, c_vp]),
    'sqc_syev': (c_i32, [c_vp, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp]),
    'sqc_hpd_inverse_c64': (c_i32, [c_vp, c_i32, c_i32, c_vp]),
    'sqc_lowband_factor': (c_i32, [c_vp, c_i32, c_vp, c_i32, c_vp, c_i32, c_i32, c_vp]),
    'sqc_lowband_apply': (c_i32, [Block, Block, Block, c_vp, c_i32, c_vp, c_i32, c_vp, c_i64, c_vp, c_i32, c_vp, c_i32,
                                  c_vp, c_i32, c_dbl, c_i32, c_vp, c_i64, c_i32, c_vp]),
    'sqc_lowblock_apply': (c_i32, [Block, Block, Block, c_vp, c_i32, c_vp, c_i32, c_vp, c_i64, c_vp, c_i32, c_dbl,
                                   c_i32, c_i64, c_i32, c_vp]),
    'sqc_hx_fused_real': (c_i32, [C.POINTER(Space), c_vp, c_vp, C.POINTER(Potential), c_vp, c_i64, c_vp, Block, Block,
                                  c_i32, c_vp]),
    'sqc_real_to_complex': (c_i32, [C.POINTER(Space), Block, Block, c_i32, c_vp]),
    'sqc_complex_to_real': (c_i32, [C.POINTER(Space), Block, Block, c_i32, c_vp]),
    'sqc_precond_residual_real': (c_i32, [Block, Block, Block, c_vp, c_i32, c_vp, c_i32, c_vp, c_i64, c_vp,
                                          c_i64, c_i32, c_vp]),
    'sqc_lowblock_assemble_real': (c_i32, [C.POINTER(Space), c_i32, c_vp, c_i64, c_vp, C.POINTER(Potential), c_vp,
                                           c_i64, c_vp, c_dbl, c_vp, c_i32, c_i32, c_vp]),
    'sqc_lowband_factor_real': (c_i32, [c_vp, c_i32, c_vp, c_i32, c_vp, c_i32, c_i32, c_vp]),
    'sqc_lowband_apply_real': (c_i32, [Block, Block, Block, c_vp, c_i32, c_vp, c_i32, c_vp, c_i64, c_vp, c_i32, c_vp,
                                       c_i32, c_vp, c_i32, c_dbl, c_i32, c_vp, c_i64, c_i32, c_vp]),
    'sqc_lowblock_apply_real': (c_i32, [Block, Block, Block, c_vp, c_i32, c_vp, c_i32, c_vp, c_i64, c_vp, c_i32,
                                        c_dbl, c_i32, c_i64, c_i32, c_vp]),
    'sqc_svqb_coeffs': (c_i32, [c_vp, c_i32, c_vp, c_dbl, c_vp, c_vp, c_vp, c_i32, c_vp]),
    'sqc_davidson_insert': (c_i32, [C.POINTER(DavidsonState), c_vp, c_i32, c_i32, c_i32, c_vp]),
}

_lib = None


def source_hash():
    """SHA-256 over csrc/*.cu, csrc/*.cuh and include/sqcircuit_b200.h (names and contents, sorted): the
    identity the build embeds in the library (csrc/buildinfo.cu) and ``load`` checks."""
    import hashlib
    csrc = os.path.join(_HERE, 'csrc')
    files = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(('.cu', '.cuh')))
    files.append(os.path.join(os.path.dirname(_HERE), 'include', 'sqcircuit_b200.h'))
    h = hashlib.sha256()
    for f in files:
        h.update(os.path.basename(f).encode() + b'\0')
        with open(f, 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()


class SqcLibraryError(RuntimeError):
    pass


def load():
    """Load the CUDA library; raise loudly if it is missing or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SqcLibraryError(
            f'{LIB_PATH} not found: the sqcircuit_b200 CUDA extension has not been built '
            '(run `python __graft_entry__.py`). There is no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise SqcLibraryError(f'{LIB_PATH} does not export {name}') from e
        fn.restype = res
        fn.argtypes = args
    if lib.sqc_abi_version() != 4:
        raise SqcLibraryError('sqcircuit_b200 ABI version mismatch')
    built, here = lib.sqc_source_hash().decode(), source_hash()
    if built != here:
        raise SqcLibraryError(
            f'{LIB_PATH} is stale: built from sources {built[:12]}, the tree holds {here[:12]} '
            '(rebuild with `python __graft_entry__.py`)')
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().sqc_error_string(rc).decode()
        raise SqcLibraryError(f'{what}: {msg} (code {rc})')
