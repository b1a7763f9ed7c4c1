"""Thin torch-tensor wrappers over the C ABI (include/sqcircuit_b200.h).

Tensors are only carriers of device memory: every call passes raw ``data_ptr()`` values
and the current CUDA stream.  Nothing here computes on the host.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import Block, DavidsonState, Potential, Space


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class BlockView:
    """A [B][nvec][N] complex128 tensor (or a window of one) described as sqc_block_t."""

    def __init__(self, tensor, cnt_fixed=None, off=None, cnt=None, first=0, stride_b=None):
        assert tensor.dtype == torch.complex128 and tensor.dim() == 3 and tensor.is_contiguous()
        self.tensor = tensor
        self.B, self.nvec, self.N = tensor.shape
        self.first = first
        self.off = off
        self.cnt = cnt
        self.cnt_fixed = int(cnt_fixed if cnt_fixed is not None else self.nvec - first)
        # stride_b = 0: every sweep point addresses the SAME pool of vectors (tensor.shape[0] == 1)
        # through its own offset -- a gather source
        self.stride_b = stride_b

    def c(self):
        base = self.tensor.data_ptr() + self.first * self.N * 16
        sb = self.nvec * self.N if self.stride_b is None else self.stride_b
        return Block(C.c_void_p(base), sb, _p(self.off), _p(self.cnt), self.cnt_fixed)


def make_space(dims, harmonic):
    sp = Space()
    sp.n_modes = len(dims)
    for i, (d, h) in enumerate(zip(dims, harmonic)):
        sp.dims[i] = int(d)
        sp.harmonic[i] = int(bool(h))
    return sp


class Kernels:
    """CUDA backend (the only product backend)."""
    name = 'cuda'
    _PIN_BYTES = 32 << 20

    def __init__(self, device):
        self.lib = _lib.load()
        self.dev = torch.device(device)
        if self.dev.type != 'cuda':
            raise _lib.SqcLibraryError('sqcircuit_b200 kernels need a CUDA device')
        self._gram_work = None
        self._ws = {}
        self._pin, self._pin_off, self._pin_down = None, 0, None
        self._free_ref, self._free_calls = None, 0
        self.graph_launches = 0        # kernel launches replayed through CUDA graphs (not seen by the C counter)
        self.use_graphs = True         # solver iterations are captured into CUDA graphs (see eigensolver.py)
        self._capture_stream = None

    def workspace(self, key, shape, dtype):
        """Persistent device scratch (grown on demand, reused by later calls): the Davidson basis
        buffers are tens of GB, and handing them back to the caching allocator after every solve
        makes a repeated sweep re-grow its segments (measured: +180 ms on a 620 ms step)."""
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            self._ws[key] = None
            del buf
            self._ws[key] = buf = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
        return buf[:nbytes].view(dtype).view(*shape)

    def upload(self, arr, dtype=None):
        """Small host array -> device tensor through a pinned staging arena and an asynchronous copy on the
        current stream.  A pageable ``torch.as_tensor(arr, device=...)`` is a synchronous driver copy; on this
        pool single copies of a few hundred bytes were measured to take up to 45 ms (normally 0.4 ms), which
        made 300 ms sweeps fluctuate by +-15 % (profiles/r02_step_time_noise.txt)."""
        a = np.ascontiguousarray(arr)
        if not a.flags.writeable:
            a = a.copy()
        t = torch.from_numpy(a)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        n = t.numel() * t.element_size()
        if n == 0 or n > self._PIN_BYTES // 4:
            return t.to(self.dev)
        if self._pin is None:
            self._pin = torch.empty(self._PIN_BYTES, dtype=torch.uint8).pin_memory()
        off = (self._pin_off + 63) & ~63
        if off + n > self._PIN_BYTES:
            torch.cuda.current_stream(self.dev).synchronize()      # every earlier staged copy has been consumed
            off = 0
        stage = self._pin[off:off + n].view(t.dtype).view(t.shape)
        stage.copy_(t)
        out = torch.empty(t.shape, dtype=t.dtype, device=self.dev)
        out.copy_(stage, non_blocking=True)
        self._pin_off = off + n
        return out

    def download(self, t):
        """Small device tensor -> numpy array through pinned staging (the counterpart of ``upload``: a pageable
        ``.cpu()`` / ``.item()`` is the same synchronous driver copy).  Blocks until the data has arrived."""
        t = t.detach()
        if not t.is_contiguous():
            t = t.contiguous()
        n = t.numel() * t.element_size()
        if n == 0 or n > self._PIN_BYTES // 4:
            return t.cpu().numpy()
        if self._pin_down is None:
            self._pin_down = torch.empty(self._PIN_BYTES // 4, dtype=torch.uint8).pin_memory()
        stage = self._pin_down[:n].view(t.dtype).view(t.shape)
        stage.copy_(t, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return stage.numpy().copy()

    def release_workspace(self):
        self._ws = {}

    def free_bytes(self):
        """Device bytes a new batched solve may claim: free device memory, the caching allocator's unused
        reserve, and the persistent workspace (the next solve reuses it).

        cudaMemGetInfo is asked once and then only every 64th call: in between, the driver's free figure is
        tracked through the caching allocator's own reserve counter (this process is the device's only tenant on
        the path).  The driver call was measured to block for up to 96 ms now and then when issued between the
        levels of a sweep (profiles/r02_step_time_noise.txt)."""
        reserved = torch.cuda.memory_reserved(self.dev)
        if self._free_ref is None or self._free_calls % 64 == 0:
            free, _ = torch.cuda.mem_get_info(self.dev)
            self._free_ref = (int(free), int(reserved))
        self._free_calls += 1
        free = self._free_ref[0] - (int(reserved) - self._free_ref[1])
        cached = reserved - torch.cuda.memory_allocated(self.dev)
        ws = sum(int(b.numel()) for b in self._ws.values() if b is not None)
        return max(0, int(free)) + int(cached) + ws

    def xbasis(self, m):
        V = torch.empty((m, m), dtype=torch.float64, device=self.dev)
        lam = torch.empty((m,), dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.sqc_xbasis(m, _p(V), _p(lam), _stream(self.dev)), 'sqc_xbasis')
        return V, lam

    def phase_tables(self, sp, n_terms, lam, s, etab):
        B = etab.shape[0]
        sb = 0 if s.shape[0] == 1 else s.stride(0)
        _lib.check(self.lib.sqc_phase_tables(C.byref(sp), n_terms, _p(lam), _p(s), sb, _p(etab),
                                             etab.stride(0) if B > 1 else 0, B, _stream(self.dev)),
                   'sqc_phase_tables')

    def lc_diagonal(self, sp, omega, q, ng, out):
        B = out.shape[0]
        per_point = 1 if B > 1 else 0
        _lib.check(self.lib.sqc_lc_diagonal(C.byref(sp), _p(omega), _p(q), _p(ng), per_point, _p(out), B,
                                            _stream(self.dev)), 'sqc_lc_diagonal')

    # ---- (2) operator application -----------------------------------------
    def diag_mul(self, X, Y, diag, accumulate):
        sb = 0 if diag.shape[0] == 1 else diag.stride(0)
        _lib.check(self.lib.sqc_diag_mul(X.c(), Y.c(), _p(diag), sb, X.N, X.B, int(accumulate),
                                         _stream(self.dev)), 'sqc_diag_mul')

    def mode_transform(self, sp, mode, mat, transpose, X, Z, ediag=None, EX=None, impl=0):
        msb = 0 if mat.dim() == 2 or mat.shape[0] == 1 else mat.stride(0)
        if ediag is not None:
            esb = 0 if ediag.shape[0] == 1 else ediag.stride(0)
            exb = EX.c()
        else:
            esb = 0
            exb = X.c()
        _lib.check(self.lib.sqc_mode_transform(C.byref(sp), mode, _p(mat), msb, int(transpose), X.c(),
                                               Z.c(), _p(ediag), esb, exb, X.B, impl, _stream(self.dev)),
                   'sqc_mode_transform')

    def potential_apply(self, sp, pot, X, Z):
        _lib.check(self.lib.sqc_potential_apply(C.byref(sp), C.byref(pot), X.c(), Z.c(), X.B,
                                                _stream(self.dev)), 'sqc_potential_apply')

    def potential_elements(self, sp, pot, Z, out):
        """out[b][i][j] = <z_i| P |z_j> for the nv = out.shape[1] leading vectors of every point (x basis)."""
        nv = out.shape[1]
        nbytes = int(self.lib.sqc_potential_elements_work_bytes(Z.N, nv, Z.B))
        work = self.workspace('potel_work', (max(nbytes, 16),), torch.uint8)
        _lib.check(self.lib.sqc_potential_elements(C.byref(sp), C.byref(pot), Z.c(), nv, _p(out), _p(work), Z.B,
                                                   _stream(self.dev)), 'sqc_potential_elements')

    def ladder_apply(self, sp, mode, c_dn, c_up, X, Y, accumulate):
        _lib.check(self.lib.sqc_ladder_apply(C.byref(sp), mode, _p(c_dn), _p(c_up), X.c(), Y.c(), X.B,
                                             int(accumulate), _stream(self.dev)), 'sqc_ladder_apply')

    # ---- (3) eigensolver blocks --------------------------------------------
    def gram(self, A, Bm, Cout, impl=0):
        """Cout[b][i][j] = <A_i | Bm_j>; Cout: [B][ldr][ldc] complex128."""
        B, ldr, ldc = Cout.shape
        need = self.lib.sqc_gram_work_bytes(B, A.cnt_fixed, Bm.cnt_fixed, A.N)
        if need and (self._gram_work is None or self._gram_work.numel() < need):
            self._gram_work = torch.empty(need, dtype=torch.uint8, device=self.dev)
        _lib.check(self.lib.sqc_gram(A.c(), Bm.c(), A.N, B, _p(Cout), ldr, ldc,
                                     _p(self._gram_work) if need else C.c_void_p(0), impl,
                                     _stream(self.dev)), 'sqc_gram')

    def gram_pair(self, A, Bm, Bm2, C1, C2, real=False):
        """C1 = A^H Bm, C2 = A^H Bm2 in one pass over A.  real: blocks of the real representation (only the
        real inner products exist; sqc_gram_pair_real)."""
        B, ldr, ldc = C1.shape
        need = self.lib.sqc_gram_work_bytes(B, A.cnt_fixed, Bm.cnt_fixed, A.N)
        if need and (self._gram_work is None or self._gram_work.numel() < need):
            self._gram_work = torch.empty(need, dtype=torch.uint8, device=self.dev)
        fn = self.lib.sqc_gram_pair_real if real else self.lib.sqc_gram_pair
        _lib.check(fn(A.c(), Bm.c(), Bm2.c(), A.N, B, _p(C1), _p(C2), ldr, ldc,
                      _p(self._gram_work) if need else C.c_void_p(0), _stream(self.dev)), 'sqc_gram_pair')

    def gen_rr(self, GA, GB, n, drop, w, Zt, real=False, n_out=None):
        """real: real symmetric pencil (sqc_gen_rr_real), only the n_out lowest Ritz vectors are written."""
        B, ld, _ = GA.shape
        if real:
            _lib.check(self.lib.sqc_gen_rr_real(_p(GA), _p(GB), _p(n), ld, drop, _p(w), _p(Zt),
                                                int(n_out or ld), B, _stream(self.dev)), 'sqc_gen_rr_real')
            return
        _lib.check(self.lib.sqc_gen_rr(_p(GA), _p(GB), _p(n), ld, drop, _p(w), _p(Zt), B, _stream(self.dev)),
                   'sqc_gen_rr')

    def davidson_insert2(self, st, SA, SB, B):
        _, ldr, ldc = SA.shape
        _lib.check(self.lib.sqc_davidson_insert2(C.byref(st), _p(SA), _p(SB), ldr, ldc, B, _stream(self.dev)),
                   'sqc_davidson_insert2')

    def block_update(self, A, Cf, Out, op=0, trans=0, impl=0):
        B, ldr, ldc = Cf.shape
        _lib.check(self.lib.sqc_block_update(A.c(), _p(Cf), ldr, ldc, trans, Out.c(), A.N, B, op, impl,
                                             _stream(self.dev)), 'sqc_block_update')

    def ritz_update(self, A, A2, Cf, Out, Out2, theta, rn2):
        """X = Cf A, HX = Cf A2 and rn2 = ||HX - theta X||^2 in one pass (sqc_ritz_update)."""
        B, ldr, ldc = Cf.shape
        need = self.lib.sqc_ritz_update_work_bytes(B, A.N)
        if getattr(self, '_ritz_work', None) is None or self._ritz_work.numel() < need:
            self._ritz_work = torch.empty(need, dtype=torch.uint8, device=self.dev)
        _lib.check(self.lib.sqc_ritz_update(A.c(), A2.c(), _p(Cf), ldr, ldc, Out.c(), Out2.c(), _p(theta),
                                            theta.shape[1], _p(rn2), _p(self._ritz_work), A.N, B,
                                            _stream(self.dev)), 'sqc_ritz_update')

    def hx_fused(self, sp, Ve, Vo, pot, fdiag, X, Y):
        """Returns False when the fused kernel does not support the layout."""
        fsb = 0 if fdiag is None or fdiag.shape[0] == 1 else fdiag.stride(0)
        rc = self.lib.sqc_hx_fused(C.byref(sp), _p(Ve), _p(Vo), C.byref(pot), _p(fdiag), fsb, X.c(), Y.c(),
                                   X.B, _stream(self.dev))
        if rc == _lib.ELIMIT:
            return False
        _lib.check(rc, 'sqc_hx_fused')
        return True

    def hx_generic(self, sp, Vcat, pot, fdiag, X, Y):
        """H.X in one kernel for layouts with leading harmonic modes (sqc_hx_generic); False when the layout is
        not supported."""
        fsb = 0 if fdiag is None or fdiag.shape[0] == 1 else fdiag.stride(0)
        rc = self.lib.sqc_hx_generic(C.byref(sp), _p(Vcat), C.byref(pot), _p(fdiag), fsb, X.c(), Y.c(), X.B,
                                     _stream(self.dev))
        if rc == _lib.ELIMIT:
            return False
        _lib.check(rc, 'sqc_hx_generic')
        return True

    def hx_fused_real(self, sp, Ve, Vo, pot, fdiag, X, Y, fdiag_sep=None):
        """Real-representation H.X (sqc_hx_fused_real); False when the layout / size is not supported.
        fdiag_sep: [m + inner] separable form of the same diagonal (used instead of fdiag when given)."""
        fsb = 0 if fdiag is None or fdiag.shape[0] == 1 else fdiag.stride(0)
        rc = self.lib.sqc_hx_fused_real(C.byref(sp), _p(Ve), _p(Vo), C.byref(pot), _p(fdiag), fsb, _p(fdiag_sep),
                                        X.c(), Y.c(), X.B, _stream(self.dev))
        if rc == _lib.ELIMIT:
            return False
        _lib.check(rc, 'sqc_hx_fused_real')
        return True

    def real_to_complex(self, sp, Xr, Zc):
        _lib.check(self.lib.sqc_real_to_complex(C.byref(sp), Xr.c(), Zc.c(), Xr.B, _stream(self.dev)),
                   'sqc_real_to_complex')

    def complex_to_real(self, sp, Zc, Xr):
        _lib.check(self.lib.sqc_complex_to_real(C.byref(sp), Zc.c(), Xr.c(), Zc.B, _stream(self.dev)),
                   'sqc_complex_to_real')

    def block_rightmul(self, T, Cq, nout):
        B, ldr, ldc = Cq.shape
        _lib.check(self.lib.sqc_block_rightmul(T.c(), _p(Cq), ldr, ldc, _p(nout), T.N, B,
                                               _stream(self.dev)), 'sqc_block_rightmul')

    def heev(self, A, n, w, Zt):
        """A: [B][ld][ld]; n: int32[B] or int; w: [B][ld]; Zt: [B][ld][ld]."""
        B, ld, _ = A.shape
        if isinstance(n, int):
            narr, nfix = None, n
        else:
            narr, nfix = n, ld
        _lib.check(self.lib.sqc_heev(_p(A), _p(narr), nfix, ld, _p(w), _p(Zt), B, _stream(self.dev)),
                   'sqc_heev')

    def residual_norms(self, X, HX, theta, rn2):
        _lib.check(self.lib.sqc_residual_norms(X.c(), HX.c(), _p(theta), theta.shape[1], X.N, X.B,
                                               _p(rn2), _stream(self.dev)), 'sqc_residual_norms')

    def davidson_state(self, **kw):
        st = DavidsonState()
        for key in ('p', 'k', 'mmax', 'ldg', 'tol_rel', 'tol_abs_floor'):
            setattr(st, key, kw[key])
        for key in ('msz', 'nact', 'act', 'restart', 'done', 'n_not_done', 'theta', 'rn2', 'resmax',
                    'den_floor', 'G'):
            setattr(st, key, kw[key].data_ptr())
        st.GB = kw['GB'].data_ptr() if kw.get('GB') is not None else None
        st.mtot = kw['mtot'].data_ptr() if kw.get('mtot') is not None else None
        st.nexp = int(kw.get('nexp') or 0)
        st.xoff = kw['xoff'].data_ptr() if kw.get('xoff') is not None else None
        st.xlast = kw['xlast'].data_ptr() if kw.get('xlast') is not None else None
        st.xrow = int(kw.get('xrow') or 0)
        st.real_basis = int(bool(kw.get('real_basis')))
        st._keep = kw
        return st

    def davidson_decide(self, st, B):
        _lib.check(self.lib.sqc_davidson_decide(C.byref(st), B, _stream(self.dev)), 'sqc_davidson_decide')

    def restart_copy(self, X, HX, V, W, restart):
        _lib.check(self.lib.sqc_restart_copy(X.c(), HX.c(), V.c(), W.c(), _p(restart), X.N, V.B,
                                             _stream(self.dev)), 'sqc_restart_copy')

    def precond_residual(self, X, HX, T, theta, act, d, den_floor, real=False):
        """real: the block holds real amplitudes (two per element) and d is [1 or B][2 N]."""
        dsb = 0 if d.shape[0] == 1 else d.stride(0)
        fn = self.lib.sqc_precond_residual_real if real else self.lib.sqc_precond_residual
        _lib.check(fn(X.c(), HX.c(), T.c(), _p(theta), theta.shape[1], _p(act), act.shape[1], _p(d), dsb,
                      _p(den_floor), X.N, X.B, _stream(self.dev)), 'sqc_precond_residual')

    def lowblock_assemble(self, sp, S, Dt, pot, d, sigma, unit, out, real=False):
        B, ns, _ = out.shape
        ssb = 0 if S.shape[0] == 1 else S.stride(0)
        dsb = 0 if d.shape[0] == 1 else d.stride(0)
        if real:
            # float storage for the banded kernels, complex64 (zero imaginary parts) for the dense ones
            _lib.check(self.lib.sqc_lowblock_assemble_real(C.byref(sp), ns, _p(S), ssb, _p(Dt), C.byref(pot), _p(d),
                                                           dsb, _p(sigma), unit, _p(out),
                                                           int(out.dtype == torch.float32), B, _stream(self.dev)),
                       'sqc_lowblock_assemble_real')
            return
        _lib.check(self.lib.sqc_lowblock_assemble(C.byref(sp), ns, _p(S), ssb, _p(Dt), C.byref(pot), _p(d), dsb,
                                                  _p(sigma), unit, _p(out), B, _stream(self.dev)),
                   'sqc_lowblock_assemble')

    def lowblock_assemble_generic(self, sp, S, Dt, pot, d, sigma, unit, out):
        """Low-energy block for any layout from per-mode displacement tables Dt [(B or 1)][T][sum m_i^2]."""
        B, ns, _ = out.shape
        ssb = 0 if S.shape[0] == 1 else S.stride(0)
        dsb = 0 if d.shape[0] == 1 else d.stride(0)
        tsb = 0 if Dt.shape[0] == 1 else Dt.stride(0)
        _lib.check(self.lib.sqc_lowblock_assemble_generic(C.byref(sp), ns, _p(S), ssb, _p(Dt), tsb, C.byref(pot), _p(d),
                                                          dsb, _p(sigma), unit, _p(out), B, _stream(self.dev)),
                   'sqc_lowblock_assemble_generic')

    def hamiltonian_dense(self, sp, Dt, pot, d, out):
        """out[b] = H(b) as a dense matrix: complex128, or float64 for layouts without charge modes."""
        B = out.shape[0]
        dsb = 0 if d.shape[0] == 1 else d.stride(0)
        tsb = 0 if Dt.shape[0] == 1 else Dt.stride(0)
        _lib.check(self.lib.sqc_hamiltonian_dense(C.byref(sp), _p(Dt), tsb, C.byref(pot), _p(d), dsb, _p(out),
                                                  int(out.dtype == torch.float64), B, _stream(self.dev)),
                   'sqc_hamiltonian_dense')

    def syev(self, A, w, Zt):
        """Batched real symmetric eigensolver: A [B][n][n] float64, w [B][n], Zt [B][n][n] (rows = eigenvectors)."""
        B, n, _ = A.shape
        _lib.check(self.lib.sqc_syev(_p(A), n, n, _p(w), _p(Zt), B, _stream(self.dev)), 'sqc_syev')

    def hpd_inverse_c64(self, A):
        B, ns, _ = A.shape
        _lib.check(self.lib.sqc_hpd_inverse_c64(_p(A), ns, B, _stream(self.dev)), 'sqc_hpd_inverse_c64')

    def lowblock_apply(self, X, HX, T, theta, act, S, Ainv, unit, group=1, real=False):
        ssb = 0 if S.shape[0] == 1 else S.stride(0)
        if Ainv.shape[0] * group < X.B:
            raise ValueError('lowblock_apply: not enough inverse slots for this group size')
        fn = self.lib.sqc_lowblock_apply_real if real else self.lib.sqc_lowblock_apply
        _lib.check(fn(X.c(), HX.c(), T.c(), _p(theta), theta.shape[1], _p(act),
                                               act.shape[1], _p(S), ssb, _p(Ainv), Ainv.shape[1], unit, group,
                                               X.N, X.B, _stream(self.dev)), 'sqc_lowblock_apply')

    def lowband_factor(self, A, blk_off, nblk, nbmax):
        """A: complex64 [B][ns][ns], or float32 (real representation, sqc_lowband_factor_real)."""
        B, ns, _ = A.shape
        fn = self.lib.sqc_lowband_factor_real if A.dtype == torch.float32 else self.lib.sqc_lowband_factor
        _lib.check(fn(_p(A), ns, _p(blk_off), blk_off.shape[1], _p(nblk), int(nbmax), B, _stream(self.dev)),
                   'sqc_lowband_factor')

    def lowband_apply(self, X, HX, T, theta, act, S, Afac, blk_off, nblk, nbmax, unit, group=1, perm=None,
                      real=False):
        ssb = 0 if S.shape[0] == 1 else S.stride(0)
        if Afac.shape[0] * group < X.B:
            raise ValueError('lowband_apply: not enough factor slots for this group size')
        if real != (Afac.dtype == torch.float32):
            raise ValueError('lowband_apply: the real representation takes a float32 factor (and only it)')
        fn = self.lib.sqc_lowband_apply_real if real else self.lib.sqc_lowband_apply
        _lib.check(fn(X.c(), HX.c(), T.c(), _p(theta), theta.shape[1], _p(act),
                                              act.shape[1], _p(S), ssb, _p(Afac), Afac.shape[1], _p(blk_off),
                                              blk_off.shape[1], _p(nblk), int(nbmax), unit, group, _p(perm), X.N, X.B,
                                              _stream(self.dev)), 'sqc_lowband_apply')

    def svqb_coeffs(self, S, nact, drop, Cq, msz=None, mtot=None):
        B, ld, _ = S.shape
        _lib.check(self.lib.sqc_svqb_coeffs(_p(S), ld, _p(nact), drop, _p(Cq), _p(msz), _p(mtot), B,
                                            _stream(self.dev)), 'sqc_svqb_coeffs')

    def davidson_insert(self, st, S, B):
        _, ldr, ldc = S.shape
        _lib.check(self.lib.sqc_davidson_insert(C.byref(st), _p(S), ldr, ldc, B, _stream(self.dev)),
                   'sqc_davidson_insert')

    def launch_count(self):
        """Kernels launched by this process: direct launches (C counter) + launches replayed by CUDA graphs."""
        return int(self.lib.sqc_launch_count()) + self.graph_launches

    def eager_ops(self):
        """Entry points that must not be captured into a CUDA graph (none; instrumented wrappers override)."""
        return frozenset()

    def capture(self, body):
        """Capture ``body()`` (kernel launches on the current stream, no allocation, no host sync) into a CUDA
        graph.  Returns (graph, launches per replay)."""
        if self._capture_stream is None:
            self._capture_stream = torch.cuda.Stream(device=self.dev)
        g = torch.cuda.CUDAGraph()
        cur = torch.cuda.current_stream(self.dev)
        s = self._capture_stream
        s.wait_stream(cur)
        n0 = int(self.lib.sqc_launch_count())
        with torch.cuda.stream(s):
            g.capture_begin(capture_error_mode='thread_local')
            try:
                body()
            finally:
                g.capture_end()
        cur.wait_stream(s)
        return g, int(self.lib.sqc_launch_count()) - n0

    def replay(self, graph_and_count):
        graph_and_count[0].replay()
        self.graph_launches += graph_and_count[1]
