"""CPU tests (no GPU): host-side logic of the product package checked against the oracle with
the kernel contracts emulated in numpy (tests/emul_kernels.py), plus the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import torch
import pytest

import circuits as cs
import sqcircuit_b200 as sq
from emul_kernels import EmulKernels
from oracle import sqc_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'tests', 'golden')


def test_abi_library_exports_every_declared_symbol():
    from sqcircuit_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'sqcircuit_b200.h')).read()
    declared = set(re.findall(r'\b(sqc_[a-z0-9_]+)\s*\(', hdr))
    declared = {d for d in declared if not d.endswith('_t')}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().sqc_abi_version() == 4
    assert _lib.load().sqc_error_string(-1).decode().startswith('sqcircuit_b200')


def test_loader_refuses_a_library_built_from_other_sources(monkeypatch):
    """The library carries the hash of the sources it was built from (csrc/buildinfo.cu); the loader
    compares it with the tree and refuses a stale binary."""
    from sqcircuit_b200 import _lib
    assert _lib.load().sqc_source_hash().decode() == _lib.source_hash()
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'source_hash', lambda: '0' * 64)
    with pytest.raises(_lib.SqcLibraryError, match='stale'):
        _lib.load()


def test_struct_layouts_match_header():
    from sqcircuit_b200 import _lib
    assert ctypes.sizeof(_lib.Block) == 40
    assert ctypes.sizeof(_lib.Space) == 4 + 32 + 32
    assert _lib.Potential.shift.offset == 4 and _lib.Potential.lam.offset == 4 + 16 * 8 * 4 + 4


def test_topology_matches_reference_goldens():
    for name, builder in [('zeropi_25_13', cs.zero_pi), ('fluxonium_100', cs.fluxonium),
                          ('fluxonium_transmon_30_10', cs.coupled_fluxonium_transmon),
                          ('ind_shunted_1_9_23', cs.inductively_shunted)]:
        g = np.load(os.path.join(G, name + '.npz'))
        cr = builder(EmulKernels())[0]
        for key in ('S', 'R', 'wTrans', 'cInvTrans', 'lTrans', 'omega', 'B'):
            ref, got = g[key], getattr(cr, key)
            assert np.allclose(got, ref, rtol=1e-12, atol=1e-12 * max(np.abs(ref).max(), 1e-300)), (name, key)
    g = np.load(os.path.join(G, 'trt_6_6_6.npz'))
    cr = cs.trt(EmulKernels())[0]
    for key in ('wTrans', 'cInvTrans', 'omega'):
        assert np.allclose(getattr(cr, key), g[key], rtol=1e-9, atol=1e-9 * np.abs(g[key]).max()), key


@pytest.mark.parametrize('case', ['zeropi', 'fluxonium', 'transmon', 'trt', 'indsh', 'cft'])
def test_matrix_free_hamiltonian_equals_reference_matrix(case):
    K = EmulKernels()
    if case == 'zeropi':
        cr, lp, oc, ol = cs.zero_pi(K)
        t = [25, 13]
        lp.set_flux(0.31)
        ol.set_flux(0.31)
    elif case == 'fluxonium':
        cr, lp, oc, ol = cs.fluxonium(K)
        t = [100]
        lp.set_flux(0.31)
        ol.set_flux(0.31)
    elif case == 'transmon':
        cr, oc = cs.transmon(K)
        t = [20]
        cr.set_charge_offset(1, 0.2)
        oc.set_charge_offset(1, 0.2)
    elif case == 'trt':
        cr, oc = cs.trt(K)
        t = [6, 6, 6]
        for m, v in ((2, 0.13), (3, 0.41)):
            cr.set_charge_offset(m, v)
            oc.set_charge_offset(m, v)
    elif case == 'indsh':
        cr, lp, oc, ol = cs.inductively_shunted(K)
        t = [1, 9, 23]
        lp.set_flux(0.2)
        ol.set_flux(0.2)
    else:
        cr, lp, oc, ol = cs.coupled_fluxonium_transmon(K)
        t = [30, 10]
        lp.set_flux(0.2)
        ol.set_flux(0.2)
    cr.set_trunc_nums(t)
    oc.set_trunc_nums(t)
    H, Ho = cr.hamiltonian(), oc.hamiltonian().toarray()
    assert np.abs(H - Ho).max() <= 1e-13 * np.abs(Ho).max()


def _check(res, oc, setter, params, k, decs):
    ref = {d: [] for d in decs}
    scl = {d: [] for d in decs}
    for i, p in enumerate(params):
        setter(p)
        e, v = oc.diag(k + 1)
        got = res.efreqs[i].numpy()
        assert np.max(np.abs(got - e[:k]) / np.abs(e[:k])) < 1e-9
        assert 1 - o.subspace_overlap(v[:, :k], res.evecs[i].numpy().T, e[:k], w_next=e[k]) < 1e-10
        oc.diag(k)
        for d in decs:
            ref[d].append(oc.dec_rate(d, (1, 0)))
            scl[d].append(oc.dec_rate_scale(d, (1, 0)))
    for d in decs:
        got, r = np.asarray(res.dec_rate(d, (1, 0))), np.asarray(ref[d])
        assert np.all(np.abs(got - r) <= 1e-8 * np.abs(r) + 1e-9 * np.asarray(scl[d])), (d, got, r, scl[d])


def test_davidson_orchestration_zero_pi():
    K = EmulKernels()
    cr, lp, oc, ol = cs.zero_pi(K)
    cr.set_trunc_nums([25, 13])
    oc.set_trunc_nums([25, 13])
    fl = np.array([0.0, 0.3, 0.5])
    res = cr.sweep_diag(10, loop_fluxes={lp: fl})
    assert res.converged and res.iterations < 80
    _check(res, oc, ol.set_flux, fl, 10, ['capacitive', 'inductive', 'quasiparticle', 'charge', 'cc', 'flux'])


def test_memory_guard_chunks_the_batch_and_n_eig_limit():
    """sweep_diag sizes every batched Davidson call from the free device memory (Kernels.free_bytes): with
    room for two points per call a five-point sweep is solved in three calls, same results; more eigenpairs
    than the block iteration supports is an error, not a silent truncation."""
    from sqcircuit_b200 import circuit as cm
    K = EmulKernels()
    cr, lp, oc, ol = cs.zero_pi(K)
    cr.set_trunc_nums([12, 7])
    oc.set_trunc_nums([12, 7])
    fl = np.linspace(0.05, 0.45, 5)
    ref = cr.sweep_diag(4, loop_fluxes={lp: fl}, real=False).efreqs.numpy()
    Nv, (p, mmax) = cr._operators().N, cm._solver_sizes(4, cr._operators().N)
    per_point = 16 * Nv * (2 * (mmax + p) + 2 * p + 4) + 16 * (3 * mmax * mmax + 2 * 16 * mmax) + 4096
    K.free_bytes_value = int(2.5 * per_point / 0.8)
    assert cm._points_per_chunk(K, Nv, 4) == 2
    st = []
    got = cr.sweep_diag(4, loop_fluxes={lp: fl}, real=False, stats=st)
    assert np.max(np.abs(got.efreqs.numpy() - ref)) < 1e-9 * np.abs(ref).max()
    assert sum(1 for it, left in st if it == 1) == 3          # three separate solves
    K.free_bytes_value = 1 << 60
    with pytest.raises(ValueError, match='at most 14'):
        cr.sweep_diag(15, loop_fluxes={lp: fl})


def test_decide_does_not_converge_on_non_finite_ritz_values():
    """A wanted Ritz value that came back +inf (dependent basis vectors dropped by the generalised
    Rayleigh-Ritz) must keep the point unfinished (emulated contract of sqc_davidson_decide; the CUDA
    kernel is checked against the same case in tests/test_gpu_kernels.py)."""
    K = EmulKernels()
    p, k, B = 4, 3, 2
    i32, f64 = torch.int32, torch.float64
    theta = torch.tensor([[1.0, 2.0, float('inf'), float('inf')], [1.0, 2.0, 3.0, 4.0]], dtype=f64)
    st = K.davidson_state(p=p, k=k, mmax=12, ldg=12, tol_rel=1e-9, tol_abs_floor=1e-300,
                          msz=torch.full((B,), 4, dtype=i32), nact=torch.zeros(B, dtype=i32),
                          act=torch.zeros((B, p), dtype=i32), restart=torch.zeros(B, dtype=i32),
                          done=torch.zeros(B, dtype=i32), n_not_done=torch.zeros(1, dtype=i32), theta=theta,
                          rn2=torch.zeros((B, p), dtype=f64), resmax=torch.zeros(B, dtype=f64),
                          den_floor=torch.zeros(B, dtype=f64), G=torch.zeros((B, 12, 12), dtype=torch.complex128))
    K.davidson_decide(st, B)
    assert st.done.tolist() == [0, 1] and int(st.n_not_done[0]) == 1
    assert np.isfinite(float(st.den_floor[0]))


def test_real_representation_host_logic():
    """Zero gate charge -> the sweep is solved in the real (time-reversal adapted) representation; cold
    path and continuation (warm starts + banded / dense low block assembled in the real basis), against
    the complex solve and the oracle."""
    from sqcircuit_b200 import operators as ops_mod
    K = EmulKernels()
    cr, lp, oc, ol = cs.zero_pi(K)
    cr.set_trunc_nums([16, 9])
    oc.set_trunc_nums([16, 9])
    assert cr._operators().supports_real(np.zeros((1, 2)))
    assert not cr._operators().supports_real(np.array([[0.0, 0.1]]))
    used = []
    orig = ops_mod.RealHamiltonianOperator.apply

    def spy(self, X, Y):
        used.append(X.B)
        return orig(self, X, Y)
    ops_mod.RealHamiltonianOperator.apply = spy
    try:
        fl = np.linspace(0.0, 0.5, 33)
        rr = cr.sweep_diag(6, loop_fluxes={lp: fl}, lowblock=200)
        assert used, 'the real representation was not used'
        n_real = len(used)
        rc = cr.sweep_diag(6, loop_fluxes={lp: fl}, real=False, lowblock=200)
        assert len(used) == n_real
    finally:
        ops_mod.RealHamiltonianOperator.apply = orig
    er, ec = rr.efreqs.numpy(), rc.efreqs.numpy()
    assert np.max(np.abs(er - ec) / np.abs(ec)) < 1e-10
    Z = rr.evecs.reshape(33, 6, 16, 17)
    assert float((Z.flip(3).conj() - Z).abs().max()) < 1e-14
    _check(rr, oc, ol.set_flux, fl, 6, ['capacitive', 'inductive', 'charge', 'flux'])


def test_dense_path_fluxonium():
    K = EmulKernels()
    cr, lp, oc, ol = cs.fluxonium(K)
    cr.set_trunc_nums([100])
    oc.set_trunc_nums([100])
    fl = np.array([0.0, 0.25, 0.5])
    res = cr.sweep_diag(10, loop_fluxes={lp: fl})
    _check(res, oc, ol.set_flux, fl, 10, ['capacitive', 'inductive', 'quasiparticle', 'cc', 'flux'])


def test_charge_sweep_transmon_and_three_modes():
    K = EmulKernels()
    cr, oc = cs.transmon(K)
    cr.set_trunc_nums([100])
    oc.set_trunc_nums([100])
    ngs = np.array([0.0, 0.25, 0.5])
    res = cr.sweep_diag(10, charge_offsets={1: ngs})
    _check(res, oc, lambda g: oc.set_charge_offset(1, float(g)), ngs, 10, ['capacitive', 'cc', 'charge'])
    cr, oc = cs.trt(K)
    cr.set_trunc_nums([6, 6, 6])
    oc.set_trunc_nums([6, 6, 6])
    res = cr.sweep_diag(10, charge_offsets={2: np.array([0.0, 0.13]), 3: np.array([0.0, 0.41])})

    def s(p):
        oc.set_charge_offset(2, p[0])
        oc.set_charge_offset(3, p[1])
    _check(res, oc, s, [(0.0, 0.0), (0.13, 0.41)], 10, ['capacitive', 'cc', 'charge'])


def test_sweep_class_and_single_point_api():
    K = EmulKernels()
    cr, lp, _, _ = cs.zero_pi(K)
    cr.set_trunc_nums([25, 13])
    ef, dec = sq.Sweep(cr, 2, properties=['efreq', 'loss']).sweepFlux([lp], [np.linspace(0, 0.5, 3)])
    assert ef.shape == (2, 3)
    assert np.allclose(ef[1] - ef[0], [0.69721321, 0.45628775, 0.0239571], rtol=1e-6)
    assert set(dec) == {'capacitive', 'inductive', 'quasiparticle', 'charge', 'cc', 'flux'}
    with pytest.raises(sq.CircuitStateError):
        cs.zero_pi(K)[0].diag(2)
    with pytest.raises(ValueError):
        cr.diag(2.0)
    efs, evs = cr.diag(3)
    assert len(evs) == 3 and evs[0].full().shape == (625, 1)
    assert np.allclose(cr.efreqs, efs)


def test_elements_api_matches_reference_values():
    cap = sq.Capacitor(10, 'GHz')
    assert cap.get_value('GHz') == 10 and np.isclose(cap.get_value(), 2e-15, rtol=2e-2)
    with pytest.raises(ValueError, match='The input unit is not valid'):
        sq.Capacitor(10, 'H')
    ind = sq.Inductor(10)
    assert np.isclose(ind.get_value(), 1.6346e-8, rtol=1e-4)
    assert np.isclose(sq.Inductor(10, 'GHz', Q='default').Q(100, 0.015), 559912673679772.0)
    assert np.isclose(sq.Capacitor(3.4, 'GHz').Q(0.56e9), 19041456.468334354)
    assert np.isclose(sq.Junction(0.142, 'GHz').Y(50e9, 0.015), 793439739057240.9)
    with pytest.raises(sq.ModeError):
        sq.Capacitor(10, requires_grad=True)


def test_cold_start_finds_every_symmetry_sector():
    """0-pi at zero external flux has an exact parity: a start block of bare basis states spans only
    some of its sectors and the Davidson iteration would 'converge' while skipping the 4th level
    (this is the scenario of __graft_entry__.smoke()).  The perturbed guard vectors must seed it."""
    K = EmulKernels()
    cr, lp, oc, ol = cs.zero_pi(K)
    cr.set_trunc_nums([25, 13])
    oc.set_trunc_nums([25, 13])
    res = cr.sweep_diag(4, loop_fluxes={lp: np.array([0.0, 0.5])})
    for i, f in enumerate((0.0, 0.5)):
        ol.set_flux(f)
        e, _ = oc.diag(4)
        assert np.max(np.abs(res.efreqs[i].numpy() - e) / np.abs(e)) < 1e-9, (f, res.efreqs[i], e)


def test_banded_low_block_equals_dense_inverse():
    """The banded two-level preconditioner must be the exact inverse of (H_SS - sigma): the index set
    in charge-major order (with merged thin blocks) has to make H_SS block tridiagonal with nothing
    left outside the three block diagonals.  Emulated factor + substitution against a dense solve of
    the separately assembled matrix."""
    from sqcircuit_b200.operators import HamiltonianOperator
    from sqcircuit_b200.kernels import BlockView
    K = EmulKernels()
    cr, lp, _, _ = cs.zero_pi(K)
    cr.set_trunc_nums([25, 13])
    ops = cr._operators()
    lv0, ng0 = cr._current_point()
    B = 3
    lv = np.repeat(lv0, B, axis=0)
    lv[:, 0] = 2 * np.pi * np.array([0.11, 0.12, 0.13])
    hop = HamiltonianOperator(ops, lv, ng0)
    sigma = torch.full((B,), -3.0 * 2 * np.pi * 1e9, dtype=torch.float64)
    lb = hop.build_lowblock(sigma, 200, group=1)
    assert lb['band'] is not None and lb['ns'] == 200
    off, nblk, nbmax, perm = lb['band']
    assert int(nblk.max()) >= 3 and nbmax <= 64
    # dense reference: same index set, assembled again, inverted in double precision
    dense = torch.empty((B, 200, 200), dtype=torch.complex64)
    pot = ops.potential_struct(hop.a, hop.g, ops.etab)
    K.lowblock_assemble(ops.sp, lb['S'], ops.displacement_tables(), pot, hop.diag, sigma, lb['unit'], dense)
    p, N = 4, ops.N
    g = torch.Generator().manual_seed(3)
    X = torch.complex(torch.randn((B, p, N), dtype=torch.float64, generator=g), torch.randn((B, p, N), dtype=torch.float64, generator=g))
    HX = torch.complex(torch.randn((B, p, N), dtype=torch.float64, generator=g), torch.randn((B, p, N), dtype=torch.float64, generator=g))
    theta = torch.randn((B, 8), dtype=torch.float64, generator=g)
    act = torch.arange(p, dtype=torch.int32).repeat(B, 1)
    T = torch.zeros((B, p, N), dtype=torch.complex128)
    K.lowband_apply(BlockView(X), BlockView(HX), BlockView(T), theta, act, lb['S'], lb['Ainv'], off, nblk, nbmax,
                    lb['unit'], 1, perm)
    for b in range(B):
        s = lb['S'][min(b, lb['S'].shape[0] - 1)].long().numpy()
        A = dense[b].numpy().astype(complex)
        for j in range(p):
            r = (HX[b, j].numpy()[s] - float(theta[b, j]) * X[b, j].numpy()[s]) / lb['unit']
            ref = np.linalg.solve(A, r)
            got = T[b, j].numpy()[s]
            assert np.abs(got - ref).max() <= 2e-4 * np.abs(ref).max(), (b, j)
            mask = np.ones(N, dtype=bool)
            mask[s] = False
            assert not T[b, j].numpy()[mask].any()        # nothing outside the low-energy set is touched


def test_warm_start_gather_and_cold_start_block():
    from sqcircuit_b200 import eigensolver as es
    K = EmulKernels()
    g = torch.Generator().manual_seed(9)
    store = torch.complex(torch.randn((5, 3, 40), dtype=torch.float64, generator=g),
                          torch.randn((5, 3, 40), dtype=torch.float64, generator=g))
    ia, ib = torch.tensor([4, 0, 2, 2]), torch.tensor([1, 3, 2, 0])
    V = torch.zeros((4, 9, 40), dtype=torch.complex128)
    es.RitzPairStart(store, ia, ib).write_into(K, V)
    assert torch.equal(V[:, :3], store[ia]) and torch.equal(V[:, 3:6], store[ib]) and not V[:, 6:].any()
    # cold start: the k wanted vectors are bare basis states, only the guards are perturbed
    diag = torch.rand((1, 300), dtype=torch.float64, generator=g)
    V0 = es.cold_start_block(diag, 2, 6, 300, torch.device('cpu'), k=4)
    low = torch.topk(diag, 6, dim=1, largest=False).indices[0]
    for j in range(4):
        e = torch.zeros(300, dtype=torch.complex128)
        e[low[j]] = 1
        assert torch.equal(V0[0, j], e) and torch.equal(V0[1, j], e)
    for j in (4, 5):
        assert int((V0[0, j] != 0).sum()) > 6 and abs(V0[0, j][low[j]] - 1) < 0.05


def test_low_block_sharing_only_between_neighbouring_points():
    """Factorisations are shared by groups of consecutive sweep points only when consecutive points
    are close in parameter space (sorted dense sweeps), never for points in arbitrary order."""
    from sqcircuit_b200.operators import HamiltonianOperator
    K = EmulKernels()
    cr, lp, _, _ = cs.zero_pi(K)
    cr.set_trunc_nums([25, 13])
    ops = cr._operators()
    lv0, ng0 = cr._current_point()

    def group_of(fluxes):
        lv = np.repeat(lv0, len(fluxes), axis=0)
        lv[:, 0] = 2 * np.pi * np.asarray(fluxes)
        return HamiltonianOperator(ops, lv, ng0)._auto_lowblock_group()
    assert group_of(np.linspace(0.1, 0.1 + 63 / 4096, 64)) == 8          # dense sorted sweep
    assert group_of(np.linspace(0.0, 0.5, 9)) == 1                        # coarse sweep
    rng = np.random.default_rng(0)
    assert group_of(rng.permutation(np.linspace(0.0, 1.0, 64))) == 1      # arbitrary order
    assert group_of([0.3]) == 1


def test_gate_charge_flux_grid_three_modes_cpu():
    """The body of the GPU test of BASELINE configs[3]'s ng x flux grid, kernels emulated, small truncation."""
    from test_gpu_parity import _grid_check
    res = _grid_check(EmulKernels(), [4, 4, 4], 5, 2)
    assert res.converged


def test_warm_start_from_nearly_dependent_neighbour_blocks_converges():
    """BASELINE configs[3] at its named size (N = 98010), one point of the gate-charge x flux grid started from
    the Ritz blocks of two neighbours that lie at the SAME gate charge: as raw basis vectors that union gave
    cond(GB) ~ 1e8 and the non-orthogonalised iteration stalled at 1e-6 for ever (found on the GPU,
    tools/stagnation_diag.py + tools/warm_point_diag.py); with the orthonormalised start block it converges in
    < 25 iterations.  Kernels emulated."""
    from sqcircuit_b200.eigensolver import RitzPairStart, solve_davidson_gen
    from sqcircuit_b200.operators import HamiltonianOperator
    K = EmulKernels()
    cr, lp, oc, ol = cs.trt_squid(K)
    cr.set_trunc_nums([10, 50, 50])
    ops = cr._operators()
    m1 = sorted(cr.charge_islands)[0]

    def hop(pts):            # (gate-charge index, flux index) on the 16 x 16 grid of bench.py's trt1e5
        lv = np.array([[2 * np.pi * (c + 0.5) / 32] for _, c in pts])
        ng = np.zeros((len(pts), ops.M))
        ng[:, m1] = [r / 32 for r, _ in pts]
        return HamiltonianOperator(ops, lv, ng)
    h = hop([(0, 1), (0, 2)])
    cold = solve_davidson_gen(K, h, 2, ops.N, 10, h.diag, lowblock_ns=512)
    assert cold.converged
    h = hop([(2, 1)])
    x0 = RitzPairStart(cold.ritz_block, torch.tensor([0]), torch.tensor([1]))
    raw = solve_davidson_gen(K, h, 1, ops.N, 10, h.diag, lowblock_ns=512, x0=x0, maxit=30, orth_start=False)
    assert not raw.converged and float(raw.resmax.max()) > 1e-7          # the failure mode this guards against
    res = solve_davidson_gen(K, h, 1, ops.N, 10, h.diag, lowblock_ns=512, x0=x0, maxit=30)
    assert res.converged and res.iterations < 25 and float(res.resmax.max()) <= 1e-9


def test_generic_low_block_is_block_tridiagonal_in_one_charge_mode():
    """Two-level preconditioner of a layout with several charge modes (transmon - resonator - SQUID transmon): the
    index set ordered by the value of ONE charge mode makes H_SS block tridiagonal (every junction term shifts a
    charge by at most one), which is what the banded factorisation relies on.  Checks the assembled block against
    the block offsets handed to sqc_lowband_factor, per slot (the gate charge, hence the index set, differs per
    point)."""
    from sqcircuit_b200.operators import HamiltonianOperator
    K = EmulKernels()
    cr, lp, oc, ol = cs.trt_squid(K)
    cr.set_trunc_nums([4, 12, 12])
    ops = cr._operators()
    m1 = sorted(cr.charge_islands)[0]
    lv = 2 * np.pi * np.array([[0.1], [0.33]])
    ng = np.zeros((2, ops.M))
    ng[:, m1] = [0.05, 0.4]
    hop = HamiltonianOperator(ops, lv, ng)
    seen = {}
    factor = K.lowband_factor

    def spy(A, off, nblk, nbmax):
        seen['A'], seen['off'], seen['nblk'], seen['nbmax'] = A.clone(), off.clone(), nblk.clone(), nbmax
        return factor(A, off, nblk, nbmax)
    K.lowband_factor = spy
    sigma = torch.full((2,), -1e12, dtype=torch.float64)
    lb = hop.build_lowblock(sigma, 512, group=1)
    assert lb['band'] is not None and lb['ns'] > 160 and seen['nbmax'] <= 64
    for b in range(2):
        A = seen['A'][b].numpy()
        off = seen['off'][b].numpy()[:int(seen['nblk'][b]) + 1]
        blk = np.searchsorted(off, np.arange(lb['ns']), side='right') - 1
        far = np.abs(blk[:, None] - blk[None, :]) > 1
        assert np.abs(A[far]).max() == 0.0 and np.abs(A[~far]).max() > 0.0
        assert len(set(lb['S'][b].tolist())) == lb['ns']
    assert not torch.equal(lb['S'][0], lb['S'][1])          # per-point index sets (different gate charge)
