"""End-to-end parity on the GPU: the product path (C ABI, CUDA kernels) against the CPU oracle
and the golden fixtures generated from the reference.

Tolerances (BASELINE.json north_star): eigenfrequencies 1e-9 relative, eigenvectors by
subspace overlap > 1 - 1e-10, |matrix elements| and decay rates 1e-8 relative.  Where a symmetry
forbids a decay the reference itself returns rounding noise, so a floor applies: 1e-9 x the size the
rate would have without the suppression (``dec_rate_scale`` of the oracle: inner products replaced by
||O v||), or 1e-9 x the channel's largest value over the sweep against stored goldens.  No absolute floor."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import circuits as cs  # noqa: E402
from oracle import sqc_oracle as o  # noqa: E402

G = os.path.join(os.path.dirname(__file__), 'golden')
DEC = ['capacitive', 'inductive', 'quasiparticle', 'charge', 'cc', 'flux']


def _rates_close(got, ref, scale=None):
    """1e-8 relative; floor: 1e-9 x the size the rate would have without symmetry suppression
    (oracle dec_rate_scale) when known, else 1e-9 x the channel's largest value over the sweep."""
    got, ref = np.asarray(got, dtype=float), np.asarray(ref, dtype=float)
    floor = 1e-9 * (np.abs(ref).max() if scale is None else np.asarray(scale, dtype=float))
    return np.all(np.abs(got - ref) <= 1e-8 * np.abs(ref) + floor)


def _scales(oc, setter, params, k, decs):
    """dec_rate_scale of the oracle at every sweep point (floor for comparisons with stored goldens)."""
    out = {d: [] for d in decs}
    for prm in params:
        setter(prm)
        oc.diag(k)
        for d in decs:
            out[d].append(oc.dec_rate_scale(d, (1, 0)))
    return out


def _check(res, oc, setter, params, k, decs, mels=(), idx=None):
    """Compare sweep points ``idx`` of ``res`` (default: all) with the oracle at ``params``."""
    ref_dec = {d: [] for d in decs}
    ref_scale = {d: [] for d in decs}
    idx = list(range(len(params))) if idx is None else list(idx)
    for i, prm in zip(idx, params):
        setter(prm)
        e, v = oc.diag(k + 1)
        got = res.efreqs[i].cpu().numpy()
        assert np.max(np.abs(got - e[:k]) / np.abs(e[:k])) < 1e-9, (i, got, e)
        ov = o.subspace_overlap(v[:, :k], res.evecs[i].cpu().numpy().T, e[:k], w_next=e[k])
        assert 1 - ov < 1e-10, (i, ov)
        oc.diag(k)
        for d in decs:
            ref_dec[d].append(oc.dec_rate(d, (1, 0)))
            ref_scale[d].append(oc.dec_rate_scale(d, (1, 0)))
        for (ct, nodes, st) in mels:
            a = abs(res.matrix_elements(ct, nodes, st)[i])
            b = abs(oc.matrix_elements(ct, nodes, st))
            # symmetry-forbidden elements are rounding noise: floor relative to ||O v||
            scale = np.linalg.norm(oc.coupling_op(ct, nodes) @ oc.evecs[:, st[1]])
            assert abs(a - b) <= 1e-8 * b + 1e-9 * scale, (ct, nodes, a, b, scale)
    for d in decs:
        got = np.asarray(res.dec_rate(d, (1, 0)))[idx]
        ref = np.asarray(ref_dec[d])
        assert _rates_close(got, ref, ref_scale[d]), (d, got, ref, ref_scale[d])


def test_hamiltonian_matrix_matches_reference_golden():
    g = np.load(os.path.join(G, 'zeropi_25_13.npz'))
    cr, lp, _, _ = cs.zero_pi()
    cr.set_trunc_nums([25, 13])
    lp.set_flux(0.31)
    H = cr.hamiltonian()
    assert np.abs(H - g['H_at_0p31']).max() <= 1e-12 * np.abs(H).max()
    g = np.load(os.path.join(G, 'fluxonium_100.npz'))
    cr, lp, _, _ = cs.fluxonium()
    cr.set_trunc_nums([100])
    lp.set_flux(0.31)
    H = cr.hamiltonian()
    assert np.abs(H - g['H_at_0p31']).max() <= 1e-12 * np.abs(H).max()
    g = np.load(os.path.join(G, 'transmon_ncut30.npz'))
    cr, _ = cs.transmon_single()
    cr.set_trunc_nums([30])
    cr.set_charge_offset(1, 0.2)
    H = cr.hamiltonian()
    assert np.abs(H - g['H_at_ng0p2']).max() <= 1e-12 * np.abs(H).max()


def test_config0_single_transmon_ncut30():
    cr, oc = cs.transmon_single()
    cr.set_trunc_nums([30])
    oc.set_trunc_nums([30])
    ngs = np.array([0.0, 0.2])
    res = cr.sweep_diag(10, charge_offsets={1: ngs})
    _check(res, oc, lambda x: oc.set_charge_offset(1, float(x)), ngs, 10, ['capacitive', 'cc', 'charge'],
           mels=[('capacitive', (0, 1), (0, 1))])
    g = np.load(os.path.join(G, 'transmon_ncut30.npz'))
    assert np.max(np.abs(res.efreqs.cpu().numpy() - g['efreqs_exact']) / np.abs(g['efreqs_exact'])) < 1e-9
    # the single-point API of the reference
    cr.set_charge_offset(1, 0.2)
    ef, ev = cr.diag(10)
    assert np.max(np.abs(ef - g['efreqs_exact'][1]) / np.abs(ef)) < 1e-9
    assert abs(cr.dec_rate('cc', (1, 0)) - g['dec_cc'][1]) <= 1e-8 * g['dec_cc'][1]


def test_config1_fluxonium_flux_sweep():
    cr, lp, oc, ol = cs.fluxonium()
    cr.set_trunc_nums([100])
    oc.set_trunc_nums([100])
    g = np.load(os.path.join(G, 'fluxonium_100.npz'))
    fl = g['params']
    res = cr.sweep_diag(10, loop_fluxes={lp: fl})
    ef = res.efreqs.cpu().numpy()
    assert np.max(np.abs(ef - g['efreqs_exact']) / np.abs(g['efreqs_exact'])) < 1e-9
    decs = ['capacitive', 'inductive', 'quasiparticle', 'cc', 'flux']
    scl = _scales(oc, ol.set_flux, fl, 10, decs)
    for d in decs:
        got, ref = res.dec_rate(d, (1, 0)), g['dec_' + d]
        assert _rates_close(got, ref, scl[d]), (d, got, ref)
    for i in (0, 4, 8):
        ov = o.subspace_overlap(g['evecs_%d' % i], res.evecs[i].cpu().numpy().T, g['efreqs_exact'][i])
        assert 1 - ov < 1e-10
    _check(res, oc, ol.set_flux, fl[:3], 10, ['capacitive', 'inductive', 'quasiparticle', 'cc', 'flux'],
           mels=[('capacitive', (0, 1), (0, 1)), ('inductive', (0, 1), (0, 1))], idx=[0, 1, 2])


def test_config2_zero_pi_flux_sweep_with_rates():
    cr, lp, oc, ol = cs.zero_pi()
    cr.set_trunc_nums([25, 13])
    oc.set_trunc_nums([25, 13])
    g = np.load(os.path.join(G, 'zeropi_25_13.npz'))
    fl = g['params']
    res = cr.sweep_diag(10, loop_fluxes={lp: fl})
    ef = res.efreqs.cpu().numpy()
    assert np.max(np.abs(ef - g['efreqs_exact']) / np.abs(g['efreqs_exact'])) < 1e-9
    scl = _scales(oc, ol.set_flux, fl, 10, DEC)
    for d in DEC:
        got, ref = res.dec_rate(d, (1, 0)), g['dec_' + d]
        assert _rates_close(got, ref, scl[d]), (d, got, ref)
    for i in (0, 3, 8):
        ov = o.subspace_overlap(g['evecs_%d' % i], res.evecs[i].cpu().numpy().T, g['efreqs_exact'][i])
        assert 1 - ov < 1e-10
    _check(res, oc, ol.set_flux, fl[[1, 5]], 10, DEC, mels=[('capacitive', (0, 1), (0, 1)),
                                                            ('inductive', (0, 2), (0, 1)),
                                                            ('capacitive', (2, 3), (1, 2))], idx=[1, 5])


def test_zero_pi_reference_known_answers():
    """SQcircuit/tests/test_qubits.py::test_zero_pi through the Sweep class."""
    import sqcircuit_b200 as sq
    cr, lp, _, _ = cs.zero_pi()
    cr.set_trunc_nums([25, 13])
    ef, _ = sq.Sweep(cr, 2).sweepFlux([lp], [np.linspace(0, 0.5, 3)])
    assert np.allclose(ef[1] - ef[0], [0.69721321, 0.45628775, 0.0239571], rtol=1e-6)


def test_config2_zero_pi_bench_size_points():
    """N = 10100 (the bench truncation): golden eigenvalues from the reference Hamiltonian."""
    g = np.load(os.path.join(G, 'zeropi_100_51.npz'))
    cr, lp, _, _ = cs.zero_pi()
    cr.set_trunc_nums([100, 51])
    st = []
    res = cr.sweep_diag(10, loop_fluxes={lp: g['params']}, stats=st)
    ef = res.efreqs.cpu().numpy()
    assert np.max(np.abs(ef - g['efreqs_exact']) / np.abs(g['efreqs_exact'])) < 1e-9, (ef, st[-1])
    # size-independent properties: orthonormal vectors, H x = theta x through the operator
    X = res.evecs
    gram = X.conj() @ X.transpose(1, 2)
    assert float((gram - torch.eye(10, device=X.device)).abs().max()) < 1e-10
    # flux -> -flux and flux -> flux + 1 leave the spectrum unchanged
    res2 = cr.sweep_diag(10, loop_fluxes={lp: np.array([-0.37, 1.37])})
    assert np.max(np.abs(res2.efreqs.cpu().numpy() - ef[1][None, :]) / np.abs(ef[1])) < 1e-9


def test_hamiltonian_columns_at_bench_truncation_match_reference_golden():
    """N = 10100: columns of the reference's own Hamiltonian matrix (tests/golden/make_golden.py, flux 0.31)
    against the fused complex H.X kernel applied to unit vectors, and against the real-representation kernel
    applied to the time-reversal symmetric combinations e(i, n) + e(i, -n) and i (e(i, n) - e(i, -n))."""
    from sqcircuit_b200.kernels import BlockView
    from sqcircuit_b200.operators import HamiltonianOperator, RealHamiltonianOperator
    g = np.load(os.path.join(G, 'zeropi_100_51.npz'))
    cols, Hc = g['H_cols'], g['H_cols_at_0p31']
    cr, lp, _, _ = cs.zero_pi()
    cr.set_trunc_nums([100, 51])
    lp.set_flux(0.31)
    ops = cr._operators()
    lv, ng = cr._current_point()
    hop = HamiltonianOperator(ops, lv, ng)
    N, dev = ops.N, ops.device
    X = torch.zeros((1, len(cols), N), dtype=torch.complex128, device=dev)
    X[0, torch.arange(len(cols)), torch.as_tensor(cols, device=dev)] = 1.0
    Y = torch.empty_like(X)
    hop.apply(BlockView(X), BlockView(Y))
    scale = np.abs(Hc).max()
    assert np.abs(Y[0].cpu().numpy().T - Hc).max() <= 1e-12 * scale
    # real representation: T-symmetric combinations of the +-n column pairs
    assert ops.supports_real(ng)
    rop = RealHamiltonianOperator(ops, lv, ng, base=hop)
    inner, nc = ops.dims[1], (ops.dims[1] - 1) // 2
    pos = {int(c): j for j, c in enumerate(cols)}
    comb, ref = [], []
    for c in cols:
        i, n = int(c) // inner, int(c) % inner - nc
        if n > 0 and (i * inner + nc - n) in pos:
            a, b = pos[int(c)], pos[i * inner + nc - n]
            comb.append((a, b, 1.0, 1.0))
            ref.append(Hc[:, a] + Hc[:, b])
            comb.append((a, b, 1j, -1j))
            ref.append(1j * Hc[:, a] - 1j * Hc[:, b])
        elif n == 0:
            comb.append((pos[int(c)], pos[int(c)], 0.5, 0.5))
            ref.append(Hc[:, pos[int(c)]])
    Zc = torch.zeros((1, len(comb), N), dtype=torch.complex128, device=dev)
    for j, (a, b, ca, cb) in enumerate(comb):
        Zc[0, j, int(cols[a])] += ca
        Zc[0, j, int(cols[b])] += cb
    Xr = torch.zeros((1, len(comb), rop.Nc), dtype=torch.complex128, device=dev)
    ops.kern.complex_to_real(ops.sp, BlockView(Zc), BlockView(Xr))
    Yr = torch.empty_like(Xr)
    rop.apply(BlockView(Xr), BlockView(Yr))
    got = ops.real_to_complex(Yr)[0].cpu().numpy().T
    assert np.abs(got - np.stack(ref, axis=1)).max() <= 1e-12 * scale


def test_config1_fluxonium_1024_point_sweep():
    """BASELINE configs[1] at its named size: fluxonium, harmonic basis trunc = 100, 1024-point flux sweep,
    k = 10.  Every 64th point against the oracle (eigenvalues, subspaces, five channels); all points:
    agreement with the 17-point reference golden where the grids coincide, orthonormal vectors, the
    flux -> 1 - flux mirror symmetry of the transition frequencies and smoothness along the sweep."""
    cr, lp, oc, ol = cs.fluxonium()
    cr.set_trunc_nums([100])
    oc.set_trunc_nums([100])
    fl = np.arange(1025) / 1024.0
    res = cr.sweep_diag(10, loop_fluxes={lp: fl[:1024]})
    ef = res.efreqs.cpu().numpy()
    assert ef.shape == (1024, 10)
    decs = ['capacitive', 'inductive', 'quasiparticle', 'cc', 'flux']
    idx = list(range(0, 1024, 64)) + [1023]
    _check(res, oc, ol.set_flux, fl[idx], 10, decs, idx=idx)
    g = np.load(os.path.join(G, 'fluxonium_100.npz'))
    for j, f in enumerate(g['params'][:-1]):                 # 0, 1/16, ... 15/16 are grid points
        i = int(round(f * 1024))
        assert np.max(np.abs(ef[i] - g['efreqs_exact'][j]) / np.abs(g['efreqs_exact'][j])) < 1e-9
    X = res.evecs
    gram = X.conj() @ X.transpose(1, 2)
    assert float((gram - torch.eye(10, device=X.device)).abs().max()) < 1e-10
    # flux -> 1 - flux mirrors the transition frequencies (flux_dist='all' puts flux on the inductor too,
    # which shifts all levels of a point by a common flux-dependent constant)
    tr = ef[1:, 1:] - ef[1:, :1]
    assert np.max(np.abs(tr - tr[::-1])) < 1e-9 * np.abs(ef).max()
    assert np.all(np.diff(ef, axis=1) >= -1e-9)
    assert np.max(np.abs(np.diff(ef, axis=0))) < 0.2


def test_config2_bench_path_continuation_at_full_size():
    """The bench path itself: N = 10100, a flux sweep long enough for the four-level continuation
    (cold level, warm starts gathered from solved neighbours, banded two-level preconditioner with
    shared factorisations, planned restarts), symmetric end points included.  Sampled points against
    the oracle: eigenfrequencies 1e-9, subspace overlap 1 - 1e-10, the six decay rates 1e-8; every
    point through size-independent properties."""
    cr, lp, oc, ol = cs.zero_pi()
    cr.set_trunc_nums([100, 51])
    oc.set_trunc_nums([100, 51])
    fl = np.linspace(0.0, 0.5, 2049)
    st = []
    res = cr.sweep_diag(10, loop_fluxes={lp: fl}, stats=st)
    levels = [s_ for s_ in st if isinstance(s_, tuple) and s_[0] == 'level']
    assert len(levels) == 4, levels                      # 64 / 8 / 2 / 1 strides all populated
    idx = [0, 77, 1024, 1555, 2048]
    _check(res, oc, ol.set_flux, fl[idx], 10, DEC, idx=idx)
    X = res.evecs
    gram = X.conj() @ X.transpose(1, 2)
    assert float((gram - torch.eye(10, device=X.device)).abs().max()) < 1e-9
    ef = res.efreqs.cpu().numpy()
    assert np.all(np.diff(ef, axis=1) >= -1e-9)            # sorted spectra
    assert np.max(np.abs(np.diff(ef, axis=0))) < 0.01       # smooth along the sweep (no skipped level)


@pytest.mark.parametrize('trunc,n', [([25, 13], 9), ([40, 21], 70)])
def test_real_representation_matches_complex_solve(trunc, n):
    """Zero gate charge: the sweep is solved as a real symmetric problem (csrc/hx_real.cu) and the
    eigenvectors are rotated back to the reference's Fock x charge basis.  Same eigenvalues, subspaces
    and rates as the complex solve and as the oracle; the cold and the continuation path."""
    cr, lp, oc, ol = cs.zero_pi()
    cr.set_trunc_nums(trunc)
    oc.set_trunc_nums(trunc)
    fl = np.linspace(0.0, 0.5, n)
    sr, sc = [], []
    rr = cr.sweep_diag(10, loop_fluxes={lp: fl}, real=True, stats=sr)
    rc = cr.sweep_diag(10, loop_fluxes={lp: fl}, real=False, stats=sc)
    er, ec = rr.efreqs.cpu().numpy(), rc.efreqs.cpu().numpy()
    assert np.max(np.abs(er - ec) / np.abs(ec)) < 1e-10
    assert rr.evecs.shape == rc.evecs.shape
    gram = rr.evecs.conj() @ rr.evecs.transpose(1, 2)
    assert float((gram - torch.eye(10, device=gram.device)).abs().max()) < 1e-10
    # z(i, -n) = conj z(i, n): the eigenvectors come back time-reversal symmetric
    Z = rr.evecs.reshape(n, 10, cr.m[0], cr.m[1])
    assert float((Z.flip(3).conj() - Z).abs().max()) < 1e-14
    idx = [0, n // 3, n - 1]
    scl = _scales(oc, ol.set_flux, fl[idx], 10, DEC)
    for d in DEC:
        a, b = np.asarray(rr.dec_rate(d, (1, 0)))[idx], np.asarray(rc.dec_rate(d, (1, 0)))[idx]
        assert _rates_close(a, b, scl[d]), (d, a, b)
    _check(rr, oc, ol.set_flux, fl[idx], 10, DEC, idx=idx)


def test_zero_pi_gate_charge_sweep_with_continuation():
    """Gate-charge sweep of the 0-pi qubit at a size that goes through Davidson: the LC diagonal, the
    low-energy index set and the banded preconditioner now differ from point to point, and the
    continuation gathers warm starts along the charge axis.  Sampled points against the oracle."""
    cr, lp, oc, ol = cs.zero_pi()
    cr.set_trunc_nums([40, 21])
    oc.set_trunc_nums([40, 21])
    mode = next(iter(cr.charge_islands)) + 1
    ngs = np.linspace(0.0, 0.5, 81)
    cr.set_charge_offset(mode, 0.0)
    lp.set_flux(0.23)
    ol.set_flux(0.23)
    res = cr.sweep_diag(8, charge_offsets={mode: ngs})
    idx = [0, 33, 80]
    _check(res, oc, lambda g: oc.set_charge_offset(mode, float(g)), ngs[idx], 8, ['capacitive', 'charge', 'cc'], idx=idx)


def test_config3_transmon_resonator_transmon():
    cr, oc = cs.trt()
    cr.set_trunc_nums([6, 6, 6])
    oc.set_trunc_nums([6, 6, 6])
    g = np.load(os.path.join(G, 'trt_6_6_6.npz'))
    m1, m2 = [int(x) for x in g['charge_modes']]
    ngs = g['params']
    res = cr.sweep_diag(10, charge_offsets={m1: ngs[:, 0], m2: ngs[:, 1]})
    ef = res.efreqs.cpu().numpy()
    assert np.max(np.abs(ef - g['efreqs_exact']) / np.abs(g['efreqs_exact'])) < 1e-9

    def setter(p):
        oc.set_charge_offset(m1, float(p[0]))
        oc.set_charge_offset(m2, float(p[1]))
    _check(res, oc, setter, ngs, 10, ['capacitive', 'cc', 'charge'])


def _grid_check(kernels, trunc, k, n_side):
    """Gate charge of the first transmon x external flux of the SQUID transmon, n_side x n_side grid, against
    the oracle point by point (eigenvalues, subspaces, four channels incl. flux dephasing)."""
    cr, lp, oc, ol = cs.trt_squid(kernels)
    cr.set_trunc_nums(trunc)
    oc.set_trunc_nums(trunc)
    m1 = sorted(cr.charge_islands)[0] + 1
    a, b = np.meshgrid(np.linspace(0.0, 0.5, n_side), np.linspace(0.05, 0.45, n_side), indexing='ij')
    ng, fl = a.ravel(), b.ravel()
    res = cr.sweep_diag(k, charge_offsets={m1: ng}, loop_fluxes={lp: fl})

    def setter(p):
        oc.set_charge_offset(m1, float(p[0]))
        ol.set_flux(float(p[1]))
    _check(res, oc, setter, np.stack([ng, fl], axis=1), k, ['capacitive', 'cc', 'charge', 'flux'])
    return res


def test_config3_gate_charge_flux_grid():
    """BASELINE configs[3], "ng/flux grid sweep": both axes at once on the flux-tunable variant."""
    _grid_check(None, [6, 6, 6], 8, 3)


def test_config3_full_size_dim_1e5():
    """BASELINE configs[3] at its named size: transmon-resonator-transmon, N = 10 x 99 x 99 = 98010,
    a 2-D gate-charge grid.  One grid point against the oracle (ARPACK on the CSR matrix), all
    points through size-independent properties (orthonormality, residuals through the operator,
    invariance of the spectrum under ng -> ng + 1 up to the truncation edge is NOT exact, so use
    ng -> -ng with both offsets flipped: H(-ng) is the complex conjugate relabelling of H(ng))."""
    from sqcircuit_b200.kernels import BlockView
    cr, oc = cs.trt()
    trunc = [10, 50, 50]
    cr.set_trunc_nums(trunc)
    g = np.load(os.path.join(G, 'trt_6_6_6.npz'))
    m1, m2 = [int(x) for x in g['charge_modes']]
    a, b = np.meshgrid(np.linspace(0.05, 0.45, 3), np.linspace(0.1, 0.4, 3), indexing='ij')
    n1, n2 = a.ravel(), b.ravel()
    n1 = np.concatenate([n1, -n1[:2]])
    n2 = np.concatenate([n2, -n2[:2]])
    res = cr.sweep_diag(10, charge_offsets={m1: n1, m2: n2})
    X = res.evecs
    N = X.shape[2]
    assert N == 98010
    gram = X.conj() @ X.transpose(1, 2)
    assert float((gram - torch.eye(10, device=X.device)).abs().max()) < 1e-10
    ef = res.efreqs.cpu().numpy()
    assert np.max(np.abs(ef[9] - ef[0]) / np.abs(ef[0])) < 1e-9 and np.max(np.abs(ef[10] - ef[1]) / np.abs(ef[1])) < 1e-9
    assert float(res.resmax.max()) < 1e-9
    oc.set_trunc_nums(trunc)
    oc.set_charge_offset(m1, float(n1[4]))
    oc.set_charge_offset(m2, float(n2[4]))
    e, v = oc.diag(11, dense=False)
    assert np.max(np.abs(ef[4] - e[:10]) / np.abs(e[:10])) < 1e-9
    assert 1 - o.subspace_overlap(v[:, :10], X[4].cpu().numpy().T, e[:10], w_next=e[10]) < 1e-10


def test_config3_bench_grid_converges_without_the_orthogonalised_fallback(monkeypatch):
    """The 16 x 16 gate-charge x flux grid of bench.py's trt1e5 at full size (N = 98010): every continuation level
    must converge in the non-orthogonalised iteration (the orthogonalised safety net costs 3x the time and used
    to be hit on every warm level: nearly dependent warm-start blocks, see eigensolver.solve_davidson_gen)."""
    import sqcircuit_b200.circuit as C
    cr, lp, oc, ol = cs.trt_squid()
    cr.set_trunc_nums([10, 50, 50])
    m1 = sorted(cr.charge_islands)[0] + 1
    cell = np.arange(256)
    ng, fl = (cell // 16) / 32.0, ((cell % 16) + 0.5) / 32.0

    def no_fallback(*a, **k):
        raise AssertionError('orthogonalised fall-back taken')
    monkeypatch.setattr(C, 'solve_davidson', no_fallback)
    res = cr.sweep_diag(10, charge_offsets={m1: ng}, loop_fluxes={lp: fl})
    assert res.converged and float(res.resmax.max()) <= 1e-9 and res.iterations < 200
    X = res.evecs
    gram = X.conj() @ X.transpose(1, 2)
    assert float((gram - torch.eye(10, device=X.device)).abs().max()) < 1e-10


def test_other_reference_circuits():
    cr, lp, oc, ol = cs.coupled_fluxonium_transmon()
    cr.set_trunc_nums([30, 10])
    oc.set_trunc_nums([30, 10])
    fl = np.array([0.0, 0.2, 0.5])
    res = cr.sweep_diag(10, loop_fluxes={lp: fl})
    _check(res, oc, ol.set_flux, fl, 10, DEC)
    ef = res.efreqs[0].cpu().numpy()
    assert np.allclose(ef - ef[0], [0., 2.07145177, 2.34187898, 3.09228726, 4.92917697, 5.19354651,
                                    5.87412504, 7.49054673, 7.74403853, 8.2146904])
    cr, lp, oc, ol = cs.inductively_shunted()
    cr.set_trunc_nums([1, 9, 23])
    oc.set_trunc_nums([1, 9, 23])
    fl = np.array([0.0, 0.25, 0.5])
    res = cr.sweep_diag(6, loop_fluxes={lp: fl})
    _check(res, oc, ol.set_flux, fl, 6, ['capacitive', 'inductive', 'cc', 'flux'])
    cr, oc = cs.transmon()
    cr.set_trunc_nums([100])
    oc.set_trunc_nums([100])
    ngs = np.linspace(0, 0.5, 5)
    res = cr.sweep_diag(10, charge_offsets={1: ngs})
    _check(res, oc, lambda x: oc.set_charge_offset(1, float(x)), ngs, 10, ['capacitive', 'cc', 'charge'])


def test_no_cpu_fallback_without_library(monkeypatch):
    from sqcircuit_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libsqcircuit_b200.so')
    with pytest.raises(_lib.SqcLibraryError):
        _lib.load()
