"""Autograd through diag on the GPU (same bodies as tests/test_autograd_host.py, CUDA kernels)."""
import pytest

pytestmark = pytest.mark.gpu

from test_autograd_host import eigvec_grad_check, grad_check  # noqa: E402


def test_autograd_transmon_gpu():
    grad_check('transmon', None)


def test_autograd_fluxonium_gpu():
    grad_check('fluxonium', None)


def test_autograd_zeropi_three_modes_gpu():
    grad_check('zeropi', None, n_eig=5, rtol=2e-2)


def test_autograd_eigenvector_path_gpu():
    eigvec_grad_check(None)


# ---- parity with the reference's own PyTorch engine (fixtures: tests/golden/gradients.npz) ----------------
from grad_golden_body import check_batch, check_single  # noqa: E402


@pytest.mark.parametrize('kind,decs', [
    ('fluxonium', ['capacitive', 'inductive', 'quasiparticle', 'cc', 'flux']),
    ('flux_transmon', ['capacitive', 'charge', 'cc', 'flux']),
    ('zeropi', ['inductive', 'charge', 'cc', 'flux'])])
def test_gradients_match_reference_torch_engine_gpu(kind, decs):
    check_single(None, kind, decs)


def test_batched_circuit_gradients_match_reference_gpu():
    check_batch(None)


def _config4_body(kernels, B, k, trunc, singles, fd_ids):
    import numpy as np
    import torch
    import sqcircuit_b200 as sq
    from grad_golden_body import G  # noqa: F401  (puts tests/golden on sys.path)
    from grad_cases import grad_circuits
    sq.set_engine('PyTorch')
    try:
        crs, elss = [], []
        for idx in range(B):
            cr, els, _, _ = grad_circuits(sq, 'discovery%d' % idx, kernels)
            cr.set_trunc_nums(trunc)
            crs.append(cr)
            elss.append(els)
        batch = sq.CircuitBatch(crs)
        ef, _ = batch.diag(k)
        assert ef.shape == (B, k)
        loss = (ef[:, 1] - ef[:, 0]).sum()
        loss.backward()
        names = list(elss[0])
        grad = np.array([[float(e[n].grad) for n in names] for e in elss])          # [B][P]
        assert np.all(np.isfinite(grad))
        ef0 = ef.detach().cpu().numpy()
        # one circuit at a time (EigenSolverFn) gives the same numbers
        for idx in singles:
            crs[idx].zero_grad()
            e1, _ = crs[idx].diag(k)
            assert np.max(np.abs(e1.detach().numpy() - ef0[idx])) < 1e-9 * np.abs(ef0[idx]).max()
            (e1[1] - e1[0]).backward()
            g1 = np.array([float(elss[idx][n].grad) for n in names])
            assert np.all(np.abs(g1 - grad[idx]) <= 1e-6 * np.abs(grad[idx]) + 1e-9 * np.abs(grad[idx]).max())
    finally:
        sq.set_engine('NumPy')
    # finite differences of the NumPy-engine batched forward for candidates ``fd_ids``, elements J01 and J30
    def f10(scale):
        out = []
        for idx in fd_ids:
            cr, els, _, _ = grad_circuits(sq, 'discovery%d' % idx, kernels, rg=False)
            for nm, sc in scale.items():
                el = els[nm]
                el.internal_value = el.internal_value * sc
            cr.update()
            cr.set_trunc_nums(trunc)
            out.append(cr)
        e, _ = sq.CircuitBatch(out).diag(k)
        e = e.cpu().numpy()
        return e[:, 1] - e[:, 0], [float(np.asarray(grad_circuits(sq, 'discovery%d' % i, kernels, rg=False)[1][nm]
                                                    .internal_value)) for i in fd_ids for nm in scale]
    d = 1e-5
    for nm in ('J01', 'J30'):      # junction energies enter H linearly: Hellmann-Feynman is exact at any truncation
        fp, x0 = f10({nm: 1 + d})
        fm, _ = f10({nm: 1 - d})
        for j, idx in enumerate(fd_ids):
            fd = (fp[j] - fm[j]) / (2 * d * x0[j])
            got = grad[idx][names.index(nm)]
            assert abs(got - fd) <= 2e-2 * abs(fd), (nm, idx, got, fd)


def test_config4_batch_of_512_four_mode_circuits():
    """BASELINE configs[4] at its named size: 512 candidates of the 4-mode discovery circuit (N = 8100) in one
    batched solve, gradient of a transition-frequency loss with respect to all 13 element values of every
    candidate.  Checked: the one-circuit-at-a-time path (EigenSolverFn) gives the same eigenfrequencies and
    gradients for three candidates; the batched gradient against central finite differences of the batched
    NumPy-engine forward for two candidates and two elements (2e-2: Hellmann-Feynman with fixed Fock operators,
    same tolerance as the reference's tests/test_grad.py)."""
    _config4_body(None, 512, 6, [10, 10, 5, 5], (0, 255, 511), (3, 400))
