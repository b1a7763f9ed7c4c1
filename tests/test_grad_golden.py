"""CPU: gradient parity with the reference's PyTorch engine, kernels emulated (see grad_golden_body.py)."""
import pytest

from emul_kernels import EmulKernels
from grad_golden_body import check_batch, check_single


@pytest.mark.parametrize('kind,decs', [
    ('fluxonium', ['capacitive', 'inductive', 'quasiparticle', 'cc', 'flux']),
    ('flux_transmon', ['capacitive', 'charge', 'cc', 'flux']),
    ('zeropi', ['inductive', 'charge', 'cc', 'flux'])])
def test_gradients_match_reference_torch_engine_cpu(kind, decs):
    check_single(EmulKernels(), kind, decs)


def test_batched_circuit_gradients_match_reference_cpu():
    check_batch(EmulKernels())


def test_config4_batch_small_cpu():
    """The body of the GPU test of BASELINE configs[4] (tests/test_gpu_autograd.py) at a small size."""
    from test_gpu_autograd import _config4_body
    _config4_body(EmulKernels(), 4, 5, [4, 4, 2, 2], (0, 3), (1, 2))
