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
