"""Engine selection and accuracy constants (reference settings.py)."""

ACC = {
    'sing_mode_detect': 1e-11,
    'Gram-Schmidt': 1e-6,
    'har_mode_elim': 1e-11,
}

SUPPORTED_ENGINES = ['NumPy', 'PyTorch']
_engine = {'name': 'numpy', 'device': 'cuda:0'}


def set_engine(eng):
    if eng.lower() not in [e.lower() for e in SUPPORTED_ENGINES]:
        raise ValueError("Invalid engine passed. Currently SQcircuit supports 'NumPy' and 'PyTorch'.")
    _engine['name'] = eng.lower()


def get_engine():
    return _engine['name']


def set_optim_mode(flag):
    set_engine('PyTorch' if flag else 'NumPy')


def get_optim_mode():
    return _engine['name'] == 'pytorch'


def set_device(device):
    """CUDA device used for the diagonalisation kernels (default cuda:0)."""
    _engine['device'] = str(device)


def get_device():
    return _engine['device']
