"""sqcircuit_b200 — B200-native drop-in for SQcircuit's diagonalisation hot path.

Same user-facing names as the reference package for this path
(``Capacitor, Inductor, Junction, Loop, Circuit, Sweep, set_engine, ...``); the numerics run in
hand-written sm_100a CUDA kernels behind a C ABI (``include/sqcircuit_b200.h``).
"""
from . import noise, units  # noqa: F401
from .circuit import Circuit, SweepResult, default_kernels  # noqa: F401
from .elements import Capacitor, Charge, Inductor, Junction, Loop  # noqa: F401
from .exceptions import CircuitStateError, ModeError  # noqa: F401
from .settings import get_device, get_engine, get_optim_mode, set_device, set_engine, set_optim_mode  # noqa: F401
from .sweep import Sweep  # noqa: F401
from .batch import CircuitBatch  # noqa: F401
from .autograd import set_max_eigenvector_grad  # noqa: F401
from .units import (get_unit_cap, get_unit_freq, get_unit_ind, get_unit_JJ, set_unit_cap,  # noqa: F401
                    set_unit_freq, set_unit_ind, set_unit_JJ)


def launch_count():
    """Kernels launched by the CUDA library in this process."""
    from . import _lib
    return int(_lib.load().sqc_launch_count())
