"""Circuit elements: Capacitor, Inductor, Junction, Loop, Charge.

Same constructor signatures, units and default loss models as the reference
(elements.py:108-726) so circuits are written the same way.  Values are kept in SI units
(farad, henry, rad/s for junctions, radians for loop fluxes).
"""
import math

import numpy as np
import torch
from scipy.special import kve

from . import units as unt
from .exceptions import raise_optim_error_if_needed, raise_unit_error
from .settings import get_optim_mode


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x


class _K0e(torch.autograd.Function):
    """exp(x) K0(x) on tensors, evaluated with scipy like the NumPy engine; d/dx = exp(x) (K0 - K1)
    (torch's own scaled_modified_bessel_k0 has no usable derivative; functions.py:141-156)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.as_tensor(kve(0, _np(x)), dtype=torch.float64)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x = _np(ctx.saved_tensors[0])
        return g * torch.as_tensor(kve(0, x) - kve(1, x), dtype=torch.float64)


class _LogK0(torch.autograd.Function):
    """log K0(x) = log(exp(x) K0(x)) - x on tensors; d/dx = -K1 / K0 (functions.py:159-186)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        xn = _np(x)
        return torch.as_tensor(np.log(kve(0, xn)) - xn, dtype=torch.float64)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x = _np(ctx.saved_tensors[0])
        return g * torch.as_tensor(-kve(1, x) / kve(0, x), dtype=torch.float64)


def _k0e(x):
    if isinstance(x, torch.Tensor):
        return _K0e.apply(x)
    return kve(0, x)


def _log_k0(x):
    if isinstance(x, torch.Tensor):
        return _LogK0.apply(x)
    return np.log(kve(0, x)) - x


def _exp(x):
    return torch.exp(x) if isinstance(x, torch.Tensor) else np.exp(x)


def _sqrt(x):
    return torch.sqrt(x) if isinstance(x, torch.Tensor) else np.sqrt(x)


def _abs(x):
    return torch.abs(x) if isinstance(x, torch.Tensor) else np.abs(x)


def _log(x):
    return torch.log(x) if isinstance(x, torch.Tensor) else np.log(x)


class _Valued:
    """Common storage: a float (NumPy engine) or a float64 tensor (PyTorch engine)."""
    _value = 0.0

    @property
    def internal_value(self):
        return self._value

    @internal_value.setter
    def internal_value(self, v):
        self._value = v

    @property
    def requires_grad(self):
        raise_optim_error_if_needed()
        return self._value.requires_grad

    @requires_grad.setter
    def requires_grad(self, flag):
        raise_optim_error_if_needed()
        self._value.requires_grad = flag

    @property
    def is_leaf(self):
        raise_optim_error_if_needed()
        return self._value.is_leaf

    @property
    def grad(self):
        raise_optim_error_if_needed()
        return self._value.grad

    @grad.setter
    def grad(self, g):
        raise_optim_error_if_needed()
        self._value.grad = g

    def _store(self, mean, error_pct):
        self.error = error_pct
        if error_pct:
            mean = float(np.random.normal(mean, mean * error_pct / 100))
        if get_optim_mode():
            self._value = torch.as_tensor(mean, dtype=torch.float64)
        else:
            self._value = float(mean)

    def si(self):
        """Plain python float of the SI value."""
        return float(_np(self._value))


class Capacitor(_Valued):
    value_unit = 'F'

    def __init__(self, value, unit=None, requires_grad=False, Q='default', error=0, id_str=None):
        unit = unt.get_unit_cap() if unit is None else unit
        self._check_unit_format(unit)
        self.set_value(value, unit, error)
        self.type = type(self)
        if requires_grad:
            self.requires_grad = True
        if isinstance(Q, str) and Q == 'default':
            self.Q = self._default_q_cap
        elif isinstance(Q, (int, float)):
            self.Q = lambda omega, _q=Q: _q
        else:
            self.Q = Q
        self.id_str = id_str if id_str is not None else f'C_{value}_{unit}'

    @staticmethod
    def _check_unit_format(u):
        if u is not None and u not in unt.freq_list and u not in unt.farad_list:
            raise_unit_error()

    def set_value(self, v, u='F', e=0.0):
        self._check_unit_format(u)
        if u in unt.farad_list:
            mean = v * unt.farad_list[u]
        else:
            mean = unt.e ** 2 / 2 / (v * unt.freq_list[u] * (2 * np.pi * unt.hbar))
        self._store(mean, e)

    def get_value(self, u='F'):
        if u in unt.farad_list:
            return self._value / unt.farad_list[u]
        if u in unt.freq_list:
            return unt.e ** 2 / 2 / self._value / (2 * np.pi * unt.hbar) / unt.freq_list[u]
        raise_unit_error()

    @staticmethod
    def _default_q_cap(omega):
        return 1e6 * (2 * np.pi * 6e9 / _abs(omega)) ** 0.7


class _TinyCap(Capacitor):
    def __init__(self):
        super().__init__(1e-20, 'F', Q=None)


class _HugeCap(Capacitor):
    def __init__(self):
        super().__init__(1e20, 'F', Q=None)


VerySmallCap = _TinyCap
VeryLargeCap = _HugeCap


class Inductor(_Valued):
    value_unit = 'H'

    def __init__(self, value, unit=None, requires_grad=False, cap=None, Q='default', error=0,
                 loops=None, id_str=None):
        unit = unt.get_unit_ind() if unit is None else unit
        self._check_unit_format(unit)
        self.set_value(value, unit, error)
        self.type = type(self)
        if requires_grad:
            self.requires_grad = True
        self.cap = cap if cap is not None else _TinyCap()
        self.loops = loops if loops is not None else []
        if isinstance(Q, str) and Q == 'default':
            self.Q = self._default_q_ind
        elif isinstance(Q, (int, float)):
            self.Q = lambda omega, T, _q=Q: _q
        else:
            self.Q = Q
        self.id_str = id_str if id_str is not None else f'L_{value}_{unit}'

    @staticmethod
    def _check_unit_format(u):
        if u is not None and u not in unt.freq_list and u not in unt.henry_list:
            raise_unit_error()

    def set_value(self, v, u='H', e=0.0):
        self._check_unit_format(u)
        if u in unt.henry_list:
            mean = v * unt.henry_list[u]
        else:
            mean = (unt.Phi0 / 2 / np.pi) ** 2 / (v * unt.freq_list[u] * (2 * np.pi * unt.hbar))
        self._store(mean, e)

    def get_value(self, u='H'):
        if u in unt.henry_list:
            return self._value / unt.henry_list[u]
        if u in unt.freq_list:
            return (unt.Phi0 / 2 / np.pi) ** 2 / self._value / (2 * np.pi * unt.hbar) / unt.freq_list[u]
        raise_unit_error()

    @staticmethod
    def _default_q_ind(omega, T):
        log_k0 = _log_k0

        def log_sinh(x):
            return x + _log(1 - _exp(-2 * x)) - math.log(2)
        alpha = unt.hbar * 2 * np.pi * 0.5e9 / (2 * unt.k_B * T)
        beta = unt.hbar * omega / (2 * unt.k_B * T)
        return 500e6 * _exp(log_k0(alpha) + log_sinh(alpha) - log_k0(beta) - log_sinh(beta))

    def get_cap_for_flux_dist(self, flux_dist):
        return {'all': self.cap.si(), 'junctions': 1e20, 'inductors': 1e-20}[flux_dist]


class Junction(_Valued):
    value_unit = 'Hz'

    def __init__(self, value, unit=None, requires_grad=False, cap=None, A=1e-7, x=3e-06,
                 delta=3.4e-4, Y='default', error=0, loops=None, id_str=None):
        unit = unt.get_unit_JJ() if unit is None else unit
        self._check_unit_format(unit)
        self.set_value(value, unit, error)
        self.type = type(self)
        self.A = A
        if requires_grad:
            self.requires_grad = True
        self.cap = cap if cap is not None else _TinyCap()
        self.loops = loops if loops is not None else []
        self.Y = self._default_y(delta, x) if (isinstance(Y, str) and Y == 'default') else Y
        self.id_str = id_str if id_str is not None else f'JJ_{value}_{unit}'

    @staticmethod
    def _check_unit_format(u):
        if u is not None and u not in unt.freq_list:
            raise_unit_error()

    def set_value(self, v, u='Hz', e=0.0):
        self._check_unit_format(u)
        self._store(v * unt.freq_list[u] * 2 * np.pi, e)

    def get_value(self, u='Hz'):
        if u in unt.freq_list:
            return self._value / unt.freq_list[u]
        raise_unit_error()

    def get_cap_for_flux_dist(self, flux_dist):
        return {'all': self.cap.si(), 'junctions': 1e-20, 'inductors': 1e20}[flux_dist]

    @staticmethod
    def _default_y(delta, x):
        gap = delta * 1.6e-19

        def y_qp(omega, T):
            alpha = unt.hbar * omega / (2 * unt.k_B * T)
            pref = np.sqrt(2 / np.pi) * (8 / gap / (unt.hbar * 2 * np.pi / unt.e ** 2))
            return (pref * (2 * gap / unt.hbar / omega) ** 1.5 * x * _sqrt(alpha) * _k0e(alpha)
                    * (1 - _exp(-2 * alpha)) / 2)
        return y_qp


class Loop(_Valued):
    """Inductive loop; the stored value is 2*pi*(external flux / Phi0)."""

    def __init__(self, value=0, A=1e-6, requires_grad=False, id_str=None):
        self.set_flux(value)
        if requires_grad:
            self.requires_grad = True
        self.A = A * 2 * np.pi
        self.indices = []
        self.k_mat = []
        self.id_str = id_str if id_str is not None else 'loop'

    def reset(self):
        self.indices, self.k_mat = [], []

    def value(self, random=False):
        if not random:
            return self._value
        return np.random.normal(_np(self._value), self.A, 1)[0]

    def set_flux(self, value):
        if get_optim_mode():
            value = torch.as_tensor(value, dtype=torch.float64)
        self._value = value * 2 * np.pi

    def set_noise(self, A):
        self.A = A

    def add_index(self, index):
        self.indices.append(index)

    def add_to_k_mat(self, w):
        self.k_mat.append(w)


class Charge:
    """Charge island: gate-charge offset and charge-noise amplitude."""

    def __init__(self, value=0, A=1e-4):
        self.chValue = value
        self.A = A

    def value(self, random=False):
        return self.chValue if not random else np.random.normal(self.chValue, self.A)

    def set_offset(self, value):
        self.chValue = value

    def set_noise(self, A):
        self.A = A

    setNoise = set_noise
