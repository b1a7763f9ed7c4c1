"""Minimal scipy-backed stand-in for the parts of QuTiP 5 that
SQcircuit/circuit.py, functions.py and torch_extensions.py touch.

TEST INFRASTRUCTURE ONLY (see README.md in this directory).
"""
import numbers

import numpy as np
import scipy.linalg
import scipy.sparse as sp

from .core.data import CSR, Dense, Dia  # noqa: F401


class _Core(dict):
    pass


class _Settings:
    core = _Core(auto_tidyup=True, auto_tidyup_atol=1e-12)


settings = _Settings()


def _tidy(m):
    """QuTiP's add_csr drops real/imag components below auto_tidyup_atol."""
    if not settings.core['auto_tidyup']:
        return m
    tol = settings.core['auto_tidyup_atol']
    m = m.tocsr()
    d = m.data
    re = np.where(np.abs(d.real) < tol, 0.0, d.real)
    im = np.where(np.abs(d.imag) < tol, 0.0, d.imag)
    m.data = re + 1j * im
    m.eliminate_zeros()
    return m


class Qobj:
    __array_ufunc__ = None
    __array_priority__ = 1000

    def __init__(self, arg=None, dims=None, copy=True):
        if isinstance(arg, Qobj):
            data = arg._data.copy()
            dims = dims if dims is not None else arg.dims
        elif sp.issparse(arg):
            data = arg.tocsr().astype(complex)
        else:
            a = np.asarray(arg, dtype=complex)
            if a.ndim == 0:
                a = a.reshape(1, 1)
            elif a.ndim == 1:
                a = a.reshape(-1, 1)
            data = sp.csr_matrix(a)
        self._data = data
        if dims is None:
            dims = [[data.shape[0]], [data.shape[1]]]
        self.dims = [list(dims[0]), list(dims[1])]
        self.dtype = CSR

    # -- representation ---------------------------------------------------
    @property
    def shape(self):
        return self._data.shape

    def to(self, _fmt):
        return self

    def copy(self):
        return Qobj(self._data.copy(), dims=self.dims)

    def full(self):
        return self._data.toarray()

    def data_as(self, fmt):
        if fmt == 'csr_matrix':
            return self._data.tocsr()
        if fmt == 'dia_matrix':
            return self._data.todia()
        raise ValueError(fmt)

    def dag(self):
        return Qobj(self._data.conj().T.tocsr(), dims=[self.dims[1], self.dims[0]])

    def contract(self):
        keep = [i for i, (a, b) in enumerate(zip(self.dims[0], self.dims[1]))
                if not (a == 1 and b == 1)] if len(self.dims[0]) == len(self.dims[1]) else None
        if keep is None:
            d0 = [x for x in self.dims[0] if x != 1] or [1]
            d1 = [x for x in self.dims[1] if x != 1] or [1]
        else:
            d0 = [self.dims[0][i] for i in keep] or [1]
            d1 = [self.dims[1][i] for i in keep] or [1]
        return Qobj(self._data.copy(), dims=[d0, d1])

    def expm(self):
        return Qobj(scipy.linalg.expm(self.full()), dims=self.dims)

    def __getitem__(self, ind):
        return self.full()[ind]

    # -- arithmetic -------------------------------------------------------
    @staticmethod
    def _is_scalar(x):
        return isinstance(x, numbers.Number) or (
            isinstance(x, np.ndarray) and x.ndim == 0) or isinstance(x, np.generic)

    def __add__(self, other):
        if self._is_scalar(other):
            if other == 0:
                return self.copy()
            return Qobj(self._data + complex(other) * sp.identity(self.shape[0], format='csr'),
                        dims=self.dims)
        if isinstance(other, Qobj):
            return Qobj(_tidy(self._data + other._data), dims=self.dims)
        return NotImplemented

    __radd__ = __add__

    def __sub__(self, other):
        return self + (-other)

    def __rsub__(self, other):
        return (-self) + other

    def __neg__(self):
        return Qobj(-self._data, dims=self.dims)

    def __mul__(self, other):
        if isinstance(other, Qobj):
            return self.__matmul__(other)
        if self._is_scalar(other):
            return Qobj(self._data * complex(other), dims=self.dims)
        return NotImplemented

    def __rmul__(self, other):
        if self._is_scalar(other):
            return Qobj(self._data * complex(other), dims=self.dims)
        return NotImplemented

    def __matmul__(self, other):
        if not isinstance(other, Qobj):
            return NotImplemented
        out = self._data @ other._data
        if out.shape == (1, 1):
            return complex(out.toarray()[0, 0])
        return Qobj(out, dims=[self.dims[0], other.dims[1]])

    def __truediv__(self, other):
        return Qobj(self._data / complex(other), dims=self.dims)

    def __pow__(self, n):
        out = self
        for _ in range(int(n) - 1):
            out = out * self
        return out

    def __eq__(self, other):
        if self is other:
            return True
        if not isinstance(other, Qobj) or self.dims != other.dims:
            return False
        return abs(self._data - other._data).max() < 1e-14

    __hash__ = object.__hash__


# -- operator factories -------------------------------------------------------

def qeye(n):
    n = int(n)
    return Qobj(sp.identity(n, format='csr', dtype=complex))


identity = qeye


def destroy(n, offset=0):
    n = int(n)
    if n < 2:
        raise ValueError("QuTiP 5 does not support destroy(N) with N < 2")
    d = np.sqrt(np.arange(offset + 1, offset + n, dtype=complex))
    return Qobj(sp.diags([d], [1], shape=(n, n), format='csr'))


def create(n, offset=0):
    return destroy(n, offset).dag()


def num(n, offset=0):
    n = int(n)
    return Qobj(sp.diags([np.arange(offset, offset + n, dtype=complex)], [0], format='csr'))


def charge(Nmax, Nmin=None, frac=1):
    if Nmin is None:
        Nmin = -Nmax
    d = np.arange(Nmin, Nmax + 1, dtype=float) * frac
    return Qobj(sp.diags([d.astype(complex)], [0], format='csr'))


def displace(n, alpha, offset=0):
    a = destroy(n, offset)
    return (alpha * a.dag() - np.conj(alpha) * a).expm()


def tensor(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = args[0]
    out = None
    d0, d1 = [], []
    for q in args:
        out = q._data if out is None else sp.kron(out, q._data, format='csr')
        d0 += q.dims[0]
        d1 += q.dims[1]
    return Qobj(out, dims=[d0, d1])
