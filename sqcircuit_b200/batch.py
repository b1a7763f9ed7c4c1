"""A batch of circuits of ONE graph with different element values, diagonalised in one batched GPU solve.

The reference has no batched entry point: a qubit-discovery loop (BASELINE configs[4]) builds one ``Circuit``
per candidate and calls ``diag`` on each (circuit.py:2016-2049), with ``torch_extensions.EigenSolver`` giving the
gradients one circuit, one element and one state at a time.  Here the candidates play the role of sweep points:
circuit b brings its own operator tables (LC diagonal, junction displacement amplitudes and energies, flux
distribution -- ``CircuitOperators`` with one Topology per circuit), the block Davidson iteration runs over all
of them in lock step, and in the PyTorch engine the eigenfrequencies come back as ONE tensor [B][k] wired to
every circuit's element tensors (``autograd.BatchEigenSolverFn``).
"""
from typing import List

import numpy as np
import torch

from . import units as unt
from .circuit import Circuit, SweepResult, _davidson, _cat_results, _points_per_chunk, default_kernels
from .eigensolver import DENSE_MAX, solve_dense
from .noise_ops import SweepNoise
from .operators import CircuitOperators, HamiltonianOperator
from .settings import get_optim_mode


class CircuitBatch:
    def __init__(self, circuits: List[Circuit], kernels=None):
        if len(circuits) == 0:
            raise ValueError('CircuitBatch needs at least one circuit')
        self.circuits = list(circuits)
        c0 = self.circuits[0]
        for c in self.circuits[1:]:
            if list(c.elements.keys()) != list(c0.elements.keys()) or c.m != c0.m or c.n != c0.n or \
                    [[type(e) for e in els] for els in c.elements.values()] != \
                    [[type(e) for e in els] for els in c0.elements.values()]:
                raise ValueError('the circuits of a batch must share one graph, element kinds and truncation')
        if len(c0.m) == 0:
            raise ValueError('Please specify the truncation number for each mode.')
        self._kern = kernels or c0._kern
        self._ops = None
        self._last = None

    def __len__(self):
        return len(self.circuits)

    def update(self):
        """Re-read the element values of every circuit (``Circuit.update`` of each)."""
        for c in self.circuits:
            c.update()
        self._ops = None
        self._last = None

    def _operators(self):
        if self._ops is None:
            kern = self._kern or default_kernels()
            self._ops = CircuitOperators(kern, [c._top for c in self.circuits], self.circuits[0].m)
        return self._ops

    def _points(self):
        pts = [c._current_point() for c in self.circuits]
        return np.concatenate([p[0] for p in pts], axis=0), np.concatenate([p[1] for p in pts], axis=0)

    def _solve(self, n_eig, tol=1e-9, maxit=300, stats=None, lowblock=512) -> SweepResult:
        ops = self._operators()
        kern = ops.kern
        lv, ng = self._points()
        hop = HamiltonianOperator(ops, lv, ng)
        B = len(self.circuits)
        if ops.N <= DENSE_MAX:
            res = solve_dense(kern, hop, B, ops.N, n_eig)
        else:
            # every circuit is a cold start (no neighbour along a sweep to continue from); the batch is cut
            # by free device memory like any sweep -- a chunk gets the tables of its own circuits
            chunk = _points_per_chunk(kern, ops.N, n_eig)
            parts = []
            for c0 in range(0, B, chunk):
                if chunk >= B:
                    op = hop
                else:
                    sub = CircuitOperators(kern, [c._top for c in self.circuits[c0:c0 + chunk]], self.circuits[0].m)
                    op = HamiltonianOperator(sub, lv[c0:c0 + chunk], ng[c0:c0 + chunk])
                parts.append(_davidson(kern, op, op.B, ops.N, n_eig, op.diag, tol=tol, maxit=maxit, stats=stats,
                                       lowblock_ns=lowblock))
            res = _cat_results(parts)
        if not res.converged:
            raise RuntimeError(f'batched Davidson did not converge after {res.iterations} iterations '
                               f'(worst relative residual {float(res.resmax.max()):.2e})')
        evecs = res.evecs if res.evecs.is_contiguous() else res.evecs.contiguous()
        out = SweepResult(res.evals / (2 * np.pi * unt.get_unit_freq()), res.evals, evecs, res.iterations,
                          res.converged, res.resmax)
        out.noise = SweepNoise(kern, ops, hop, res.evals, evecs, self.circuits[0])
        self._last = out
        return out

    def parameters(self):
        """[B][P] tensor of the optimisation parameters of every circuit (PyTorch engine)."""
        return torch.stack([torch.stack(c.parameters) for c in self.circuits])

    def diag(self, n_eig, **kw):
        """(efreqs [B][k], evecs [B][k][N]) of every circuit, device tensors.  In the PyTorch engine they carry
        the autograd graph to every circuit's element tensors; in the NumPy engine they are plain results."""
        if not isinstance(n_eig, int):
            raise ValueError('n_eig (number of eigenvalues) should be an integer.')
        if get_optim_mode() and all(len(c._parameters) for c in self.circuits):
            from .autograd import BatchEigenSolverFn
            evals, evecs = BatchEigenSolverFn.apply(self.parameters(), self, n_eig)
            return evals / (2 * np.pi * unt.get_unit_freq()), evecs
        res = self._solve(n_eig, **kw)
        return res.efreqs, res.evecs
