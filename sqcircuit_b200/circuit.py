"""Circuit: the reference's user-facing class for the diagonalisation path, backed by the
batched GPU solver.

API parity (reference circuit.py): ``Circuit(elements, flux_dist)``, ``set_trunc_nums``,
``set_charge_offset``, ``set_charge_noise``, ``trunc_nums``, ``truncate_circuit``, ``diag``,
``efreqs``, ``evecs``, ``hamiltonian``, ``matrix_elements``, ``dec_rate``,
``check_convergence``, ``update`` and the transformation attributes (``S, R, C, L, W, B,
wTrans, cInvTrans, lTrans, omega, loops``).  New: ``sweep_diag`` solves all points of a flux /
gate-charge sweep in one batched call (what ``Sweep.sweepFlux`` loops over in the reference).
"""
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import units as unt
from .eigensolver import DENSE_MAX, solve_davidson, solve_davidson_gen, solve_dense

# iteration used by sweep_diag: 'gen' (non-orthogonalised basis, generalised Rayleigh-Ritz; 3 passes
# over the basis per iteration) or 'orth' (CGS2 + SVQB; 7 passes)
SOLVER = {'kind': 'gen'}
# the Davidson kernels work on blocks of <= 16 vectors, two of which are guard vectors
MAX_EIG_DAVIDSON = 14


class _RealRepresentationFailed(RuntimeError):
    """The solve in the real representation did not converge; the caller redoes it with complex vectors."""


def _davidson(*a, **kw):
    if SOLVER['kind'] == 'gen':
        kw.pop('impl', None)
        res = solve_davidson_gen(*a, **kw)
        if res.converged:
            return res
        if getattr(a[1], 'real', False):
            # the orthogonalised iteration below works on complex vectors only
            raise _RealRepresentationFailed()
        # Safety net: the non-orthogonalised iteration can stagnate a little above the tolerance
        # when expansion vectors keep arriving almost inside the span of the basis (seen on
        # charge-dominated Hamiltonians with ||H|| >> the wanted eigenvalues).  The orthogonalised
        # iteration (CGS2 + SVQB: seven basis passes instead of three) does not have that floor.
        kw = {k: v for k, v in kw.items() if k not in ('drop', 'cold_gate', 'expand')}
        kw['lowblock_ns'] = 0          # plain diagonal preconditioner: slower, nothing to go wrong
        a[0].release_workspace()       # the fallback allocates its own (larger) basis buffers
    return solve_davidson(*a, **kw)


def _solver_sizes(k, N, block_size=None, max_basis=None):
    """(p, mmax) exactly as the solvers choose them (eigensolver.solve_davidson_gen)."""
    p = block_size if block_size is not None else min(16, max(k + 2, (k * 6 + 4) // 5))
    p = min(p, 16, N)
    mmax = max_basis if max_basis is not None else 3 * p + 4
    return p, min(mmax, 64, N)


def _points_per_chunk(kern, Nv, k, block_size=None, max_basis=None, fraction=0.8):
    """Sweep points one batched Davidson call may take: the basis buffers V and W = H V (2 (mmax + p) vectors
    per point) plus the gathered Ritz block, the returned vectors and the Gram / coefficient matrices have to
    fit into ``fraction`` of what the device has free right now (kern.free_bytes)."""
    p, mmax = _solver_sizes(k, Nv, block_size, max_basis)
    per_point = 16 * Nv * (2 * (mmax + p) + 2 * p + k) + 16 * (3 * mmax * mmax + 2 * 16 * mmax) + 4096
    return max(1, int(fraction * kern.free_bytes() // per_point))


def _cat_results(parts):
    """Concatenate the EigResults of consecutive chunks of one level."""
    from .eigensolver import EigResult
    if len(parts) == 1:
        return parts[0]
    cat = lambda name: torch.cat([getattr(r, name) for r in parts], dim=0)
    return EigResult(cat('evals'), cat('evecs'), max(r.iterations for r in parts), cat('resmax'),
                     all(r.converged for r in parts),
                     ritz_block=cat('ritz_block') if parts[0].ritz_block is not None else None)
from .elements import Capacitor, Charge, Inductor, Junction, Loop
from .exceptions import CircuitStateError
from .noise_ops import SweepNoise
from .operators import CircuitOperators, HamiltonianOperator, RealHamiltonianOperator
from .settings import get_device, get_optim_mode
from .topology import assemble, transform

_KERNELS = {}


def default_kernels(device=None):
    """The CUDA kernel backend for ``device`` (raises if the extension is not built)."""
    from .kernels import Kernels
    device = device or get_device()
    if device not in _KERNELS:
        _KERNELS[device] = Kernels(device)
    return _KERNELS[device]


class Ket(np.ndarray):
    """Column state vector with the few Qobj-like accessors the reference's users touch."""

    def __new__(cls, data, dims=None):
        obj = np.asarray(data, dtype=complex).reshape(-1, 1).view(cls)
        obj.dims = dims
        return obj

    def __array_finalize__(self, obj):
        self.dims = getattr(obj, 'dims', None)

    def full(self):
        return np.asarray(self)

    def dag(self):
        return np.asarray(self).conj().T


@dataclass
class SweepResult:
    efreqs: torch.Tensor        # [B][k] in the package frequency unit (GHz by default)
    evals_rad: torch.Tensor     # [B][k] rad/s
    evecs: torch.Tensor         # [B][k][N]
    iterations: int
    converged: bool
    resmax: torch.Tensor
    noise: SweepNoise = None

    def matrix_elements(self, ctype, nodes, states):
        return self.noise.matrix_elements(ctype, nodes, states)

    def dec_rate(self, dec_type, states, total=True):
        return self.noise.dec_rate(dec_type, states, total)


class Circuit:
    def __init__(self, elements, flux_dist='junctions', kernels=None):
        self.elements = OrderedDict((k, elements[k]) for k in elements.keys())
        if flux_dist not in ('junctions', 'inductors', 'all'):
            raise ValueError("flux_dist option must either be 'junctions', 'inductors', or 'all'.")
        self.flux_dist = flux_dist
        self._kern = kernels
        self._analyse()
        self.charge_islands = {i: Charge() for i in range(self.n) if self._is_charge_mode(i)}
        self.m: List[int] = []
        self.ms: List[int] = []
        self._ops = None
        self._efreqs = np.array([])
        self._evecs = []
        self._last = None

    # ---- optimisation parameters (circuit.py:498-591) ---------------------------------------------
    def _collect_parameters(self):
        from .exceptions import raise_optim_error_if_needed
        self._parameters = OrderedDict()
        if not get_optim_mode():
            return

        def add(el):
            v = el.internal_value
            if isinstance(v, torch.Tensor) and (v.requires_grad or not v.is_leaf) and el not in self._parameters:
                self._parameters[el] = v
        for edge, els in self.elements.items():
            for el in els:
                if hasattr(el, 'loops'):
                    add(el.cap)
                    for lp in el.loops:
                        add(lp)
                add(el)

    @property
    def parameters(self):
        from .exceptions import raise_optim_error_if_needed
        raise_optim_error_if_needed()
        return list(self._parameters.values())

    @property
    def parameters_elems(self):
        from .exceptions import raise_optim_error_if_needed
        raise_optim_error_if_needed()
        return list(self._parameters.keys())

    @property
    def parameters_grad(self):
        grads = [v.grad for v in self.parameters]
        if any(g is None for g in grads):
            return grads
        return torch.stack(grads).detach().clone()

    def zero_grad(self):
        for v in self.parameters:
            v.grad = None

    # ---- topology ---------------------------------------------------------------------------
    def _analyse(self):
        top = assemble(self.elements, self.flux_dist)
        top = transform(top, remove_uncoupled=not get_optim_mode())
        self._top = top
        self.loops = top.loops
        self.C, self.L, self.W, self.B = top.C, top.L, top.W, top.B
        self.S, self.R = top.S, top.R
        self.cInvTrans, self.lTrans, self.wTrans = top.cInvTrans, top.lTrans, top.wTrans
        self.omega = top.omega
        self.n = top.n
        self.num_jun_without_ind = top.n_jj_without_ind
        self.partial_mats = top.partial_mats
        self.elem_keys = {Inductor: top.ind_keys, Junction: top.jj_keys}
        self._collect_parameters()

    def update(self):
        """Re-read element values after in-place changes (circuit.py:397-440)."""
        self._analyse()
        self._ops = None
        self._efreqs = np.array([])
        self._evecs = []
        self._last = None

    def _is_charge_mode(self, i):
        if i >= self.n:
            raise ValueError(f'The circuit only has {self.n} modes!')
        return self.omega[i] == 0

    # ---- truncation / offsets (circuit.py:1340-1422) ------------------------------------------
    def set_trunc_nums(self, nums):
        if not isinstance(nums, list):
            raise ValueError('The input must be be a python list')
        if len(nums) != self.n:
            raise ValueError('The number of modes (length of the input) must be equal to the number of nodes.')
        self.m = [2 * nums[i] - 1 if self._is_charge_mode(i) else nums[i] for i in range(self.n)]
        self.ms = [x for x in self.m if x != 1]
        self._ops = None

    @property
    def trunc_nums(self):
        return [int((self.m[i] + 1) / 2) if self._is_charge_mode(i) else self.m[i] for i in range(self.n)]

    def truncate_circuit(self, K):
        if not isinstance(K, int):
            raise ValueError('The total truncation number must be an integer.')
        if K < 1:
            raise ValueError('The total truncation number must be >=1.')
        avg = K ** (1 / len(self.omega))
        nums = [int(np.floor(avg)) if self.omega[i] != 0 else int(np.floor((avg + 1) / 2))
                for i in range(len(self.omega))]
        self.set_trunc_nums(nums)
        return nums

    def set_charge_offset(self, mode, ng):
        if not isinstance(mode, int):
            raise ValueError('Mode number should be an integer.')
        if mode - 1 not in self.charge_islands:
            raise ValueError('The specified mode is not a charge mode.')
        self.charge_islands[mode - 1].set_offset(ng)

    def set_charge_noise(self, mode, amp):
        if not isinstance(mode, int):
            raise ValueError('The mode number should be an integer.')
        if mode - 1 not in self.charge_islands:
            raise ValueError('The specified mode is not a charge mode.')
        self.charge_islands[mode - 1].set_noise(amp)

    # ---- GPU plan ---------------------------------------------------------------------------------
    def _operators(self):
        if len(self.m) == 0:
            raise CircuitStateError('Please specify the truncation number for each mode.')
        if self._ops is None:
            kern = self._kern or default_kernels()
            self._ops = CircuitOperators(kern, self._top, self.m)
        return self._ops

    def _current_point(self):
        lv = np.array([[float(np.asarray(_to_float(lp.value()))) for lp in self.loops]], dtype=float) \
            if self.loops else np.zeros((1, 0))
        ng = np.array([[float(self.charge_islands[i].value()) if i in self.charge_islands else 0.0
                        for i in range(self.n)]], dtype=float)
        return lv, ng

    def sweep_diag(self, n_eig, loop_fluxes: Optional[Dict] = None,
                   charge_offsets: Optional[Dict[int, Sequence[float]]] = None, tol=1e-9,
                   block_size=None, max_basis=None, maxit=300, stats=None,
                   warm_start=True, coarse_stride=8, lowblock=512, real=None) -> SweepResult:
        """Diagonalise at every point of a sweep in one batched GPU solve.

        loop_fluxes: {Loop: array[B]} external fluxes in units of Phi0 (as Loop.set_flux);
        charge_offsets: {mode (1-based): array[B]} gate charges (as set_charge_offset).
        Parameters not listed keep their current value.  Arrays must share the length B.
        real: None (default) uses the real representation whenever the circuit admits it, False forces
        complex vectors.
        """
        if not isinstance(n_eig, int):
            raise ValueError('n_eig (number of eigenvalues) should be an integer.')
        ops = self._operators()
        lv0, ng0 = self._current_point()
        sizes = [len(v) for v in (loop_fluxes or {}).values()] + [len(v) for v in (charge_offsets or {}).values()]
        B = sizes[0] if sizes else 1
        if any(s != B for s in sizes):
            raise ValueError('all sweep arrays must have the same length')
        lv = np.repeat(lv0, B, axis=0)
        for lp, vals in (loop_fluxes or {}).items():
            lv[:, self.loops.index(lp)] = 2 * np.pi * np.asarray(vals, dtype=float)
        if charge_offsets:
            ng = np.repeat(ng0, B, axis=0)
            for mode, vals in charge_offsets.items():
                if mode - 1 not in self.charge_islands:
                    raise ValueError('The specified mode is not a charge mode.')
                ng[:, mode - 1] = np.asarray(vals, dtype=float)
        else:
            ng = ng0
        hop = HamiltonianOperator(ops, lv, ng)
        kern = ops.kern
        if isinstance(coarse_stride, (tuple, list)):
            if len(coarse_stride) < 1:
                raise ValueError('coarse_stride: at least one stride is needed')
            min_warm = 4 * int(coarse_stride[-2] if len(coarse_stride) >= 2 else coarse_stride[-1])
        else:
            min_warm = 4 * coarse_stride
        if ops.N <= DENSE_MAX:
            res = solve_dense(kern, hop, B, ops.N, n_eig)
        else:
            if n_eig > MAX_EIG_DAVIDSON:
                raise ValueError(f'n_eig = {n_eig}: the batched block Davidson iteration (Hilbert dimension > '
                                 f'{DENSE_MAX}) finds at most {MAX_EIG_DAVIDSON} eigenpairs per call')
            # zero gate charge on a harmonic x charge circuit: H commutes with a time reversal and is
            # solved as a real symmetric problem (half the bytes and flops, csrc/hx_real.cu)
            use_real = real if real is not None else ops.supports_real(ng)
            res = None
            if use_real:
                try:
                    res = self._solve(ops, hop, lv, ng, n_eig, tol, block_size, max_basis, maxit, stats,
                                      coarse_stride, lowblock, warm_start and B >= min_warm, True)
                except _RealRepresentationFailed:
                    res = None
            if res is None:
                res = self._solve(ops, hop, lv, ng, n_eig, tol, block_size, max_basis, maxit, stats,
                                  coarse_stride, lowblock, warm_start and B >= min_warm, False)
        if not res.converged:
            why = 'iteration limit' if res.iterations >= maxit else 'stagnation guard'
            raise RuntimeError(f'batched Davidson did not converge: stopped after {res.iterations} lock-step '
                               f'iterations ({why}, limit {maxit}), worst relative residual '
                               f'{float(res.resmax.max()):.2e}')
        evecs = res.evecs if res.evecs.is_contiguous() else res.evecs.contiguous()
        out = SweepResult(res.evals / (2 * np.pi * unt.get_unit_freq()), res.evals, evecs,
                          res.iterations, res.converged, res.resmax)
        out.noise = SweepNoise(kern, ops, hop, res.evals, evecs, self)
        return out

    def _solve(self, ops, hop, lv, ng, n_eig, tol, block_size, max_basis, maxit, stats, stride, lowblock,
               continuation, real):
        """One batched solve of all points, with complex vectors or in the real representation (the
        eigenvectors are converted back to the reference's Fock x charge basis at the end)."""
        from .eigensolver import EigResult
        kern = ops.kern
        if continuation:
            res = self._solve_continuation(ops, lv, ng, n_eig, tol, block_size, max_basis, maxit, stats,
                                           stride, lowblock, real)
        else:
            B = lv.shape[0]
            Nv = (ops.N + 1) // 2 if real else ops.N
            chunk = _points_per_chunk(kern, Nv, n_eig, block_size, max_basis)
            parts = []
            for c0 in range(0, B, chunk):
                if chunk >= B:
                    op = RealHamiltonianOperator(ops, lv, ng, base=hop) if real else hop
                else:
                    cls = RealHamiltonianOperator if real else HamiltonianOperator
                    op = cls(ops, lv[c0:c0 + chunk], ng if ng.shape[0] == 1 else ng[c0:c0 + chunk])
                parts.append(_davidson(kern, op, op.B, Nv, n_eig, op.diag, p=block_size, mmax=max_basis,
                                       tol=tol, maxit=maxit, stats=stats, lowblock_ns=lowblock))
                if not parts[-1].converged:
                    break
            res = _cat_results(parts)
        if real and res.converged:
            out = torch.empty(res.evecs.shape[:2] + (ops.N,), dtype=torch.complex128, device=ops.device)
            step = max(1, int(2 ** 31 // (16 * ops.N * res.evecs.shape[1])))
            for c0 in range(0, out.shape[0], step):
                ops.real_to_complex(res.evecs[c0:c0 + step], out=out[c0:c0 + step])
            res = EigResult(res.evals, out, res.iterations, res.resmax, res.converged)
        return res

    def _solve_continuation(self, ops, lv, ng, n_eig, tol, block_size, max_basis, maxit, stats, stride,
                            lowblock=512, real=False):
        """Sweep continuation: eigenvectors vary smoothly along a sweep, so a coarse subset of the
        points is solved from a cold start and every other point starts from the union of the Ritz
        blocks of its two nearest already-solved neighbours (Rayleigh-Ritz in that subspace is
        second order accurate in the parameter distance).  Levels: every stride^2-th point cold,
        then every stride-th point, then the rest."""
        from scipy.spatial import cKDTree
        from .eigensolver import EigResult, RitzPairStart
        kern = ops.kern
        B = lv.shape[0]
        ngb = np.broadcast_to(ng, (B, ng.shape[1]))
        prm = np.concatenate([lv / (2 * np.pi), ngb], axis=1)
        span = prm.max(axis=0) - prm.min(axis=0)
        vary = span > 0

        Nv = (ops.N + 1) // 2 if real else ops.N       # complex elements per stored vector

        def sub(idx):
            cls = RealHamiltonianOperator if real else HamiltonianOperator
            return cls(ops, lv[idx], ng if ng.shape[0] == 1 else ng[idx])

        def solve(idx, x0, st):
            hop = sub(idx)
            return _davidson(kern, hop, len(idx), Nv, n_eig, hop.diag, p=block_size,
                             mmax=max_basis, tol=tol, maxit=maxit, x0=x0, stats=st,
                             lowblock_ns=lowblock)

        def solve_cold(idx, st):
            chunk = _points_per_chunk(kern, Nv, n_eig, block_size, max_basis)
            return _cat_results([solve(idx[c0:c0 + chunk], None, st) for c0 in range(0, len(idx), chunk)])
        if not vary.any():
            return solve_cold(np.arange(B), stats)
        q = prm[:, vary] / span[vary]
        order = np.lexsort(q.T[::-1])
        # level sets: strides stride^2, stride, 1 over the sorted sweep (coarsest kept >= 32 points)
        # levels: every stride^2-th point cold, then every stride-th, every 2nd, the rest (measured on the
        # 0-pi bench: 64/8/2/1 is 3-4 % faster than 64/8/1; the last level starts from both direct
        # neighbours); an explicit tuple of strides overrides
        if isinstance(stride, (tuple, list)):
            strides = [int(x) for x in stride]
        else:
            strides = [stride * stride, stride, 2, 1] if stride > 2 else [stride * stride, stride, 1]
        while len(strides) > 1 and len(order[::strides[0]]) < 32:
            strides.pop(0)
        assigned = np.zeros(B, dtype=bool)
        levels = []
        for sd in strides:
            idx = order[::sd]
            if sd != 1 and order[-1] not in idx:
                idx = np.append(idx, order[-1])
            idx = idx[~assigned[idx]]
            assigned[idx] = True
            if len(idx):
                levels.append(idx)

        dev = ops.device
        evals = evecs = resmax = ritz = None
        solved = np.zeros(0, dtype=int)
        iters, conv = 0, True
        for li, idx in enumerate(levels):
            st = [] if stats is not None else None
            if li == 0:
                r = solve_cold(idx, st)
                p = r.ritz_block.shape[1]
                k = r.evals.shape[1]
                evals = torch.empty((B, k), dtype=torch.float64, device=dev)
                evecs = torch.empty((B, k, Nv), dtype=torch.complex128, device=dev)
                resmax = torch.empty((B,), dtype=torch.float64, device=dev)
                # Ritz blocks (all p vectors) of solved points, kept until the last level starts
                ritz = {int(i): j for j, i in enumerate(idx)}
                ritz_store = [r.ritz_block]
                ii = kern.upload(idx)
                evals[ii], evecs[ii], resmax[ii] = r.evals, r.evecs, r.resmax
                iters, conv = r.iterations, r.converged
            else:
                # two nearest solved neighbours of every point of this level
                # (k-d tree: the dense distance matrix costs ~100 ms of host time at 2000 x 2000)
                if len(solved) == 1:
                    nb = np.zeros((len(idx), 2), dtype=int)
                else:
                    nb = cKDTree(q[solved]).query(q[idx], k=2)[1]
                store = torch.cat(ritz_store, dim=0) if len(ritz_store) > 1 else ritz_store[0]
                ritz_store = [store]
                pos = np.array([ritz[int(i)] for i in solved])
                chunk = min(len(idx), _points_per_chunk(kern, Nv, n_eig, block_size, max_basis))
                last = li == len(levels) - 1
                new_blocks = []
                for c0 in range(0, len(idx), chunk):
                    sel = slice(c0, c0 + chunk)
                    x0 = RitzPairStart(store, kern.upload(pos[nb[sel, 0]]), kern.upload(pos[nb[sel, 1]]))
                    r = solve(idx[sel], x0, st)
                    del x0
                    ii = kern.upload(idx[sel])
                    evals[ii], evecs[ii], resmax[ii] = r.evals, r.evecs, r.resmax
                    iters += r.iterations
                    conv = conv and r.converged
                    if not last:
                        new_blocks.append(r.ritz_block)
                if not last:
                    base = store.shape[0]
                    for j, i in enumerate(idx):
                        ritz[int(i)] = base + j
                    ritz_store.extend(new_blocks)
            solved = np.concatenate([solved, idx])
            if stats is not None:
                stats.append(('level', li, len(idx), st))
        return EigResult(evals, evecs, iters, resmax, conv)

    # ---- single-point API of the reference -----------------------------------------------------------
    def diag(self, n_eig):
        """circuit.py:2016-2049."""
        if len(self.m) == 0:
            raise CircuitStateError('Please specify the truncation number for each mode.')
        if not isinstance(n_eig, int):
            raise ValueError('n_eig (number of eigenvalues) should be an integer.')
        if get_optim_mode():
            from .autograd import EigenSolverFn
            params = torch.stack(self.parameters) if self._parameters else torch.tensor([])
            sol = EigenSolverFn.apply(params, self, n_eig)
            self._efreqs = torch.real(sol[:, 0])
            self._evecs = sol[:, 1:]
            return self._efreqs / (2 * np.pi * unt.get_unit_freq()), self._evecs
        res = self.sweep_diag(n_eig)
        self._last = res
        ev = res.evecs[0]
        dims = [self.ms, len(self.ms) * [1]]
        self._efreqs = res.evals_rad[0].cpu().numpy()
        vecs = ev.cpu().numpy()
        self._evecs = [Ket(vecs[i], dims) for i in range(vecs.shape[0])]
        return self._efreqs / (2 * np.pi * unt.get_unit_freq()), self._evecs

    @property
    def efreqs(self):
        if len(self._efreqs) == 0:
            raise CircuitStateError('Please diagonalize the circuit first.')
        return self._efreqs / (2 * np.pi * unt.get_unit_freq())

    @property
    def evecs(self):
        if len(self._evecs) == 0:
            raise CircuitStateError('Please diagonalize the circuit first.')
        return self._evecs

    def hamiltonian(self):
        """Dense Hamiltonian matrix (rad/s) at the current parameter point, produced by applying
        the matrix-free operator to the identity (debug / parity helper)."""
        ops = self._operators()
        lv, ng = self._current_point()
        hop = HamiltonianOperator(ops, lv, ng)
        from .kernels import BlockView
        N = ops.N
        eye = torch.eye(N, dtype=torch.complex128, device=ops.device).unsqueeze(0).contiguous()
        out = torch.empty_like(eye)
        hop.apply(BlockView(eye), BlockView(out))
        return out[0].cpu().numpy().T

    def matrix_elements(self, ctype, nodes, states):
        if self._last is None:
            raise CircuitStateError('Please diagonalize the circuit first.')
        if ctype not in ('capacitive', 'inductive'):
            raise ValueError("The coupling type must be either 'capacitive' or 'inductive'.")
        if get_optim_mode():
            return self._matrix_elements_torch(ctype, nodes, states)
        return complex(self._last.matrix_elements(ctype, nodes, states)[0])

    def dec_rate(self, dec_type, states, total=True):
        if self._last is None:
            raise CircuitStateError('Please diagonalize the circuit first.')
        if get_optim_mode():
            return self._dec_rate_torch(dec_type, states, total)
        return float(self._last.dec_rate(dec_type, states, total)[0])

    # ---- PyTorch engine: rates as differentiable tensors (circuit.py:2529-2647 in optim mode) -------------
    def _inner_torch(self, states, apply_fn):
        from .autograd import operator_inner_product
        return operator_inner_product(self._evecs[states[0]], apply_fn, self._evecs[states[1]], self._last.noise.dev)

    def _matrix_elements_torch(self, ctype, nodes, states):
        """<m| coupling_op |n>: the node-to-mode coefficients K are torch expressions of the element values
        (K = C^-1 R depends on the capacitances, circuit.py:2356-2393), the mode operators are constants."""
        nz = self._last.noise
        ops = nz.ops
        top = self._top
        if ctype == 'capacitive':
            K = torch.linalg.inv(self._C_torch()) @ torch.as_tensor(np.real(top.R), dtype=torch.float64)
        else:
            K = torch.as_tensor(np.real(top.S), dtype=torch.float64)
        n1, n2 = nodes
        if 0 in nodes:
            row = K[n1 + n2 - 1, :]
        elif ctype == 'capacitive':
            row = K[n2 - 1, :] - K[n1 - 1, :]
        else:
            row = K[n1 - 1, :] - K[n2 - 1, :]
        out = 0
        for i in range(ops.M):
            unit = np.zeros((1, ops.M))
            unit[0, i] = 1.0
            if ctype == 'capacitive':
                fn = lambda X, unit=unit: nz._apply_charge_comb(X, unit)
            else:
                fn = lambda X, unit=unit: nz._apply_pot(X, nz._zeros_a(), ops.phi_lin_b(unit))
            out = out + row[i] * self._inner_torch(states, fn)
        return out

    def _C_torch(self):
        """Capacitance matrix as a torch expression of the element tensors (circuit.py:743-835)."""
        from .topology import edge_matrix
        n = self._top.n_nodes
        Cm = torch.zeros((n, n), dtype=torch.float64)
        for edge, els in self.elements.items():
            em = torch.as_tensor(edge_matrix(edge, n), dtype=torch.float64)
            for el in els:
                cap = el if isinstance(el, Capacitor) else el.cap
                Cm = Cm + cap.get_value() * em
        return Cm

    def _dec_rate_torch(self, dec_type, states, total=True):
        from . import autograd as ag
        from .noise import ENV
        nz = self._last.noise
        ops, t = nz.ops, self._top
        omega1, omega2 = self._efreqs[states[0]], self._efreqs[states[1]]
        omega = torch.abs(omega2 - omega1)
        decay = torch.zeros((), dtype=torch.float64)
        alpha = unt.hbar * omega / (unt.k_B * ENV['T'])
        if float(alpha.detach()) > 709:
            down, up = 2, 0
        else:
            down = 1 + 1 / torch.tanh(alpha / 2)
            up = down * torch.exp(-alpha)
        if not total:
            tempS = down if states[0] > states[1] else up
        else:
            tempS = down + up
        if dec_type == 'capacitive':
            for edge, els in self.elements.items():
                for el in els:
                    cap = el if isinstance(el, Capacitor) else el.cap
                    if cap.Q:
                        me = self._matrix_elements_torch('capacitive', edge, states)
                        decay = decay + tempS * cap.get_value() / cap.Q(omega) * torch.abs(me) ** 2
        elif dec_type == 'inductive':
            for edge, el, b_id in t.ind_keys:
                Q = el.Q(omega, ENV['T'])
                if np.isnan(float(Q.detach()) if isinstance(Q, torch.Tensor) else float(Q)):
                    continue
                # the stored 'ind_hamil' operator of the reference is a constant (numpy coefficients)
                fn = lambda X, edge=edge: nz.apply_coupling(X, 'inductive', edge)
                decay = decay + tempS / Q / el.get_value() * torch.abs(self._inner_torch(states, fn)) ** 2
        elif dec_type == 'quasiparticle':
            for _, el, b_id, w_id in t.jj_keys:
                Y = el.Y(omega, ENV['T'])
                if np.isnan(float(Y.detach()) if isinstance(Y, torch.Tensor) else float(Y)):
                    continue
                # circuit.py:1899-1936: sin_half = (e^{i phi/2} root_exp - h.c.) / 2i with phi a torch
                # expression of the loop values; split into the cos(phi/2) and sin(phi/2) parts
                phi = self._ext_flux_torch(b_id)
                fc = lambda X, w_id=w_id: self._apply_half(nz, X, w_id, 0.5 / 1j)       # (root_exp - h.c.) / 2i
                fs = lambda X, w_id=w_id: self._apply_half(nz, X, w_id, 0.5)            # (root_exp + h.c.) / 2
                me = torch.cos(phi / 2) * self._inner_torch(states, fc) + torch.sin(phi / 2) * self._inner_torch(states, fs)
                decay = decay + tempS * Y * omega * el.get_value() * unt.hbar * torch.abs(me) ** 2
        elif dec_type == 'charge':
            decay = decay + ag.dec_rate_charge_torch(self, states)
        elif dec_type == 'cc':
            decay = decay + ag.dec_rate_cc_torch(self, states)
        elif dec_type == 'flux':
            decay = decay + ag.dec_rate_flux_torch(self, states)
        else:
            raise ValueError(f'The decoherence type {dec_type} is not supported.')
        return decay

    @staticmethod
    def _apply_half(nz, X, w_id, coef):
        """(coef root_exp_w + conj(coef) root_exp_w^dag) X: half-angle junction operator without flux phase."""
        a = nz._zeros_a()
        a[:, w_id] = coef
        return nz._apply_pot(X, a, np.zeros((1, 1 + nz.ops.M)), etab=nz.ops.etab_half(),
                             shift=np.zeros_like(nz.ops.shift))

    def _ext_flux_torch(self, b_id):
        """phi_ext at inductive element b_id as a torch expression of the loop values (circuit.py:1759-1786)."""
        phi = torch.zeros((), dtype=torch.float64)
        for li, lp in enumerate(self.loops):
            v = lp.internal_value
            v = v if isinstance(v, torch.Tensor) else torch.as_tensor(float(v), dtype=torch.float64)
            phi = phi + v * float(self._top.B[b_id, li])
        return phi

    def charge_op(self, mode, basis='original'):
        """Charge operator of a mode as a dense matrix (circuit.py:1831-1875); host-side helper for
        user expressions such as the paper's g_10 = |<1|Q|0>| / C.  Tensor in the PyTorch engine."""
        if len(self.m) == 0:
            raise CircuitStateError('Please specify the truncation number for each mode.')
        if basis not in ('original', 'eig'):
            raise ValueError(f"Invalid basis '{basis}' passed. Permitted bases are 'original' and 'eig'.")
        i = mode - 1
        ops = self._operators()
        mats = []
        for j in range(self.n):
            mj = self.m[j]
            if j != i:
                mats.append(np.eye(mj))
            elif self._is_charge_mode(j):
                nmax = (mj - 1) // 2
                ng = self.charge_islands[j].value()
                mats.append((2 * unt.e / np.sqrt(unt.hbar)) * (np.diag(np.arange(-nmax, nmax + 1.0)) - ng * np.eye(mj)))
            elif mj == 1:
                mats.append(np.zeros((1, 1)))
            else:
                a = np.diag(np.sqrt(np.arange(1, mj)), 1)
                mats.append(-1j * np.sqrt(0.5 / ops.Z[j]) * (a - a.T))
        Q = mats[0]
        for mt in mats[1:]:
            Q = np.kron(Q, mt)
        Q = Q.astype(complex)
        if basis == 'eig':
            if get_optim_mode():
                V = self._evecs.detach().cpu().numpy().T
            else:
                V = np.concatenate([np.asarray(v) for v in self._evecs], axis=1)
            Q = V.conj().T @ Q @ V
        return torch.as_tensor(Q, dtype=torch.complex128) if get_optim_mode() else Q

    def check_convergence(self, eig_vec_idx=1, t=10, threshold=1e-5):
        """circuit.py:2090-2118."""
        if self._last is None:
            raise CircuitStateError('Must call circuit.diag before testing convergence.')
        v = self._last.evecs[0, eig_vec_idx].cpu().numpy().reshape(self.m)
        sub = v[(slice(-t),) * len(self.m)]
        eps = 1 - np.sum(np.abs(sub) ** 2)
        return (eps < threshold), eps

    def coord_transform(self, var_type):
        var_type = var_type.lower()
        if var_type == 'charge':
            return np.linalg.inv(self.R)
        if var_type == 'flux':
            return np.linalg.inv(self.S)
        raise ValueError("The input must be either 'charge' or 'flux'.")


def _to_float(v):
    return v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v
