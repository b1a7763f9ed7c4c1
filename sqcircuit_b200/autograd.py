"""Autograd through ``Circuit.diag`` and the decoherence rates (PyTorch engine).

Mirrors the reference's ``torch_extensions`` (torch_extensions.py:135-1016): ``EigenSolver`` (forward:
diagonalise, detached; once-differentiable backward from

    d omega_m / dx = <m| dH/dx |m>                                  (circuit.py:2777-2826)
    d |m> / dx     = sum_{n != m} <n| dH/dx |m> / (omega_m - omega_n + eps) |n>   (circuit.py:2875-2921)

over the computed eigenpairs) and ``DecRateCharge / DecRateCC / DecRateFlux`` (second derivatives of the
transition frequency).  All operator work runs on the GPU and is shared by every element and state
(``gradients.GradEngine``); ``BatchEigenSolverFn`` does the same for a batch of circuits of one graph in ONE
batched solve (``batch.CircuitBatch``).  T1-type rates (capacitive, inductive, quasiparticle) are composed from
torch operations on the eigenvectors like in the reference, with the operators applied matrix-free by
``LinearOpFn``.
"""
import numpy as np
import torch

from .gradients import GradEngine

EIGENVECTOR_MAX_GRAD = {'n': None}


def set_max_eigenvector_grad(n):
    EIGENVECTOR_MAX_GRAD['n'] = n


def _engine_of(circuit):
    """GradEngine of the circuit's last diagonalisation (cached on its SweepResult)."""
    res = circuit._last
    if getattr(res, '_grad_engine', None) is None:
        res._grad_engine = GradEngine(res.noise, [circuit])
    return res._grad_engine


class EigenSolverFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, element_tensors, circuit, n_eig):
        res = circuit.sweep_diag(n_eig, warm_start=False)
        circuit._last = res
        ctx.circuit = circuit
        ctx.res = res
        ctx.n_params = len(circuit._parameters)
        ctx.out_shape = element_tensors.shape
        ev = res.evals_rad[0].cpu().to(torch.complex128).unsqueeze(-1)
        return torch.cat([ev, res.evecs[0].cpu()], dim=-1)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        res, cr = ctx.res, ctx.circuit
        if getattr(res, '_grad_engine', None) is None:
            res._grad_engine = GradEngine(res.noise, [cr])
        eng = res._grad_engine
        dev = eng.dev
        g_val = grad_output[:, 0].unsqueeze(0)                                   # [1][k]
        g_vec = grad_output[:, 1:].to(dev).contiguous().unsqueeze(0)            # [1][k][N]
        if not bool((g_vec != 0).any()):
            g_vec = None
        out = eng.eigensolver_backward(ctx.n_params, g_val, g_vec, EIGENVECTOR_MAX_GRAD['n'])
        return out[0].cpu().view(ctx.out_shape), None, None


class BatchEigenSolverFn(torch.autograd.Function):
    """Eigenpairs of B circuits of one graph in one batched GPU solve; ``element_tensors`` [B][P] stacks the
    optimisation parameters of every circuit.  Outputs: eigenvalues [B][k] (rad/s) and eigenvectors [B][k][N],
    both on the device."""

    @staticmethod
    def forward(ctx, element_tensors, batch, n_eig):
        res = batch._solve(n_eig)
        ctx.batch = batch
        ctx.res = res
        ctx.shape = element_tensors.shape
        ctx.in_device = element_tensors.device
        # new tensor objects: the outputs get this node as grad_fn, and the node keeps ``res`` alive -- returning
        # res.evals_rad / res.evecs themselves would close a reference cycle through the C++ graph that
        # Python's collector cannot see (measured: 13 GB of device memory leaked per call at B = 512)
        return res.evals_rad.clone(), res.evecs.view(res.evecs.shape)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_val, g_vec):
        res, batch = ctx.res, ctx.batch
        if getattr(res, '_grad_engine', None) is None:
            res._grad_engine = GradEngine(res.noise, batch.circuits)
        eng = res._grad_engine
        if g_vec is not None and not bool((g_vec != 0).any()):
            g_vec = None
        if g_val is None:
            g_val = torch.zeros_like(res.evals_rad)
        out = eng.eigensolver_backward(ctx.shape[1], g_val, g_vec, EIGENVECTOR_MAX_GRAD['n'])
        eng.prim.release_vectors()       # 2 M blocks of the size of the eigenvector block; the k x k matrices stay
        return out.to(ctx.in_device).view(ctx.shape), None, None


class LinearOpFn(torch.autograd.Function):
    """Y = O X for a Hermitian operator O applied matrix-free on the GPU to a [k][N] block of (CPU or device)
    state vectors; backward: grad_X = O grad_Y.  ``apply_fn`` maps a device block [1][k][N] to O times it."""

    @staticmethod
    def forward(ctx, X, apply_fn, dev):
        ctx.apply_fn, ctx.dev, ctx.in_dev = apply_fn, dev, X.device
        Y = apply_fn(X.detach().to(dev).contiguous().unsqueeze(0))[0]
        return Y.to(X.device)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gY):
        gX = ctx.apply_fn(gY.to(ctx.dev).contiguous().unsqueeze(0))[0]
        return gX.to(ctx.in_dev), None, None


def operator_inner_product(state_m, apply_fn, state_n, dev):
    """<m| O |n> with autograd through both states (functions.py operator_inner_product)."""
    Y = LinearOpFn.apply(state_n.unsqueeze(0), apply_fn, dev)[0]
    return torch.sum(state_m.conj() * Y)


class _DecRateFn(torch.autograd.Function):
    """Common wrapper of the three dephasing channels (torch_extensions.py:620-657, 789-827, 979-1016):
    forward = the rate of the NumPy path, backward = grad_output x d rate / d element."""
    kind = ''

    @staticmethod
    def _run(ctx, cls, circuit_parameters, circuit, states):
        rate = circuit._last.dec_rate(cls.kind, states)[0]
        ctx.circuit, ctx.states, ctx.shape = circuit, tuple(states), circuit_parameters.shape
        return torch.as_tensor(float(rate), dtype=torch.float64)

    @staticmethod
    def _back(ctx, cls, grad_output):
        eng = _engine_of(ctx.circuit)
        n = len(ctx.circuit._parameters)
        g = getattr(eng, 'grad_dec_' + cls.kind)(ctx.states, n)[0].cpu()
        return (grad_output * g).view(ctx.shape), None, None


class DecRateCharge(_DecRateFn):
    kind = 'charge'

    @staticmethod
    def forward(ctx, circuit_parameters, circuit, states):
        return _DecRateFn._run(ctx, DecRateCharge, circuit_parameters, circuit, states)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        return _DecRateFn._back(ctx, DecRateCharge, grad_output)


class DecRateCC(_DecRateFn):
    kind = 'cc'

    @staticmethod
    def forward(ctx, circuit_parameters, circuit, states):
        return _DecRateFn._run(ctx, DecRateCC, circuit_parameters, circuit, states)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        return _DecRateFn._back(ctx, DecRateCC, grad_output)


class DecRateFlux(_DecRateFn):
    kind = 'flux'

    @staticmethod
    def forward(ctx, circuit_parameters, circuit, states):
        return _DecRateFn._run(ctx, DecRateFlux, circuit_parameters, circuit, states)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        return _DecRateFn._back(ctx, DecRateFlux, grad_output)


def _params(circuit):
    return torch.stack(circuit.parameters) if circuit._parameters else torch.tensor([])


def dec_rate_charge_torch(circuit, states):
    return DecRateCharge.apply(_params(circuit), circuit, states)


def dec_rate_cc_torch(circuit, states):
    return DecRateCC.apply(_params(circuit), circuit, states)


def dec_rate_flux_torch(circuit, states):
    return DecRateFlux.apply(_params(circuit), circuit, states)
