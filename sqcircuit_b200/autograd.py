"""Autograd through ``Circuit.diag`` (PyTorch engine).

Mirrors the reference's ``torch_extensions.EigenSolver`` (torch_extensions.py:135-279): the forward
pass diagonalises on the GPU (detached); the once-differentiable backward uses

    d omega_m / dx = <m| dH/dx |m>                                  (circuit.py:2777-2826)
    d |m> / dx     = sum_{n != m} <n| dH/dx |m> / (omega_m - omega_n + eps) |n>   (circuit.py:2875-2921)

over the computed eigenpairs, with dH/dx of ``Circuit._get_partial_H`` (circuit.py:2704-2775)
applied matrix-free to the eigenvector block (ladder / diagonal / potential kernels) and the
k x k matrices <n| dH/dx |m> from the tensor-core Gram kernel.
"""
import numpy as np
import torch

from . import units as unt
from .elements import Capacitor, Inductor, Junction, Loop
from .kernels import BlockView

EIGENVECTOR_MAX_GRAD = {'n': None}


def set_max_eigenvector_grad(n):
    EIGENVECTOR_MAX_GRAD['n'] = n


def gram_blocks(kern, A, Bm):
    """<A_i | Bm_j> for [1][ka][N] x [1][kb][N] blocks of any size (the tensor-core Gram kernel takes
    <= 64 left and <= 16 right vectors per call)."""
    ka, kb = A.shape[1], Bm.shape[1]
    out = torch.empty((1, ka, kb), dtype=torch.complex128, device=A.device)
    for r0 in range(0, ka, 64):
        ra = A[:, r0:r0 + 64].contiguous()
        for c0 in range(0, kb, 16):
            cb = Bm[:, c0:c0 + 16].contiguous()
            tmp = torch.empty((1, ra.shape[1], cb.shape[1]), dtype=torch.complex128, device=A.device)
            kern.gram(BlockView(ra), BlockView(cb), tmp)
            out[:, r0:r0 + ra.shape[1], c0:c0 + cb.shape[1]] = tmp
    return out


class PartialH:
    """dH/dx applied to a block of state vectors [B=1][k][N] for one circuit at one point."""

    def __init__(self, circuit, noise):
        self.cr = circuit
        self.nz = noise                 # SweepNoise of the diagonalisation (operators, phases, kernels)
        self.ops = noise.ops
        self.kern = noise.kern
        self.dev = noise.dev

    # --- single-mode operators (circuit.py:1471-1524) -------------------------------------------
    def _apply_mode_op(self, kind, i, X):
        ops = self.ops
        k = np.zeros(ops.M)
        k[i] = 1.0
        if kind == 'Q':
            return self.nz._apply_charge_comb(X, k)
        return self.nz._apply_pot(X, self.nz._zeros_a(), ops._phi_lin(k))

    def _quadratic(self, kind, A, X):
        """(1/2 sum_i A_ii O_i^2 + sum_{i<j} A_ij O_i O_j) X  (circuit.py:2649-2702)."""
        M = self.ops.M
        Z = [self._apply_mode_op(kind, j, X) for j in range(M)]
        Y = torch.zeros_like(X)
        for i in range(M):
            comb = torch.zeros_like(X)
            used = False
            for j in range(i, M):
                c = 0.5 * A[i, i] if i == j else A[i, j]
                if c != 0:
                    comb += c * Z[j]
                    used = True
            if used:
                Y += self._apply_mode_op(kind, i, comb)
        return Y

    def apply(self, el, X):
        cr, t, ops = self.cr, self.cr._top, self.ops
        M = ops.M
        if isinstance(el, Capacitor):
            cinv = np.linalg.inv(t.C)
            A = -t.R.T @ cinv @ t.partial_mats[el] @ cinv @ t.R
            return self._quadratic('Q', A[:M, :M], X)
        if isinstance(el, Inductor):
            A = -t.S.T @ t.partial_mats[el] @ t.S
            Y = self._quadratic('phi', A[:M, :M], X)
            for edge, li, b_id in t.ind_keys:
                if li is el:
                    phi = self.ops.ext_flux(self.nz.hop.loop_vals, b_id)[0]
                    coef = -phi / el.si() ** 2 / np.sqrt(unt.hbar) * (unt.Phi0 / 2 / np.pi)
                    if coef != 0:
                        Y += coef * self.nz.apply_coupling(X, 'inductive', edge)
            return Y
        if isinstance(el, Loop):
            li_ = t.loops.index(el)
            g = np.zeros(1 + M)
            for edge, li, b_id in t.ind_keys:
                g += (t.B[b_id, li_] / li.si() * unt.Phi0 / np.sqrt(unt.hbar) / 2 / np.pi
                      * ops._phi_lin(ops._kcoef('inductive', edge)))
            a = self.nz._zeros_a()
            for _, jj, b_id, w_id in t.jj_keys:
                a[:, w_id] += t.B[b_id, li_] * jj.si() * self.nz._phases(b_id) / 2j
            return self.nz._apply_pot(X, a, g)
        if isinstance(el, Junction):
            a = self.nz._zeros_a()
            for _, jj, b_id, w_id in t.jj_keys:
                if jj is el:
                    a[:, w_id] += -self.nz._phases(b_id) / 2
            return self.nz._apply_pot(X, a, np.zeros(1 + M))
        raise NotImplementedError(type(el))


class EigenSolverFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, element_tensors, circuit, n_eig):
        res = circuit.sweep_diag(n_eig, warm_start=False)
        circuit._last = res
        ctx.circuit = circuit
        ctx.res = res
        ctx.elems = list(circuit._parameters.keys())
        ctx.out_shape = element_tensors.shape
        ev = res.evals_rad[0].cpu().to(torch.complex128).unsqueeze(-1)
        return torch.cat([ev, res.evecs[0].cpu()], dim=-1)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        res, cr = ctx.res, ctx.circuit
        nz = res.noise
        dev = nz.dev
        k = res.evecs.shape[1]
        g_val = grad_output[:, 0]
        g_vec = grad_output[:, 1:].to(dev).contiguous().unsqueeze(0)          # [1][k][N]
        psi = res.evecs[:1].contiguous()                                       # [1][k][N]
        w = res.evals_rad[0].cpu().numpy()
        nmax = EIGENVECTOR_MAX_GRAD['n'] if EIGENVECTOR_MAX_GRAD['n'] is not None else k
        # <g_m | psi_n>
        ov = gram_blocks(nz.kern, g_vec, psi)[0].cpu().numpy()                 # ov[m][n]
        ph = PartialH(cr, nz)
        grads = []
        for el in ctx.elems:
            Y = ph.apply(el, psi)
            Mx = gram_blocks(nz.kern, psi, Y)[0].cpu().numpy()                 # Mx[n][m] = <n|dH|m>
            gval = np.sum(np.real(np.diag(Mx)) * np.conj(g_val.numpy()))
            gvec = 0.0
            for m in range(min(nmax, k)):
                for n in range(k):
                    if n != m:
                        gvec += Mx[n, m] / (w[m] - w[n] + 1e-12) * ov[m, n]
            grads.append(float(np.real(gval + gvec)))
        out = torch.tensor(grads, dtype=torch.float64).view(ctx.out_shape)
        return out, None, None
