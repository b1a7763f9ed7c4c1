"""Derivatives of eigenfrequencies, eigenvectors and dephasing rates with respect to element values,
batched over sweep points / circuits on the GPU.

Reference: ``Circuit._get_partial_H / get_partial_omega / get_partial_vec`` (circuit.py:2649-2921) and the
``torch_extensions`` helpers ``partial_squared_omega``, ``partial_charge_dec``, ``partial_cc_dec``,
``partial_flux_dec`` (torch_extensions.py:286-976).  The reference materialises dH/dx as a sparse matrix per
element and takes inner products one (element, state) pair at a time.  Every dH/dx and d2H/dx dy of the
reference is a fixed linear combination of a few elementary operators -- Q_i Q_j, phi_i phi_j, phi_i, Q_i and
the junction exponentials D_t -- whose coefficients are a handful of numbers per circuit.  So the GPU work is
done ONCE per diagonalisation, independent of the number of elements: the elementary operators are applied
matrix-free to the eigenvector block of every point (``SweepNoise`` kernels), their k x k matrices between
the eigenvectors are formed by the tensor-core Gram kernel (``Primitives``), and every derivative is a
[B][k][k] combination of those with per-point coefficient arrays built on the host from the circuit
matrices.
"""
import numpy as np
import torch

from . import units as unt
from .elements import Capacitor, Inductor, Junction, Loop
from .kernels import BlockView
from .noise import ENV

EPS_VEC = 1e-12          # circuit.py:2878 (epsilon of get_partial_vec)


class Primitives:
    """k x k matrices <psi_n| O |psi_m> of the elementary operators, per sweep point: [B][k][k] complex128
    tensors on the device, computed on first use.

      GQ(i, j) = Q_i Q_j      GP(i, j) = phi_i phi_j      FQ(i) = Q_i      FP(i) = phi_i
      JC(t) = D_t + D_t^dag   JS(t) = i D_t - i D_t^dag   (D_t: junction exponential of term t, no flux phase)

    with Q_i, phi_i the mode operators of the reference's ``_memory_ops`` (normalised by sqrt(hbar); phi of a
    charge mode is the identity) -- circuit.py:1471-1524, 1673-1714."""

    def __init__(self, noise):
        self.nz = noise
        self.ops = noise.ops
        self.kern = noise.kern
        self.psi = noise.evecs if noise.evecs.is_contiguous() else noise.evecs.contiguous()
        self.B, self.k, self.N = self.psi.shape
        self.dev = self.psi.device
        self._z = {}
        self._m = {}

    def gram(self, X, Y):
        """<X_i | Y_j> per point, [B][ka][kb]; the tensor-core Gram kernel takes <= 64 left and <= 16 right
        vectors per call, larger blocks go through it in tiles."""
        ka, kb = X.shape[1], Y.shape[1]
        out = torch.empty((self.B, ka, kb), dtype=torch.complex128, device=self.dev)
        if ka <= 64 and kb <= 16:
            self.kern.gram(BlockView(X), BlockView(Y), out)
            return out
        for r0 in range(0, ka, 64):
            xa = X[:, r0:r0 + 64].contiguous()
            for c0 in range(0, kb, 16):
                yb = Y[:, c0:c0 + 16].contiguous()
                tmp = torch.empty((self.B, xa.shape[1], yb.shape[1]), dtype=torch.complex128, device=self.dev)
                self.kern.gram(BlockView(xa), BlockView(yb), tmp)
                out[:, r0:r0 + xa.shape[1], c0:c0 + yb.shape[1]] = tmp
        return out

    def _unit(self, i):
        k = np.zeros((1, self.ops.M))
        k[0, i] = 1.0
        return k

    def ZQ(self, i):
        if ('Q', i) not in self._z:
            self._z['Q', i] = self.nz._apply_charge_comb(self.psi, self._unit(i))
        return self._z['Q', i]

    def ZP(self, i):
        if ('P', i) not in self._z:
            self._z['P', i] = self.nz._apply_pot(self.psi, self.nz._zeros_a(), self.ops.phi_lin_b(self._unit(i)))
        return self._z['P', i]

    def _cached(self, key, make):
        if key not in self._m:
            self._m[key] = make()
        return self._m[key]

    def GQ(self, i, j):
        return self._cached(('GQ', i, j), lambda: self.gram(self.ZQ(i), self.ZQ(j)))

    def GP(self, i, j):
        return self._cached(('GP', i, j), lambda: self.gram(self.ZP(i), self.ZP(j)))

    def FQ(self, i):
        return self._cached(('FQ', i), lambda: self.gram(self.psi, self.ZQ(i)))

    def FP(self, i):
        return self._cached(('FP', i), lambda: self.gram(self.psi, self.ZP(i)))

    def _J(self, t, val):
        a = self.nz._zeros_a()
        a[:, t] = val
        return self.gram(self.psi, self.nz._apply_pot(self.psi, a, np.zeros((1, 1 + self.ops.M))))

    def JC(self, t):
        return self._cached(('JC', t), lambda: self._J(t, 1.0))

    def JS(self, t):
        return self._cached(('JS', t), lambda: self._J(t, 1.0j))

    def release_vectors(self):
        """Drop the cached operator-times-eigenvector blocks (the k x k matrices stay)."""
        self._z = {}


class OpCoeffs:
    """Coefficients of one operator on the elementary operators, for every point: AQ, AP [B][M][M] (upper
    triangle; the diagonal enters with 1/2 as in ``_get_quadratic_Q``), kF, kQ [B][M] (on phi_i, Q_i) and
    aJ [B][T] complex (sum_t a_t D_t + h.c.)."""

    def __init__(self, B, M, T):
        self.AQ = np.zeros((B, M, M))
        self.AP = np.zeros((B, M, M))
        self.kF = np.zeros((B, M))
        self.kQ = np.zeros((B, M))
        self.aJ = np.zeros((B, max(T, 1)), dtype=complex)


class GradEngine:
    """Derivatives for B points.  ``circuits``: the Circuit of every point (one entry: a single circuit whose
    B sweep points share it).  ``noise``: the SweepNoise of the diagonalisation (operators, eigenpairs)."""

    def __init__(self, noise, circuits):
        self.nz = noise
        self.ops = noise.ops
        self.circuits = list(circuits)
        self.prim = Primitives(noise)
        self.B, self.k = self.prim.B, self.prim.k
        self.M, self.T = self.ops.M, self.ops.T
        self.dev = self.prim.dev
        self.loop_vals = np.asarray(noise.hop.loop_vals, dtype=float)
        self.w = noise.evals                                   # [B][k] rad/s, device

    # ---- host side: which circuit / topology / element a point uses -------------------------------
    def _circuit(self, b):
        return self.circuits[b if len(self.circuits) > 1 else 0]

    def _phase(self, c, b, b_id):
        top = self._circuit(c)._top
        if not top.loops:
            return 1.0
        return np.exp(1j * float(self.loop_vals[b] @ np.asarray(top.B, dtype=float)[b_id, :]))

    def element(self, b, index):
        """The ``index``-th optimisation parameter (element or loop) of the circuit of point b."""
        return self._circuit(b).parameters_elems[index]

    # ---- coefficients of dH/dx (circuit.py:2704-2775) -------------------------------------------------
    def partial_H_coeffs(self, el_of_point, b_sel=None):
        """``el_of_point(b)`` -> the element of point b's circuit to differentiate with respect to."""
        M, T = self.M, self.T
        co = OpCoeffs(self.B, M, T)
        for b in range(self.B):
            cr = self._circuit(b)
            top = cr._top
            el = el_of_point(b)
            kc = self.ops.kcoef_b
            c = b if self.ops.nb > 1 else 0
            if isinstance(el, Capacitor):
                cinv = np.linalg.inv(top.C)
                A = -top.R.T @ cinv @ top.partial_mats[el] @ cinv @ top.R
                co.AQ[b] = np.real(A[:M, :M])
            elif isinstance(el, Inductor):
                A = -top.S.T @ top.partial_mats[el] @ top.S
                co.AP[b] = np.real(A[:M, :M])
                for edge, li, b_id in top.ind_keys:
                    if li is el:
                        phi = float(self.loop_vals[b] @ np.asarray(top.B, dtype=float)[b_id, :]) if top.loops else 0.0
                        coef = -phi / el.si() ** 2 / np.sqrt(unt.hbar) * (unt.Phi0 / 2 / np.pi)
                        co.kF[b] += coef * kc('inductive', edge)[c]
            elif isinstance(el, Loop):
                li_ = top.loops.index(el)
                for edge, li, b_id in top.ind_keys:
                    co.kF[b] += (top.B[b_id, li_] / li.si() * unt.Phi0 / np.sqrt(unt.hbar) / 2 / np.pi
                                 * kc('inductive', edge)[c])
                for _, jj, b_id, w_id in top.jj_keys:
                    co.aJ[b, w_id] += top.B[b_id, li_] * jj.si() * self._phase(b, b, b_id) / 2j
            elif isinstance(el, Junction):
                for _, jj, b_id, w_id in top.jj_keys:
                    if jj is el and (b_sel is None or b_sel == b_id):
                        co.aJ[b, w_id] += -self._phase(b, b, b_id) / 2
            else:
                raise NotImplementedError(type(el))
        return co

    # ---- [B][k][k] matrix of an operator given by its coefficients -----------------------------------
    def matrix(self, co: OpCoeffs):
        pr = self.prim
        out = torch.zeros((self.B, self.k, self.k), dtype=torch.complex128, device=self.dev)

        def add(coef, prim):
            if np.any(coef != 0):
                c = torch.as_tensor(np.ascontiguousarray(coef), device=self.dev).to(torch.complex128)
                out.add_(c.reshape(self.B, 1, 1) * prim())
        for i in range(self.M):
            for j in range(i, self.M):
                f = 0.5 if i == j else 1.0
                add(f * co.AQ[:, i, j], lambda: pr.GQ(i, j))
                add(f * co.AP[:, i, j], lambda: pr.GP(i, j))
            add(co.kF[:, i], lambda: pr.FP(i))
            add(co.kQ[:, i], lambda: pr.FQ(i))
        for t in range(self.T):
            add(co.aJ[:, t].real, lambda: pr.JC(t))
            add(co.aJ[:, t].imag, lambda: pr.JS(t))
        return out

    def partial_H_matrix(self, el_of_point, b_sel=None):
        return self.matrix(self.partial_H_coeffs(el_of_point, b_sel))

    # ---- first derivatives of the eigenpairs ------------------------------------------------------------
    def vec_coeffs(self, Mx):
        """c[b][n][m] = <n|dH|m> / (w_m - w_n + eps) (n != m), 0 on the diagonal: d|m> = sum_n c[n][m] |n>
        (circuit.py:2875-2921)."""
        w = self.w
        den = (w.unsqueeze(1) - w.unsqueeze(2)) + EPS_VEC           # [b][n][m] = w_m - w_n + eps
        c = Mx / den
        eye = torch.eye(self.k, dtype=torch.bool, device=self.dev)
        return c.masked_fill(eye, 0.0)

    # ---- second derivatives of a transition frequency (torch_extensions.py:286-350) ------------------------
    def partial_squared_omega(self, Mx, P, P2, states):
        """d/dx [ <m|P|m> - <n|P|n> ]: 2 Re(<dm|P|m> - <dn|P|n>) + <m|P2|m> - <n|P2|n>; Mx = k x k matrix of
        dH/dx, P of dH/dy, P2 of d2H/dx dy (None: zero).  [B] real tensor."""
        m, n = states
        c = self.vec_coeffs(Mx)

        def first(s):          # <d psi_s | P | psi_s> = sum_l conj(c[l][s]) P[l][s]
            return (c[:, :, s].conj() * P[:, :, s]).sum(dim=1)
        out = 2 * (first(m) - first(n)).real
        if P2 is not None:
            out = out + (P2[:, m, m] - P2[:, n, n]).real
        return out

    # ---- dephasing-rate derivatives (torch_extensions.py:353-383, 575-617, 737-786, 933-976) -------------------
    @staticmethod
    def _deph_factor():
        return np.sqrt(2 * np.abs(np.log(ENV['omega_low'] * ENV['t_exp'])))

    def _each_param(self, n_params):
        for e in range(n_params):
            yield e, (lambda b, e=e: self.element(b, e))

    def grad_dec_charge(self, states, n_params):
        """[B][P] derivative of the charge dephasing rate."""
        m, n = states
        out = torch.zeros((self.B, n_params), dtype=torch.float64, device=self.dev)
        mats = {e: self.partial_H_matrix(f) for e, f in self._each_param(n_params)}
        for i in range(self.M):
            if self.ops.harm[i]:
                continue
            co = OpCoeffs(self.B, self.M, self.T)
            A = np.zeros(self.B)
            for b in range(self.B):
                cr = self._circuit(b)
                co.kQ[b] = np.real(cr._top.cInvTrans[i, :self.M]) / np.sqrt(unt.hbar)
                A[b] = cr.charge_islands[i].A * 2 * unt.e
            P = self.matrix(co)
            p_omega = (P[:, m, m] - P[:, n, n]).real
            At = torch.as_tensor(A, device=self.dev)
            for e, f in self._each_param(n_params):
                co2 = OpCoeffs(self.B, self.M, self.T)
                any2 = False
                for b in range(self.B):
                    el = f(b)
                    if isinstance(el, Capacitor):
                        top = self._circuit(b)._top
                        cinv = np.linalg.inv(top.C)
                        # torch_extensions.py:454-490 (the reference uses C^-1 dC C^-1 without the mode
                        # transformation here)
                        A2 = cinv @ top.partial_mats[el] @ cinv
                        co2.kQ[b] = -np.real(A2[i, :self.M]) / np.sqrt(unt.hbar)
                        any2 = True
                P2 = self.matrix(co2) if any2 else None
                d2 = self.partial_squared_omega(mats[e], P, P2, states)
                out[:, e] += torch.sign(p_omega) * self._deph_factor() * At * d2
        return out

    def grad_dec_cc(self, states, n_params):
        """[B][P] derivative of the critical-current dephasing rate."""
        m, n = states
        out = torch.zeros((self.B, n_params), dtype=torch.float64, device=self.dev)
        mats = {e: self.partial_H_matrix(f) for e, f in self._each_param(n_params)}
        top0 = self._circuit(0)._top
        for j, (_, _, b_id, w_id) in enumerate(top0.jj_keys):
            jj_of = lambda b, j=j: self._circuit(b)._top.jj_keys[j][1]
            P = self.partial_H_matrix(jj_of, b_sel=b_id)
            p_omega = (P[:, m, m] - P[:, n, n]).real
            A = torch.as_tensor(np.array([jj_of(b).A * jj_of(b).si() for b in range(self.B)]), device=self.dev)
            for e, f in self._each_param(n_params):
                co2 = OpCoeffs(self.B, self.M, self.T)
                any2 = False
                pA = np.zeros(self.B)
                for b in range(self.B):
                    el = f(b)
                    if el is jj_of(b):
                        pA[b] = jj_of(b).A
                    if isinstance(el, Loop):
                        top = self._circuit(b)._top
                        # B[b_id, loop] sin(el, b_id)  (torch_extensions.py:664-694)
                        co2.aJ[b, w_id] += top.B[b_id, top.loops.index(el)] * self._phase(b, b, b_id) / 2j
                        any2 = True
                P2 = self.matrix(co2) if any2 else None
                d2 = self.partial_squared_omega(mats[e], P, P2, states)
                pAt = torch.as_tensor(pA, device=self.dev)
                out[:, e] += torch.sign(p_omega) * self._deph_factor() * (pAt * p_omega + A * d2)
        return out

    def grad_dec_flux(self, states, n_params):
        """[B][P] derivative of the flux dephasing rate."""
        m, n = states
        out = torch.zeros((self.B, n_params), dtype=torch.float64, device=self.dev)
        mats = {e: self.partial_H_matrix(f) for e, f in self._each_param(n_params)}
        top0 = self._circuit(0)._top
        for li_ in range(len(top0.loops)):
            loop_of = lambda b, li_=li_: self._circuit(b)._top.loops[li_]
            P = self.partial_H_matrix(loop_of)
            p_omega = (P[:, m, m] - P[:, n, n]).real
            A = torch.as_tensor(np.array([loop_of(b).A for b in range(self.B)]), device=self.dev)
            for e, f in self._each_param(n_params):
                co2 = OpCoeffs(self.B, self.M, self.T)
                any2 = False
                for b in range(self.B):
                    el = f(b)
                    top = self._circuit(b)._top
                    c = b if self.ops.nb > 1 else 0
                    # torch_extensions.py:834-893
                    if isinstance(el, Junction):
                        for _, jj, b_id, w_id in top.jj_keys:
                            if jj is el:
                                co2.aJ[b, w_id] += top.B[b_id, li_] * self._phase(b, b, b_id) / 2j
                                any2 = True
                    elif isinstance(el, Inductor):
                        for edge, li, b_id in top.ind_keys:
                            if li is el:
                                co2.kF[b] += (top.B[b_id, li_] / -(el.si() ** 2) * unt.Phi0 / np.sqrt(unt.hbar)
                                              / 2 / np.pi * self.ops.kcoef_b('inductive', edge)[c])
                                any2 = True
                    elif isinstance(el, Loop):
                        l2 = top.loops.index(el)
                        for _, jj, b_id, w_id in top.jj_keys:
                            co2.aJ[b, w_id] += jj.si() * top.B[b_id, l2] * top.B[b_id, li_] * self._phase(b, b, b_id) / 2
                            any2 = True
                P2 = self.matrix(co2) if any2 else None
                d2 = self.partial_squared_omega(mats[e], P, P2, states)
                out[:, e] += torch.sign(p_omega) * self._deph_factor() * A * d2
        return out

    # ---- backward of the eigensolver (torch_extensions.py:213-279) -------------------------------------------
    def eigensolver_backward(self, n_params, g_val, g_vec=None, nmax=None):
        """[B][P] gradient of a scalar whose derivatives w.r.t. the outputs are g_val [B][k] (rad/s
        eigenvalues) and g_vec [B][k][N] (eigenvectors; None: eigenvalues only)."""
        out = torch.zeros((self.B, n_params), dtype=torch.float64, device=self.dev)
        k = self.k
        nmax = k if nmax is None else min(nmax, k)
        ov = None
        if g_vec is not None:
            g_vec = g_vec if g_vec.is_contiguous() else g_vec.contiguous()
            ov = self.prim.gram(g_vec, self.prim.psi)                 # ov[m][n] = <g_m | psi_n>
        gv = g_val.to(device=self.dev, dtype=torch.complex128)
        for e, f in self._each_param(n_params):
            Mx = self.partial_H_matrix(f)
            diag = torch.diagonal(Mx, dim1=1, dim2=2).real
            val = (diag * gv.conj()).sum(dim=1)
            if ov is not None:
                c = self.vec_coeffs(Mx)                               # c[n][m]
                c = c[:, :, :nmax]
                val = val + (c * ov.transpose(1, 2)[:, :, :nmax]).sum(dim=(1, 2))
            out[:, e] = val.real
        return out
