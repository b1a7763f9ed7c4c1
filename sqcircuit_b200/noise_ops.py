"""Batched transition matrix elements and decoherence rates on the eigenvectors of a sweep.

Reference: Circuit.matrix_elements / dec_rate / _dec_rate_*_np / _get_partial_H
(circuit.py:2397-2873).  Operators are applied matrix-free on the GPU to the two states of
interest of every sweep point at once; the inner products use the tensor-core Gram kernel;
the closed-form rate expressions (a handful of flops per sweep point) are evaluated on the host.
"""
import itertools

import numpy as np
import torch

from . import units as unt
from .elements import Capacitor
from .kernels import BlockView
from .noise import ENV


_TOKENS = itertools.count(1)


class SweepNoise:
    def __init__(self, kern, ops, hop, evals_rad, evecs, circuit):
        self._token = next(_TOKENS)       # identity of this result in the shared workspace (ids are reused)
        self.kern, self.ops, self.hop = kern, ops, hop
        self.evals = evals_rad            # [B][k] tensor
        self.evecs = evecs                # [B][k][N] tensor
        self.cr = circuit
        self.B = evecs.shape[0]
        self.dev = evecs.device

    # ---- helpers ------------------------------------------------------------------------
    def _buffer(self, name):
        """[B][2][N] scratch from the backend's persistent workspace (GB-sized at bench scale: allocating it per
        query makes the caching allocator grow new segments between sweeps, measured +170 ms on a 300 ms step).
        The workspace is shared by every SweepNoise of the backend; ``owner`` tells whose data it holds."""
        buf = self.kern.workspace('noise_' + name, (self.B, 2, self.evecs.shape[2]), torch.complex128)
        owners = self.kern.__dict__.setdefault('_noise_owner', {})
        mine = owners.get(name) == (self._token, getattr(self, '_pair_key', None))
        return buf, mine, owners

    def _pair(self, states):
        key = tuple(int(x) for x in states)
        buf, mine, owners = self._buffer('pair')
        if getattr(self, '_pair_key', None) != key or not mine:      # the six channels of a query share it
            self._pair_key = key
            torch.stack([self.evecs[:, key[0]], self.evecs[:, key[1]]], dim=1, out=buf)        # [B][2][N]
            owners['pair'] = (self._token, key)
        return buf

    def _gram2(self, X, Y):
        out = torch.empty((self.B, 2, 2), dtype=torch.complex128, device=self.dev)
        self.kern.gram(BlockView(X), BlockView(Y), out)
        return self.kern.download(out)

    def _phases(self, b_id, scale=1.0):
        return np.exp(1j * scale * self.ops.ext_flux(self.hop.loop_vals, b_id))

    def _zeros_a(self):
        return np.zeros((self.B, max(self.ops.T, 1)), dtype=complex)

    def _apply_pot(self, X, a, g, etab=None, shift=None):
        ops = self.ops
        Y = torch.empty_like(X)
        g = np.broadcast_to(np.asarray(g, dtype=float).reshape(-1, 1 + ops.M), (self.B, 1 + ops.M))
        ops.apply_potential_op(BlockView(X), BlockView(Y), ops.to_device(a, torch.complex128),
                               ops.to_device(g, torch.float64), etab=etab, shift=shift)
        return Y

    # ---- potential-type operators on the pair: matrix elements in the x basis ---------------------------
    def _pair_x(self, states):
        """The pair of ``_pair`` rotated to the x basis of every harmonic mode (once per pair): every
        potential-type operator O = F^T P F then has <a|O|b> = <Fa| P |Fb> with P element-wise + shifts, so
        the six channels of a T1 / Tphi query share ONE forward transform instead of paying a forward and a
        backward transform per operator."""
        X = self._pair(states)
        ops = self.ops
        if not ops.tmodes:
            return X
        buf, mine, owners = self._buffer('pairx')
        if not mine:
            src = BlockView(X)
            for i in ops.tmodes:
                self.kern.mode_transform(ops.sp, i, ops.V[i], 1, src, BlockView(buf), impl=ops.impl)
                src = BlockView(buf)
            owners['pairx'] = (self._token, self._pair_key)
        return buf

    def _pot_elements(self, states, a, g, etab=None, shift=None):
        """[B][2][2] matrix of the potential-type operator (a, g) between the two states (host array)."""
        ops = self.ops
        Z = self._pair_x(states)
        g = np.broadcast_to(np.asarray(g, dtype=float).reshape(-1, 1 + ops.M), (self.B, 1 + ops.M))
        pot = ops.potential_struct(ops.to_device(a, torch.complex128), ops.to_device(g, torch.float64),
                                   ops.etab if etab is None else etab, shift, None)
        out = torch.empty((self.B, 2, 2), dtype=torch.complex128, device=self.dev)
        self.kern.potential_elements(ops.sp, pot, BlockView(Z), out)      # one read of the pair, nothing written
        return self.kern.download(out)

    def _junction_elements(self, states, kind, w_id, b_id, weight=1.0):
        """Matrix of the cos / sin / sin_half operator of one junction (circuit.py:1809-1825) on the pair."""
        a = self._zeros_a()
        zero_g = np.zeros(1 + self.ops.M)
        if kind == 'cos':
            a[:, w_id] = weight * self._phases(b_id) / 2
            return self._pot_elements(states, a, zero_g)
        if kind == 'sin':
            a[:, w_id] = weight * self._phases(b_id) / 2j
            return self._pot_elements(states, a, zero_g)
        if kind == 'sin_half':
            a[:, w_id] = weight * self._phases(b_id, 0.5) / 2j
            return self._pot_elements(states, a, zero_g, etab=self.ops.etab_half(),
                                      shift=np.zeros_like(self.ops.shift))
        raise ValueError(kind)

    def _coupling_elements(self, states, ctype, nodes):
        if ctype == 'capacitive':
            X = self._pair(states)
            return self._gram2(X, self.apply_coupling(X, ctype, nodes))
        k = self.ops.kcoef_b(ctype, nodes)
        return self._pot_elements(states, self._zeros_a(), self.ops.phi_lin_b(k))

    def _apply_charge_comb(self, X, k):
        """Y = (sum_i k_i Q_i) X  with Q_i of circuit.py:1471-1497 (normalised by sqrt(hbar)); k: [M], or one
        row per point [B][M] (batched circuits)."""
        ops = self.ops
        Y = torch.zeros_like(X)
        ng = self.hop.ng
        k = np.asarray(k, dtype=float).reshape(-1, ops.M)
        Bn = max(ng.shape[0], k.shape[0])
        c = 2 * unt.e / np.sqrt(unt.hbar)
        dvec = torch.zeros((Bn,) + tuple(ops.dims), dtype=torch.float64, device=self.dev)
        any_charge = False
        for i in range(ops.M):
            if ops.harm[i] or not np.any(k[:, i] != 0):
                continue
            any_charge = True
            nmax = (ops.dims[i] - 1) // 2
            vals = (np.arange(-nmax, nmax + 1)[None, :] - ng[:, i:i + 1]) * (k[:, i:i + 1] * c)
            shape = [vals.shape[0]] + [1] * ops.M
            shape[1 + i] = ops.dims[i]
            dvec += self.kern.upload(vals, torch.float64).reshape(shape)
        first = True
        if any_charge:
            self.kern.diag_mul(BlockView(X), BlockView(Y), dvec.reshape(Bn, ops.N).contiguous(), False)
            first = False
        for i in range(ops.M):
            if not ops.harm[i] or ops.dims[i] == 1 or not np.any(k[:, i] != 0):
                continue
            qzp = -1j * np.sqrt(0.5 / ops.Zb[:, i]) * k[:, i]          # Q = qzp (a - a^dag), per point
            qzp = np.broadcast_to(qzp, (self.B,))
            cdn = ops.to_device(qzp, torch.complex128)
            cup = ops.to_device(-qzp, torch.complex128)
            self.kern.ladder_apply(ops.sp, i, cdn, cup, BlockView(X), BlockView(Y), not first)
            first = False
        return Y

    # ---- operators of the reference ----------------------------------------------------------
    def apply_coupling(self, X, ctype, nodes):
        k = self.ops.kcoef_b(ctype, nodes)
        if ctype == 'capacitive':
            return self._apply_charge_comb(X, k)
        return self._apply_pot(X, self._zeros_a(), self.ops.phi_lin_b(k))

    def apply_junction_fn(self, X, kind, w_id, b_id, weight=1.0):
        """cos / sin / sin_half operator of one junction (circuit.py:1809-1825)."""
        a = self._zeros_a()
        if kind == 'cos':
            a[:, w_id] = weight * self._phases(b_id) / 2
            return self._apply_pot(X, a, np.zeros(1 + self.ops.M))
        if kind == 'sin':
            a[:, w_id] = weight * self._phases(b_id) / 2j
            return self._apply_pot(X, a, np.zeros(1 + self.ops.M))
        if kind == 'sin_half':
            a[:, w_id] = weight * self._phases(b_id, 0.5) / 2j
            return self._apply_pot(X, a, np.zeros(1 + self.ops.M), etab=self.ops.etab_half(),
                                   shift=np.zeros_like(self.ops.shift))
        raise ValueError(kind)

    def matrix_elements(self, ctype, nodes, states):
        return self._coupling_elements(states, ctype, nodes)[:, 0, 1]

    # ---- decoherence ----------------------------------------------------------------------------
    @staticmethod
    def _dephasing(A, p_omega):
        return np.abs(p_omega * A) * np.sqrt(2 * np.abs(np.log(ENV['omega_low'] * ENV['t_exp'])))

    def _partial_omega_loop(self, states, loop_idx):
        """<m|dH/dphi_loop|m> - <n|...|n>  (circuit.py:2749-2763)."""
        ops, t = self.ops, self.ops.topo
        g = np.zeros(1 + ops.M)
        for edge, el, b_id in t.ind_keys:
            g += (t.B[b_id, loop_idx] / el.si() * unt.Phi0 / np.sqrt(unt.hbar) / 2 / np.pi
                  * ops._phi_lin(ops._kcoef('inductive', edge)))
        a = self._zeros_a()
        for _, jj, b_id, w_id in t.jj_keys:
            a[:, w_id] += t.B[b_id, loop_idx] * jj.si() * self._phases(b_id) / 2j
        G = self._pot_elements(states, a, g)
        return np.real(G[:, 0, 0] - G[:, 1, 1])

    def dec_rate(self, dec_type, states, total=True):
        cr, ops, t = self.cr, self.ops, self.ops.topo
        w = self.kern.download(self.evals)
        omega = np.abs(w[:, states[1]] - w[:, states[0]])
        X = self._pair(states)
        alpha = unt.hbar * omega / (unt.k_B * ENV['T'])
        big = alpha > 709
        al = np.where(big, 1.0, alpha)
        down = np.where(big, 2.0, 1 + 1 / np.tanh(al / 2))
        up = np.where(big, 0.0, down * np.exp(-al))
        if total:
            tempS = down + up
        else:
            tempS = down if states[0] > states[1] else up
        decay = np.zeros(self.B)
        if dec_type == 'capacitive':
            for edge, els in cr.elements.items():
                for el in els:
                    cap = el if isinstance(el, Capacitor) else el.cap
                    if cap.Q:
                        me = self._coupling_elements(states, 'capacitive', edge)[:, 0, 1]
                        decay += tempS * cap.si() / cap.Q(omega) * np.abs(me) ** 2
        elif dec_type == 'inductive':
            for edge, el, b_id in t.ind_keys:
                Q = np.asarray(el.Q(omega, ENV['T']), dtype=float) * np.ones(self.B)
                me = self._coupling_elements(states, 'inductive', edge)[:, 0, 1]
                term = tempS / Q / el.si() * np.abs(me) ** 2
                decay += np.where(np.isnan(Q), 0.0, term)
        elif dec_type == 'quasiparticle':
            for _, el, b_id, w_id in t.jj_keys:
                Y = np.asarray(el.Y(omega, ENV['T']), dtype=float) * np.ones(self.B)
                me = self._junction_elements(states, 'sin_half', w_id, b_id)[:, 0, 1]
                term = tempS * Y * omega * el.si() * unt.hbar * np.abs(me) ** 2
                decay += np.where(np.isnan(Y), 0.0, term)
        elif dec_type == 'charge':
            for i in range(ops.M):
                if ops.harm[i]:
                    continue
                k = t.cInvTrans[i, :ops.M] / np.sqrt(unt.hbar)
                G = self._gram2(X, self._apply_charge_comb(X, k))
                p_omega = np.abs(G[:, 0, 0] - G[:, 1, 1])
                decay += self._dephasing(cr.charge_islands[i].A * 2 * unt.e, p_omega)
        elif dec_type == 'cc':
            for _, el, b_id, w_id in t.jj_keys:
                G = self._junction_elements(states, 'cos', w_id, b_id, weight=-1.0)
                p_omega = np.real(G[:, 0, 0] - G[:, 1, 1])
                decay += self._dephasing(el.A * el.si(), p_omega)
        elif dec_type == 'flux':
            for li, lp in enumerate(t.loops):
                decay += self._dephasing(lp.A, self._partial_omega_loop(states, li))
        else:
            raise ValueError(f'The decoherence type {dec_type} is not supported.')
        return decay
