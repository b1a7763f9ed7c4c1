"""Device-resident description of the circuit operators and their matrix-free application.

The truncated-basis Hamiltonian of the reference (circuit.py:1731-1829, built there as a scipy
CSR matrix from Kronecker products) is applied here without ever being formed:

    H X = d_LC .* X  +  (x_m V_m) [ P ( (x_m V_m^T) X ) ]

``d_LC`` is the LC part (diagonal in the Fock x charge basis), ``V_m`` rotates harmonic mode m
to the eigenbasis of its truncated a + a^dag, where every junction displacement operator
``exp(i s (a + a^dag))`` -- qt.displace(...).expm() in the reference -- is diagonal, and ``P``
holds the junction cosines (diagonal phases x unit shifts on charge modes) and the external
flux inductor terms (diagonal).  The same machinery applies the cos / sin / sin_half and
coupling operators needed for matrix elements and decoherence rates.
"""
import ctypes as C

import os
import numpy as np
import torch

from . import units as unt
from ._lib import Potential
from .kernels import BlockView, make_space

_XBASIS_CACHE = {}


def xbasis(kern, m):
    key = (str(kern.dev), kern.name, m)
    if key not in _XBASIS_CACHE:
        _XBASIS_CACHE[key] = kern.xbasis(m)
    return _XBASIS_CACHE[key]


class CircuitOperators:
    use_fused = True      # class-level switch (tests compare the fused and unfused paths)

    def __init__(self, kern, topo, dims, transform_impl=0):
        self.kern = kern
        self.device = kern.dev
        self.topo = topo
        self.dims = [int(d) for d in dims]
        self.M = len(dims)
        self.harm = [bool(topo.omega[i] != 0) for i in range(self.M)]
        self.N = int(np.prod(self.dims))
        self.sp = make_space(self.dims, self.harm)
        self.impl = transform_impl
        self.tmodes = [i for i in range(self.M) if self.harm[i] and self.dims[i] > 1]
        self.toff = np.concatenate(([0], np.cumsum(self.dims)))[:-1]
        self.tlen = int(np.sum(self.dims))
        dev = self.device
        lam = torch.zeros(self.tlen, dtype=torch.float64, device=dev)
        self.V = {}
        for i in self.tmodes:
            Vi, li = xbasis(kern, self.dims[i])
            self.V[i] = Vi
            lam[self.toff[i]:self.toff[i] + self.dims[i]] = li
        self.lam = lam
        # impedance / zero-point flux of the harmonic modes (circuit.py:1455-1469, 1585-1601)
        self.Z = np.zeros(self.M)
        self.phi_zp = np.zeros(self.M)
        for i in range(self.M):
            if self.harm[i]:
                self.Z[i] = np.sqrt(topo.cInvTrans[i, i] / topo.lTrans[i, i])
                self.phi_zp[i] = (2 * np.pi / unt.Phi0) * np.sqrt(unt.hbar / 2 * self.Z[i])
        # junction terms: one per row of W (junctions on the same edge share the operator)
        self.T = topo.wTrans.shape[0] if len(topo.W) else 0
        self.s = np.zeros((max(self.T, 1), self.M))
        self.shift = np.zeros((max(self.T, 1), self.M), dtype=int)
        for t in range(self.T):
            for i in range(self.M):
                w = topo.wTrans[t, i]
                if self.harm[i]:
                    self.s[t, i] = self.phi_zp[i] * w if self.dims[i] > 1 else 0.0
                else:
                    self.shift[t, i] = int(np.sign(w))
        self.etab = self._tables(self.s)
        self._etab_half = None
        # parity halves of the x basis for the fused kernel (one harmonic x one charge mode)
        self.fused = (self.M == 2 and self.harm[0] and not self.harm[1] and self.dims[0] > 1
                      and self.T <= 4)
        if self.fused:
            Vfull = self.V[0]
            he, ho = (self.dims[0] + 1) // 2, self.dims[0] // 2
            self.Ve = Vfull[0::2, :he].contiguous()
            self.Vo = Vfull[1::2, :ho].contiguous()
        self._scratch = {}
        self._keep = []

    # ---- tables --------------------------------------------------------------------------
    def _tables(self, s):
        nt = max(self.T, 1)
        etab = torch.empty((1, nt, self.tlen), dtype=torch.complex128, device=self.device)
        sd = torch.as_tensor(s, dtype=torch.float64, device=self.device).reshape(1, nt, self.M).contiguous()
        self.kern.phase_tables(self.sp, nt, self.lam, sd, etab)
        return etab

    def etab_half(self):
        if self._etab_half is None:
            self._etab_half = self._tables(self.s / 2)
        return self._etab_half

    # ---- per-sweep-point coefficients (host, tiny) ------------------------------------------
    def ext_flux(self, loop_vals, b_id):
        """phi_ext at inductive element b_id for every sweep point (circuit.py:1759-1786)."""
        if not self.topo.loops:
            return np.zeros(loop_vals.shape[0])
        return loop_vals @ self.topo.B[b_id, :]

    def _kcoef(self, ctype, nodes):
        """Node-to-mode coefficients of a coupling operator (circuit.py:2356-2393; like the
        reference, the first n columns of the full transformation matrix are used)."""
        t = self.topo
        K = np.linalg.inv(t.C) @ t.R if ctype == 'capacitive' else t.S
        n1, n2 = nodes
        if 0 in nodes:
            row = K[n1 + n2 - 1, :]
        elif ctype == 'capacitive':
            row = K[n2 - 1, :] - K[n1 - 1, :]
        else:
            row = K[n1 - 1, :] - K[n2 - 1, :]
        return row[:self.M]

    def hamiltonian_coeffs(self, loop_vals):
        """a[b][t], g[b][1+M] of the Hamiltonian's potential part for loop values [B][nloops]."""
        B = loop_vals.shape[0]
        a = np.zeros((B, max(self.T, 1)), dtype=complex)
        for _, el, b_id, w_id in self.topo.jj_keys:
            a[:, w_id] += -el.si() / 2 * np.exp(1j * self.ext_flux(loop_vals, b_id))
        g = np.zeros((B, 1 + self.M))
        for edge, el, b_id in self.topo.ind_keys:
            coef = (1 / el.si()) * self.ext_flux(loop_vals, b_id) * (unt.Phi0 / 2 / np.pi) / np.sqrt(unt.hbar)
            g += coef[:, None] * self._phi_lin(self._kcoef('inductive', edge))[None, :]
        return a, g

    def _phi_lin(self, k):
        """Linear x-basis coefficients [1+M] of sum_i k_i phi_i (circuit.py:1499-1524):
        phi_i = sqrt(Z_i/2)(a+a^dag) for harmonic modes (0 when truncated to one level),
        identity for charge modes."""
        g = np.zeros(1 + self.M)
        for i in range(self.M):
            if self.harm[i]:
                if self.dims[i] > 1:
                    g[1 + i] = k[i] * np.sqrt(self.Z[i] / 2)
            else:
                g[0] += k[i]
        return g

    def lc_params(self, ng):
        """omega, q, ng arrays for sqc_lc_diagonal (circuit.py:1716-1757)."""
        t = self.topo
        c2 = (2 * unt.e) ** 2 / unt.hbar
        q = np.zeros((self.M, self.M))
        for i in range(self.M):
            for j in range(i, self.M):
                cij = t.cInvTrans[i, j]
                if i == j:
                    if not self.harm[i]:
                        q[i, i] = c2 * cij / 2
                elif cij != 0:
                    small = abs(cij) <= 1e-10 * np.sqrt(abs(t.cInvTrans[i, i] * t.cInvTrans[j, j]))
                    if self.harm[i] or self.harm[j]:
                        if not small and self.dims[i] > 1 and self.dims[j] > 1:
                            raise NotImplementedError(
                                'capacitive cross coupling that involves a harmonic mode survived the '
                                'mode transformation; not supported by the matrix-free operator')
                    else:
                        q[i, j] = c2 * cij
        omega = np.array([t.omega[i] if self.harm[i] else 0.0 for i in range(self.M)])
        return omega, q, np.asarray(ng, dtype=float).reshape(-1, self.M)

    def lc_diagonal(self, ng):
        omega, q, ng = self.lc_params(ng)
        B = ng.shape[0]
        dev = self.device
        out = torch.empty((B, self.N), dtype=torch.float64, device=dev)
        om = torch.as_tensor(np.broadcast_to(omega, (B, self.M)).copy(), dtype=torch.float64, device=dev)
        qq = torch.as_tensor(np.broadcast_to(q, (B, self.M, self.M)).copy(), dtype=torch.float64, device=dev)
        ngd = torch.as_tensor(ng, dtype=torch.float64, device=dev).contiguous()
        self.kern.lc_diagonal(self.sp, om, qq, ngd, out)
        return out

    # ---- application ---------------------------------------------------------------------------
    def _scr(self, name, like: BlockView):
        key = (name, like.B, like.cnt_fixed, like.N)
        if key not in self._scratch:
            self._scratch[key] = torch.empty((like.B, like.cnt_fixed, like.N), dtype=torch.complex128,
                                             device=self.device)
        return BlockView(self._scratch[key], cnt_fixed=like.cnt_fixed, cnt=like.cnt)

    def potential_struct(self, a, g, etab, shift=None, diag=None):
        pot = Potential()
        nt = a.shape[1]
        pot.n_terms = nt
        sh = self.shift if shift is None else shift
        for t in range(nt):
            for i in range(self.M):
                pot.shift[t][i] = int(sh[t, i])
        pot.lam = self.lam.data_ptr()
        pot.etab = etab.data_ptr()
        pot.etab_stride_b = 0 if etab.shape[0] == 1 else etab.stride(0)
        pot.g = g.data_ptr()
        pot.a = a.data_ptr()
        if diag is not None:
            pot.diag = diag.data_ptr()
            pot.diag_stride_b = 0 if diag.shape[0] == 1 else diag.stride(0)
        else:
            pot.diag = None
            pot.diag_stride_b = 0
        pot._keep = (a, g, etab, diag)
        return pot

    def apply_potential_op(self, X: BlockView, Y: BlockView, a, g, etab=None, shift=None,
                           fock_diag=None):
        """Y = fock_diag .* X + back( P(a, g) fwd(X) )   (a: [B][T] complex, g: [B][1+M])."""
        etab = self.etab if etab is None else etab
        k = self.kern
        if self.fused and self.use_fused:
            pot = self.potential_struct(a, g, etab, shift, None)
            if k.hx_fused(self.sp, self.Ve, self.Vo, pot, fock_diag, X, Y):
                return
        if not self.tmodes:
            k.potential_apply(self.sp, self.potential_struct(a, g, etab, shift, fock_diag), X, Y)
            return
        s1 = self._scr('s1', X)
        s2 = self._scr('s2', X)
        src = X
        for i in self.tmodes:
            k.mode_transform(self.sp, i, self.V[i], 1, src, s1, impl=self.impl)
            src = s1
        k.potential_apply(self.sp, self.potential_struct(a, g, etab, shift, None), s1, s2)
        for i in self.tmodes[:-1]:
            k.mode_transform(self.sp, i, self.V[i], 0, s2, s2, impl=self.impl)
        last = self.tmodes[-1]
        if fock_diag is not None:
            k.mode_transform(self.sp, last, self.V[last], 0, s2, Y, ediag=fock_diag, EX=X, impl=self.impl)
        else:
            k.mode_transform(self.sp, last, self.V[last], 0, s2, Y, impl=self.impl)

    def displacement_tables(self):
        """Dt[t] = V diag(etab_t) V^T: Fock-basis matrix of junction term t (the matrix the reference
        gets from qt.displace), only needed to assemble the low-energy preconditioner block."""
        if getattr(self, '_Dt', None) is None:
            V = self.V[0].to(torch.complex128)
            m = self.dims[0]
            e = self.etab[0, :, :m]                              # [T][m]
            self._Dt = torch.einsum('ni,ti,ki->tnk', V, e, V).contiguous()
        return self._Dt

    def to_device(self, arr, dtype):
        return torch.as_tensor(np.ascontiguousarray(arr), dtype=dtype, device=self.device).contiguous()


# sweep points per shared low-block inverse of the two-level preconditioner (see build_lowblock)
LOWBLOCK_GROUP = int(os.environ.get('SQC_LOWBLOCK_GROUP', '0'))     # 0 = choose from the sweep spacing


class HamiltonianOperator:
    """H(b) for a batch of sweep points: loop values [B][nloops] (radians), ng [B or 1][M]."""

    def __init__(self, ops: CircuitOperators, loop_vals, ng):
        self.ops = ops
        self.device = ops.device
        self.B = loop_vals.shape[0]
        a, g = ops.hamiltonian_coeffs(loop_vals)
        self.a = ops.to_device(a, torch.complex128)
        self.g = ops.to_device(g, torch.float64)
        self.diag = ops.lc_diagonal(ng)          # [1 or B][N]
        self.loop_vals = loop_vals
        self.ng = np.asarray(ng, dtype=float).reshape(-1, ops.M)

    def apply(self, X: BlockView, Y: BlockView):
        self.ops.apply_potential_op(X, Y, self.a, self.g, fock_diag=self.diag)

    def supports_lowblock(self):
        return bool(self.ops.fused and self.ops.T >= 1)

    def _auto_lowblock_group(self, max_group=8, budget=0.05):
        """Largest power-of-two group whose members differ by less than ``budget`` (relative) in the
        junction phase factors and the LC diagonal between first and last member: consecutive
        points of a sorted sweep qualify, points in arbitrary order do not."""
        if self.B < 2:
            return 1
        step = 0.0
        if self.a.numel():
            da = (self.a[1:] - self.a[:-1]).abs().amax(dim=0)
            step = float((da / self.a.abs().amax(dim=0).clamp_min(1e-300)).max())
        if self.g.numel():
            dg = (self.g[1:] - self.g[:-1]).abs().amax(dim=0)
            step = max(step, float((dg / self.g.abs().amax(dim=0).clamp_min(1e-300)).max()))
        if self.diag.shape[0] > 1:
            dd = (self.diag[1:] - self.diag[:-1]).abs().amax()
            step = max(step, float(dd / self.diag.abs().amax().clamp_min(1e-300)))
        group = max_group
        while group > 1 and step * group > budget:
            group //= 2
        return group

    def build_lowblock(self, sigma, ns, group=None):
        """Inverse of (H_SS - sigma)/unit on the ns basis states of smallest LC energy
        (two-level preconditioner, see csrc/lowblock.cu).  sigma: [B] float64 (rad/s).

        ``group`` consecutive sweep points (they are neighbours along the sweep) share the block of
        the middle point of the group: the low block varies slowly along a sweep and only has to
        precondition, so one inversion per ``group`` points is enough."""
        ops, kern = self.ops, self.ops.kern
        ns = int(min(ns, 1024, ops.N // 2))
        if group is None:
            group = LOWBLOCK_GROUP or self._auto_lowblock_group()
        group = int(max(1, min(group, self.B)))
        unit = 2 * np.pi * 1e9
        if group == 1:
            a, g, diag, sig, nslot = self.a, self.g, self.diag, sigma.contiguous(), self.B
        else:
            nslot = (self.B + group - 1) // group
            rep = torch.clamp(torch.arange(nslot, device=self.device) * group + group // 2, max=self.B - 1)
            a, g = self.a[rep].contiguous(), self.g[rep].contiguous()
            diag = self.diag if self.diag.shape[0] == 1 else self.diag[rep].contiguous()
            sig = sigma[rep].contiguous()
        pot = ops.potential_struct(a, g, ops.etab)
        band = None
        if ns > 160:
            # large block: charge-major ordering makes H_SS block tridiagonal (csrc/lowblock.cu)
            m, inner = ops.dims
            S = torch.topk(diag, ns, dim=1, largest=False).indices
            key = torch.sort((S % inner) * m + S // inner, dim=1).values
            S = ((key % m) * inner + key // m).to(torch.int32).contiguous()
            qv = (key // m).cpu().numpy()
            rows = qv.shape[0]
            starts = [np.concatenate(([0], np.nonzero(np.diff(r))[0] + 1, [ns])) for r in qv]
            # neighbouring small blocks (the thin ends of the low-energy ellipse) are merged up to one
            # warp's width: still block tridiagonal, fewer sequential substitution steps
            cap = max(32, int(max(np.diff(st).max() for st in starts)))
            merged = []
            for st in starts:
                out = [0]
                for e in st[1:]:
                    if len(out) >= 2 and e - out[-2] <= cap:
                        out[-1] = e
                    else:
                        out.append(e)
                merged.append(np.array(out))
            starts = merged
            nblk = np.array([len(st) - 1 for st in starts], dtype=np.int32)
            nbmax = int(max(np.diff(st).max() for st in starts))
            if nbmax <= 64:
                off = np.zeros((rows, int(nblk.max()) + 1), dtype=np.int32)
                for r, st in enumerate(starts):
                    off[r, :len(st)] = st
                if rows == 1 and nslot > 1:      # LC diagonal shared by all sweep points (flux sweeps)
                    off, nblk = np.repeat(off, nslot, axis=0), np.repeat(nblk, nslot)
                band = (torch.as_tensor(off, device=self.device), torch.as_tensor(nblk, device=self.device), nbmax,
                        torch.argsort(S, dim=1).to(torch.int32).contiguous())
            else:
                ns = 160          # blocks too large for the banded kernels: dense 160 x 160 instead
        if band is None:
            ns = min(ns, 160)
            S = torch.topk(diag, ns, dim=1, largest=False).indices.to(torch.int32).contiguous()
        Ainv = torch.empty((nslot, ns, ns), dtype=torch.complex64, device=self.device)
        kern.lowblock_assemble(ops.sp, S, ops.displacement_tables(), pot, diag, sig, unit, Ainv)
        if band is None:
            kern.hpd_inverse_c64(Ainv)
        else:
            kern.lowband_factor(Ainv, band[0], band[1], band[2])
        return {'S': S, 'Ainv': Ainv, 'unit': unit, 'ns': ns, 'group': group, 'band': band}
