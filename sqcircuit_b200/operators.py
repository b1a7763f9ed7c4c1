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
    use_real = True       # class-level switch: real representation for zero-gate-charge sweeps

    def __init__(self, kern, topo, dims, transform_impl=0):
        """``topo``: the transformed Topology of one circuit, or a list of them -- circuits of ONE graph (same
        edges, element kinds, loops, truncation) with different element values, which then play the role of
        sweep points: point b uses the tables of circuit b (``nb`` = number of circuits)."""
        topos = list(topo) if isinstance(topo, (list, tuple)) else [topo]
        topo = topos[0]
        self.kern = kern
        self.device = kern.dev
        self._scratch = {}
        self._keep = []
        self.topos = topos
        self.nb = len(topos)
        self.topo = topo
        self.dims = [int(d) for d in dims]
        self.M = len(dims)
        self.harm = [bool(np.real(topo.omega[i]) != 0) for i in range(self.M)]
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
        # junction terms: one per row of W (junctions on the same edge share the operator)
        self.T = topo.wTrans.shape[0] if len(topo.W) else 0
        nt = max(self.T, 1)
        # per circuit: impedance / zero-point flux of the harmonic modes (circuit.py:1455-1469, 1585-1601),
        # displacement amplitudes s[t][i] of the junction terms; shared: the charge shifts
        self.Zb = np.zeros((self.nb, self.M))
        self.phi_zp_b = np.zeros((self.nb, self.M))
        self.s_b = np.zeros((self.nb, nt, self.M))
        self.shift = np.zeros((nt, self.M), dtype=int)
        for c, tp in enumerate(topos):
            if [bool(np.real(tp.omega[i]) != 0) for i in range(self.M)] != self.harm or \
                    (tp.wTrans.shape[0] if len(tp.W) else 0) != self.T:
                raise ValueError('the circuits of a batch must share one graph (same harmonic / charge modes '
                                 'and junction terms)')
            shift = np.zeros((nt, self.M), dtype=int)
            for i in range(self.M):
                if self.harm[i]:
                    self.Zb[c, i] = np.sqrt(np.real(tp.cInvTrans[i, i]) / np.real(tp.lTrans[i, i]))
                    self.phi_zp_b[c, i] = (2 * np.pi / unt.Phi0) * np.sqrt(unt.hbar / 2 * self.Zb[c, i])
            for t in range(self.T):
                for i in range(self.M):
                    w = float(np.real(tp.wTrans[t, i]))
                    if self.harm[i]:
                        self.s_b[c, t, i] = self.phi_zp_b[c, i] * w if self.dims[i] > 1 else 0.0
                    else:
                        shift[t, i] = int(np.sign(w))
            if c == 0:
                self.shift = shift
            elif not np.array_equal(shift, self.shift):
                raise ValueError('the circuits of a batch must share the charge shifts of their junction terms')
        self.Z, self.phi_zp, self.s = self.Zb[0], self.phi_zp_b[0], self.s_b[0]
        # element values and flux distribution per circuit (same element order in every circuit)
        self.EJ_b = np.array([[el.si() for _, el, _, _ in tp.jj_keys] for tp in topos]).reshape(self.nb, -1)
        self.Lind_b = np.array([[el.si() for _, el, _ in tp.ind_keys] for tp in topos]).reshape(self.nb, -1)
        self.Bmat_b = np.stack([np.asarray(tp.B, dtype=float).reshape(-1, max(len(tp.loops), 1))
                                if len(tp.loops) else np.zeros((0, 1)) for tp in topos])
        self.etab = self._tables(self.s_b)
        self._etab_half = None
        # parity halves of the x basis for the fused kernel (one harmonic x one charge mode)
        self.fused = (self.M == 2 and self.harm[0] and not self.harm[1] and self.dims[0] > 1
                      and self.T <= 4)
        if self.fused:
            Vfull = self.V[0]
            he, ho = (self.dims[0] + 1) // 2, self.dims[0] // 2
            self.Ve = Vfull[0::2, :he].contiguous()
            self.Vo = Vfull[1::2, :ho].contiguous()
        # generic one-kernel H.X (csrc/hx_generic.cu): harmonic modes lead the Kronecker order, <= 4 of them,
        # <= 32 levels each (circuits of several small modes: transmon-resonator-transmon, 4-mode candidates)
        big = [i for i in range(self.M) if self.dims[i] > 1]
        nh = len(self.tmodes)
        self.generic = (not self.fused and 1 <= nh <= 4 and big[:nh] == self.tmodes
                        and all(self.dims[i] <= 32 for i in self.tmodes) and self.T <= 16)
        if self.generic:
            self.Vcat = torch.cat([self.V[i].reshape(-1) for i in self.tmodes]).contiguous()

    # ---- tables --------------------------------------------------------------------------
    def _tables(self, s):
        """Phase tables exp(i s lam) of every junction term: [nb][T][tlen] for s of shape [nb][T][M]."""
        nt = max(self.T, 1)
        s = np.asarray(s, dtype=float).reshape(-1, nt, self.M)
        etab = torch.empty((s.shape[0], nt, self.tlen), dtype=torch.complex128, device=self.device)
        sd = self.kern.upload(s, torch.float64)
        self.kern.phase_tables(self.sp, nt, self.lam, sd, etab)
        return etab

    def etab_half(self):
        if self._etab_half is None:
            self._etab_half = self._tables(self.s_b / 2)
        return self._etab_half

    # ---- per-sweep-point coefficients (host, tiny) ------------------------------------------
    def ext_flux(self, loop_vals, b_id):
        """phi_ext at inductive element b_id for every sweep point (circuit.py:1759-1786)."""
        if not self.topo.loops:
            return np.zeros(loop_vals.shape[0])
        return np.sum(loop_vals * self.Bmat_b[:, b_id, :], axis=1)

    def kcoef_b(self, ctype, nodes):
        """Node-to-mode coefficients of a coupling operator, one row per circuit [nb][M] (circuit.py:2356-2393;
        like the reference, the first n columns of the full transformation matrix are used)."""
        key = ('kcoef', ctype, tuple(nodes))
        if key in self._scratch:
            return self._scratch[key]
        out = np.zeros((self.nb, self.M))
        n1, n2 = nodes
        for c, t in enumerate(self.topos):
            K = np.linalg.inv(t.C) @ t.R if ctype == 'capacitive' else t.S
            if 0 in nodes:
                row = K[n1 + n2 - 1, :]
            elif ctype == 'capacitive':
                row = K[n2 - 1, :] - K[n1 - 1, :]
            else:
                row = K[n1 - 1, :] - K[n2 - 1, :]
            out[c] = np.real(row[:self.M])
        self._scratch[key] = out
        return out

    def _kcoef(self, ctype, nodes):
        return self.kcoef_b(ctype, nodes)[0]

    def hamiltonian_coeffs(self, loop_vals):
        """a[b][t], g[b][1+M] of the Hamiltonian's potential part for loop values [B][nloops]."""
        B = loop_vals.shape[0]
        a = np.zeros((B, max(self.T, 1)), dtype=complex)
        for j, (_, el, b_id, w_id) in enumerate(self.topo.jj_keys):
            a[:, w_id] += -self.EJ_b[:, j] / 2 * np.exp(1j * self.ext_flux(loop_vals, b_id))
        g = np.zeros((B, 1 + self.M))
        for j, (edge, el, b_id) in enumerate(self.topo.ind_keys):
            coef = (1 / self.Lind_b[:, j]) * self.ext_flux(loop_vals, b_id) * (unt.Phi0 / 2 / np.pi) / np.sqrt(unt.hbar)
            g = g + coef[:, None] * self.phi_lin_b(self.kcoef_b('inductive', edge))
        return a, g

    def phi_lin_b(self, k):
        """Linear x-basis coefficients [nb][1+M] of sum_i k_i phi_i for k [nb or 1][M] (circuit.py:1499-1524):
        phi_i = sqrt(Z_i/2)(a+a^dag) for harmonic modes (0 when truncated to one level), identity for charge
        modes."""
        k = np.asarray(k, dtype=float).reshape(-1, self.M)
        g = np.zeros((max(k.shape[0], self.nb), 1 + self.M))
        for i in range(self.M):
            if self.harm[i]:
                if self.dims[i] > 1:
                    g[:, 1 + i] = k[:, i] * np.sqrt(self.Zb[:, i] / 2)
            else:
                g[:, 0] += k[:, i]
        return g

    def _phi_lin(self, k):
        return self.phi_lin_b(k)[0]

    def lc_params(self, ng):
        """omega [nb][M], q [nb][M][M], ng arrays for sqc_lc_diagonal (circuit.py:1716-1757)."""
        c2 = (2 * unt.e) ** 2 / unt.hbar
        q = np.zeros((self.nb, self.M, self.M))
        omega = np.zeros((self.nb, self.M))
        for c, t in enumerate(self.topos):
            for i in range(self.M):
                if self.harm[i]:
                    omega[c, i] = np.real(t.omega[i])
                for j in range(i, self.M):
                    cij = float(np.real(t.cInvTrans[i, j]))
                    if i == j:
                        if not self.harm[i]:
                            q[c, i, i] = c2 * cij / 2
                    elif cij != 0:
                        small = abs(cij) <= 1e-10 * np.sqrt(abs(t.cInvTrans[i, i] * t.cInvTrans[j, j]))
                        if self.harm[i] or self.harm[j]:
                            if not small and self.dims[i] > 1 and self.dims[j] > 1:
                                raise NotImplementedError(
                                    'capacitive cross coupling that involves a harmonic mode survived the '
                                    'mode transformation; not supported by the matrix-free operator')
                        else:
                            q[c, i, j] = c2 * cij
        return omega, q, np.asarray(ng, dtype=float).reshape(-1, self.M)

    def lc_diagonal(self, ng):
        omega, q, ng = self.lc_params(ng)
        B = max(ng.shape[0], self.nb)
        dev = self.device
        out = torch.empty((B, self.N), dtype=torch.float64, device=dev)
        om = self.kern.upload(np.broadcast_to(omega, (B, self.M)), torch.float64)
        qq = self.kern.upload(np.broadcast_to(q, (B, self.M, self.M)), torch.float64)
        ngd = self.kern.upload(np.broadcast_to(ng, (B, self.M)), torch.float64)
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
        if self.generic and self.use_fused:
            pot = self.potential_struct(a, g, etab, shift, None)
            if k.hx_generic(self.sp, self.Vcat, pot, fock_diag, X, Y):
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

    # ---- real (time-reversal adapted) representation, see csrc/hx_real.cu ------------------------
    def supports_real(self, ng):
        """True when H(b) can be solved in the real representation: one harmonic x one charge mode
        (the fused layout), zero gate charge on the charge mode, sizes the kernel accepts."""
        if not (self.fused and self.use_fused and self.use_real and self.nb == 1):
            return False
        ng = np.asarray(ng, dtype=float).reshape(-1, self.M)
        if np.any(ng[:, 1] != 0.0):
            return False
        m, inner = self.dims
        if inner % 2 == 0 or inner > 127 or m > 112:
            return False
        # shared memory of one team of sqc_hx_fused_real (must mirror the launch code)
        kp = ((m + 1) // 2 + 3) & ~3
        vld = kp if kp % 8 == 4 else kp + 4
        pp = 2 * inner
        while pp % 16 not in (4, 12):
            pp += 2
        team = kp * pp + 16 + ((m + 1) & ~1) + 2 * 4 * m + 8
        nc = (inner - 1) // 2
        return (2 * kp * vld + team + ((m + 1) & ~1) + ((nc + 2) & ~1) + 4) * 8 <= 227 * 1024

    def real_index_map(self):
        """Complex charge index c = nc + n of every real column j (j = 0: n = 0, 1..nc: Re plane,
        nc+1..2nc: Im plane): the LC diagonal of real amplitude (i, j) is d_LC[i][c(j)]."""
        inner = self.dims[1]
        nc = (inner - 1) // 2
        return np.concatenate(([nc], nc + np.arange(1, nc + 1), nc + np.arange(1, nc + 1)))

    def real_to_complex(self, Xr, out=None):
        """[B][k][Nc] real-representation vectors -> [B][k][N] complex vectors in the reference's basis."""
        Xr = Xr if Xr.is_contiguous() else Xr.contiguous()
        if out is None:
            out = torch.empty((Xr.shape[0], Xr.shape[1], self.N), dtype=torch.complex128, device=self.device)
        self.kern.real_to_complex(self.sp, BlockView(Xr), BlockView(out))
        return out

    def displacement_tables(self):
        """Dt[t] = V diag(etab_t) V^T: Fock-basis matrix of junction term t (the matrix the reference
        gets from qt.displace), only needed to assemble the low-energy preconditioner block."""
        if getattr(self, '_Dt', None) is None:
            V = self.V[0].to(torch.complex128)
            m = self.dims[0]
            e = self.etab[0, :, :m]                              # [T][m]
            self._Dt = torch.einsum('ni,ti,ki->tnk', V, e, V).contiguous()
        return self._Dt

    def displacement_tables_generic(self):
        """Per harmonic mode (more than one level) the Fock-basis matrix of every junction term on that mode,
        V_i diag(etab_t^(i)) V_i^T, concatenated in mode order: [nb][T][sum m_i^2] (sqc_lowblock_assemble_generic)."""
        if getattr(self, '_Dtg', None) is None:
            nt = max(self.T, 1)
            parts = []
            for i in self.tmodes:
                m, o = self.dims[i], int(self.toff[i])
                V = self.V[i].to(torch.complex128)
                e = self.etab[:, :, o:o + m]                                  # [nb][T][m]
                parts.append(torch.einsum('ni,bti,ki->btnk', V, e, V).reshape(e.shape[0], nt, m * m))
            if not parts:
                parts = [torch.zeros((self.etab.shape[0], nt, 1), dtype=torch.complex128, device=self.device)]
            self._Dtg = torch.cat(parts, dim=2).contiguous()
        return self._Dtg

    def to_device(self, arr, dtype):
        return self.kern.upload(arr, dtype)


# sweep points per shared low-block inverse of the two-level preconditioner (see build_lowblock)
LOWBLOCK_GROUP = 0     # 0 = choose from the sweep spacing (tests / tools may set a fixed group size)


class HamiltonianOperator:
    """H(b) for a batch of sweep points: loop values [B][nloops] (radians), ng [B or 1][M]."""

    def __init__(self, ops: CircuitOperators, loop_vals, ng):
        self.ops = ops
        self.device = ops.device
        self.B = loop_vals.shape[0]
        a, g = ops.hamiltonian_coeffs(loop_vals)
        self.a_host, self.g_host = a, g          # host copies: decisions on them cost no device round trip
        self.a = ops.to_device(a, torch.complex128)
        self.g = ops.to_device(g, torch.float64)
        self.diag = ops.lc_diagonal(ng)          # [1 or B][N]
        self.loop_vals = loop_vals
        self.ng = np.asarray(ng, dtype=float).reshape(-1, ops.M)

    def apply(self, X: BlockView, Y: BlockView):
        self.ops.apply_potential_op(X, Y, self.a, self.g, fock_diag=self.diag)

    def dense_matrix(self):
        """H(b) of every point as a dense [B][N][N] matrix, H[b][r][q] = <r|H(b)|q> (small Hilbert spaces: the
        dense eigensolver path): float64 for layouts without charge modes, else complex128."""
        ops = self.ops
        # without charge modes H is real symmetric in the Fock basis (a_t D_t + h.c. = 2 Re(a_t D_t))
        real = not any((not h) and d > 1 for h, d in zip(ops.harm, ops.dims))
        H = torch.empty((self.B, ops.N, ops.N), dtype=torch.float64 if real else torch.complex128, device=self.device)
        pot = ops.potential_struct(self.a, self.g, ops.etab)
        ops.kern.hamiltonian_dense(ops.sp, ops.displacement_tables_generic(), pot, self.diag, H)
        return H

    def supports_lowblock(self):
        # harmonic x charge layouts of one circuit: the banded block of up to 512 states; every other layout
        # (several modes, batches of circuits): the dense block of up to 160 states from per-mode tables
        return bool(self.ops.T >= 1)

    def _auto_lowblock_group(self, max_group=8, budget=0.05):
        """Largest power-of-two group whose members differ by less than ``budget`` (relative) in the
        junction phase factors and the LC diagonal between first and last member: consecutive
        points of a sorted sweep qualify, points in arbitrary order do not."""
        if self.B < 2:
            return 1
        step = 0.0
        a, g = getattr(self, 'a_host', None), getattr(self, 'g_host', None)
        if a is None:
            a, g = self.a.cpu().numpy(), self.g.cpu().numpy()
        if a.size:
            da = np.abs(a[1:] - a[:-1]).max(axis=0)
            step = float((da / np.maximum(np.abs(a).max(axis=0), 1e-300)).max())
        if g.size:
            dg = np.abs(g[1:] - g[:-1]).max(axis=0)
            step = max(step, float((dg / np.maximum(np.abs(g).max(axis=0), 1e-300)).max()))
        if self.diag.shape[0] > 1:
            dd = (self.diag[1:] - self.diag[:-1]).abs().amax()
            step = max(step, float(dd / self.diag.abs().amax().clamp_min(1e-300)))
        group = max_group
        while group > 1 and step * group > budget:
            group //= 2
        return group

    def _band_layout(self, qv, S, nslot, nbmax_limit=64):
        """Block offsets of a low-energy index set ordered by the value ``qv`` [rows][ns] (ascending) of one
        charge mode: every junction term shifts that charge by at most one, so H_SS is block tridiagonal in
        the blocks of equal qv.  Returns (off [nslot][nblk + 1], nblk [nslot], largest block, gather
        permutation) or None when a block exceeds what the banded kernels hold in shared memory."""
        rows, ns = qv.shape
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
        if nbmax > nbmax_limit:
            return None
        off = np.zeros((rows, int(nblk.max()) + 1), dtype=np.int32)
        for r, st in enumerate(starts):
            off[r, :len(st)] = st
        if rows == 1 and nslot > 1:      # LC diagonal shared by all sweep points (flux sweeps)
            off, nblk = np.repeat(off, nslot, axis=0), np.repeat(nblk, nslot)
        return (self.ops.kern.upload(off), self.ops.kern.upload(nblk), nbmax,
                torch.argsort(S, dim=1).to(torch.int32).contiguous())

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
        self._last_group = group
        unit = 2 * np.pi * 1e9
        if not (ops.fused and ops.nb == 1):
            return self._build_lowblock_generic(sigma, min(ns, 512), group, unit)
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
            band = self._band_layout(kern.download(key // m), S, nslot)
            if band is None:
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


    def _build_lowblock_generic(self, sigma, ns, group, unit):
        """Low block of any layout: per-mode displacement tables, generic assembly in complex64.  Per slot: the
        index set of its own LC diagonal, the tables of its own circuit.

        With a charge mode in the layout the index set is ordered by the value of one charge mode: every
        junction term shifts a charge by at most one, so H_SS is block tridiagonal in the blocks of equal
        charge and the banded kernels (block LDL^H, csrc/lowblock.cu) take up to 512 states -- measured on the
        transmon - resonator - transmon grid: a fifth fewer iterations than the dense 160-state block.  Layouts
        of harmonic modes only, or blocks too large for shared memory, keep the dense inverse of 160 states."""
        ops, kern = self.ops, self.ops.kern
        ns = int(min(ns, ops.N // 2))
        nslot = (self.B + group - 1) // group
        if group == 1:
            rep = None
            a, g, diag, sig = self.a, self.g, self.diag, sigma.contiguous()
        else:
            rep = torch.clamp(torch.arange(nslot, device=self.device) * group + group // 2, max=self.B - 1)
            a, g = self.a[rep].contiguous(), self.g[rep].contiguous()
            diag = self.diag if self.diag.shape[0] == 1 else self.diag[rep].contiguous()
            sig = sigma[rep].contiguous()
        Dt = ops.displacement_tables_generic()
        if Dt.shape[0] > 1 and rep is not None:
            Dt = Dt[rep].contiguous()
        band, S = None, None
        cmodes = [i for i in range(ops.M) if not ops.harm[i] and ops.dims[i] > 1]
        if ns > 160 and cmodes:
            low = torch.topk(diag, ns, dim=1, largest=False).indices          # [rows][ns], ascending energy
            for nb in range(ns, 160, -64):
                # the blocking mode whose largest block is smallest; fewer states until the blocks fit
                best = None
                for i in cmodes:
                    stride = int(np.prod(ops.dims[i + 1:]))
                    key = torch.sort(((low[:, :nb] // stride) % ops.dims[i]) * ops.N + low[:, :nb], dim=1).values
                    Sb = (key % ops.N).to(torch.int32).contiguous()
                    lay = self._band_layout(kern.download(key // ops.N), Sb, nslot)
                    if lay is not None and (best is None or lay[2] < best[0][2]):
                        best = (lay, Sb)
                if best is not None:
                    (band, S), ns = best, nb
                    break
        if band is None:
            ns = min(ns, 160)
            S = torch.topk(diag, ns, dim=1, largest=False).indices.to(torch.int32).contiguous()
        pot = ops.potential_struct(a, g, ops.etab)
        Ainv = torch.empty((nslot, ns, ns), dtype=torch.complex64, device=self.device)
        kern.lowblock_assemble_generic(ops.sp, S, Dt, pot, diag, sig, unit, Ainv)
        if band is None:
            kern.hpd_inverse_c64(Ainv)
        else:
            kern.lowband_factor(Ainv, band[0], band[1], band[2])
        return {'S': S, 'Ainv': Ainv, 'unit': unit, 'ns': ns, 'group': group, 'band': band}


class RealHamiltonianOperator(HamiltonianOperator):
    """H(b) in the real (time-reversal adapted) representation (csrc/hx_real.cu): at zero gate charge H
    commutes with T = (complex conjugation) x (n -> -n), its eigenvectors satisfy z(i, -n) = conj z(i, n)
    and are stored as N real amplitudes -- half the bytes and half the H.X flops, and real Gram /
    Rayleigh-Ritz arithmetic.  Blocks are complex128 tensors [B][nvec][Nc], Nc = ceil(N / 2): two real
    amplitudes per element.  ``diag`` is the LC diagonal in the real layout, [1][2 Nc]."""
    real = True

    def __init__(self, ops: CircuitOperators, loop_vals, ng, base: HamiltonianOperator = None):
        if base is None:
            super().__init__(ops, loop_vals, ng)
        else:           # share the coefficient tables of an existing complex operator of the same points
            self.ops, self.device, self.B = ops, ops.device, base.B
            self.a, self.g, self.diag, self.loop_vals, self.ng = base.a, base.g, base.diag, base.loop_vals, base.ng
            self.a_host, self.g_host = getattr(base, 'a_host', None), getattr(base, 'g_host', None)
        m, inner = ops.dims
        self.N = ops.N
        self.Nc = (ops.N + 1) // 2
        self.diag_c = self.diag
        cmap = ops.kern.upload(ops.real_index_map())
        dr = self.diag_c.reshape(-1, m, inner)[:, :, cmap].reshape(-1, ops.N)
        if 2 * self.Nc != ops.N:       # padding amplitude: never selected as a low-energy state, stays zero
            dr = torch.cat([dr, torch.full((dr.shape[0], 1), 1e300, dtype=dr.dtype, device=self.device)], dim=1)
        self.diag = dr.contiguous()
        # separable form d(i, j) = dh[i] + dc[j] (the LC energies of a harmonic and a charge mode add up): the
        # kernel then applies the charge part in the x basis and needs one scalar per row in its epilogue
        self.diag_sep = None
        if self.diag.shape[0] == 1:
            d2 = dr[0, :ops.N].reshape(m, inner)
            dh, dc = d2[:, 0] - d2[0, 0], d2[0, :]
            chk = ops.kern.download(torch.stack([(d2 - dh[:, None] - dc[None, :]).abs().max(), d2.abs().max()]))
            if float(chk[0]) <= 4e-16 * float(chk[1]):
                self.diag_sep = torch.cat([dh, dc]).contiguous()

    def apply(self, X: BlockView, Y: BlockView):
        ops = self.ops
        pot = ops.potential_struct(self.a, self.g, ops.etab)
        if not ops.kern.hx_fused_real(ops.sp, ops.Ve, ops.Vo, pot, self.diag, X, Y, fdiag_sep=self.diag_sep):
            raise RuntimeError('sqc_hx_fused_real rejected a layout that supports_real() accepted')

    def to_complex(self, Xr, out=None):
        return self.ops.real_to_complex(Xr, out)

    def supports_lowblock(self):
        return bool(self.ops.T >= 1)

    def build_lowblock(self, sigma, ns, group=None):
        """Two-level preconditioner block in the real representation.  The index set is the ns real basis
        states of smallest LC energy, taken in whole (Fock i, |charge| n) units so that the cos- and sin-type
        partners of a charge stay together, ordered charge-major: H_SS is block tridiagonal in n."""
        ops, kern = self.ops, self.ops.kern
        m, inner = ops.dims
        nc = (inner - 1) // 2
        ns = int(min(ns, 1024, ops.N // 2))
        if group is None:
            group = LOWBLOCK_GROUP or self._auto_lowblock_group()
        group = int(max(1, min(group, self.B)))
        unit = 2 * np.pi * 1e9
        nslot = (self.B + group - 1) // group
        if group == 1:
            a, g, sig = self.a, self.g, sigma.contiguous()
        else:
            rep = torch.clamp(torch.arange(nslot, device=self.device) * group + group // 2, max=self.B - 1)
            a, g, sig = self.a[rep].contiguous(), self.g[rep].contiguous(), sigma[rep].contiguous()
        key = ('lowset', ns)
        if key not in ops._scratch:
            ops._scratch[key] = self._real_low_set(ns, m, inner, nc)
        S, off, nblk, nbmax, perm = ops._scratch[key]
        ns = S.shape[1]
        pot = ops.potential_struct(a, g, ops.etab)
        # banded: float storage (sqc_lowband_*_real); dense: complex64 with zero imaginary parts
        Ainv = torch.empty((nslot, ns, ns), dtype=torch.complex64 if off is None else torch.float32,
                           device=self.device)
        kern.lowblock_assemble(ops.sp, S, ops.displacement_tables(), pot, self.diag, sig, unit, Ainv, real=True)
        if off is None:
            kern.hpd_inverse_c64(Ainv)
            band = None
        else:
            offs = off.expand(nslot, -1).contiguous()
            nb = nblk.expand(nslot).contiguous()
            kern.lowband_factor(Ainv, offs, nb, nbmax)
            band = (offs, nb, nbmax, perm)
        return {'S': S, 'Ainv': Ainv, 'unit': unit, 'ns': ns, 'group': group, 'band': band, 'real': True}

    def _real_low_set(self, ns, m, inner, nc):
        """Index set (device int32 [1][ns']), block offsets, block count, largest block, address-sorting
        permutation.  Host work, done once per circuit and size (the LC diagonal does not depend on the
        sweep point at zero gate charge)."""
        d = self.ops.kern.download(self.diag_c[0].reshape(m, inner))
        units = sorted(((d[i, nc + n], n, i) for i in range(m) for n in range(nc + 1)))
        chosen, cnt = [], 0
        for e, n, i in units:
            w = 1 if n == 0 else 2
            if cnt + w > ns:
                break
            chosen.append((n, i))
            cnt += w
        chosen.sort()                                   # charge-major, Fock index inside
        idx, blocks = [], []
        for n, i in chosen:
            if not blocks or blocks[-1][0] != n:
                blocks.append([n, len(idx)])
            idx.append(i * inner + n)
            if n:
                idx.append(i * inner + nc + n)
        starts = np.array([b[1] for b in blocks] + [len(idx)])
        S = self.ops.kern.upload(np.array(idx, dtype=np.int32)).reshape(1, -1).contiguous()
        if len(idx) <= 160:
            return S, None, None, 0, None
        # neighbouring small blocks (the thin ends of the low-energy ellipse) are merged up to one warp's
        # width: still block tridiagonal, fewer sequential substitution steps
        cap = max(32, int(np.diff(starts).max()))
        out = [0]
        for e in starts[1:]:
            if len(out) >= 2 and e - out[-2] <= cap:
                out[-1] = e
            else:
                out.append(e)
        starts = np.array(out)
        nbmax = int(np.diff(starts).max())
        if nbmax > 64:
            # blocks too large for the banded kernels: dense block on the 160 lowest states instead
            return self._real_low_set(160, m, inner, nc)
        off = self.ops.kern.upload(starts.astype(np.int32)).reshape(1, -1)
        nblk = self.ops.kern.upload(np.array([len(starts) - 1], dtype=np.int32))
        perm = torch.argsort(S, dim=1).to(torch.int32).contiguous()
        return S, off, nblk, nbmax, perm
