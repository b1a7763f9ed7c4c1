"""CPU oracle for the SQcircuit diagonalisation hot path.

THIS FILE IS TEST INFRASTRUCTURE.  It is a plain numpy/scipy restatement of
the reference algorithm (stanfordLINQS/SQcircuit) used only as the checker in
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.  Nothing under ``sqcircuit_b200/`` imports it.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this oracle
against (a) the known-answer numbers inside the reference's own tests
(``SQcircuit/tests/test_qubits.py``: zero-pi, fluxonium, transmon, coupled
fluxonium-transmon, inductively shunted, loop-issue circuits, incl. the
fluxonium/transmon decay-rate tables), (b) ``tests/data/spectra/zeropi_paper.npy``
and (c) fixtures produced by running the reference's unmodified ``circuit.py``
in this container (``tests/golden/make_golden.py``; QuTiP is replaced by the
scipy stand-in in ``tests/golden/ref_shims`` because it is not installed).

Every function cites the reference lines it follows (paths relative to
``/root/reference/SQcircuit``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg
import scipy.sparse as sp
import scipy.sparse.linalg
from scipy.special import kve

# ---------------------------------------------------------------------------
# constants -- units.py:17-29, settings.py:4-8, noise.py:10-15
# ---------------------------------------------------------------------------
HBAR = 1.0545718e-34
PHI0 = 2.067833e-15
E_CH = 1.6021766e-19
K_B = 1.38e-23
FARAD = {'F': 1, 'mF': 1.0e-3, 'uF': 1.0e-6, 'nF': 1.0e-9, 'pF': 1.0e-12, 'fF': 1.0e-15}
HENRY = {'H': 1.0, 'mH': 1.0e-3, 'uH': 1.0e-6, 'nH': 1.0e-9, 'pH': 1.0e-12, 'fH': 1.0e-15}
FREQ = {'Hz': 1.0, 'kHz': 1.0e3, 'MHz': 1.0e6, 'GHz': 1.0e9, 'THz': 1.0e12}
UNIT_FREQ = 1.0e9
ACC_SING = 1e-11
ACC_GS = 1e-6
ACC_HAR = 1e-11
ENV = {'T': 0.015, 'omega_low': 2 * np.pi, 'omega_high': 2 * np.pi * 3e9, 't_exp': 10e-6}


# ---------------------------------------------------------------------------
# elements -- elements.py:108-726
# ---------------------------------------------------------------------------
class OLoop:
    """elements.py:599-690 (value stored as 2*pi*flux; A stored as 2*pi*A)."""

    def __init__(self, value: float = 0.0, A: float = 1e-6):
        self.val = value * 2 * np.pi
        self.A = A * 2 * np.pi
        self.indices: List[int] = []
        self.k_rows: List[list] = []

    def set_flux(self, value: float) -> None:
        self.val = value * 2 * np.pi

    def reset(self) -> None:
        self.indices, self.k_rows = [], []

    def g_vector(self) -> np.ndarray:
        # elements.py:683-690
        k = remove_dependent_columns(np.array(self.k_rows))
        b = np.zeros((1, k.shape[0]))
        b[0, 0] = 1
        return (np.linalg.inv(np.concatenate((b, k.T), axis=0)) @ b.T).T


def _default_q_cap(omega):
    # elements.py:228-232
    return 1e6 * (2 * np.pi * 6e9 / np.abs(omega)) ** 0.7


def _log_k0(x):
    # functions.py:135-138
    return np.log(kve(0, x)) - x


def _log_sinh(x):
    # functions.py:52-56
    return x + np.log(1 - np.exp(-2 * x)) - np.log(2)


def _default_q_ind(omega, T):
    # elements.py:379-391
    alpha = HBAR * 2 * np.pi * 0.5e9 / (2 * K_B * T)
    beta = HBAR * omega / (2 * K_B * T)
    return 500e6 * np.exp(_log_k0(alpha) + _log_sinh(alpha) - _log_k0(beta) - _log_sinh(beta))


def _default_y_junc(delta, x):
    # elements.py:573-596
    def y(omega, T):
        alpha = HBAR * omega / (2 * K_B * T)
        return (np.sqrt(2 / np.pi)
                * (8 / (delta * 1.6e-19) / (HBAR * 2 * np.pi / E_CH ** 2))
                * (2 * (delta * 1.6e-19) / HBAR / omega) ** 1.5
                * x * np.sqrt(alpha) * kve(0, alpha)
                * (1 - np.exp(-2 * alpha)) / 2)
    return y


class OCap:
    """elements.py:108-232.  ``value`` kept in farad."""
    kind = 'C'

    def __init__(self, value, unit='GHz', Q='default'):
        if unit in FARAD:
            self.value = value * FARAD[unit]
        else:
            self.value = E_CH ** 2 / 2 / (value * FREQ[unit] * (2 * np.pi * HBAR))
        if Q == 'default':
            self.Q = _default_q_cap
        elif isinstance(Q, (int, float)):
            self.Q = lambda omega, _q=Q: _q
        else:
            self.Q = Q


class OInd:
    """elements.py:247-424.  ``value`` kept in henry."""
    kind = 'L'

    def __init__(self, value, unit='GHz', loops=None, cap=None, Q='default'):
        if unit in HENRY:
            self.value = value * HENRY[unit]
        else:
            self.value = (PHI0 / 2 / np.pi) ** 2 / (value * FREQ[unit] * (2 * np.pi * HBAR))
        self.loops = loops or []
        self.cap = cap if cap is not None else OCap(1e-20, 'F', Q=None)
        if Q == 'default':
            self.Q = _default_q_ind
        elif isinstance(Q, (int, float)):
            self.Q = lambda omega, T, _q=Q: _q
        else:
            self.Q = Q

    def cap_for_flux_dist(self, flux_dist):
        # elements.py:406-413
        return {'all': self.cap.value, 'junctions': 1e20, 'inductors': 1e-20}[flux_dist]


class OJJ:
    """elements.py:427-596.  ``value`` kept in rad/s."""
    kind = 'J'

    def __init__(self, value, unit='GHz', loops=None, cap=None, A=1e-7, x=3e-06,
                 delta=3.4e-4, Y='default'):
        self.value = value * FREQ[unit] * 2 * np.pi
        self.loops = loops or []
        self.cap = cap if cap is not None else OCap(1e-20, 'F', Q=None)
        self.A = A
        self.Y = _default_y_junc(delta, x) if Y == 'default' else Y

    def cap_for_flux_dist(self, flux_dist):
        # elements.py:564-571
        return {'all': self.cap.value, 'junctions': 1e-20, 'inductors': 1e20}[flux_dist]


# ---------------------------------------------------------------------------
# small linear-algebra helpers -- functions.py:309-336
# ---------------------------------------------------------------------------
def remove_dependent_columns(matrix, tol=1e-5):
    nz = matrix[:, np.any(matrix != 0, axis=0)]
    _, s, _ = np.linalg.svd(nz)
    return nz[:, np.where(np.abs(s) > tol)[0]]


def _edge_w(edge, n):
    # circuit.py:78-92
    i1, i2 = edge
    w = (n + 1) * [0]
    if i1 == 0 or i2 == 0:
        w[i1 + i2] += 1
    else:
        w[i1] += 1
        w[i2] -= 1
    return w[1:]


def _edge_mat(edge, n):
    # circuit.py:94-113
    m = np.zeros((n, n))
    if 0 in edge:
        i = edge[0] + edge[1] - 1
        m[i, i] = 1
    else:
        i, j = edge[0] - 1, edge[1] - 1
        m[i, i] = m[j, j] = 1
        m[i, j] = m[j, i] = -1
    return m


# ---------------------------------------------------------------------------
# circuit
# ---------------------------------------------------------------------------
class OCircuit:
    """Restatement of ``Circuit`` (NumPy engine) -- circuit.py:248-2921."""

    def __init__(self, elements: Dict[Tuple[int, int], list], flux_dist: str = 'junctions',
                 rng: Optional[np.random.Generator] = None):
        self.elements = dict(elements)
        self.flux_dist = flux_dist
        self.n = max(max(self.elements))
        self.rng = rng or np.random.default_rng(0)
        self.loops: List[OLoop] = []
        self.ind_keys: list = []   # (edge, el, b_id)
        self.jj_keys: list = []    # (edge, el, b_id, w_id)
        self._lcwb()
        self.R, self.S = np.eye(self.n), np.eye(self.n)
        self.cInvTrans = np.linalg.inv(self.C)
        self.lTrans = self.L.copy()
        self.wTrans = self.W.copy()
        self.omega = np.zeros(self.n)
        self._transform()
        if len(self.W) != 0:
            self._remove_uncoupled()
        self.ng = {i: 0.0 for i in range(self.n) if self.omega[i] == 0}
        self.ng_A = {i: 1e-4 for i in self.ng}
        self.m: List[int] = []
        self.efreqs_rad = None
        self.evecs = None

    # -- circuit.py:743-835 --------------------------------------------------
    def _lcwb(self):
        n = self.n
        C = np.zeros((n, n))
        L = np.zeros((n, n))
        W, K, c_edge = [], [], []
        b_id = w_id = 0
        n_j_no_l = 0
        self.partial_mats = {}
        for edge, els in self.elements.items():
            w = _edge_w(edge, n)
            mat = _edge_mat(edge, n)
            caps, inds, jjs = [], [], []
            for el in els:
                if el.kind == 'C':
                    caps.append(el)
                    self.partial_mats[el] = self.partial_mats.get(el, 0) + mat
                else:
                    (inds if el.kind == 'L' else jjs).append(el)
                    caps.append(el.cap)
                    if el.kind == 'L':
                        self.ind_keys.append((edge, el, b_id))
                        # elements.py:415-424
                        self.partial_mats[el] = self.partial_mats.get(el, 0) + mat / el.value ** 2
                    else:
                        self.jj_keys.append((edge, el, b_id, w_id))
                    for lp in el.loops:
                        if lp not in self.loops:
                            lp.reset()
                            self.loops.append(lp)
                        lp.indices.append(b_id)
                        lp.k_rows.append(w)
                    b_id += 1
                    K.append(w)
                    c_edge.append(el.cap_for_flux_dist(self.flux_dist))
            if jjs and not inds:
                n_j_no_l += 1
            C += sum(c.value for c in caps) * mat
            L += sum(1 / i.value for i in inds) * mat
            if jjs:
                W.append(w)
                w_id += 1
        self.C, self.L, self.W = C, L, np.array(W)
        self.num_jun_without_ind = n_j_no_l
        self.B = np.array([])
        K = remove_dependent_columns(np.array(K)) if K else np.array(K)
        if len(K):
            x = K.T @ np.diag(c_edge)
            for lp in self.loops:
                g = np.zeros((1, b_id))
                g[0, lp.indices] = lp.g_vector()
                x = np.concatenate((x, g), axis=0)
            nl = len(self.loops)
            if nl:
                imat = np.concatenate((np.zeros((b_id - nl, nl)), np.eye(nl)), axis=0)
                self.B = np.around(np.linalg.inv(x) @ imat, 5)

    # -- circuit.py:882-900 --------------------------------------------------
    def _apply(self, s, r):
        self.cInvTrans = r.T @ self.cInvTrans @ r
        self.lTrans = s.T @ self.lTrans @ s
        if len(self.W) != 0:
            self.wTrans = self.wTrans @ s
        self.S = self.S @ s
        self.R = self.R @ r

    def is_charge(self, i):
        return self.omega[i] == 0

    # -- circuit.py:902-939 --------------------------------------------------
    def _t1(self):
        c_root = scipy.linalg.sqrtm(self.C)
        c_root_inv = np.linalg.inv(c_root)
        l_root = scipy.linalg.sqrtm(self.L)
        _, D, U = np.linalg.svd(l_root @ c_root_inv)
        if np.max(D) == 0:
            D = np.diag(np.eye(self.n))
            sing = list(range(self.n))
        else:
            ev, _ = np.linalg.eig(self.L)
            ns = len(ev[ev / np.max(ev) < ACC_SING])
            sing = list(range(self.n - ns, self.n))
            D[sing] = np.max(D)
        s1 = c_root_inv @ U.T @ np.diag(np.sqrt(D))
        r1 = np.linalg.inv(s1).T
        self._apply(s1, r1)
        self.lTrans[sing, sing] = 0
        self.S1 = s1

    # -- circuit.py:941-1008 -------------------------------------------------
    @staticmethod
    def _indep_rows(a):
        an = a / (np.linalg.norm(a, axis=1).reshape(a.shape[0], 1) + 1e-9)
        order = np.argsort(-np.linalg.norm(a, axis=1))
        basis, idx = [], []
        for i in order:
            v = an[i, :]
            vp = v - sum([np.dot(v, e) * e for e in basis])
            if (np.abs(vp) > ACC_GS).any():
                idx.append(i)
                basis.append(vp / np.linalg.norm(vp))
        return idx, basis

    def _round01(self, w):
        r = w.copy()
        if self.num_jun_without_ind == 0:
            r[:, self.omega == 0] = 0
        cw = r[:, self.omega == 0]
        cw[np.abs(cw) < ACC_GS] = 0
        cw[np.abs(cw - 1) < ACC_GS] = 1
        cw[np.abs(cw + 1) < ACC_GS] = -1
        r[:, self.omega == 0] = cw
        return r

    # -- circuit.py:1047-1121 ------------------------------------------------
    def _t2(self):
        w_red = remove_dependent_columns(self.W.T).T @ self.S1
        w_charge = w_red[:, self.omega == 0].copy()
        nq = w_charge.shape[1]
        for _ in range(50):
            if nq != 0 and self.num_jun_without_ind != 0:
                idx, basis = self._indep_rows(w_charge)
                rnd: list = []
                while len(basis) != nq:
                    rnd = list(np.linalg.norm(w_charge, 'fro')
                               * self.rng.standard_normal((nq - len(basis), nq)))
                    _, basis = self._indep_rows(np.array(basis + rnd))
                f = np.array(list(w_charge[idx, :]) + rnd)
                s2 = scipy.linalg.block_diag(np.eye(self.n - nq), np.linalg.inv(f))
                r2 = np.linalg.inv(s2.T)
            else:
                s2 = np.eye(self.n)
                r2 = s2
            cur = self._round01(self.wTrans @ s2)
            ok = all(a in (0, 1) for j in range(self.n) if self.is_charge(j)
                     for a in np.abs(cur[:, j]))
            if ok:
                self._apply(s2, r2)
                self.wTrans = self._round01(self.wTrans)
                return
        raise RuntimeError('Gram-Schmidt failed')

    def _impedance(self, i):
        return np.sqrt(self.cInvTrans[i, i] / self.lTrans[i, i])

    def _phi_zp(self, i):
        # circuit.py:1585-1601
        return (2 * np.pi / PHI0) * np.sqrt(HBAR / 2 * self._impedance(i))

    # -- circuit.py:1123-1168 ------------------------------------------------
    def _t3(self):
        s3 = np.eye(self.n)
        for j in range(self.n):
            if self.is_charge(j):
                continue
            if len(self.W) != 0:
                al = np.abs(self._phi_zp(j) * self.wTrans[:, j])
                self.wTrans[:, j][al < ACC_HAR] = 0
                if np.max(al) > ACC_HAR:
                    s3[j, j] = 1 / np.abs(self.wTrans[np.argmax(al), j])
                else:
                    s3[j, j] = 1 / np.max(np.abs(self.S[:, j]))
            else:
                s3[j, j] = 1 / np.max(np.abs(self.S[:, j]))
        self._apply(s3, np.linalg.inv(s3.T))

    # -- circuit.py:1170-1186 ------------------------------------------------
    def _transform(self):
        self._t1()
        self.omega = np.sqrt(np.diag(self.cInvTrans) * np.diag(self.lTrans))
        if len(self.W) != 0:
            self._t2()
        self._t3()

    # -- circuit.py:837-863 --------------------------------------------------
    def _remove_uncoupled(self):
        nJ = self.wTrans.shape[0]
        unc = list(np.where(np.sum(self.wTrans == 0, axis=0) == nJ)[0])
        self.n -= len(unc)
        for ax in (0, 1):
            self.cInvTrans = np.delete(self.cInvTrans, unc, axis=ax)
            self.lTrans = np.delete(self.lTrans, unc, axis=ax)
        self.wTrans = np.delete(self.wTrans, unc, axis=1)
        self.omega = np.delete(self.omega, unc)

    # -- circuit.py:1340-1379 ------------------------------------------------
    def set_trunc_nums(self, nums: Sequence[int]):
        assert len(nums) == self.n
        self.m = [2 * nums[i] - 1 if self.is_charge(i) else nums[i] for i in range(self.n)]
        self._build_ops()

    def set_charge_offset(self, mode: int, ng: float):
        assert mode - 1 in self.ng
        self.ng[mode - 1] = ng
        if self.m:
            self._build_ops()

    # -- isolated operators: circuit.py:1471-1583 ------------------------------
    def _q_iso(self, i):
        m = self.m[i]
        if self.is_charge(i):
            nmax = (m - 1) / 2
            return (2 * E_CH / np.sqrt(HBAR)) * (np.diag(np.arange(-nmax, nmax + 1)) - self.ng[i] * np.eye(m)) + 0j
        if m == 1:
            return np.zeros((1, 1), complex)
        a = np.diag(np.sqrt(np.arange(1, m)), 1)
        return -1j * np.sqrt(0.5 / self._impedance(i)) * (a - a.T)

    def _phi_iso(self, i):
        m = self.m[i]
        if self.is_charge(i):
            return np.eye(m, dtype=complex)
        if m == 1:
            return np.zeros((1, 1), complex)
        a = np.diag(np.sqrt(np.arange(1, m)), 1)
        return np.sqrt(0.5 * self._impedance(i)) * (a + a.T) + 0j

    def _n_iso(self, i):
        m = self.m[i]
        if self.is_charge(i):
            nmax = (m - 1) / 2
            return np.diag(np.arange(-nmax, nmax + 1)) + 0j
        return np.diag(np.arange(m)) + 0j

    def _d_iso(self, i, w):
        m = self.m[i]
        if w == 0:
            return np.eye(m, dtype=complex)
        d = np.diag(np.ones(m - 1), 1) + 0j     # d[k, k+1] = 1
        return d if w < 0 else d.conj().T

    @staticmethod
    def _displace(m, alpha):
        # qutip.displace: expm(alpha a^dag - conj(alpha) a) on the TRUNCATED ladder ops
        a = np.diag(np.sqrt(np.arange(1, m)), 1).astype(complex)
        return scipy.linalg.expm(alpha * a.conj().T - np.conj(alpha) * a)

    def _kron(self, mats):
        out = None
        for x in mats:
            x = sp.csr_matrix(x)
            out = x if out is None else sp.kron(out, x, format='csr')
        return out

    def _embed(self, i, op):
        return self._kron([op if j == i else np.eye(self.m[j]) for j in range(self.n)])

    # -- circuit.py:1621-1714 --------------------------------------------------
    def _build_ops(self):
        n = self.n
        self.Q = [self._embed(i, self._q_iso(i)) for i in range(n)]
        self.PHI = [self._embed(i, self._phi_iso(i)) for i in range(n)]
        self.NUM = [self._embed(i, self._n_iso(i)) for i in range(n)]
        self.EXP, self.ROOT_EXP = [], []
        for j in range(self.wTrans.shape[0] if len(self.W) else 0):
            e, eh = [], []
            for i in range(n):
                if self.is_charge(i):
                    e.append(self._d_iso(i, self.wTrans[j, i]))
                    eh.append(np.eye(self.m[i]))
                elif self.m[i] == 1:
                    e.append(np.eye(1))
                    eh.append(np.eye(1))
                else:
                    al = 1j * self._phi_zp(i) * self.wTrans[j, i]
                    e.append(self._displace(self.m[i], al))
                    eh.append(self._displace(self.m[i], al / 2))
            self.EXP.append(self._kron(e))
            self.ROOT_EXP.append(self._kron(eh))

    # -- circuit.py:1716-1757 --------------------------------------------------
    def lc_hamil(self):
        qc = self.cInvTrans.copy()
        qc[np.diag_indices_from(qc)] /= 2
        H = 0
        for i in range(self.n):
            if self.is_charge(i):
                H = H + qc[i, i] * (self.Q[i] @ self.Q[i])
            else:
                H = H + self.omega[i] * self.NUM[i]
            for j in range(i + 1, self.n):
                if self.cInvTrans[i, j] != 0:
                    H = H + qc[i, j] * (self.Q[i] @ self.Q[j])
        return H

    def ext_flux(self, b_id):
        # circuit.py:1759-1786
        return sum(lp.val * self.B[b_id, k] for k, lp in enumerate(self.loops)) if self.loops else 0.0

    def coupling_op(self, ctype, nodes):
        # circuit.py:2314-2395
        K = np.linalg.inv(self.C) @ self.R if ctype == 'capacitive' else self.S
        ops = self.Q if ctype == 'capacitive' else self.PHI
        n1, n2 = nodes
        op = 0
        for i in range(self.n):
            if 0 in nodes:
                c = K[n1 + n2 - 1, i]
            elif ctype == 'capacitive':
                c = K[n2 - 1, i] - K[n1 - 1, i]
            else:
                c = K[n1 - 1, i] - K[n2 - 1, i]
            op = op + c * ops[i]
        return op

    # -- circuit.py:1788-1829, 2147-2164 ---------------------------------------
    def hamiltonian(self):
        H = self.lc_hamil()
        self.ind_ops, self.cos_ops, self.sin_ops, self.sin_half_ops = {}, {}, {}, {}
        for edge, el, b_id in self.ind_keys:
            phi = self.ext_flux(b_id)
            op = self.coupling_op('inductive', edge)
            H = H + (1 / el.value) * phi * (PHI0 / 2 / np.pi) * op / np.sqrt(HBAR)
            self.ind_ops[(el, b_id)] = op
        for _, el, b_id, w_id in self.jj_keys:
            phi = self.ext_flux(b_id)
            ex = np.exp(1j * phi) * self.EXP[w_id]
            rex = np.exp(1j * phi / 2) * self.ROOT_EXP[w_id]
            self.cos_ops[(el, b_id)] = (ex + ex.conj().T) / 2
            self.sin_ops[(el, b_id)] = (ex - ex.conj().T) / 2j
            self.sin_half_ops[(el, b_id)] = (rex - rex.conj().T) / 2j
            H = H - el.value * self.cos_ops[(el, b_id)]
        return sp.csr_matrix(H)

    # -- circuit.py:1938-1986 (dense LAPACK instead of ARPACK: same spectrum) ---
    def diag(self, n_eig: int, dense: Optional[bool] = None):
        H = self.hamiltonian()
        N = H.shape[0]
        if dense is None:
            dense = N <= 3000
        if dense:
            w, v = scipy.linalg.eigh(H.toarray(), subset_by_index=[0, min(n_eig, N) - 1])
        else:
            w, v = scipy.sparse.linalg.eigsh(H, k=n_eig, which='SA', tol=0, ncv=max(4 * n_eig, 40))
            o = np.argsort(w)
            w, v = w[o], v[:, o]
        self.efreqs_rad = w
        self.evecs = v
        return w / (2 * np.pi * UNIT_FREQ), v

    # -- circuit.py:2397-2433 ---------------------------------------------------
    _scale = False

    def _ip(self, a, op, b):
        if self._scale:          # dec_rate_scale: ||op b|| bounds |<a|op|b>|
            return np.linalg.norm(op @ b)
        return np.vdot(a, op @ b)

    def _diff(self, x, y):
        return x + y if self._scale else x - y

    def matrix_elements(self, ctype, nodes, states):
        op = self.coupling_op(ctype, nodes)
        return self._ip(self.evecs[:, states[0]], op, self.evecs[:, states[1]])

    # -- circuit.py:2649-2775 ---------------------------------------------------
    def partial_H(self, el, b_sel=None):
        if isinstance(el, OLoop):
            k = self.loops.index(el)
            out = 0
            for _, li, b_id in self.ind_keys:
                out = out + (self.B[b_id, k] * self.ind_ops[(li, b_id)] / li.value
                             * PHI0 / np.sqrt(HBAR) / 2 / np.pi)
            for _, jj, b_id, _w in self.jj_keys:
                out = out + self.B[b_id, k] * jj.value * self.sin_ops[(jj, b_id)]
            return out
        if isinstance(el, OJJ):
            out = 0
            for _, jj, b_id, _w in self.jj_keys:
                if jj is el and (b_sel is None or b_sel == b_id):
                    out = out - self.cos_ops[(el, b_id)]
            return out
        M = len(self.m)
        if isinstance(el, OCap):
            cinv = np.linalg.inv(self.C)
            A = -self.R.T @ cinv @ self.partial_mats[el] @ cinv @ self.R
            return self._quadratic(self.Q, A[:M, :M])
        if isinstance(el, OInd):
            A = -self.S.T @ self.partial_mats[el] @ self.S
            out = self._quadratic(self.PHI, A[:M, :M])
            for _, li, b_id in self.ind_keys:
                if li is el:
                    out = out - (self.ind_ops[(el, b_id)] / el.value ** 2 / np.sqrt(HBAR)
                                 * (PHI0 / 2 / np.pi) * self.ext_flux(b_id))
            return out
        raise NotImplementedError

    def _quadratic(self, ops, A):
        # circuit.py:2649-2702: 1/2 sum_i A_ii O_i^2 + sum_{i<j} A_ij O_i O_j
        out = 0
        for i in range(len(ops)):
            for j in range(i, len(ops)):
                if A[i, j] != 0:
                    out = out + (0.5 if i == j else 1.0) * A[i, j] * (ops[i] @ ops[j])
        return out

    def partial_omega(self, el, m, subtract_ground=False):
        # circuit.py:2777-2826 (Hellmann-Feynman on the computed eigenvectors; hamiltonian() must have run)
        pH = self.partial_H(el)
        v = self.evecs[:, m]
        out = np.real(np.vdot(v, pH @ v))
        if subtract_ground:
            v0 = self.evecs[:, 0]
            out -= np.real(np.vdot(v0, pH @ v0))
        return out

    def _partial_omega_mn(self, el, states, b_sel=None):
        # circuit.py:2828-2873
        pH = self.partial_H(el, b_sel)
        a, b = self.evecs[:, states[0]], self.evecs[:, states[1]]
        return np.real(self._diff(self._ip(a, pH, a), self._ip(b, pH, b)))

    @staticmethod
    def _dephasing(A, p_omega):
        # circuit.py:2435-2454
        return np.abs(p_omega * A) * np.sqrt(2 * np.abs(np.log(ENV['omega_low'] * ENV['t_exp'])))

    # -- circuit.py:2456-2647 ---------------------------------------------------
    def dec_rate(self, dec_type, states, total=True):
        w = self.efreqs_rad
        omega = np.abs(w[states[1]] - w[states[0]])
        s1, s2 = self.evecs[:, states[0]], self.evecs[:, states[1]]
        alpha = HBAR * omega / (K_B * ENV['T'])
        if alpha > 709:
            down, up = 2, 0
        else:
            down = 1 + 1 / np.tanh(alpha / 2)
            up = down * np.exp(-alpha)
        tempS = (down + up) if total else (down if states[0] > states[1] else up)
        decay = 0.0
        if dec_type == 'capacitive':
            for edge, els in self.elements.items():
                for el in els:
                    cap = el if el.kind == 'C' else el.cap
                    if cap.Q:
                        decay += tempS * cap.value / cap.Q(omega) * np.abs(
                            self.matrix_elements('capacitive', edge, states)) ** 2
        elif dec_type == 'inductive':
            for (el, b_id), op in self.ind_ops.items():
                Q = el.Q(omega, ENV['T'])
                if not np.isnan(Q):
                    decay += tempS / Q / el.value * np.abs(self._ip(s1, op, s2)) ** 2
        elif dec_type == 'quasiparticle':
            for (el, b_id), op in self.sin_half_ops.items():
                Y = el.Y(omega, ENV['T'])
                if not np.isnan(Y):
                    decay += tempS * Y * omega * el.value * HBAR * np.abs(self._ip(s1, op, s2)) ** 2
        elif dec_type == 'charge':
            for i in range(self.n):
                if self.is_charge(i):
                    op = 0
                    for j in range(self.n):
                        op = op + self.cInvTrans[i, j] * self.Q[j] / np.sqrt(HBAR)
                    po = np.abs(self._diff(self._ip(s1, op, s1), self._ip(s2, op, s2)))
                    decay += self._dephasing(self.ng_A[i] * 2 * E_CH, po)
        elif dec_type == 'cc':
            for (el, b_id) in self.cos_ops:
                decay += self._dephasing(el.A * el.value, self._partial_omega_mn(el, states, b_id))
        elif dec_type == 'flux':
            for lp in self.loops:
                decay += self._dephasing(lp.A, self._partial_omega_mn(lp, states))
        else:
            raise ValueError(dec_type)
        return decay


    def dec_rate_scale(self, dec_type, states, total=True):
        """TEST HELPER (not in the reference): the same closed forms as ``dec_rate`` with every inner
        product <a|O|b> replaced by its bound ||O b|| (and differences of expectation values by sums).
        This is the size a rate would have if no symmetry suppressed it; rounding noise in a rate is
        machine epsilon times this number, so parity tests use 1e-9 x it as the floor below which a
        symmetry-forbidden rate (the reference itself returns noise there) is not compared."""
        self._scale = True
        try:
            return self.dec_rate(dec_type, states, total)
        finally:
            self._scale = False


# ---------------------------------------------------------------------------
# named circuits used by BASELINE.json's configs (builders shared by tests/bench)
# ---------------------------------------------------------------------------
def zero_pi(loop: OLoop, EC=0.15, ECJ=10.0, EJ=5.0, EL=0.13):
    """tests/test_qubits.py:52-81 (paper zero-pi)."""
    C, CJ = OCap(EC, 'GHz'), OCap(ECJ, 'GHz')
    JJ, L = OJJ(EJ, 'GHz', loops=[loop]), OInd(EL, 'GHz', loops=[loop])
    return OCircuit({(0, 1): [CJ, JJ], (0, 2): [L], (0, 3): [C], (1, 2): [C], (1, 3): [L],
                     (2, 3): [CJ, JJ]})


DISCOVERY_NAMES = ['C01', 'J01', 'C12', 'L12', 'C20', 'J20', 'C23', 'C30', 'J30', 'C34', 'C40', 'L40']


def discovery(values, loop: OLoop):
    """BASELINE configs[4]: the 4-mode circuit (two harmonic, two charge modes) of tests/golden/grad_cases.py,
    element values in GHz in the order of DISCOVERY_NAMES.  Returns (circuit, {name: element})."""
    mk = {'C': OCap, 'J': OJJ, 'L': OInd}
    els = {}
    for nm, val in zip(DISCOVERY_NAMES, values):
        extra = {'loops': [loop]} if nm in ('J01', 'L12', 'J20') else {}
        els[nm] = mk[nm[0]](float(val), 'GHz', **extra)
    cr = OCircuit({(0, 1): [els['C01'], els['J01']], (1, 2): [els['C12'], els['L12']],
                   (2, 0): [els['C20'], els['J20']], (2, 3): [els['C23']],
                   (3, 0): [els['C30'], els['J30']], (3, 4): [els['C34']],
                   (4, 0): [els['C40'], els['L40']]})
    els['loop'] = loop
    return cr, els


def fluxonium(loop: OLoop):
    """tests/test_qubits.py:145-175."""
    cap = OCap(3.6, 'GHz', Q=1e6)
    ind = OInd(0.46, 'GHz', Q=500e6, loops=[loop])
    junc = OJJ(10.2, 'GHz', cap=cap, A=1e-7, x=3e-06, loops=[loop])
    return OCircuit({(0, 1): [ind, junc]}, flux_dist='all')


def transmon():
    """tests/test_qubits.py:178-203."""
    return OCircuit({(0, 1): [OCap(20, 'GHz'), OJJ(15.0, 'GHz', A=1e-7)]})


def subspace_overlap(v_ref: np.ndarray, v: np.ndarray, w_ref: np.ndarray, rel_gap=1e-7,
                     w_next=None):
    """Phase/degeneracy-invariant eigenvector comparison.

    For each reference eigenvector i, the squared norm of its projection on the
    span of the computed vectors whose reference eigenvalues lie in the same
    (near-)degenerate cluster.  Returns the minimum over i.  If ``w_next`` (the first
    eigenvalue that was NOT requested) is given and it is within ``rel_gap`` of the last
    requested one, the trailing cluster is cut by the request and is skipped.
    """
    k = v_ref.shape[1]
    if w_next is not None:
        sc = max(1.0, float(np.max(np.abs(w_ref))))
        while k > 1 and abs(w_next - w_ref[k - 1]) <= rel_gap * sc:
            w_next = w_ref[k - 1]
            k -= 1
        v_ref, v, w_ref = v_ref[:, :k], v[:, :k], w_ref[:k]
    scale = max(1.0, float(np.max(np.abs(w_ref))))
    clusters = []
    start = 0
    for i in range(1, k + 1):
        if i == k or abs(w_ref[i] - w_ref[i - 1]) > rel_gap * scale:
            clusters.append(range(start, i))
            start = i
    worst = 1.0
    for cl in clusters:
        idx = list(cl)
        Q, _ = np.linalg.qr(v[:, idx])
        for i in idx:
            p = Q.conj().T @ v_ref[:, i]
            worst = min(worst, float(np.vdot(p, p).real / np.vdot(v_ref[:, i], v_ref[:, i]).real))
    return worst
