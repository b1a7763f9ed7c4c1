"""Host-side circuit analysis: capacitance / susceptance matrices, external-flux
distribution and the coordinate transformation into harmonic and charge modes.

This is small dense linear algebra on (#nodes x #nodes) matrices and runs once per circuit
(not per sweep point); it follows the construction of the SQcircuit paper as implemented in
the reference (circuit.py:743-1186) step for step, with the same LAPACK primitives, so mode
order and sign conventions -- and therefore eigenvectors -- agree with the reference.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np
import scipy.linalg

from .elements import Capacitor, Inductor, Junction
from .settings import ACC


def _psd_sqrt(a):
    """Principal square root of a symmetric PSD matrix.  scipy.linalg.sqrtm is what the
    reference calls (circuit.py:911-913), and using the same routine keeps the sign / order
    conventions of the modes identical; for singular matrices some scipy versions return
    inf/nan, in which case the spectral square root is used instead."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        r = scipy.linalg.sqrtm(a)
    if np.all(np.isfinite(r)):
        return np.real(r) if np.iscomplexobj(r) and np.abs(np.imag(r)).max() == 0 else r
    w, v = np.linalg.eigh((a + a.T) / 2)
    w = np.where(w > 0, w, 0.0)
    return (v * np.sqrt(w)) @ v.T


def independent_columns(mat, tol=1e-5):
    """functions.py:309-336: drop zero columns, keep the first rank(mat) of the rest."""
    nz = mat[:, np.any(mat != 0, axis=0)]
    s = np.linalg.svd(nz, compute_uv=False)
    return nz[:, np.where(np.abs(s) > tol)[0]]


def edge_vector(edge, n):
    a, b = edge
    w = np.zeros(n + 1)
    if a == 0 or b == 0:
        w[a + b] += 1
    else:
        w[a] += 1
        w[b] -= 1
    return list(w[1:].astype(int))


def edge_matrix(edge, n):
    w = np.array(edge_vector(edge, n), dtype=float)
    return np.outer(w, w)


@dataclass
class Topology:
    n_nodes: int
    C: np.ndarray
    L: np.ndarray
    W: np.ndarray
    B: np.ndarray
    loops: list
    ind_keys: List[Tuple]                  # (edge, inductor, b_id)
    jj_keys: List[Tuple]                   # (edge, junction, b_id, w_id)
    partial_mats: Dict
    n_jj_without_ind: int
    S: np.ndarray = None
    R: np.ndarray = None
    S1: np.ndarray = None
    cInvTrans: np.ndarray = None
    lTrans: np.ndarray = None
    wTrans: np.ndarray = None
    omega: np.ndarray = None
    n: int = 0                             # number of modes kept


def loop_weights(loop):
    """elements.py:683-690."""
    k = independent_columns(np.array(loop.k_mat))
    b = np.zeros((1, k.shape[0]))
    b[0, 0] = 1
    return (np.linalg.inv(np.concatenate((b, k.T), axis=0)) @ b.T).T


def assemble(elements, flux_dist) -> Topology:
    """C, L, W and the external-flux distribution B (circuit.py:743-835)."""
    n = max(max(elements))
    C = np.zeros((n, n))
    L = np.zeros((n, n))
    W, K, c_branch = [], [], []
    loops, ind_keys, jj_keys = [], [], []
    partial = {}
    b_id = w_id = 0
    n_jj_no_l = 0
    for edge, els in elements.items():
        w = edge_vector(edge, n)
        em = edge_matrix(edge, n)
        c_sum = 0.0
        l_inv_sum = 0.0
        has_jj = has_l = False
        for el in els:
            if isinstance(el, Capacitor):
                c_sum += el.si()
                partial[el] = partial.get(el, 0) + em
                continue
            c_sum += el.cap.si()
            if isinstance(el, Inductor):
                has_l = True
                l_inv_sum += 1.0 / el.si()
                ind_keys.append((edge, el, b_id))
                partial[el] = partial.get(el, 0) + em / el.si() ** 2
            elif isinstance(el, Junction):
                has_jj = True
                jj_keys.append((edge, el, b_id, w_id))
            for lp in el.loops:
                if lp not in loops:
                    lp.reset()
                    loops.append(lp)
                lp.add_index(b_id)
                lp.add_to_k_mat(w)
            b_id += 1
            K.append(w)
            c_branch.append(el.get_cap_for_flux_dist(flux_dist))
        C += c_sum * em
        L += l_inv_sum * em
        if has_jj:
            W.append(w)
            w_id += 1
            if not has_l:
                n_jj_no_l += 1
    Bm = np.array([])
    try:
        if K:
            Kr = independent_columns(np.array(K))
            x = Kr.T @ np.diag(c_branch)
            for lp in loops:
                row = np.zeros((1, b_id))
                row[0, lp.indices] = loop_weights(lp)
                x = np.concatenate((x, row), axis=0)
            if loops:
                sel = np.concatenate((np.zeros((b_id - len(loops), len(loops))), np.eye(len(loops))), axis=0)
                Bm = np.around(np.linalg.inv(x) @ sel, 5)
    except ValueError:
        raise ValueError('The edge list does not specify a connected graph or all inductive loops '
                         'of the circuit are not specified.') from None
    return Topology(n, C, L, np.array(W), Bm, loops, ind_keys, jj_keys, partial, n_jj_no_l)


def _gs_independent_rows(a):
    """circuit.py:941-978: Gram-Schmidt selection of independent rows, largest norm first."""
    norms = np.linalg.norm(a, axis=1)
    unit = a / (norms.reshape(-1, 1) + 1e-9)
    picked, basis = [], []
    for i in np.argsort(-norms):
        v = unit[i, :]
        rest = v - sum([np.dot(v, e) * e for e in basis])
        if (np.abs(rest) > ACC['Gram-Schmidt']).any():
            picked.append(i)
            basis.append(rest / np.linalg.norm(rest))
    return picked, basis


class _Transformer:
    def __init__(self, top: Topology, rng):
        self.t = top
        self.rng = rng
        n = top.n_nodes
        self.n = n
        self.S = np.eye(n)
        self.R = np.eye(n)
        self.cInv = np.linalg.inv(top.C)
        self.lT = top.L.copy()
        self.wT = top.W.copy()
        self.omega = np.zeros(n)
        self.has_jj = len(top.W) != 0

    def apply(self, s, r):
        self.cInv = r.T @ self.cInv @ r
        self.lT = s.T @ self.lT @ s
        if self.has_jj:
            self.wT = self.wT @ s
        self.S = self.S @ s
        self.R = self.R @ r

    def charge_mask(self):
        return self.omega == 0

    def step1(self):
        """Simultaneous diagonalisation of C^-1 and L (circuit.py:902-939)."""
        c_root = _psd_sqrt(self.t.C)
        c_root_inv = np.linalg.inv(c_root)
        l_root = _psd_sqrt(self.t.L)
        _, d, u = np.linalg.svd(l_root @ c_root_inv)
        if np.max(d) == 0:
            d = np.ones(self.n)
            sing = list(range(self.n))
        else:
            ev = np.linalg.eigvals(self.t.L)
            n_sing = int(np.sum(ev / np.max(ev) < ACC['sing_mode_detect']))
            sing = list(range(self.n - n_sing, self.n))
            d[sing] = np.max(d)
        s1 = c_root_inv @ u.T @ np.diag(np.sqrt(d))
        r1 = np.linalg.inv(s1).T
        self.apply(s1, r1)
        self.lT[sing, sing] = 0
        self.S1 = s1
        self.omega = np.sqrt(np.diag(self.cInv) * np.diag(self.lT))

    def round_charge(self, w):
        """circuit.py:980-1008."""
        out = w.copy()
        cm = self.charge_mask()
        if self.t.n_jj_without_ind == 0:
            out[:, cm] = 0
        cw = out[:, cm]
        tol = ACC['Gram-Schmidt']
        cw[np.abs(cw) < tol] = 0
        cw[np.abs(cw - 1) < tol] = 1
        cw[np.abs(cw + 1) < tol] = -1
        out[:, cm] = cw
        return out

    def step2(self):
        """Integer-valued junction vectors on the charge modes (circuit.py:1047-1121)."""
        cm = self.charge_mask()
        w_red = independent_columns(self.t.W.T).T @ self.S1
        w_ch = w_red[:, cm].copy()
        nq = w_ch.shape[1]
        for _ in range(100):
            if nq != 0 and self.t.n_jj_without_ind != 0:
                picked, basis = _gs_independent_rows(w_ch)
                extra = []
                while len(basis) != nq:
                    extra = list(np.linalg.norm(w_ch, 'fro')
                                 * self.rng.standard_normal((nq - len(basis), nq)))
                    _, basis = _gs_independent_rows(np.array(basis + extra))
                f = np.array(list(w_ch[picked, :]) + extra)
                s2 = scipy.linalg.block_diag(np.eye(self.n - nq), np.linalg.inv(f))
                r2 = np.linalg.inv(s2.T)
            else:
                s2 = np.eye(self.n)
                r2 = s2
            trial = np.abs(self.round_charge(self.wT @ s2))[:, cm]
            if np.all((trial == 0) | (trial == 1)):
                self.apply(s2, r2)
                self.wT = self.round_charge(self.wT)
                return
        raise RuntimeError('Gram-Schmidt process failed to produce integer charge vectors')

    def impedance(self, i):
        return np.sqrt(self.cInv[i, i] / self.lT[i, i])

    def step3(self):
        """Scale the harmonic modes (circuit.py:1123-1168)."""
        from . import units as unt
        s3 = np.eye(self.n)
        for j in range(self.n):
            if self.omega[j] == 0:
                continue
            by_s = 1 / np.max(np.abs(self.S[:, j]))
            if self.has_jj:
                phi_zp = (2 * np.pi / unt.Phi0) * np.sqrt(unt.hbar / 2 * self.impedance(j))
                alphas = np.abs(phi_zp * self.wT[:, j])
                self.wT[:, j][alphas < ACC['har_mode_elim']] = 0
                if np.max(alphas) > ACC['har_mode_elim']:
                    s3[j, j] = 1 / np.abs(self.wT[np.argmax(alphas), j])
                else:
                    s3[j, j] = by_s
            else:
                s3[j, j] = by_s
        self.apply(s3, np.linalg.inv(s3.T))


def transform(top: Topology, rng=None, remove_uncoupled=True) -> Topology:
    tr = _Transformer(top, rng or np.random.default_rng(0))
    tr.step1()
    if tr.has_jj:
        tr.step2()
    tr.step3()
    cInv, lT, wT, omega = tr.cInv, tr.lT, tr.wT, tr.omega
    n = tr.n
    if tr.has_jj and remove_uncoupled:
        # circuit.py:837-863
        dead = list(np.where(np.sum(wT == 0, axis=0) == wT.shape[0])[0])
        n -= len(dead)
        for ax in (0, 1):
            cInv = np.delete(cInv, dead, axis=ax)
            lT = np.delete(lT, dead, axis=ax)
        wT = np.delete(wT, dead, axis=1)
        omega = np.delete(omega, dead)
    top.S, top.R, top.S1 = tr.S, tr.R, tr.S1
    top.cInvTrans, top.lTrans, top.wTrans, top.omega, top.n = cInv, lT, wT, omega, n
    return top
