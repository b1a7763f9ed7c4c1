"""Prototype: cost model vs restart size for the warm-started non-orthogonalised Davidson."""
import sys
import numpy as np
sys.path.insert(0, '.')
from oracle import sqc_oracle as o
from tools.proto_eig import build
from tools.proto_precond import make_prec, davidson as davidson_orth


def davidson_gen_cost(H, d, k, p, mmax, tol, prec, V0, drop=1e-8, maxit=100, keep=None):
    V = V0 / np.linalg.norm(V0, axis=0)
    W = H @ V
    GA = V.conj().T @ W; GB = V.conj().T @ V
    cost = 0.0
    for it in range(maxit):
        m = V.shape[1]
        GA = (GA + GA.conj().T) / 2; GB = (GB + GB.conj().T) / 2
        s, U = np.linalg.eigh(GB)
        kp = s > drop * s.max()
        Tm = U[:, kp] / np.sqrt(s[kp])
        th, y = np.linalg.eigh(Tm.conj().T @ GA @ Tm)
        c = Tm @ y
        nk = p if keep is None else keep
        X = V @ c[:, :nk]; HX = W @ c[:, :nk]
        cost += 2 * m * p            # X, HX (12 columns)
        R = HX[:, :p] - X[:, :p] * th[:p]
        rn = np.linalg.norm(R, axis=0) / np.max(np.abs(th[:k]))
        if (rn[:k] < tol).all():
            return it + 1, cost
        act = np.where(rn[:p] >= tol * 0.5)[0]
        T = prec(R[:, act], th[act])
        T = T / np.linalg.norm(T, axis=0)
        if m + len(act) > mmax:
            if nk > p:
                cost += 2 * m * (nk - p)
            V, W = X, HX
            GA = np.diag(th[:nk]).astype(complex); GB = np.eye(nk, dtype=complex)
            m = nk
        HT = H @ T
        cost += 45 * len(act)        # H.X (relative units: ~400 flop vs 8 flop per row-col pair... calibrated)
        cost += (m + len(act)) * 2 * len(act)     # paired gram
        VT = V.conj().T @ T; TT = T.conj().T @ T
        VHT = V.conj().T @ HT; THT = T.conj().T @ HT
        GB = np.block([[GB, VT], [VT.conj().T, TT]])
        GA = np.block([[GA, VHT], [VHT.conj().T, THT]])
        V = np.concatenate([V, T], axis=1); W = np.concatenate([W, HT], axis=1)
    return maxit, cost


for flux in (0.37, 0.499, 0.12):
    c, H, d = build(flux)
    nbr = []
    for off in (-4, 4):
        c2, H2, d2 = build(flux + off / 4096)
        th2, X2, _ = davidson_orth(H2, d2, 12, 12, 48, 1e-10, make_prec(H2, d2, 0))
        nbr.append(X2)
    V0 = np.concatenate(nbr, axis=1)
    pr = make_prec(H, d, 160)
    out = []
    for mmax, keep in ((36, None), (40, None), (48, None), (60, None), (48, 24), (60, 24), (36, 24)):
        its, cost = davidson_gen_cost(H, d, 10, 12, mmax, 1e-9, pr, V0, keep=keep)
        out.append('mmax %d keep %s: its %d cost %.0f' % (mmax, keep, its, cost))
    print('flux %.3f | ' % flux + ' | '.join(out))
