"""Prototype: Davidson WITHOUT orthogonalising the expansion block (generalised Rayleigh-Ritz
with the Gram matrix of the basis) vs the CGS2+SVQB version.  CPU study only."""
import sys, time
import numpy as np, scipy.linalg as sla
sys.path.insert(0, '.')
from oracle import sqc_oracle as o
from tools.proto_eig import build, ortho_against, svqb
from tools.proto_precond import make_prec


def davidson_gen(H, d, k, p, mmax, tol, prec, V0=None, maxit=200, drop=1e-10, log=False):
    N = H.shape[0]
    if V0 is None:
        idx = np.argsort(d)[:p]
        V0 = np.zeros((N, p), complex); V0[idx, np.arange(p)] = 1
    V = V0 / np.linalg.norm(V0, axis=0)
    W = H @ V
    GA = V.conj().T @ W; GB = V.conj().T @ V
    condmax = 0
    for it in range(maxit):
        m = V.shape[1]
        GA = (GA + GA.conj().T) / 2; GB = (GB + GB.conj().T) / 2
        s, U = np.linalg.eigh(GB)
        keep = s > drop * s.max()
        condmax = max(condmax, s.max() / s[keep].min())
        Tm = U[:, keep] / np.sqrt(s[keep])
        th, y = np.linalg.eigh(Tm.conj().T @ GA @ Tm)
        c = Tm @ y
        X = V @ c[:, :p]; HX = W @ c[:, :p]
        R = HX - X * th[:p]
        rn = np.linalg.norm(R, axis=0) / np.max(np.abs(th[:k]))
        if log: print(it, m, ' '.join('%.1e' % x for x in rn[:k]))
        if (rn[:k] < tol).all():
            return th[:k], X[:, :k], it + 1, condmax
        act = np.where(rn[:p] >= tol * 0.5)[0]
        T = prec(R[:, act], th[act])
        T = T / np.linalg.norm(T, axis=0)
        if m + len(act) > mmax:
            V, W = X, HX
            GA = np.diag(th[:p]).astype(complex); GB = np.eye(p, dtype=complex)
        HT = H @ T
        VT = V.conj().T @ T; TT = T.conj().T @ T
        VHT = V.conj().T @ HT; THT = T.conj().T @ HT
        GB = np.block([[GB, VT], [VT.conj().T, TT]])
        GA = np.block([[GA, VHT], [VHT.conj().T, THT]])
        V = np.concatenate([V, T], axis=1); W = np.concatenate([W, HT], axis=1)
    return th[:k], X[:, :k], maxit, condmax


if __name__ == '__main__':
    from tools.proto_precond import davidson as davidson_orth
    tolr = 1e-9
    for flux in (0.37, 0.499, 0.0):
        c, H, d = build(flux)
        wref, vref = c.diag(10, dense=False)
        Xn = []
        for df in (-4 / 4096, 4 / 4096):
            c2, H2, d2 = build(flux + df)
            th2, X2, _ = davidson_orth(H2, d2, 12, 12, 48, 1e-10, make_prec(H2, d2, 0))
            Xn.append(X2)
        V0 = np.concatenate(Xn, axis=1)
        for ns in (0, 160):
            pr = make_prec(H, d, ns)
            for name, fn in (('orth', davidson_orth), ('gen ', davidson_gen)):
                r1 = fn(H, d, 10, 12, 48, tolr, pr)
                r2 = fn(H, d, 10, 12, 48, tolr, pr, V0=V0)
                err = np.max(np.abs(r2[0] / (2e9 * np.pi) - wref) / np.abs(wref))
                errc = np.max(np.abs(r1[0] / (2e9 * np.pi) - wref) / np.abs(wref))
                print('flux %.3f ns %3d %s: cold its %3d (err %.1e ovl %.1e)  warm its %3d (err %.1e ovl %.1e) %s' % (
                    flux, ns, name, r1[2], errc, 1 - o.subspace_overlap(vref, r1[1], wref), r2[2], err,
                    1 - o.subspace_overlap(vref, r2[1], wref), ('cond %.1e/%.1e' % (r1[3], r2[3])) if len(r1) > 3 else ''))
