"""CPU study: block size p, which columns get expanded, and basis cap in the non-orthogonal Davidson
(warm start from two neighbours, two-level preconditioner).  Cost model ~ vectors streamed."""
import sys
import numpy as np
sys.path.insert(0, '.')
from tools.proto_eig import build
from tools.proto_precond import make_prec, davidson as davidson_orth
from tools.proto_gdk import rr


def dav(H, d, k, p, mmax, tol, prec, V0, nexp, maxit=100, drop=1e-8):
    V = V0 / np.linalg.norm(V0, axis=0)
    W = H @ V
    GA = V.conj().T @ W; GB = V.conj().T @ V
    cost = dict(hx=V.shape[1], upd=0, gram=0, its=0)
    for it in range(maxit):
        m = V.shape[1]
        th, c = rr(GA, GB, drop)
        c = c[:, :p]; th = th[:p]
        X = V @ c; HX = W @ c
        cost['upd'] += 2 * m
        R = HX - X * th
        rn = np.linalg.norm(R, axis=0) / np.max(np.abs(th[:k]))
        cost['its'] = it + 1
        if (rn[:k] < tol).all():
            return cost
        act = np.where(rn[:nexp] >= tol * 0.5)[0]
        T = prec(R[:, act], th[act]); T = T / np.linalg.norm(T, axis=0)
        if m + len(act) > mmax:
            V, W = X, HX
            GA = np.diag(th).astype(complex); GB = np.eye(p, dtype=complex)
        HT = H @ T
        cost['hx'] += len(act); cost['gram'] += V.shape[1] + 3 * len(act)
        VT = V.conj().T @ T; TT = T.conj().T @ T; VHT = V.conj().T @ HT; THT = T.conj().T @ HT
        GB = np.block([[GB, VT], [VT.conj().T, TT]]); GA = np.block([[GA, VHT], [VHT.conj().T, THT]])
        V = np.concatenate([V, T], axis=1); W = np.concatenate([W, HT], axis=1)
    return cost


if __name__ == '__main__':
    tot = {}
    for flux in (0.37, 0.499, 0.05, 0.21):
        c, H, d = build(flux, (100, 50))
        pr = make_prec(H, d, 160)
        nbr = {}
        for pp in (12, 14, 16):
            Xn = []
            for df in (-3 / 4096, 5 / 4096):
                c2, H2, d2 = build(flux + df, (100, 50))
                th2, X2, _ = davidson_orth(H2, d2, pp, pp, 4 * pp, 1e-10, make_prec(H2, d2, 0))
                Xn.append(X2)
            nbr[pp] = np.concatenate(Xn, axis=1)
        for (p, nexp, mmax) in ((12, 12, 40), (12, 10, 40), (12, 10, 36), (14, 14, 46), (14, 10, 40), (14, 12, 42), (16, 10, 44), (16, 16, 52), (12, 12, 52), (12, 10, 48)):
            r = dav(H, d, 10, p, mmax, 1e-9, pr, nbr[p], nexp)
            # cost model in units of N-vectors of HBM traffic / tensor work (rough weights from the bench)
            model = 1.0 * r['upd'] + 1.3 * r['gram'] + 5.0 * r['hx'] + 30 * r['its']
            key = (p, nexp, mmax)
            tot[key] = tot.get(key, 0) + model
            print('flux %.3f p %2d expand %2d mmax %2d : its %2d  hx %3d  upd %4d  gram %4d  model %6.0f' % (
                flux, p, nexp, mmax, r['its'], r['hx'], r['upd'], r['gram'], model))
    for kk, v in sorted(tot.items(), key=lambda kv: kv[1]):
        print(kk, round(v))
