"""Prototype: thick restart that also keeps the 'P' block (LOBPCG / GD+k style) in the
non-orthogonal Davidson.  CPU study only."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from oracle import sqc_oracle as o
from tools.proto_eig import build
from tools.proto_precond import make_prec, davidson as davidson_orth
from tools.proto_nonortho import davidson_gen


def rr(GA, GB, drop):
    GA = (GA + GA.conj().T) / 2; GB = (GB + GB.conj().T) / 2
    s, U = np.linalg.eigh(GB)
    keep = s > drop * s.max()
    Tm = U[:, keep] / np.sqrt(s[keep])
    th, y = np.linalg.eigh(Tm.conj().T @ GA @ Tm)
    return th, Tm @ y


def davidson_gdk(H, d, k, p, mmax, tol, prec, V0=None, maxit=200, drop=1e-8, log=False, keep_p=True):
    N = H.shape[0]
    if V0 is None:
        idx = np.argsort(d)[:p]
        V0 = np.zeros((N, p), complex); V0[idx, np.arange(p)] = 1
    V = V0 / np.linalg.norm(V0, axis=0)
    W = H @ V
    GA = V.conj().T @ W; GB = V.conj().T @ V
    nx = 0            # leading basis vectors that are the Ritz block of the last restart
    work = 0
    for it in range(maxit):
        m = V.shape[1]
        th, c = rr(GA, GB, drop)
        c = c[:, :p]; th = th[:p]
        X = V @ c; HX = W @ c
        work += m
        R = HX - X * th
        rn = np.linalg.norm(R, axis=0) / np.max(np.abs(th[:k]))
        if log: print(it, m, ' '.join('%.1e' % x for x in rn[:k]))
        if (rn[:k] < tol).all():
            return th[:k], X[:, :k], it + 1, work
        act = np.where(rn[:p] >= tol * 0.5)[0]
        T = prec(R[:, act], th[act])
        T = T / np.linalg.norm(T, axis=0)
        if m + len(act) > mmax:
            if keep_p and nx > 0:
                cp = c[nx:, :][:, act]                 # the part of the ACTIVE Ritz vectors outside span(X_old)
                nrm = np.linalg.norm(V[:, nx:] @ cp, axis=0)
                cp = cp / nrm
                C = np.concatenate([c, np.concatenate([np.zeros((nx, len(act))), cp], axis=0)], axis=1)
            else:
                C = c
            Vn = V @ C; Wn = W @ C
            GA = C.conj().T @ GA @ C; GB = C.conj().T @ GB @ C
            V, W = Vn, Wn
            nx = p
        elif nx == 0 and False:
            pass
        HT = H @ T
        VT = V.conj().T @ T; TT = T.conj().T @ T
        VHT = V.conj().T @ HT; THT = T.conj().T @ HT
        GB = np.block([[GB, VT], [VT.conj().T, TT]])
        GA = np.block([[GA, VHT], [VHT.conj().T, THT]])
        V = np.concatenate([V, T], axis=1); W = np.concatenate([W, HT], axis=1)
        if nx == 0:
            nx = 0
    return th[:k], X[:, :k], maxit, work


if __name__ == '__main__':
    tolr = 1e-9
    for flux in (0.37, 0.499, 0.05):
        c, H, d = build(flux, (100, 50))
        wref, vref = c.diag(10, dense=False)
        Xn = []
        for df in (-3 / 4096, 5 / 4096):
            c2, H2, d2 = build(flux + df, (100, 50))
            th2, X2, _ = davidson_orth(H2, d2, 12, 12, 48, 1e-10, make_prec(H2, d2, 0))
            Xn.append(X2)
        V0 = np.concatenate(Xn, axis=1)
        pr = make_prec(H, d, 160)
        r = davidson_gen(H, d, 10, 12, 40, tolr, pr, V0=V0, drop=1e-8)
        print('flux %.3f  current (mmax 40)        : warm its %3d' % (flux, r[2]))
        for mm in (36, 40, 48):
            for kp in (False, True):
                r1 = davidson_gdk(H, d, 10, 12, mm, tolr, pr, V0=V0, keep_p=kp)
                r0 = davidson_gdk(H, d, 10, 12, mm, tolr, pr, keep_p=kp)
                err = np.max(np.abs(r1[0] / (2e9 * np.pi) - wref) / np.abs(wref))
                print('flux %.3f  mmax %d keep_p %d : warm its %3d (basis vectors streamed %4d, err %.1e ovl %.1e)  cold its %3d (%4d)' % (
                    flux, mm, kp, r1[2], r1[3], err, 1 - o.subspace_overlap(vref, r1[1], wref), r0[2], r0[3]))
