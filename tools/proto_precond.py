"""Prototype: two-level preconditioner (exact low-energy block + diagonal) for the Davidson
solver, cold and warm started.  CPU study only."""
import sys, time
import numpy as np, scipy.linalg as sla
sys.path.insert(0, '.')
from oracle import sqc_oracle as o
from tools.proto_eig import build, ortho_against, svqb


def davidson(H, d, k, p, mmax, tol, prec, V0=None, maxit=200):
    N = H.shape[0]
    if V0 is None:
        idx = np.argsort(d)[:p]
        V0 = np.zeros((N, p), complex); V0[idx, np.arange(p)] = 1
    V = svqb(V0); W = H @ V
    for it in range(maxit):
        G = V.conj().T @ W; G = (G + G.conj().T) / 2
        th, c = np.linalg.eigh(G)
        X = V @ c[:, :p]; HX = W @ c[:, :p]
        R = HX - X * th[:p]
        rn = np.linalg.norm(R, axis=0) / np.max(np.abs(th[:k]))
        if (rn[:k] < tol).all():
            return th[:k], X[:, :k], it + 1
        act = np.where(rn[:p] >= tol * 0.5)[0]
        T = prec(R[:, act], th[act])
        if V.shape[1] + len(act) > mmax:
            V, W = X, HX
        T = svqb(ortho_against(T, V)); T = svqb(ortho_against(T, V))
        V = np.concatenate([V, T], axis=1); W = np.concatenate([W, H @ T], axis=1)
    return th[:k], X[:, :k], maxit


def make_prec(H, d, ns):
    fl = None
    def diagp(R, th):
        den = d[:, None] - th[None, :]
        f = 1e-3 * max(abs(th[0]), abs(th[-1]))
        den = np.where(np.abs(den) < f, np.where(den < 0, -f, f), den)
        return R / den
    if ns == 0:
        return diagp
    S = np.argsort(d)[:ns]
    Hss = H[S][:, S].toarray()
    def twolevel(R, th):
        T = diagp(R, th)
        sig = th[0] - 0.05 * abs(th[0])          # one shift below the lowest Ritz value
        A = Hss - sig * np.eye(ns)
        T[S, :] = np.linalg.solve(A, R[S, :])
        return T
    return twolevel


if __name__ == '__main__':
    tolr = 1e-9
    for flux in (0.37, 0.499):
        c, H, d = build(flux)
        wref, vref = c.diag(10, dense=False)
        # neighbours for the warm start (spacing 4/4096 on each side)
        Xn = []
        for df in (-4 / 4096, 4 / 4096):
            c2, H2, d2 = build(flux + df)
            th2, X2, _ = davidson(H2, d2, 12, 12, 48, 1e-10, make_prec(H2, d2, 0))
            Xn.append(X2)
        V0 = np.concatenate(Xn, axis=1)
        for ns in (0, 128, 256, 512, 1024):
            pr = make_prec(H, d, ns)
            t0 = time.time(); th, X, itc = davidson(H, d, 10, 12, 48, tolr, pr); tc = time.time() - t0
            th2, X2, itw = davidson(H, d, 10, 12, 48, tolr, pr, V0=V0)
            err = np.max(np.abs(th2 / (2e9 * np.pi) - wref) / np.abs(wref))
            print('flux %.3f  ns %4d : cold its %3d  warm its %3d   (eig err %.1e, ovl err %.1e)' % (
                flux, ns, itc, itw, err, 1 - o.subspace_overlap(vref, X2, wref)))
