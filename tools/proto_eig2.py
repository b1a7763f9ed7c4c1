"""Davidson parameter study (algorithm prototype; not product code)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from oracle import sqc_oracle as o
from tools.proto_eig import build, ortho_against, svqb


def davidson(H, d, k, p, mmax, tol, keep_prev=True, X0=None, maxit=200, log=False, sigma_floor=1e-3,
             lock=True):
    N = H.shape[0]
    if X0 is None:
        idx = np.argsort(d)[:p]
        X0 = np.zeros((N, p), complex)
        X0[idx, np.arange(p)] = 1
    V = svqb(X0)
    W = H @ V
    nmv = V.shape[1]
    units = 0.0   # tall-skinny GEMM work in units of N*8 flops (complex MAC = 8 flops) * cols*cols
    cprev = None
    hist = []
    for it in range(maxit):
        m = V.shape[1]
        G = V.conj().T @ W
        units += m * m
        G = (G + G.conj().T) / 2
        th, c = np.linalg.eigh(G)
        X = V @ c[:, :p]
        HX = W @ c[:, :p]
        units += 2 * m * p
        R = HX - X * th[:p]
        rn = np.linalg.norm(R, axis=0)
        hist.append(rn[:k].max())
        if log:
            print(it, m, nmv, ' '.join('%.1e' % x for x in rn[:p]))
        if (rn[:k] < tol).all():
            break
        if lock:
            act = np.where(rn[:p] >= tol * 0.5)[0]
        else:
            act = np.arange(p)
        den = d[:, None] - th[act][None, :]
        fl = sigma_floor * np.abs(th[0])
        den = np.where(np.abs(den) < fl, fl * np.where(den < 0, -1.0, 1.0), den)
        T = R[:, act] / den
        if m + len(act) > mmax:
            cc = [c[:, :p]]
            if keep_prev and cprev is not None and cprev.shape[0] <= m:
                cp = np.zeros((m, cprev.shape[1]), complex)
                cp[:cprev.shape[0]] = cprev
                cc.append(cp)
            cc = np.concatenate(cc, axis=1)
            q, _ = np.linalg.qr(cc)
            V = V @ q
            W = W @ q
            units += 2 * m * q.shape[1]
            cprev = None
            m = V.shape[1]
        else:
            cprev = c[:, :p]
        T = ortho_against(T, V)
        units += 4 * m * T.shape[1]
        T = svqb(T)
        units += 4 * T.shape[1] ** 2
        T = ortho_against(T, V)
        units += 4 * m * T.shape[1]
        T = svqb(T)
        units += 4 * T.shape[1] ** 2
        V = np.concatenate([V, T], axis=1)
        W = np.concatenate([W, H @ T], axis=1)
        nmv += T.shape[1]
    return th[:k], X[:, :k], it + 1, nmv, units, hist


if __name__ == '__main__':
    for flux in [0.37, 0.5, 0.0]:
        c, H, d = build(flux)
        w, v = np.linalg.eigh(H.toarray()) if False else (None, None)
        wref, vref = c.diag(10, dense=False)
        lam0 = abs(wref[0]) * 2 * np.pi * 1e9
        print('flux', flux, 'wref', wref)
        for tolrel in [1e-8, 1e-10]:
            for (p, mmax, kp) in [(12, 36, True), (12, 48, True), (12, 48, False), (12, 60, True), (14, 56, True), (16, 64, True), (10, 40, True)]:
                t0 = time.time()
                th, X, its, nmv, units, hist = davidson(H, d, 10, p, mmax, tol=tolrel * lam0, keep_prev=kp)
                hx_flops = nmv * 10100 * 850
                gemm_flops = units * 10100 * 8
                print(' tol %.0e p %d mmax %d prev %d: its %d matvecs %d  GFLOP hx %.2f gemm %.2f tot %.2f | eig err %.1e ovl err %.1e' % (
                    tolrel, p, mmax, kp, its, nmv, hx_flops / 1e9, gemm_flops / 1e9, (hx_flops + gemm_flops) / 1e9,
                    np.max(np.abs(th / (2 * np.pi * 1e9) - wref) / np.abs(wref)), 1 - o.subspace_overlap(vref, X, wref)))
