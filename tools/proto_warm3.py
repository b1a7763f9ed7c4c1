"""Prototype: warm start from 2, 3 or 4 neighbour Ritz blocks (non-orthogonalised Davidson, two-level
preconditioner).  CPU study only."""
import sys
import numpy as np
sys.path.insert(0, '.')
from oracle import sqc_oracle as o
from tools.proto_eig import build
from tools.proto_precond import make_prec, davidson as davidson_orth
from tools.proto_nonortho import davidson_gen

tolr = 1e-9
for flux in (0.37, 0.499, 0.12):
    c, H, d = build(flux)
    wref, vref = c.diag(10, dense=False)
    nbr = {}
    for off in (-12, -4, 4, 12):
        c2, H2, d2 = build(flux + off / 4096)
        th2, X2, _ = davidson_orth(H2, d2, 12, 12, 48, 1e-10, make_prec(H2, d2, 0))
        nbr[off] = X2
    pr = make_prec(H, d, 160)
    for name, offs, mmax in (('2 nb', (-4, 4), 48), ('3 nb', (-4, 4, 12), 48), ('3 nb mmax60', (-4, 4, 12), 60),
                             ('4 nb mmax60', (-12, -4, 4, 12), 60), ('1 nb', (4,), 48)):
        V0 = np.concatenate([nbr[k] for k in offs], axis=1)
        th, X, its, cond = davidson_gen(H, d, 10, 12, mmax, tolr, pr, V0=V0, drop=1e-8)
        err = np.max(np.abs(th / (2e9 * np.pi) - wref) / np.abs(wref))
        # cost model in units of one 12-vector H.X + projection: initial block count + iterations
        print('flux %.3f %-12s its %2d  (initial blocks %d)  err %.1e ovl %.1e' % (
            flux, name, its, len(offs), err, 1 - o.subspace_overlap(vref, X, wref)))
