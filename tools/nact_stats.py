"""Dev tool: per-iteration active-column statistics of the generalised Davidson on the bench sweep."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sqcircuit_b200 as sq
from sqcircuit_b200 import eigensolver

points = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sq.set_device('cuda:0')
loop = sq.Loop()
C, CJ = sq.Capacitor(0.15, 'GHz'), sq.Capacitor(10, 'GHz')
JJ, L = sq.Junction(5, 'GHz', loops=[loop]), sq.Inductor(0.13, 'GHz', loops=[loop])
cr = sq.Circuit({(0, 1): [CJ, JJ], (0, 2): [L], (0, 3): [C], (1, 2): [C], (1, 3): [L], (2, 3): [CJ, JJ]})
cr.set_trunc_nums([100, 50])
fl = np.linspace(0.0, 1.0, points)
eigensolver.DETAIL_STATS = []
stats = []
LB = int(sys.argv[2]) if len(sys.argv) > 2 else 160
res = cr.sweep_diag(10, loop_fluxes={loop: fl}, tol=1e-9, stats=stats, lowblock=LB)
tot = dict(cols=0, colmax=0, tiles=0, tilemax=0, rows=0)
print('it B left nact_sum msz_sum tiles restarts')
for (it, B, left, na, ms, tl, rs) in eigensolver.DETAIL_STATS:
    print(it, B, left, na, ms, tl, rs)
    tot['cols'] += na; tot['colmax'] += 12 * left; tot['tiles'] += tl; tot['tilemax'] += 3 * left
print(tot, 'active col frac', tot['cols'] / tot['colmax'], 'tile frac', tot['tiles'] / tot['tilemax'])

# ---- wall time per continuation level (no detail stats: no extra syncs) ----
import time
from sqcircuit_b200 import circuit as _c
eigensolver.DETAIL_STATS = None
orig = _c._davidson
times = []
def timed(kern, hop, B, *a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig(kern, hop, B, *a, **k)
    torch.cuda.synchronize(); times.append((B, r.iterations, (time.perf_counter() - t0) * 1e3))
    return r
_c._davidson = timed
for rep in range(2):
    times.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = cr.sweep_diag(10, loop_fluxes={loop: fl}, tol=1e-9, lowblock=LB)
    torch.cuda.synchronize(); tt = (time.perf_counter() - t0) * 1e3
    print('sweep total ms', round(tt, 1), 'levels (B, its, ms):', [(b, i, round(t, 1)) for b, i, t in times],
          'outside solver ms', round(tt - sum(t for _, _, t in times), 1))
