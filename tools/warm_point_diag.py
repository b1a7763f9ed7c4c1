import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import circuits as cs
from emul_kernels import EmulKernels
import sqcircuit_b200.circuit as C
import sqcircuit_b200.operators as O
from sqcircuit_b200.eigensolver import RitzPairStart
drop = float(sys.argv[1]); lb = int(sys.argv[2])
trunc = [10, 50, 50]
pts = eval(sys.argv[3])   # [(r,c) nb1, (r,c) nb2, (r,c) target]
import sqcircuit_b200 as sq
if torch.cuda.is_available():
    from sqcircuit_b200.kernels import Kernels
    sq.set_device('cuda:0'); kern = Kernels('cuda:0'); kern.use_graphs = False
else:
    kern = EmulKernels()
hist = []
od = kern.davidson_decide
def dec(st, B):
    r = od(st, B); hist.append((st._keep['resmax'] if hasattr(st, '_keep') else st.resmax).clone()); return r
kern.davidson_decide = dec
cr, lp, oc, ol = cs.trt_squid(kern); cr.set_trunc_nums(trunc)
ops = cr._operators()
m1 = sorted(cr.charge_islands)[0]
side = 16
def prm(row, col):
    return row / side * 0.5, (col + 0.5) / side * 0.5
def hop_of(pts):
    lv = np.array([[2 * np.pi * prm(r, c)[1]] for r, c in pts])
    ng = np.zeros((len(pts), ops.M)); ng[:, m1] = [prm(r, c)[0] for r, c in pts]
    return O.HamiltonianOperator(ops, lv, ng)
N = ops.N
import os
cache = '/tmp/trt_cold_%s.pt' % '_'.join(map(str, pts[0] + pts[1]))
if os.path.exists(cache):
    store = torch.load(cache).to(ops.device)
else:
    h = hop_of(pts[:2])
    t = time.time()
    r = C.solve_davidson_gen(kern, h, 2, N, 10, h.diag, lowblock_ns=512, drop=1e-5)
    print('cold', r.iterations, r.converged, time.time() - t)
    store = r.ritz_block.clone(); torch.save(store, cache)
hist.clear()
h = hop_of(pts[2:])
x0 = RitzPairStart(store, torch.tensor([0], device=ops.device), torch.tensor([1], device=ops.device))
t = time.time()
r = C.solve_davidson_gen(kern, h, 1, N, 10, h.diag, lowblock_ns=lb, drop=drop, x0=x0, maxit=70)
print('warm drop', drop, 'lb', lb, 'its', r.iterations, r.converged, 'time %.0f' % (time.time() - t))
print(' '.join('%.1e' % float(x[0]) for x in hist))
