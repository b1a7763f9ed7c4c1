"""Dev tool: torch.profiler (CUPTI) kernel table of one full bench step (sweep + decay rates)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sqcircuit_b200 as sq
from torch.profiler import profile, ProfilerActivity

points = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sq.set_device('cuda:0')
loop = sq.Loop()
C, CJ = sq.Capacitor(0.15, 'GHz'), sq.Capacitor(10, 'GHz')
JJ, L = sq.Junction(5, 'GHz', loops=[loop]), sq.Inductor(0.13, 'GHz', loops=[loop])
cr = sq.Circuit({(0, 1): [CJ, JJ], (0, 2): [L], (0, 3): [C], (1, 2): [C], (1, 3): [L], (2, 3): [CJ, JJ]})
cr.set_trunc_nums([100, 51])
fl = (np.arange(points) + 0.5) / points
DEC = ['capacitive', 'inductive', 'quasiparticle', 'charge', 'cc', 'flux']


def step():
    res = cr.sweep_diag(10, loop_fluxes={loop: fl}, tol=1e-9, lowblock=512)
    t0 = time.perf_counter(); torch.cuda.synchronize(); t1 = time.perf_counter()
    rates = [res.dec_rate(d, states=(1, 0)) for d in DEC]
    torch.cuda.synchronize()
    return (time.perf_counter() - t1) * 1e3


for _ in range(2):
    step()
torch.cuda.synchronize(); t0 = time.perf_counter()
tr = step()
torch.cuda.synchronize(); print('step wall ms', (time.perf_counter() - t0) * 1e3, ' dec_rate part ms', tr)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.self_device_time_total for e in ev)
print('sum of device time ms', tot / 1e3)
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:40]:
    print('%8.2f ms  %5d  %s' % (e.self_device_time_total / 1e3, e.count, e.key[:100]))
