"""Dev tool: torch.profiler table of the decay-rate part of a bench step only."""
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
cr.set_trunc_nums([100, 50])
fl = np.linspace(0.0, 1.0, points)
DEC = ['capacitive', 'inductive', 'quasiparticle', 'charge', 'cc', 'flux']
res = cr.sweep_diag(10, loop_fluxes={loop: fl}, tol=1e-9)
for _ in range(2):
    [res.dec_rate(d, states=(1, 0)) for d in DEC]
torch.cuda.synchronize()
for d in DEC:
    t0 = time.perf_counter(); res.dec_rate(d, states=(1, 0)); torch.cuda.synchronize()
    print(d, round((time.perf_counter() - t0) * 1e3, 2), 'ms')
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    [res.dec_rate(d, states=(1, 0)) for d in DEC]
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:16]:
    print('%8.2f ms  %5d  %s' % (e.self_device_time_total / 1e3, e.count, e.key[:90]))
