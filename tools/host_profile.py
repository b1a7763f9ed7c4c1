"""Dev tool: where the HOST time of a bench step goes (cProfile of the Python thread over a few steps of a
workload, GPU running asynchronously; blocking waits show up as synchronize / item / cpu).

    python tools/host_profile.py [zeropi4096] [steps] [points]
"""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
import sqcircuit_b200 as sq                    # noqa: E402
from sqcircuit_b200.kernels import Kernels     # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'zeropi4096'
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    sq.set_device('cuda:0')
    kern = Kernels('cuda:0')
    wl = bench.WORKLOADS[name]()
    cr = wl.build(sq, kern)
    prm = wl.grid(int(sys.argv[3]) if len(sys.argv) > 3 else wl.points, 0, 1)

    def step():
        res = cr.sweep_diag(10, **wl.sweep_args(prm))
        return res, [res.dec_rate(d, states=(1, 0)) for d in wl.dec_types]
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    for _ in range(steps):
        out = step()
        del out
    pr.disable()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f'{steps} steps: host returned after {1e3 * t_host / steps:.1f} ms per step, GPU done after '
          f'{1e3 * t_all / steps:.1f} ms per step')
    for key in ('cumulative', 'tottime'):
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats(key).print_stats(45)
        print(buf.getvalue()[:9000])


if __name__ == '__main__':
    main()
