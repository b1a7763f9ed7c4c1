"""Dev tool: residual histories of the sweep points that make the non-orthogonalised Davidson iteration give up
(stagnation guard -> orthogonalised fall-back).  Runs one bench workload on the GPU, eagerly (no CUDA graphs), and
prints per level: points, iterations, and for the points still unfinished when the guard fired their worst relative
residual per iteration.

    python tools/stagnation_diag.py trt1e5 [drop|-] [low-block group, 0 = auto] [low-block states] [max basis|-] [zero|nan: fill the persistent workspace before every solve]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
import sqcircuit_b200 as sq                    # noqa: E402
import sqcircuit_b200.circuit as C             # noqa: E402
from sqcircuit_b200.kernels import Kernels     # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'trt1e5'
    drop = float(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != '-' else None
    import sqcircuit_b200.operators as O
    O.LOWBLOCK_GROUP = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    lowblock = int(sys.argv[4]) if len(sys.argv) > 4 else 512
    max_basis = int(sys.argv[5]) if len(sys.argv) > 5 and sys.argv[5] != '-' else None
    fill = sys.argv[6] if len(sys.argv) > 6 else None
    dev = 'cuda:0'
    sq.set_device(dev)
    kern = Kernels(dev)
    kern.use_graphs = False
    wl = bench.WORKLOADS[name]()
    cr = wl.build(sq, kern)
    hist = []
    decide = kern.davidson_decide

    def spy_decide(st, B):
        r = decide(st, B)
        kw = st._keep
        hist.append((kw['resmax'].clone(), kw['msz'].clone(), kw['nact'].clone()))
        return r
    kern.davidson_decide = spy_decide
    gen = C.solve_davidson_gen
    orth = C.solve_davidson

    def spy_gen(*a, **k):
        hist.clear()
        if fill:
            for buf in kern._ws.values():
                if buf is not None:
                    buf.view(torch.float64).fill_(float('nan') if fill == 'nan' else 0.0)
        if drop is not None:
            k['drop'] = drop
        r = gen(*a, **k)
        rm = r.resmax.cpu().numpy()
        print(f'level: {a[2]} points, {r.iterations} iterations, converged {r.converged}, group '
              f'{getattr(a[1], "_last_group", "?")}, '
              f'worst residual {rm.max():.2e}', flush=True)
        if not r.converged:
            bad = np.where(rm > 1e-9)[0]
            print('  unfinished points:', bad.tolist())
            for b in bad[:6]:
                print(f'  point {b}: ' + ' '.join(f'{float(h[0][b]):.1e}' for h in hist[-60:]))
                print(f'     basis: ' + ' '.join(str(int(h[1][b])) for h in hist[-60:]))
                print(f'     nact : ' + ' '.join(str(int(h[2][b])) for h in hist[-60:]))
        return r

    def spy_orth(*a, **k):
        print('  -> orthogonalised fall-back on', a[2], 'points', flush=True)
        return orth(*a, **k)
    C.solve_davidson_gen = spy_gen
    C.solve_davidson = spy_orth
    prm = wl.grid(wl.points, 0, 1)
    res = cr.sweep_diag(10, lowblock=lowblock, max_basis=max_basis, **wl.sweep_args(prm))
    print('total iterations', res.iterations, 'worst residual', float(res.resmax.max()))


if __name__ == '__main__':
    main()
