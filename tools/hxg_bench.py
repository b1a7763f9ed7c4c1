"""sqc_hx_generic alone at the sizes of the two benches that use it (CUDA events, warm, best of 5).  Dev tool.
usage: python tools/hxg_bench.py [trt|disc] [points]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sqcircuit_b200._lib import Potential
from sqcircuit_b200.kernels import BlockView, Kernels, make_space

dev = 'cuda:0'
k = Kernels(dev)
which = sys.argv[1] if len(sys.argv) > 1 else 'trt'
LAYOUTS = {'trt': ([10, 99, 99], [1, 0, 0], [[0, 1, 0], [0, 0, 1]], 64),
           'disc': ([10, 10, 9, 9], [1, 1, 0, 0], [[0, 0, 1, 0], [0, 0, 1, 0], [0, 0, 0, 1]], 512)}
dims, harm, shifts, B = LAYOUTS[which]
B = int(sys.argv[2]) if len(sys.argv) > 2 else B
p = 12
M, N, T = len(dims), int(np.prod(dims)), len(shifts)
sp = make_space(dims, harm)
toff = np.concatenate(([0], np.cumsum(dims)))[:-1]
tlen = int(np.sum(dims))
lam = torch.zeros(tlen, dtype=torch.float64, device=dev)
Vs = []
for i in range(M):
    if harm[i]:
        V, l = k.xbasis(dims[i])
        Vs.append(V.reshape(-1))
        lam[toff[i]:toff[i] + dims[i]] = l
Vcat = torch.cat(Vs).contiguous()
torch.manual_seed(0)
s = (0.4 * torch.randn((1, T, M), dtype=torch.float64, device=dev) * torch.tensor(harm, dtype=torch.float64, device=dev)).contiguous()
etab = torch.empty((1, T, tlen), dtype=torch.complex128, device=dev)
k.phase_tables(sp, T, lam, s, etab)
a = torch.randn((B, T), dtype=torch.complex128, device=dev)
g = torch.randn((B, 1 + M), dtype=torch.float64, device=dev)
fd = torch.randn((1, N), dtype=torch.float64, device=dev)
pot = Potential()
pot.n_terms = T
for t in range(T):
    for i in range(M):
        pot.shift[t][i] = shifts[t][i]
pot.lam, pot.etab, pot.etab_stride_b = lam.data_ptr(), etab.data_ptr(), 0
pot.g, pot.a, pot.diag, pot.diag_stride_b = g.data_ptr(), a.data_ptr(), None, 0
X = torch.randn((B, p, N), dtype=torch.complex128, device=dev)
Y = torch.zeros_like(X)


def run():
    assert k.hx_generic(sp, Vcat, pot, fd, BlockView(X), BlockView(Y))


import ctypes
force = int(os.environ.get('HXG_THREADS', '0'))
k.lib.sqc_debug_set_hxg_threads.argtypes = [ctypes.c_int]
k.lib.sqc_debug_set_hxg_threads(force)
run()
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    run()
    e0.record()
    torch.cuda.synchronize()
    best = min(best, s0.elapsed_time(e0))
flops = 4.0 * sum(d for d, h in zip(dims, harm) if h) * N * p * B
print(json.dumps({'layout': which, 'threads': force or 'auto', 'dims': dims, 'vectors': B * p, 'ms': round(best, 3),
                  'GBps': round(32.0 * N * p * B / best / 1e6, 1), 'TFLOPs': round(flops / best / 1e9, 2),
                  'us_per_vector_per_sm': round(best * 1e3 * 148 / (B * p), 2)}))
