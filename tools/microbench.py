"""Per-kernel timings at bench-like sizes (CUDA events, warm, best of 5).  Dev tool."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sqcircuit_b200.kernels import Kernels, BlockView, make_space
from sqcircuit_b200._lib import Potential

dev = 'cuda:0'
k = Kernels(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
N, mm, p = 10100, 48, 12
only = sys.argv[2].split(',') if len(sys.argv) > 2 else None
torch.manual_seed(0)
V = torch.randn((B, mm, N), dtype=torch.complex128, device=dev)
X = torch.randn((B, p, N), dtype=torch.complex128, device=dev)
Y = torch.zeros_like(X)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


def report(name, ms, bytes_, flops):
    print(json.dumps({'kernel': name, 'ms': round(ms, 3), 'GBps': round(bytes_ / ms / 1e6, 1),
                      'TFLOPs': round(flops / ms / 1e9, 2)}))


import ctypes
for split in ((1, 2, 4, 8, 16) if (only and 'gramsplit' in only) else (None,)):
  if split: k.lib.sqc_debug_set_gram_split(split); print('gram split', split)
  for msz in (12, 30, 48):
    if only and 'gram' not in only and 'gramsplit' not in only: break
    cm = torch.full((B,), msz, dtype=torch.int32, device=dev)
    cn = torch.full((B,), p, dtype=torch.int32, device=dev)
    S = torch.zeros((B, mm, 16), dtype=torch.complex128, device=dev)
    ms = timeit(lambda: k.gram(BlockView(V, cnt=cm), BlockView(X, cnt=cn), S))
    report(f'gram msz={msz} nb=12', ms, 16.0 * (msz + p) * N * B, 8.0 * msz * p * N * B)
for msz in (12, 30, 48):
    if only and 'update' not in only: break
    cm = torch.full((B,), msz, dtype=torch.int32, device=dev)
    cn = torch.full((B,), p, dtype=torch.int32, device=dev)
    Cf = torch.randn((B, mm, mm), dtype=torch.complex128, device=dev)
    for op in (0, 1):
        ms = timeit(lambda: k.block_update(BlockView(V, cnt=cm), Cf, BlockView(X, cnt=cn), op=op))
        report(f'update msz={msz} nb=12 op={op}', ms, 16.0 * (msz + p * (1 + op)) * N * B, 8.0 * msz * p * N * B)
if not only or 'hx' in only:
    m, inner = 100, 101
    sp = make_space([m, inner], [1, 0])
    Vx, lam = k.xbasis(m)
    lamcat = torch.zeros(m + inner, dtype=torch.float64, device=dev); lamcat[:m] = lam
    s = torch.tensor([[[0.61, 0.0], [-0.61, 0.0]]], dtype=torch.float64, device=dev)
    etab = torch.empty((1, 2, m + inner), dtype=torch.complex128, device=dev)
    k.phase_tables(sp, 2, lamcat, s, etab)
    a = torch.randn((B, 2), dtype=torch.complex128, device=dev)
    g = torch.randn((B, 3), dtype=torch.float64, device=dev)
    fd = torch.randn((1, N), dtype=torch.float64, device=dev)
    pot = Potential(); pot.n_terms = 2; pot.shift[0][1] = 1; pot.shift[1][1] = -1
    pot.lam, pot.etab, pot.etab_stride_b = lamcat.data_ptr(), etab.data_ptr(), 0
    pot.g, pot.a, pot.diag, pot.diag_stride_b = g.data_ptr(), a.data_ptr(), None, 0
    Ve, Vo = Vx[0::2, :50].contiguous(), Vx[1::2, :50].contiguous()
    ms = timeit(lambda: k.hx_fused(sp, Ve, Vo, pot, fd, BlockView(X), BlockView(Y)))
    report('hx_fused 12 vec/pt', ms, 32.0 * N * p * B, 2.0 * 2 * m * N * p * B)
if not only or 'hxr' in only:
    # real-representation H.X at the bench size: m = 100, inner = 101 (trunc_nums [100, 51])
    m, inner = 100, 101
    Nc = (m * inner + 1) // 2
    sp = make_space([m, inner], [1, 0])
    Vx, lam = k.xbasis(m)
    lamcat = torch.zeros(m + inner, dtype=torch.float64, device=dev); lamcat[:m] = lam
    s = torch.tensor([[[0.61, 0.0], [-0.61, 0.0]]], dtype=torch.float64, device=dev)
    etab = torch.empty((1, 2, m + inner), dtype=torch.complex128, device=dev)
    k.phase_tables(sp, 2, lamcat, s, etab)
    a = torch.randn((B, 2), dtype=torch.complex128, device=dev)
    g = torch.randn((B, 3), dtype=torch.float64, device=dev)
    fd = torch.randn((1, 2 * Nc), dtype=torch.float64, device=dev)
    pot = Potential(); pot.n_terms = 2; pot.shift[0][1] = 1; pot.shift[1][1] = -1
    pot.lam, pot.etab, pot.etab_stride_b = lamcat.data_ptr(), etab.data_ptr(), 0
    pot.g, pot.a, pot.diag, pot.diag_stride_b = g.data_ptr(), a.data_ptr(), None, 0
    Ve, Vo = Vx[0::2, :50].contiguous(), Vx[1::2, :50].contiguous()
    Xr = torch.randn((B, p, Nc), dtype=torch.complex128, device=dev)
    Yr = torch.zeros_like(Xr)
    ms = timeit(lambda: k.hx_fused_real(sp, Ve, Vo, pot, fd, BlockView(Xr), BlockView(Yr)))
    report('hx_fused_real 12 vec/pt (full diagonal)', ms, 16.0 * N * p * B, 2.0 * m * N * p * B)
    sep = torch.randn(m + inner, dtype=torch.float64, device=dev)
    ms = timeit(lambda: k.hx_fused_real(sp, Ve, Vo, pot, None, BlockView(Xr), BlockView(Yr), fdiag_sep=sep))
    report('hx_fused_real 12 vec/pt (separable diagonal)', ms, 16.0 * N * p * B, 2.0 * m * N * p * B)
    ms = timeit(lambda: k.hx_fused_real(sp, Ve, Vo, pot, None, BlockView(Xr), BlockView(Yr)))
    report('hx_fused_real 12 vec/pt (no diagonal)', ms, 16.0 * N * p * B, 2.0 * m * N * p * B)
if not only or 'pair' in only:
    W = torch.randn((B, mm, N), dtype=torch.complex128, device=dev)
    for msz in (12, 24, 28):
        cm = torch.full((B,), msz, dtype=torch.int32, device=dev)
        ct = torch.full((B,), msz + p, dtype=torch.int32, device=dev)
        cn = torch.full((B,), p, dtype=torch.int32, device=dev)
        S1 = torch.zeros((B, mm, 16), dtype=torch.complex128, device=dev); S2 = torch.zeros_like(S1)
        ms = timeit(lambda: k.gram_pair(BlockView(V, cnt=ct), BlockView(V, cnt_fixed=p, off=cm, cnt=cn), BlockView(W, cnt_fixed=p, off=cm, cnt=cn), S1, S2))
        report(f'gram_pair rows={msz + p} nb=12+12', ms, 16.0 * (msz + p + 2 * p) * N * B, 8.0 * (msz + p) * 2 * p * N * B)
    for n in (24, 32, 40):
        GA = torch.randn((B, mm, mm), dtype=torch.complex128, device=dev); GA = GA + GA.conj().transpose(1, 2)
        Q = torch.randn((B, mm, 80), dtype=torch.complex128, device=dev); GB = Q @ Q.conj().transpose(1, 2) / 80
        w = torch.zeros((B, mm), dtype=torch.float64, device=dev); Zt = torch.zeros_like(GA)
        nn = torch.full((B,), n, dtype=torch.int32, device=dev)
        ms = timeit(lambda: k.gen_rr(GA, GB, nn, 1e-8, w, Zt))
        report(f'gen_rr n={n}', ms, 0, 0)
if only and 'rrr' in only:
    # real generalised Rayleigh-Ritz: latency of one wave (<= 740 resident problems) and of several
    for Bq in (64, 740, 2048):
        for n in (12, 24, 36, 40):
            GAr = torch.randn((Bq, 40, 40), dtype=torch.float64, device=dev); GAr = GAr + GAr.transpose(1, 2)
            Q = torch.randn((Bq, 40, 80), dtype=torch.float64, device=dev); GBr = Q @ Q.transpose(1, 2) / 80
            GAc, GBc = GAr.to(torch.complex128), GBr.to(torch.complex128)
            w = torch.zeros((Bq, 40), dtype=torch.float64, device=dev); Zt = torch.zeros_like(GAc)
            nn = torch.full((Bq,), n, dtype=torch.int32, device=dev)
            ms = timeit(lambda: k.gen_rr(GAc, GBc, nn, 1e-8, w, Zt, real=True, n_out=12))
            ms2 = timeit(lambda: k.gen_rr(GAc, GBc, nn, 1e-8, w, Zt))
            print(json.dumps({'kernel': f'gen_rr B={Bq} n={n}', 'real_ms': round(ms, 3), 'complex_ms': round(ms2, 3)}))
if not only or 'heev' in only:
    for n in (24, 36, 48):
        G = torch.randn((B, mm, mm), dtype=torch.complex128, device=dev); G = G + G.conj().transpose(1, 2)
        w = torch.zeros((B, mm), dtype=torch.float64, device=dev); Zt = torch.zeros_like(G)
        nn = torch.full((B,), n, dtype=torch.int32, device=dev)
        ms = timeit(lambda: k.heev(G, nn, w, Zt))
        report(f'heev n={n}', ms, 0, 0)
if not only or 'misc' in only:
    cn = torch.full((B,), p, dtype=torch.int32, device=dev)
    Cq = torch.randn((B, 16, 16), dtype=torch.complex128, device=dev)
    ms = timeit(lambda: k.block_rightmul(BlockView(X, cnt=cn), Cq, cn))
    report('rightmul 12', ms, 32.0 * p * N * B, 8.0 * p * p * N * B)
    th = torch.randn((B, mm), dtype=torch.float64, device=dev); rn = torch.zeros((B, p), dtype=torch.float64, device=dev)
    ms = timeit(lambda: k.residual_norms(BlockView(X), BlockView(Y), th, rn))
    report('residual_norms', ms, 32.0 * p * N * B, 0)
