#!/usr/bin/env python
"""bench.py — diagonalisations/sec (k=10) on the 0-pi flux sweep (BASELINE.json configs[2]).

Workload (one "step"): the 0-pi qubit, harmonic x charge basis with trunc_nums [100, 51]
(Hilbert dim N = 10100), a 4096-point external-flux sweep per GPU, lowest k = 10 eigenpairs at
every point plus the six T1/Tphi decoherence channels for the (1, 0) transition.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, C ABI)
  python bench.py --impl reference --gpus N ...            reference arm: the reference's CPU
        algorithm (oracle port: scipy CSR Hamiltonian + ARPACK eigs 'SR', circuit.py:1938-1986)
        on the host cores, a bounded sample of the same sweep per step.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRUNC = [100, 51]
N_EIG = 10
POINTS = 4096
DEC_TYPES = ['capacitive', 'inductive', 'quasiparticle', 'charge', 'cc', 'flux']


def flux_grid(points, rank, world):
    """The global grid has world*points fluxes in [0, 1); rank r owns a contiguous shard."""
    total = points * world
    return (np.arange(total, dtype=np.float64)[rank * points:(rank + 1) * points] + 0.5) / total


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: oracle port on host cores
# --------------------------------------------------------------------------------------------
def _ref_init():
    # one BLAS/OpenMP thread per worker process: the workers already cover every core
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    try:
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass


def _ref_one(flux):
    import warnings
    warnings.filterwarnings('ignore')
    import scipy.sparse.linalg as spla
    from oracle import sqc_oracle as o
    lp = o.OLoop(float(flux))
    c = o.zero_pi(lp)
    c.set_trunc_nums(TRUNC)
    H = c.hamiltonian()
    w, v = spla.eigs(H, k=N_EIG, which='SR')      # the reference's call (circuit.py:1963)
    order = np.argsort(w.real)
    c.efreqs_rad, c.evecs = w.real[order], v[:, order]
    rates = [float(np.real(c.dec_rate(d, (1, 0)))) for d in DEC_TYPES]
    return c.efreqs_rad[:2].tolist(), rates


def run_reference_sample(n_points, procs):
    """Diagonalise n_points sweep points with `procs` worker processes; returns seconds."""
    import multiprocessing as mp
    fl = (np.arange(n_points) + 0.5) / n_points
    t0 = time.perf_counter()
    if procs <= 1:
        for f in fl:
            _ref_one(f)
    else:
        with mp.get_context('fork').Pool(procs, initializer=_ref_init) as pool:
            pool.map(_ref_one, list(fl), chunksize=1)
    return time.perf_counter() - t0


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    for var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[var] = '1'
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    procs = max(1, min(cores, 64))
    n_points = procs          # one point per worker per step: ~10 s of wall clock per step
    for _ in range(args.warmup if args.warmup < 1 else 1):
        run_reference_sample(n_points, procs)
    times = [run_reference_sample(n_points, procs) for _ in range(max(1, args.steps))]
    t = float(np.mean(times))
    value = n_points / t
    line = {
        'impl': 'reference', 'metric': 'diagonalisations/sec (k=10), 0-pi flux sweep', 'value': value,
        'unit': 'diag/s', 'n_gpus': args.gpus, 'steps': max(1, args.steps), 'warmup': min(args.warmup, 1),
        'ms_per_step': 1e3 * t, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64 (complex128)', 'data': 'synthetic',
        'config': {'workload': '0-pi qubit, trunc_nums [100,51] (N=10100), flux sweep, k=10, 6 T1/Tphi channels',
                   'points_per_step': n_points, 'note': 'bounded sample of the 4096-point sweep'},
        'cpu_baseline': {'value': value, 'unit': 'diag/s', 'cores': procs, 'kind': 'port',
                         'sample': f'{n_points} sweep points per step, one process per core, scipy CSR build + '
                                   'ARPACK eigs(k=10, which=SR) + dec_rate x6 (oracle/sqc_oracle.py)'},
        'e2e': {'value': value, 'unit': 'diag/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                                       '-i', str(self.idx), '-lms', '200'], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.f.read().strip().splitlines():
            parts = [x.strip() for x in ln.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        if sm:
            out = {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(np.max(mx)), 'reasons': sorted(reasons),
                   'samples': len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


class TimedKernels:
    """Wraps the CUDA backend: CUDA events on the launching stream around every C-ABI call,
    so per-kernel device time is measured live inside the timed region (no profiler)."""
    TIMED = ['hx_fused', 'hx_fused_real', 'real_to_complex', 'complex_to_real', 'mode_transform', 'potential_apply', 'gram', 'gram_pair', 'gen_rr', 'davidson_insert2',
             'lowblock_assemble', 'hpd_inverse_c64', 'lowblock_apply', 'lowband_factor', 'lowband_apply', 'block_update', 'ritz_update', 'block_rightmul', 'heev',
             'residual_norms', 'precond_residual', 'restart_copy', 'svqb_coeffs', 'davidson_decide',
             'davidson_insert', 'diag_mul', 'ladder_apply']

    def __init__(self, kern):
        import torch
        self._k = kern
        self._torch = torch
        self.enabled = False
        self.only = None       # restrict the instrumentation to these entry points
        self.records = []      # (name, start_event, stop_event, meta)
        for name in self.TIMED:
            setattr(self, name, self._wrap(name))

    def __getattr__(self, name):
        return getattr(self._k, name)

    def eager_ops(self):
        """Instrumented entry points carry CUDA events and must stay outside the solver's CUDA graphs."""
        if not self.enabled:
            return frozenset()
        return frozenset(self.TIMED if self.only is None else self.only)

    def _wrap(self, name):
        fn = getattr(self._k, name)
        torch = self._torch

        def call(*a, **kw):
            if not self.enabled or (self.only is not None and name not in self.only):
                return fn(*a, **kw)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **kw)
            e.record()
            self.records.append((name, s, e))
            return r
        return call

    def summary(self):
        tot = {}
        for name, s, e in self.records:
            ms = s.elapsed_time(e)
            t = tot.setdefault(name, [0.0, 0])
            t[0] += ms
            t[1] += 1
        return tot


def our_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = f'cuda:{local}'
    # stdout must carry the single JSON line only: NCCL prints its version banner to fd 1 when the
    # communicator is created (NCCL_DEBUG=VERSION / WARN / INFO), so fd 1 points at stderr until the
    # first collective has run
    saved_stdout = None
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=torch.device(dev))

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    import sqcircuit_b200 as sq
    from sqcircuit_b200.kernels import Kernels
    from sqcircuit_b200.sharding import gather_to_root
    sq.set_device(dev)
    kern = TimedKernels(Kernels(dev))
    if args.no_graphs:
        kern._k.use_graphs = False

    loop = sq.Loop()
    C, CJ = sq.Capacitor(0.15, 'GHz'), sq.Capacitor(10, 'GHz')
    JJ, L = sq.Junction(5, 'GHz', loops=[loop]), sq.Inductor(0.13, 'GHz', loops=[loop])
    cr = sq.Circuit({(0, 1): [CJ, JJ], (0, 2): [L], (0, 3): [C], (1, 2): [C], (1, 3): [L], (2, 3): [CJ, JJ]},
                    kernels=kern)
    cr.set_trunc_nums(TRUNC)
    points = args.points
    fluxes_host = torch.from_numpy(flux_grid(points, rank, world)).pin_memory()
    N = int(np.prod(cr.m))
    lib_launch0 = kern.launch_count()

    def step(e2e):
        """One pass of the hot path over this rank's shard.  e2e=True: host flux array in,
        host eigenfrequencies + rates out (the user-facing call); otherwise device-resident."""
        if e2e:
            fl = fluxes_host.to(dev, non_blocking=True).cpu().numpy()   # H2D of the step's inputs
        else:
            fl = fluxes_host.numpy()
        res = cr.sweep_diag(N_EIG, loop_fluxes={loop: fl}, tol=args.tol, max_basis=args.mmax, lowblock=args.lowblock, block_size=args.block, real=(False if args.complex else None), **({'coarse_stride': tuple(int(x) for x in args.strides.split(','))} if args.strides else {}))
        rates = [res.dec_rate(d, states=(1, 0)) for d in DEC_TYPES]
        if world > 1:
            # the only collective of the path: final gather of the per-point results to rank 0
            gather_to_root(res.efreqs.contiguous(), points * world)
        if e2e:
            return res.efreqs.cpu(), rates, res
        return res.efreqs, rates, res

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, e2e):
        sync()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        last = None
        for _ in range(nsteps):
            if DEBUG:
                torch.cuda.synchronize()
                ts = time.perf_counter()
            last = step(e2e)
            if DEBUG:
                torch.cuda.synchronize()
                st_ = torch.cuda.memory_stats()
                print(f'[debug] step e2e={e2e} {1e3 * (time.perf_counter() - ts):.1f} ms  cudaMalloc segments '
                      f'{st_["segment.all.allocated"]} retries {st_["num_alloc_retries"]} reserved '
                      f'{st_["reserved_bytes.all.current"] / 2**30:.1f} GiB peak-alloc '
                      f'{st_["allocated_bytes.all.peak"] / 2**30:.1f} GiB', file=sys.stderr)
        t1.record()
        sync()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), last

    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # timed region: only the dominant kernel (the fused H.X) carries CUDA events -- bracketing all
    # ~900 C-ABI calls of a step costs ~60 ms of host time per step (measured), which would be
    # charged to `value`.  The full per-kernel breakdown comes from one extra, untimed step below.
    kern.only = {'hx_fused', 'hx_fused_real', 'mode_transform', 'potential_apply'}
    kern.enabled = True
    l0 = kern.launch_count()
    ms_dev, last = timed(args.steps, False)
    launches = kern.launch_count() - l0
    kern.enabled = False
    hx_kernel = kern.summary()
    nvec = HX_COUNTER.get('vectors', 0) + (int(HX_COUNTER['vectors_dev'].item()) if 'vectors_dev' in HX_COUNTER else 0)
    nvec_real = HX_COUNTER.get('real_vectors', 0) + (int(HX_COUNTER['real_vectors_dev'].item())
                                                      if 'real_vectors_dev' in HX_COUNTER else 0)
    ms_e2e, last_e2e = timed(args.steps, True)
    clocks = sampler.stop() if rank == 0 else None
    iters = last[2].iterations
    kern.records = []
    kern.only = None
    kern.enabled = True
    step(False)
    sync()
    kern.enabled = False
    per_kernel = kern.summary()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_points = points * world
    value = total_points * args.steps / (ms_dev / 1e3)
    e2e_value = total_points * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the H.X apply (transform -> potential -> transform), measured live ----
    peaks = {}
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'MEASURED_PEAKS.json (sustained copy)' if peaks else 'fallback 6.65 TB/s'
    # the dominant H.X kernel of this run: the real-representation kernel (zero gate charge) or the complex one
    real_ms = hx_kernel.get('hx_fused_real', [0.0, 0])[0]
    cplx_ms = sum(hx_kernel.get(k, [0, 0])[0] for k in ('hx_fused', 'mode_transform', 'potential_apply'))
    use_real = real_ms > cplx_ms
    total_ms = sum(v[0] for v in per_kernel.values())
    shares = {k: {'ms': round(v[0], 3), 'calls': v[1], 'share': round(v[0] / total_ms, 4)}
              for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][0])}
    if use_real:
        hx_ms, hx_calls, nv = real_ms, hx_kernel['hx_fused_real'][1], nvec_real
        bytes_per_vec = 16 * N        # read N doubles, write N doubles
        kname = 'H.X apply, real representation (sqc_hx_fused_real: TMA load, transform + potential + transform, TMA store)'
    else:
        hx_ms, nv = cplx_ms, nvec
        hx_calls = hx_kernel.get('hx_fused', hx_kernel.get('potential_apply', [0, 1]))[1]
        bytes_per_vec = 32 * N        # read N complex, write N complex
        kname = 'H.X apply (sqc_hx_fused: transform + potential + transform in one kernel)'
    roof = {'bound': 'hbm', 'kernel': kname, 'unit': 'GB/s', 'peak': hbm_peak, 'peak_source': peak_src,
            'traffic': None, 'achieved': None, 'frac': None}
    if hx_ms > 0 and nv > 0:
        alg_bytes = float(bytes_per_vec) * nv
        roof['achieved'] = alg_bytes / (hx_ms / 1e3) / 1e9
        roof['frac'] = roof['achieved'] / hbm_peak
        roof['algorithmic_bytes_per_vector'] = bytes_per_vec
        roof['algorithmic_bytes_per_launch'] = alg_bytes / max(hx_calls, 1)
        ratio, src = NCU_HX_TRAFFIC.get('real' if use_real else 'complex', (None, None))
        if ratio is not None:
            # DRAM bytes of the kernel from its ncu --set full capture, as a ratio to the algorithmic bytes of
            # the captured launch; reported per launch, scaled to this run's average launch
            roof['traffic'] = ratio * alg_bytes / max(hx_calls, 1)
            roof['traffic_source'] = src
        roof['vectors'] = nv
        roof['ms'] = hx_ms
        # FP64 work with the parity-split transforms: 2 m flop per real amplitude (forward + backward), twice
        # that per complex amplitude
        flops = (2.0 if use_real else 4.0) * cr.m[0] * N * nv
        roof['fp64_tflops'] = flops / (hx_ms / 1e3) / 1e12
        # the kernel is FP64-tensor-bound (12.5 flop/B in both representations): context for the HBM fraction.
        # Peak = measured DMMA issue rate (tools/cu/dmma_peak.cu, profiles/r01_dmma_peak.txt): 37.1 TFLOP/s
        roof['fp64_tensor_peak_tflops'] = 37.1
        roof['fp64_tensor_frac'] = roof['fp64_tflops'] / 37.1
        roof['hbm_frac_ceiling_fp64_bound'] = (37.1e12 / (flops / alg_bytes)) / 1e9 / hbm_peak
    roof['calls'] = hx_calls

    # ---- cpu baseline: bounded sample of the same workload on the host cores ----
    cpu = None
    if not args.no_cpu_baseline:
        for var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
            os.environ[var] = '1'
        cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
        procs = max(1, min(cores, 64))
        t = run_reference_sample(procs, procs)
        cpu = {'value': procs / t, 'unit': 'diag/s', 'cores': procs, 'kind': 'port',
               'sample': f'{procs} sweep points of the same sweep, one process per core (oracle port of the '
                         f'reference: scipy CSR build + ARPACK eigs SR + 6 rates), {t:.1f} s'}

    h2d = points * 8
    d2h = points * N_EIG * 8 + len(DEC_TYPES) * points * 8
    line = {
        'metric': 'diagonalisations/sec (k=10), 0-pi flux sweep', 'value': value, 'unit': 'diag/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_dev / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': '0-pi qubit, trunc_nums [100,51] (N=10100), 4096-point flux sweep per GPU, k=10, '
                               '6 T1/Tphi channels on (1,0)',
                   'points_per_gpu': points, 'n_eig': N_EIG, 'hilbert_dim': N, 'tol_rel_residual': args.tol,
                   'davidson_iterations': iters, 'representation': 'real (time-reversal adapted)' if use_real else 'complex', 'l2': 'working set (>60 GB) larger than L2',
                   'parallelism': f'sweep points sharded over {world} GPU(s), final NCCL gather'},
        'e2e': {'value': e2e_value, 'unit': 'diag/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'roofline': roof,
        'kernel_time_breakdown': shares,
        'kernel_time_breakdown_note': 'ms of ONE extra fully instrumented step run after the timed region',
        'cpu_baseline': cpu,
        'clocks': clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


NCU_HX_TRAFFIC = {
    'complex': (1.0855, 'ncu dram__bytes_read+write / algorithmic = 1.0855 (profiles/r01_ncu_full_v28_hx_fused.txt)'),
}
HX_COUNTER = {'vectors': 0}
DEBUG = bool(os.environ.get('SQC_BENCH_DEBUG'))


def _install_hx_counter():
    """Count the state vectors that go through H.X inside the timed region (for the roofline)."""
    from sqcircuit_b200 import operators

    def hook(cls, key):
        orig = cls.apply

        def apply(self, X, Y):
            if getattr(self.ops.kern, 'enabled', False):
                if X.cnt is not None:      # device-side accumulation: no host sync in the timed region
                    HX_COUNTER[key + '_dev'] = HX_COUNTER.get(key + '_dev', 0) + X.cnt.sum()
                else:
                    HX_COUNTER[key] = HX_COUNTER.get(key, 0) + X.B * X.cnt_fixed
            return orig(self, X, Y)
        cls.apply = apply
    hook(operators.HamiltonianOperator, 'vectors')
    hook(operators.RealHamiltonianOperator, 'real_vectors')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--points', type=int, default=POINTS, help='sweep points per GPU')
    ap.add_argument('--tol', type=float, default=1e-9)
    ap.add_argument('--mmax', type=int, default=None, help='maximum Davidson basis size (default 4 x block)')
    ap.add_argument('--lowblock', type=int, default=512)
    ap.add_argument('--block', type=int, default=None, help='Davidson block size p (default k + 2)')
    ap.add_argument('--strides', default='', help='continuation level strides, e.g. 64,16,4,1 (default 64,8,1)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--complex', action='store_true', help='force complex vectors (no real representation)')
    ap.add_argument('--no-graphs', action='store_true', help='enqueue every solver iteration kernel by kernel')
    args = ap.parse_args()
    if args.impl == 'reference':
        reference_arm(args)
        return
    _install_hx_counter()
    our_arm(args)


if __name__ == '__main__':
    main()
