#!/usr/bin/env python
"""bench.py — diagonalisations/sec (k=10) on the 0-pi flux sweep (BASELINE.json configs[2]).

Workload (one "step"): the 0-pi qubit, harmonic x charge basis with trunc_nums [100, 51]
(Hilbert dim N = 10100), a 4096-point external-flux sweep per GPU, lowest k = 10 eigenpairs at
every point plus the six T1/Tphi decoherence channels for the (1, 0) transition.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, C ABI)
  python bench.py --impl reference --gpus N ...            reference arm: the reference's CPU
        algorithm (oracle port: scipy CSR Hamiltonian + ARPACK eigs 'SR', circuit.py:1938-1986)
        on the host cores, a bounded sample of the same sweep per step.

  --config {zeropi4096 (default, the headline), fluxonium1024, trt1e5, discovery512}   the other named
        circuits of BASELINE.json (configs[1], configs[3], configs[4]); same JSON line, same arms.
Multi-GPU (torchrun): weak scaling with IDENTICAL work per GPU (`value`, `per_rank`), and in the same
invocation the strong scaling of the one named sweep split over the GPUs (`strong`).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
# A sweep step allocates and frees GB-sized result tensors of varying sizes (levels of the continuation).  With the
# default caching allocator that fragments into new multi-GB cudaMalloc segments now and then (80-170 ms each,
# serialised across the ranks of a box: measured 378 instead of 298 ms per step at 2 GPUs); expandable segments
# grow one mapping instead.  Deployment knob, set before torch initialises CUDA; an explicit user setting wins.
os.environ.setdefault('PYTORCH_CUDA_ALLOC_CONF', 'expandable_segments:True')
sys.path.insert(0, ROOT)

N_EIG = 10
STATES = (1, 0)


# --------------------------------------------------------------------------------------------
# workloads: BASELINE.json configs[2] (the headline, default), configs[1] and configs[3]
# --------------------------------------------------------------------------------------------
class Workload:
    """One named sweep: circuit builders for the product package and for the oracle, the per-rank grid."""
    name = ''
    text = ''
    points = 0
    dec_types = []

    def build(self, sq, kern):
        raise NotImplementedError

    def oracle(self, o):
        """(oracle circuit, setter(param)) -- used by the reference arm / cpu baseline"""
        raise NotImplementedError

    def grid(self, points, rank, world, strong=False):
        """Per-rank parameter array [n][d].  Weak scaling: EVERY rank sweeps the same range with the same
        number of points and the same spacing (offset by a fraction of the spacing, so no two ranks solve
        the same point): identical work per GPU.  strong=True: `points` is the size of the ONE global
        grid and the rank owns a contiguous shard of it."""
        raise NotImplementedError

    def sweep_args(self, prm):
        raise NotImplementedError

    @staticmethod
    def _line(points, rank, world, strong):
        from sqcircuit_b200.sharding import shard_bounds
        if strong:
            lo, hi = shard_bounds(points, rank, world)
            return (np.arange(lo, hi, dtype=np.float64) + 0.5) / points
        return (np.arange(points, dtype=np.float64) + (rank + 0.5) / world) / points


class ZeroPi(Workload):
    name = 'zeropi4096'
    text = '0-pi qubit, trunc_nums [100,51] (N=10100), 4096-point flux sweep per GPU, k=10, 6 T1/Tphi channels on (1,0)'
    metric = 'diagonalisations/sec (k=10), 0-pi flux sweep'
    points = 4096
    trunc = [100, 51]
    dec_types = ['capacitive', 'inductive', 'quasiparticle', 'charge', 'cc', 'flux']

    def build(self, sq, kern):
        self.loop = sq.Loop()
        C, CJ = sq.Capacitor(0.15, 'GHz'), sq.Capacitor(10, 'GHz')
        JJ, L = sq.Junction(5, 'GHz', loops=[self.loop]), sq.Inductor(0.13, 'GHz', loops=[self.loop])
        cr = sq.Circuit({(0, 1): [CJ, JJ], (0, 2): [L], (0, 3): [C], (1, 2): [C], (1, 3): [L], (2, 3): [CJ, JJ]},
                        kernels=kern)
        cr.set_trunc_nums(self.trunc)
        return cr

    def oracle(self, o):
        lp = o.OLoop()
        c = o.zero_pi(lp)
        c.set_trunc_nums(self.trunc)
        return c, lambda prm: lp.set_flux(float(prm[0]))

    def grid(self, points, rank, world, strong=False):
        return self._line(points, rank, world, strong)[:, None]

    def sweep_args(self, prm):
        return {'loop_fluxes': {self.loop: prm[:, 0]}}


class Fluxonium(Workload):
    name = 'fluxonium1024'
    text = 'fluxonium, harmonic basis trunc=100 (N=100), 1024-point flux sweep per GPU, k=10, 5 T1/Tphi channels on (1,0)'
    metric = 'diagonalisations/sec (k=10), fluxonium flux sweep'
    points = 1024
    dec_types = ['capacitive', 'inductive', 'quasiparticle', 'cc', 'flux']

    def build(self, sq, kern):
        self.loop = sq.Loop()
        cap = sq.Capacitor(3.6, 'GHz', Q=1e6)
        ind = sq.Inductor(0.46, 'GHz', Q=500e6, loops=[self.loop])
        junc = sq.Junction(10.2, 'GHz', cap=cap, A=1e-7, x=3e-06, loops=[self.loop])
        cr = sq.Circuit({(0, 1): [ind, junc]}, flux_dist='all', kernels=kern)
        cr.set_trunc_nums([100])
        return cr

    def oracle(self, o):
        lp = o.OLoop()
        c = o.fluxonium(lp)
        c.set_trunc_nums([100])
        return c, lambda prm: lp.set_flux(float(prm[0]))

    def grid(self, points, rank, world, strong=False):
        return self._line(points, rank, world, strong)[:, None]

    def sweep_args(self, prm):
        return {'loop_fluxes': {self.loop: prm[:, 0]}}


class TransmonResonatorTransmon(Workload):
    name = 'trt1e5'
    text = ('transmon-resonator-(flux-tunable transmon), trunc_nums [10,50,50] (N=98010), 16x16 grid over (gate charge '
            'of transmon 1) x (external flux of the SQUID of transmon 2) per GPU, k=10, 4 T1/Tphi channels on (1,0)')
    metric = 'diagonalisations/sec (k=10), transmon-resonator-transmon gate-charge x flux grid'
    points = 256
    trunc = [10, 50, 50]
    dec_types = ['capacitive', 'charge', 'cc', 'flux']

    def build(self, sq, kern):
        self.loop = sq.Loop()
        cr = sq.Circuit({(0, 1): [sq.Capacitor(0.5, 'GHz'), sq.Inductor(8.0, 'GHz')],
                         (0, 2): [sq.Capacitor(0.30, 'GHz'), sq.Junction(14.0, 'GHz')],
                         (0, 3): [sq.Capacitor(0.27, 'GHz'), sq.Junction(9.0, 'GHz', loops=[self.loop]),
                                  sq.Junction(8.0, 'GHz', loops=[self.loop])],
                         (1, 2): [sq.Capacitor(12.0, 'GHz')], (1, 3): [sq.Capacitor(15.0, 'GHz')]}, kernels=kern)
        cr.set_trunc_nums(self.trunc)
        self.modes = [i + 1 for i in sorted(cr.charge_islands)]
        return cr

    def oracle(self, o):
        lp = o.OLoop()
        c = o.OCircuit({(0, 1): [o.OCap(0.5, 'GHz'), o.OInd(8.0, 'GHz')],
                        (0, 2): [o.OCap(0.30, 'GHz'), o.OJJ(14.0, 'GHz')],
                        (0, 3): [o.OCap(0.27, 'GHz'), o.OJJ(9.0, 'GHz', loops=[lp]), o.OJJ(8.0, 'GHz', loops=[lp])],
                        (1, 2): [o.OCap(12.0, 'GHz')], (1, 3): [o.OCap(15.0, 'GHz')]})
        c.set_trunc_nums(self.trunc)
        modes = [i + 1 for i in range(c.n) if c.is_charge(i)]

        def setter(prm):
            c.set_charge_offset(modes[0], float(prm[0]))
            lp.set_flux(float(prm[1]))
        return c, setter

    def grid(self, points, rank, world, strong=False):
        # square grid over [0, 0.5)^2 (gate charge, flux), row-major; the line index of _line walks the grid
        side = int(round(np.sqrt(points)))
        t = self._line(side * side, rank, world, strong) * side * side      # fractional cell index
        cell = np.floor(t).astype(int)
        frac = t - cell
        return np.stack([((cell // side) + frac) / side * 0.5, ((cell % side) + 0.5) / side * 0.5], axis=1)

    def sweep_args(self, prm):
        return {'charge_offsets': {self.modes[0]: prm[:, 0]}, 'loop_fluxes': {self.loop: prm[:, 1]}}


class Discovery(Workload):
    """BASELINE configs[4]: qubit-discovery batch -- 512 candidates of one 4-mode circuit graph (two harmonic and
    two charge modes after the transformation, three junction terms, one flux loop; element values = a base
    design times a deterministic +-20 % spread), every candidate diagonalised (k = 10) and the gradient of its
    01 transition frequency taken with respect to its 12 element values and its loop flux.  One step = new
    element values in (host), eigenfrequencies [B][10] and gradients [B][13] out (host)."""
    name = 'discovery512'
    text = ('4-mode qubit-discovery circuit, trunc_nums [10,10,5,5] (N=8100), batch of 512 element-value sets per GPU, '
            'k=10, torch autograd gradients of the 01 transition frequency w.r.t. 13 parameters per circuit')
    metric = 'diagonalisations/sec (k=10) with element-value gradients, 4-mode discovery batch'
    points = 512
    trunc = [10, 10, 5, 5]
    dec_types = []
    NAMES = ['C01', 'J01', 'C12', 'L12', 'C20', 'J20', 'C23', 'C30', 'J30', 'C34', 'C40', 'L40']
    BASE = np.array([3.0, 8.0, 5.0, 0.6, 4.0, 6.0, 20.0, 1.2, 12.0, 25.0, 0.8, 1.5])

    def values(self, idx):
        rng = np.random.default_rng(1000 + int(idx))
        return self.BASE * (1 + 0.2 * rng.uniform(-1, 1, len(self.BASE))) if idx else self.BASE.copy()

    def grid(self, points, rank, world, strong=False):
        from sqcircuit_b200.sharding import shard_bounds
        lo, hi = shard_bounds(points, rank, world) if strong else (rank * points, (rank + 1) * points)
        return np.stack([self.values(i) for i in range(lo, hi)])

    def _circuit(self, sq, kern, v):
        lp = sq.Loop(0.37, requires_grad=True)
        mk = {'C': sq.Capacitor, 'J': sq.Junction, 'L': sq.Inductor}
        els = {}
        for nm, val in zip(self.NAMES, v):
            extra = {'loops': [lp]} if nm in ('J01', 'L12', 'J20') else {}
            els[nm] = mk[nm[0]](float(val), 'GHz', requires_grad=True, **extra)
        cr = sq.Circuit({(0, 1): [els['C01'], els['J01']], (1, 2): [els['C12'], els['L12']],
                         (2, 0): [els['C20'], els['J20']], (2, 3): [els['C23']],
                         (3, 0): [els['C30'], els['J30']], (3, 4): [els['C34']],
                         (4, 0): [els['C40'], els['L40']]}, kernels=kern)
        cr.set_trunc_nums(self.trunc)
        els['loop'] = lp
        return cr, els

    def build(self, sq, kern):
        """The Circuit returned only carries the truncation for the bench's bookkeeping; batches are built per
        grid by make_batch."""
        sq.set_engine('PyTorch')
        self.sq, self.kern = sq, kern
        self._batches = {}
        return self._circuit(sq, kern, self.BASE)[0]

    def make_batch(self, n):
        if n not in self._batches:
            pairs = [self._circuit(self.sq, self.kern, self.BASE) for _ in range(n)]
            self._batches[n] = (self.sq.CircuitBatch([c for c, _ in pairs], kernels=self.kern), [e for _, e in pairs])
        return self._batches[n]

    def step(self, prm, torch):
        """prm: host array [n][12] of element values (GHz).  Returns (efreqs [n][k] device tensor, grads [n][13]
        host array, SweepResult)."""
        batch, elss = self.make_batch(prm.shape[0])
        with torch.no_grad():
            for els, v in zip(elss, prm):
                for nm, val in zip(self.NAMES, v):
                    el = els[nm]
                    el.set_value(float(val), 'GHz')
                    el.requires_grad = True
                    el.grad = None
                els['loop'].grad = None
        batch.update()
        ef, _ = batch.diag(N_EIG)
        (ef[:, 1] - ef[:, 0]).sum().backward()
        names = self.NAMES + ['loop']
        grads = np.array([[float(els[nm].grad) for nm in names] for els in elss])
        return ef.detach(), grads, batch._last

    def oracle(self, o):
        raise NotImplementedError('built per candidate in _ref_one')


WORKLOADS = {w.name: w for w in (ZeroPi, Fluxonium, TransmonResonatorTransmon, Discovery)}
# bounded CPU sample: sweep points per host core and step (~10-30 s of CPU work per step)
REF_POINTS_PER_CORE = {'zeropi4096': 1, 'fluxonium1024': 512, 'trt1e5': 1, 'discovery512': 1}


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


_REF_CACHE = {}


def _ref_one(job):
    """One sweep point the way the reference does it: CSR Hamiltonian from Kronecker products, ARPACK
    eigs(which='SR') (circuit.py:1962-1969), then the decoherence rates with those eigenvectors."""
    import warnings
    warnings.filterwarnings('ignore')
    import scipy.sparse.linalg as spla
    from oracle import sqc_oracle as o
    name, prm = job
    if name == 'discovery512':
        # one candidate the way the reference's optimisation loop does it: a new Circuit (transformation,
        # operator memory), diag, then torch_extensions.EigenSolver.backward = get_partial_omega per
        # (element, state) -- circuit.py:2777-2826 -- for the two states of the loss
        lp = o.OLoop(0.37)
        c, els = o.discovery(prm, lp)
        c.set_trunc_nums(Discovery.trunc)
        H = c.hamiltonian()
        w, v = spla.eigs(H, k=N_EIG, which='SR')
        order = np.argsort(w.real)
        c.efreqs_rad, c.evecs = w.real[order], v[:, order]
        g = [c.partial_omega(els[nm], 1) - c.partial_omega(els[nm], 0) for nm in Discovery.NAMES + ['loop']]
        return c.efreqs_rad[:2].tolist(), g
    if name not in _REF_CACHE:          # one circuit per worker process, like a Sweep over one Circuit object
        wl = WORKLOADS[name]()
        _REF_CACHE[name] = (wl,) + tuple(wl.oracle(o))
    wl, c, setter = _REF_CACHE[name]
    setter(prm)
    H = c.hamiltonian()
    w, v = spla.eigs(H, k=N_EIG, which='SR')      # the reference's call (circuit.py:1963)
    order = np.argsort(w.real)
    c.efreqs_rad, c.evecs = w.real[order], v[:, order]
    rates = [float(np.real(c.dec_rate(d, STATES))) for d in wl.dec_types]
    return c.efreqs_rad[:2].tolist(), rates


def run_reference_sample(wl, n_points, procs):
    """Diagonalise n_points sweep points with `procs` worker processes; returns seconds."""
    import multiprocessing as mp
    prm = wl.grid(n_points, 0, 1)
    jobs = [(wl.name, prm[i]) for i in range(n_points)]
    t0 = time.perf_counter()
    if procs <= 1:
        for jb in jobs:
            _ref_one(jb)
    else:
        with mp.get_context('fork').Pool(procs, initializer=_ref_init) as pool:
            pool.map(_ref_one, jobs, chunksize=1)
    return time.perf_counter() - t0


def _host_procs():
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    return max(1, min(cores, 64))


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    for var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[var] = '1'
    wl = WORKLOADS[args.config]()
    procs = _host_procs()
    # bounded sample: the small circuit is cheap per point, so it gets more points per worker
    n_points = procs * (REF_POINTS_PER_CORE.get(wl.name, 1))
    # a CPU loop has no clocks or caches to warm beyond the first pass: one warm-up pass whatever W asks for
    for _ in range(1 if args.warmup >= 1 else 0):
        run_reference_sample(wl, n_points, procs)
    times = [run_reference_sample(wl, n_points, procs) for _ in range(max(1, args.steps))]
    t = float(np.mean(times))
    value = n_points / t
    line = {
        'impl': 'reference', 'metric': wl.metric, 'value': value,
        'unit': 'diag/s', 'n_gpus': args.gpus, 'steps': max(1, args.steps), 'warmup': min(args.warmup, 1),
        'warmup_note': 'CPU loop: one warm-up pass is run whatever --warmup asks for',
        'ms_per_step': 1e3 * t, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64 (complex128)', 'data': 'synthetic',
        'config': {'workload': wl.text, 'points_per_step': n_points,
                   'note': f'bounded sample of the {wl.points}-point sweep'},
        'cpu_baseline': {'value': value, 'unit': 'diag/s', 'cores': procs, 'kind': 'port',
                         'sample': f'{n_points} sweep points per step, one process per core, scipy CSR build + '
                                   'ARPACK eigs(k=10, which=SR) + dec_rate per channel (oracle/sqc_oracle.py)'},
        'e2e': {'value': value, 'unit': 'diag/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of the timed regions, in-process through NVML.

    Measured on this pool (profiles/r02_clock_sampler_stall.txt): a polling ``nvidia-smi -lms 200`` child, or NVML
    queries from a polling thread, stall single 300 ms steps of the bench by 30-200 ms (the query serialises with
    kernel launches in the driver; nvmlDeviceGetCurrentClocksEventReasons is the worst).  So the timed loop itself
    reads the SM clock once per step, right after enqueueing it; the event reasons are read right before the
    first and right after the last timed step, and throttling in between is caught by the cumulative NVML
    violation counters (thermal / power / reliability), whose deltas over the regions are reported."""
    REASONS = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20, 'sw_power_cap': 0x4}
    POLICIES = {'sw_power_cap': 0, 'sw_thermal_slowdown': 1, 'hw_slowdown': 5}   # NVML_PERF_POLICY_POWER / THERMAL / RELIABILITY

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.sm, self.mx, self.reasons = [], [], set()
        self._nvml = self._h = None
        self._viol0 = {}
        self._cost = 0.0

    def _physical_index(self):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
        ids = [x.strip() for x in vis.split(',') if x.strip()]
        if ids and self.idx < len(ids) and ids[self.idx].isdigit():
            return int(ids[self.idx])
        return self.idx

    def _read_reasons(self):
        try:
            mask = int(self._nvml.nvmlDeviceGetCurrentClocksEventReasons(self._h))
            for nm, bit in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(nm)
        except Exception:
            pass

    def _violations(self):
        out = {}
        for nm, pol in self.POLICIES.items():
            try:
                out[nm] = int(self._nvml.nvmlDeviceGetViolationStatus(self._h, pol).violationTime)
            except Exception:
                pass
        return out

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None
            return
        self._read_reasons()
        self._viol0 = self._violations()
        self._cost = 0.0

    def sample(self):
        """One SM-clock reading, called by the timed loop right after it has enqueued a step (inside the timed
        region, from the launching thread itself: nothing races with the launches; the GPU is busy with the
        step just enqueued, so the reading is taken under load)."""
        if self._nvml is None:
            return
        t = time.perf_counter()
        try:
            self.sm.append(float(self._nvml.nvmlDeviceGetClockInfo(self._h, self._nvml.NVML_CLOCK_SM)))
            self.mx.append(self._max)
        except Exception:
            pass
        self._cost += time.perf_counter() - t

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self._nvml is None:
            return out
        self._read_reasons()
        viol = {}
        for nm, t1 in self._violations().items():
            dt = t1 - self._viol0.get(nm, t1)
            viol[nm] = dt
            if dt > 0:
                self.reasons.add(nm)
        if self.sm:
            out = {'sm_mhz': float(np.median(self.sm)), 'sm_max_mhz': float(np.max(self.mx)),
                   'reasons': sorted(self.reasons), 'samples': len(self.sm),
                   'source': 'NVML: SM clock read by the timed loop after every enqueued step (GPU under load); '
                             'event reasons before / after the timed regions + violation-counter deltas (ns) '
                             'over them', 'violation_ns': viol, 'sampling_ms_total': round(1e3 * self._cost, 2)}
        return out


class TimedKernels:
    """Wraps the CUDA backend: CUDA events on the launching stream around every C-ABI call,
    so per-kernel device time is measured live inside the timed region (no profiler)."""
    TIMED = ['hx_fused', 'hx_fused_real', 'hx_generic', 'real_to_complex', 'complex_to_real', 'mode_transform', 'potential_apply', 'potential_elements', 'gram', 'gram_pair', 'gen_rr', 'davidson_insert2',
             'lowblock_assemble', 'lowblock_assemble_generic', 'hamiltonian_dense', 'hpd_inverse_c64', 'lowblock_apply', 'lowband_factor', 'lowband_apply', 'block_update', 'ritz_update', 'block_rightmul', 'heev', 'syev',
             'residual_norms', 'precond_residual', 'restart_copy', 'svqb_coeffs', 'davidson_decide',
             'davidson_insert', 'diag_mul', 'ladder_apply']

    def __init__(self, kern):
        import torch
        self._k = kern
        self._torch = torch
        self.enabled = False
        self.only = None       # restrict the instrumentation to these entry points
        self.records = []      # (name, start_event, stop_event, meta)
        self._pool = []        # pre-created event pairs
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
            if self._pool:
                s, e = self._pool.pop()
            else:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **kw)
            e.record()
            self.records.append((name, s, e))
            return r
        return call

    def reserve(self, n):
        """Create (and record once, so that the driver objects exist) n event pairs outside the timed region:
        cudaEventCreate inside it would be charged to `value`."""
        torch = self._torch
        while len(self._pool) < n:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            e.record()
            self._pool.append((s, e))

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
    _debug_level_times()
    if args.no_graphs:
        kern._k.use_graphs = False

    wl = WORKLOADS[args.config]()
    cr = wl.build(sq, kern)
    points = args.points or wl.points
    N = int(np.prod(cr.m))
    solver_kw = dict(tol=args.tol, max_basis=args.mmax, lowblock=args.lowblock, block_size=args.block,
                     real=(False if args.complex else None))
    if args.strides:
        solver_kw['coarse_stride'] = tuple(int(x) for x in args.strides.split(','))

    host_out = {}
    stage_wall = []

    def to_host(t):
        """D2H of a step's result into pinned host memory (the read of the result the e2e contract asks for)."""
        key = (tuple(t.shape), t.dtype)
        if key not in host_out:
            host_out[key] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
        host_out[key].copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host_out[key]

    def make_step(prm_host, n_total):
        """One pass of the hot path over this rank's points.  e2e=True: host parameter array in, host
        eigenfrequencies + rates out (the user-facing call); otherwise device-resident."""
        def step(e2e):
            if e2e:
                prm = prm_host.to(dev, non_blocking=True).cpu().numpy()   # H2D of the step's inputs
            else:
                prm = prm_host.numpy()
            if hasattr(wl, 'step'):
                # batched circuits with gradients: the element values ARE the step's input (operator tables are
                # rebuilt from them every step, e2e or not), gradients come back to the host
                ef, grads, res = wl.step(prm, torch)
                if world > 1:
                    gather_to_root(ef.contiguous(), n_total)
                return (to_host(ef) if e2e else ef), grads, res
            if DEBUG:
                torch.cuda.synchronize()
            t0_ = time.perf_counter()
            res = cr.sweep_diag(N_EIG, **wl.sweep_args(prm), **solver_kw)
            ta_ = time.perf_counter()
            if DEBUG:
                torch.cuda.synchronize()
                t1_ = time.perf_counter()
            rates = [res.dec_rate(d, states=STATES) for d in wl.dec_types]
            if DEBUG:
                torch.cuda.synchronize()
                print(f'[debug] sweep_diag {1e3 * (t1_ - t0_):.1f} ms (levels {sum(_LEVEL_MS):.1f}), rates '
                      f'{1e3 * (time.perf_counter() - t1_):.1f} ms', file=sys.stderr)
            if world > 1:
                # the only collective of the path: final gather of the per-point results to rank 0
                gather_to_root(res.efreqs.contiguous(), n_total)
            if e2e:
                tb_ = time.perf_counter()
                out = to_host(res.efreqs)
                # host wall clock of the stages of this step: sweep_diag (returns once its last level has been
                # judged on the host), decay rates (every channel ends with a host read), final D2H
                stage_wall.append([round(1e3 * (ta_ - t0_), 1), round(1e3 * (tb_ - ta_), 1),
                                   round(1e3 * (time.perf_counter() - tb_), 1),
                                   [round(x, 1) for x in _LEVEL_MS[-4:]]])
                return out, rates, res
            return res.efreqs, rates, res
        return step

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = None
    step_wall = []         # host wall time of every step of the last timed region (an e2e step ends with a host read)

    def timed(step, nsteps, e2e):
        """K steps bracketed by barrier + synchronize; returns (max over ranks of the device time, this
        rank's own device time, last result)."""
        sync()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        last = None
        step_wall.clear()
        for _ in range(nsteps):
            tw = time.perf_counter()
            last = None          # a sweep loop keeps the newest result only (two resident results double the peak)
            if DEBUG:
                torch.cuda.synchronize()
                ts = time.perf_counter()
            _LEVEL_MS.clear()
            last = step(e2e)
            if sampler is not None:
                sampler.sample()
            step_wall.append(round(1e3 * (time.perf_counter() - tw), 1))
            if DEBUG:
                torch.cuda.synchronize()
                st_ = torch.cuda.memory_stats()
                print(f'[debug] levels (ms, host wall): {[round(x, 1) for x in _LEVEL_MS]}', file=sys.stderr)
                print(f'[debug] step e2e={e2e} {1e3 * (time.perf_counter() - ts):.1f} ms  cudaMalloc segments '
                      f'{st_["segment.all.allocated"]} retries {st_["num_alloc_retries"]} reserved '
                      f'{st_["reserved_bytes.all.current"] / 2**30:.1f} GiB peak-alloc '
                      f'{st_["allocated_bytes.all.peak"] / 2**30:.1f} GiB', file=sys.stderr)
        t1.record()
        torch.cuda.synchronize()
        own = t0.elapsed_time(t1)          # this rank alone, before the closing barrier
        sync()
        ms = torch.tensor([own], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), own, last

    def per_rank(values):
        """[world] list of a per-rank scalar, on every rank"""
        t = torch.tensor([float(values)], device=dev, dtype=torch.float64)
        if world == 1:
            return [float(values)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(x.item()) for x in out]

    prm_weak = torch.from_numpy(np.ascontiguousarray(wl.grid(points, rank, world))).pin_memory()
    step = make_step(prm_weak, points * world)
    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get('SQC_BENCH_NO_SAMPLER'):      # (dev switch: is the sampling perturbing the run?)
        sampler.start()
    # timed region: only the dominant kernel (the H.X apply) carries CUDA events -- bracketing all
    # ~900 C-ABI calls of a step costs ~60 ms of host time per step (measured), which would be
    # charged to `value`.  The full per-kernel breakdown comes from one extra, untimed step below.
    kern.only = {'hx_fused', 'hx_fused_real', 'hx_generic', 'mode_transform', 'potential_apply', 'heev', 'syev'}
    kern.reserve(512 * (max(args.steps, 1) + 1))
    kern.enabled = True
    # one more untimed step in exactly the mode of the timed region (H.X outside the iteration's CUDA graphs, which
    # are then captured in two pieces): its first-time costs showed up in `value` on a fresh box otherwise
    step(False)
    torch.cuda.synchronize()
    kern.records = []
    for key_ in list(HX_COUNTER):
        HX_COUNTER[key_] = 0
    l0 = kern.launch_count()
    ms_dev, own_dev, last = timed(step, args.steps, False)
    launches = kern.launch_count() - l0
    kern.enabled = False
    hx_kernel = kern.summary()
    nvec = int(HX_COUNTER.get('vectors', 0)) + int(HX_COUNTER.get('vectors_dev', 0))
    nvec_real = int(HX_COUNTER.get('real_vectors', 0)) + int(HX_COUNTER.get('real_vectors_dev', 0))
    iters = last[2].iterations
    del last                      # (its GB of eigenvectors would stay resident during the next region)
    step(True)                    # one untimed pass of the end-to-end path (pinned-memory copies, host reads)
    ms0_ = torch.cuda.memory_stats()
    ms_e2e, own_e2e, last_e2e = timed(step, args.steps, True)
    ms1_ = torch.cuda.memory_stats()
    alloc_events = {k: int(ms1_.get(k, 0) - ms0_.get(k, 0)) for k in ('num_device_alloc', 'num_device_free',
                                                                       'num_alloc_retries', 'num_ooms')}
    e2e_step_wall = list(step_wall)
    del last_e2e
    clocks = sampler.stop() if rank == 0 else None
    rank_ms = per_rank(own_dev / args.steps)
    rank_iters = per_rank(iters)

    # ---- strong scaling of the same sweep: ONE grid of `points` fluxes split over the ranks ----
    strong = None
    if world > 1 and not args.no_strong:
        prm_strong = torch.from_numpy(np.ascontiguousarray(wl.grid(points, rank, world, strong=True))).pin_memory()
        sstep = make_step(prm_strong, points)
        for _ in range(max(2, min(args.warmup, 3))):
            sstep(False)
        ms_s, own_s, last_s = timed(sstep, args.steps, False)
        s_rank_ms = per_rank(own_s / args.steps)
        s_rank_iters = per_rank(last_s[2].iterations)
        del last_s
        strong = {'points_total': points, 'points_per_gpu': int(prm_strong.shape[0]),
                  'value': points * args.steps / (ms_s / 1e3), 'unit': 'diag/s', 'ms_per_step': ms_s / args.steps,
                  'per_rank_ms_per_step': [round(x, 2) for x in s_rank_ms],
                  'per_rank_iterations': [int(x) for x in s_rank_iters],
                  'note': 'the named 4096-point sweep split into contiguous shards, same timing rules; '
                          'efficiency = value / (n_gpus x the 1-GPU value of this bench)'}

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

    # ---- roofline of the dominant kernel, measured live ----
    peaks = {}
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'MEASURED_PEAKS.json (sustained copy)' if peaks else 'fallback 6.65 TB/s'
    total_ms = sum(v[0] for v in per_kernel.values())
    shares = {k: {'ms': round(v[0], 3), 'calls': v[1], 'share': round(v[0] / total_ms, 4)}
              for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][0])}
    real_ms = hx_kernel.get('hx_fused_real', [0.0, 0])[0]
    cplx_ms = sum(hx_kernel.get(k, [0, 0])[0] for k in ('hx_fused', 'hx_generic', 'mode_transform', 'potential_apply'))
    heev_ms = hx_kernel.get('heev', [0.0, 0])[0]
    syev_ms = hx_kernel.get('syev', [0.0, 0])[0]
    use_real = real_ms > cplx_ms
    m0 = cr.m[0]
    if max(heev_ms, syev_ms) > max(real_ms, cplx_ms):
        # dense path (N <= 112): the batched Jacobi eigensolver dominates; judged against the FP64 pipe.
        # Algorithmic work of a Hermitian eigendecomposition with vectors: ~9 n^3 multiply-adds
        # (tridiagonalisation 4/3 + QR with vectors ~6 + back-transformation 2): 72 n^3 real flop for a complex
        # matrix (sqc_heev), 18 n^3 for a real symmetric one (sqc_syev: circuits without charge modes)
        sym = syev_ms > heev_ms
        name, ev_ms, per_mat = ('syev', syev_ms, 18.0 * N ** 3) if sym else ('heev', heev_ms, 72.0 * N ** 3)
        calls = hx_kernel[name][1]
        nmat = points * args.steps
        flops = per_mat * nmat
        roof = {'bound': 'tensor', 'kernel': ('sqc_syev (batched real symmetric Jacobi eigensolver, matrix and '
                                              'eigenvectors in shared memory' if sym else
                                              'sqc_heev (batched Hermitian Jacobi eigensolver') +
                ', one CTA per sweep point; SIMT FP64 -- there is no FP64 tcgen05, the FP64 pipe is the bound)',
                'unit': 'TFLOP/s',
                'peak': 37.1, 'peak_source': 'measured FP64 DMMA/DFMA issue rate (tools/cu/dmma_peak.cu, '
                'profiles/r01_dmma_peak.txt); MEASURED_PEAKS.json has no FP64 figure',
                'achieved': flops / (ev_ms / 1e3) / 1e12, 'traffic': None, 'calls': calls, 'ms': ev_ms,
                'algorithmic_flops_per_matrix': per_mat, 'matrices': nmat}
        roof['frac'] = roof['achieved'] / roof['peak']
    else:
        if use_real:
            hx_ms, hx_calls, nv = real_ms, hx_kernel['hx_fused_real'][1], nvec_real
            bytes_per_vec = 16 * N        # read N doubles, write N doubles
            kname = ('H.X apply, real representation (sqc_hx_fused_real: TMA load, transform + potential + '
                     'transform, TMA store)')
            flops_per_vec = 2.0 * m0 * N
        elif hx_kernel.get('hx_generic', [0.0, 0])[0] > 0:
            hx_ms, nv = hx_kernel['hx_generic'][0], nvec
            hx_calls = hx_kernel['hx_generic'][1]
            bytes_per_vec = 32 * N
            kname = ('H.X apply, generic layout in one kernel (sqc_hx_generic: TMA-staged tile of all harmonic '
                     'indices x a window of charge columns, transforms + potential + transforms in shared memory)')
            flops_per_vec = 4.0 * sum(d for d, h in zip(cr.m, cr._operators().harm) if h and d > 1) * N
        elif hx_kernel.get('hx_fused', [0.0, 0])[0] > 0:
            hx_ms, nv = cplx_ms, nvec
            hx_calls = hx_kernel['hx_fused'][1]
            bytes_per_vec = 32 * N        # read N complex, write N complex
            kname = 'H.X apply (sqc_hx_fused: transform + potential + transform in one kernel)'
            flops_per_vec = 4.0 * m0 * N
        else:
            hx_ms, nv = cplx_ms, nvec
            hx_calls = hx_kernel.get('potential_apply', [0, 1])[1]
            bytes_per_vec = 32 * N
            kname = ('H.X apply, generic layout (sqc_mode_transform per harmonic mode + sqc_potential_apply: '
                     'several passes over the vector; algorithmic bytes = one read + one write)')
            flops_per_vec = 4.0 * sum(d for d, h in zip(cr.m, cr._operators().harm) if h and d > 1) * N
        roof = {'bound': 'hbm', 'kernel': kname, 'unit': 'GB/s', 'peak': hbm_peak, 'peak_source': peak_src,
                'traffic': None, 'achieved': None, 'frac': None}
        if hx_ms > 0 and nv > 0:
            alg_bytes = float(bytes_per_vec) * nv
            roof['achieved'] = alg_bytes / (hx_ms / 1e3) / 1e9
            roof['frac'] = roof['achieved'] / hbm_peak
            roof['algorithmic_bytes_per_vector'] = bytes_per_vec
            roof['algorithmic_bytes_per_launch'] = alg_bytes / max(hx_calls, 1)
            ratio, src = NCU_HX_TRAFFIC.get((wl.name, 'real' if use_real else 'complex'), (None, None))
            if ratio is not None:
                # DRAM bytes of the kernel from its ncu --set full capture, as a ratio to the algorithmic bytes
                # of the captured launch; reported per launch, scaled to this run's average launch
                roof['traffic'] = ratio * alg_bytes / max(hx_calls, 1)
                roof['traffic_source'] = src
            roof['vectors'] = nv
            roof['ms'] = hx_ms
            flops = flops_per_vec * nv
            roof['fp64_tflops'] = flops / (hx_ms / 1e3) / 1e12
            # context for the HBM fraction: FP64 issue is the real bound of the harmonic-mode transforms.
            # Peak = measured DMMA issue rate (tools/cu/dmma_peak.cu, profiles/r01_dmma_peak.txt)
            roof['fp64_tensor_peak_tflops'] = 37.1
            roof['fp64_tensor_frac'] = roof['fp64_tflops'] / 37.1
            roof['hbm_frac_ceiling_fp64_bound'] = min(1.0, (37.1e12 / (flops / alg_bytes)) / 1e9 / hbm_peak)
        roof['calls'] = hx_calls

    # ---- cpu baseline: bounded sample of the same workload on the host cores ----
    cpu = None
    if not args.no_cpu_baseline:
        for var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
            os.environ[var] = '1'
        procs = _host_procs()
        npts = procs * (REF_POINTS_PER_CORE.get(wl.name, 1))
        t = run_reference_sample(wl, npts, procs)
        cpu = {'value': npts / t, 'unit': 'diag/s', 'cores': procs, 'kind': 'port',
               'sample': f'{npts} sweep points of the same sweep, one process per core (oracle port of the '
                         f'reference: scipy CSR build + ARPACK eigs SR + the rates), {t:.1f} s'}

    h2d = int(prm_weak.numel()) * 8
    d2h = points * N_EIG * 8 + len(wl.dec_types) * points * 8 + (points * 13 * 8 if hasattr(wl, 'step') else 0)
    line = {
        'metric': wl.metric, 'value': value, 'unit': 'diag/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_dev / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': wl.text, 'name': wl.name,
                   'points_per_gpu': points, 'n_eig': N_EIG, 'hilbert_dim': N, 'tol_rel_residual': args.tol,
                   'davidson_iterations': iters,
                   'representation': 'real (time-reversal adapted)' if use_real else 'complex',
                   'l2': 'working set (tens of GB of basis vectors) larger than L2' if N > 112 else
                         'per-point matrices live in shared memory; inputs are a few KB',
                   'parallelism': f'every one of the {world} GPU(s) sweeps the same range with {points} points '
                                  '(offset by a fraction of the spacing): identical work per GPU, final NCCL gather'},
        'e2e': {'value': e2e_value, 'unit': 'diag/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': ms_e2e / args.steps, 'step_wall_ms': e2e_step_wall, 'allocator_events': alloc_events,
                'stage_wall_ms(sweep_diag, rates, d2h)': stage_wall[-args.steps:]},
        'gpu_launches': int(launches),
        'per_rank': {'ms_per_step': [round(x, 2) for x in rank_ms], 'iterations': [int(x) for x in rank_iters]},
        'strong': strong,
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
    ('zeropi4096', 'complex'): (1.0855, 'ncu dram__bytes_read+write / algorithmic = 1.0855 '
                                        '(profiles/r01_ncu_full_v28_hx_fused.txt)'),
    ('zeropi4096', 'real'): (1.28, 'ncu dram__bytes_read+write / algorithmic = 1.28 on the launches that carry most of a '
                                   "step's bytes: 2.27 GB read + 2.82 GB written by one launch on 24576 vectors of 161.6 kB "
                                   '(largest continuation level, profiles/r02_ncu_full_hx_real_lag2.txt; the excess over '
                                   'the 1.99 GB of results is write-back traffic the kernel does not ask for by count -- '
                                   'presumably its 268 B of register-spill stores per thread); 0.87 on a 4140-vector '
                                   'launch whose input still sits in L2 (r02_ncu_full_hx_real_final.txt), 0.98 on a cold '
                                   '12288-vector launch (r02_ncu_full_hx_real_v1.txt)'),
    ('trt1e5', 'complex'): (1.09, 'ncu dram__bytes_read+write / algorithmic = 1.09: 1.459 GB + 1.171 GB for 768 vectors '
                                  'of 3.136 MB, halo columns read twice (profiles/r02_hx_generic.txt)'),
    ('discovery512', 'complex'): (0.99, 'ncu dram__bytes_read+write / algorithmic = 0.99: 0.800 GB + 0.776 GB for 6144 '
                                        'vectors of 259.2 kB (profiles/r02_hx_generic.txt)'),
}
HX_COUNTER = {'vectors': 0}
DEBUG = bool(os.environ.get('SQC_BENCH_DEBUG'))
_LEVEL_MS = []


def _debug_level_times():
    """SQC_BENCH_DEBUG: host wall time of every batched Davidson call of a step (one per continuation level)."""
    import sqcircuit_b200.circuit as C
    inner = C._davidson

    def timed_davidson(*a, **kw):
        t = time.perf_counter()
        r = inner(*a, **kw)
        _LEVEL_MS.append(1e3 * (time.perf_counter() - t))
        return r
    C._davidson = timed_davidson


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
    ap.add_argument('--points', type=int, default=0, help='sweep points per GPU (default: the workload\'s)')
    ap.add_argument('--config', default='zeropi4096', choices=sorted(WORKLOADS), help='named workload (default: the headline)')
    ap.add_argument('--no-strong', action='store_true', help='skip the strong-scaling section of a multi-GPU run')
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
