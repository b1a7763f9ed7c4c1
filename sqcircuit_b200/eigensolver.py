"""Batched lowest-k Hermitian eigensolver over the points of a parameter sweep.

Replaces ``scipy.sparse.linalg.eigs(H, k, which='SR')`` in ``Circuit._diag_np``
(reference circuit.py:1938-1986), one call per sweep point in ``Sweep.sweepFlux``
(sweep.py:71-108), by ONE lock-step batched solve on the GPU:

* N <= DENSE_MAX : H(b) is materialised by applying the matrix-free operator to the identity
  and handed to the batched Jacobi eigensolver (``sqc_heev``).
* otherwise      : block Davidson with the LC diagonal as preconditioner.  Per iteration and
  sweep point:  Rayleigh-Ritz on the projected matrix G = V^H H V (``sqc_heev``), Ritz
  vectors / residuals (``sqc_block_update``, ``sqc_residual_norms``), preconditioned residuals
  T = (d_LC - theta)^-1 r written into the free tail of the basis, two passes of block
  Gram-Schmidt against V plus SVQB inside the block (``sqc_gram`` on the FP64 tensor cores,
  ``sqc_svqb_coeffs``), W_new = H T (matrix-free) and the new columns of G.  Basis sizes are
  per-point device counters; the host only reads one int per iteration (#unconverged points).

All numerics run in the CUDA library; torch is used for allocation and trivial glue.
"""
from dataclasses import dataclass

import torch

from .kernels import BlockView

DENSE_MAX = 112


@dataclass
class EigResult:
    evals: torch.Tensor      # [B][k] float64 (rad/s)
    evecs: torch.Tensor      # [B][k][N] complex128
    iterations: int
    resmax: torch.Tensor     # [B] final relative residual of the wanted pairs
    converged: bool
    ritz_block: torch.Tensor = None   # [B][p][N] all Ritz vectors of the last iteration


def solve_dense(kern, op, B, N, k):
    dev = op.device
    if hasattr(op, 'dense_matrix'):
        # H(b) assembled directly from the per-mode Fock-basis tables (sqc_hamiltonian_dense): one small kernel
        # instead of pushing the identity through the matrix-free operator (measured on the 1024-point fluxonium
        # sweep: 23 ms of transform passes on 102 400 unit vectors)
        H = op.dense_matrix()
        w = torch.empty((B, N), dtype=torch.float64, device=dev)
        Zt = torch.empty((B, N, N), dtype=H.dtype, device=dev)
        if H.dtype == torch.float64:
            kern.syev(H, w, Zt)           # real symmetric: a quarter of the flops, eigenvectors in shared memory
        else:
            kern.heev(H, N, w, Zt)
        k = min(k, N)
        return EigResult(w[:, :k].contiguous(), Zt[:, :k, :].to(torch.complex128).contiguous(), 0,
                         torch.zeros(B, dtype=torch.float64, device=dev), True)
    eye = torch.eye(N, dtype=torch.complex128, device=dev).unsqueeze(0).expand(B, N, N).contiguous()
    HX = torch.empty_like(eye)
    op.apply(BlockView(eye), BlockView(HX))
    # HX[b][j][:] is column j of H, i.e. the array holds H^T = conj(H): eigenvectors come out
    # conjugated.
    w = torch.empty((B, N), dtype=torch.float64, device=dev)
    Zt = torch.empty((B, N, N), dtype=torch.complex128, device=dev)
    kern.heev(HX, N, w, Zt)
    k = min(k, N)
    return EigResult(w[:, :k].contiguous(), Zt[:, :k, :].conj().resolve_conj().contiguous(), 0,
                     torch.zeros(B, dtype=torch.float64, device=dev), True)


def cold_start_block(diag, B, p, N, dev, k=None, start_noise=0.03, real=False):
    """Start block of a cold solve: unit vectors on the p smallest entries of the LC diagonal plus a
    small deterministic perturbation on the 8p smallest.  The perturbation matters at symmetric
    operating points (e.g. the 0-pi qubit at zero flux): H then commutes with a parity, every
    Davidson vector stays inside the symmetry sectors of the start block, and a wanted eigenstate
    whose sector is not represented is never found (the iteration "converges" to the wrong set).

    real: ``diag`` has one entry per REAL amplitude (2 N of them, real representation); the block is
    built as real amplitudes and returned as the [B][p][N] complex view of them."""
    f64 = torch.float64
    na = diag.shape[1]                                                    # amplitudes: N, or 2 N (real)
    nl = min(8 * p, na)
    dlow, low = torch.topk(diag, nl, dim=1, largest=False)               # [1 or B][nl], ascending energy
    # damped with the LC energy above the start block, so that the Rayleigh quotients of the start
    # vectors are not swamped by high-energy basis states (charge-dominated circuits)
    spread = (dlow[:, p - 1:p] - dlow[:, :1]).clamp_min(1e-300)
    damp = spread / (spread + (dlow - dlow[:, :1]))                      # [1 or B][nl]
    V0 = torch.zeros((B, p, na), dtype=f64, device=dev)
    V0.scatter_(2, low[:, :p].expand(B, p).unsqueeze(2), torch.ones((B, p, 1), dtype=f64, device=dev))
    i = torch.arange(nl, device=dev, dtype=f64).unsqueeze(0)
    j = torch.arange(p, device=dev, dtype=f64).unsqueeze(1)
    xi = start_noise * torch.cos(2.399963229728653 * (i + 1.0) * (j + 1.0) + 0.5 * j)      # [p][nl]
    if k is not None and k < p:
        # only the guard vectors carry the perturbation: the k wanted start vectors stay clean (they
        # are already eigenvectors of charge-dominated circuits), the guards seed the other sectors
        xi = xi * (torch.arange(p, device=dev) >= k).to(xi.dtype).unsqueeze(1)
    pert = (xi.unsqueeze(0) * damp.unsqueeze(1)).expand(B, p, nl).contiguous()
    V0.scatter_add_(2, low.unsqueeze(1).expand(B, p, nl), pert)
    if real:
        return torch.view_as_complex(V0.reshape(B, p, na // 2, 2))
    return V0.to(torch.complex128)


def solve_davidson(kern, op, B, N, k, diag, p=None, mmax=None, tol=1e-9, maxit=300, x0=None,
                   impl=0, stats=None, lowblock_ns=0):
    """diag: [1 or B][N] real preconditioner diagonal (the LC part of H)."""
    dev = op.device
    if p is None:
        p = min(16, max(k + 2, (k * 6 + 4) // 5))
    p = min(p, 16, N)
    k = min(k, p)
    if mmax is None:
        mmax = 4 * p
    mmax = min(mmax, 64, N)
    ld = mmax
    c128, f64, i32 = torch.complex128, torch.float64, torch.int32

    V = torch.zeros((B, mmax, N), dtype=c128, device=dev)
    W = torch.zeros((B, mmax, N), dtype=c128, device=dev)
    Xb = torch.zeros((B, p, N), dtype=c128, device=dev)
    HXb = torch.zeros((B, p, N), dtype=c128, device=dev)
    G = torch.zeros((B, ld, ld), dtype=c128, device=dev)
    Zt = torch.zeros((B, ld, ld), dtype=c128, device=dev)
    theta = torch.zeros((B, ld), dtype=f64, device=dev)
    S1 = torch.zeros((B, mmax, 16), dtype=c128, device=dev)
    S2 = torch.zeros((B, 16, 16), dtype=c128, device=dev)
    Cq = torch.zeros((B, 16, 16), dtype=c128, device=dev)
    msz = torch.full((B,), p, dtype=i32, device=dev)      # per-point basis size
    mtot = torch.full((B,), p, dtype=i32, device=dev)
    nact = torch.zeros((B,), dtype=i32, device=dev)
    nin = torch.zeros((B,), dtype=i32, device=dev)     # block size before SVQB drops directions
    act = torch.zeros((B, p), dtype=i32, device=dev)
    restart = torch.zeros((B,), dtype=i32, device=dev)
    done = torch.zeros((B,), dtype=i32, device=dev)
    n_not_done = torch.zeros((1,), dtype=i32, device=dev)
    rn2 = torch.zeros((B, p), dtype=f64, device=dev)
    resmax = torch.zeros((B,), dtype=f64, device=dev)
    den_floor = torch.zeros((B,), dtype=f64, device=dev)

    Vp = BlockView(V, cnt_fixed=p)
    Wp = BlockView(W, cnt_fixed=p)
    warm = x0 is not None
    if x0 is None:
        x0 = cold_start_block(diag, B, p, N, dev, k)    # orthonormalised below like a warm start
    if True:
        # warm start: x0 holds one or two blocks of p vectors per point (Ritz vectors of solved
        # neighbouring sweep points).  Block 0 is orthonormalised in place (SVQB x2); block 1 is
        # orthogonalised against block 0 and itself exactly like a Davidson expansion block.
        nblk = x0.shape[1] // p
        if hasattr(x0, 'write_into'):
            x0.write_into(kern, V)
        else:
            V[:, :nblk * p, :] = x0[:, :nblk * p, :]
        msz.zero_()
        for blk in range(nblk):
            Tb = BlockView(V, cnt_fixed=p, off=msz, cnt=nact)
            Tbin = BlockView(V, cnt_fixed=p, off=msz, cnt=nin)
            Vcur = BlockView(V, cnt_fixed=mmax, cnt=msz)
            nact.fill_(p)
            for _ in range(2):
                if blk > 0:
                    kern.gram(Vcur, Tb, S1, impl)
                    kern.block_update(Vcur, S1, Tb, op=1, trans=1)
                kern.gram(Tb, Tb, S2, impl)
                nin.copy_(nact)
                kern.svqb_coeffs(S2, nact, 1e-10, Cq, msz, mtot)
                kern.block_rightmul(Tbin, Cq, nact)
            if blk > 0:
                # compact: dropped directions leave zero vectors at the tail of the block, which
                # the counters simply do not include
                pass
            msz.copy_(mtot)
        Vinit = BlockView(V, cnt_fixed=mmax, cnt=msz)
        Winit = BlockView(W, cnt_fixed=mmax, cnt=msz)
        op.apply(Vinit, Winit)
        # projected matrix of the initial basis, <= 16 columns at a time
        for c0 in range(0, nblk * p, p):
            ccnt = torch.clamp(msz - c0, min=0, max=p).to(i32)
            coff = torch.full((B,), c0, dtype=i32, device=dev)
            kern.gram(Vinit, BlockView(W, cnt_fixed=p, off=coff, cnt=ccnt), S1, impl)
            G[:, :, c0:c0 + p] = S1[:, :, :p]
        G.copy_(0.5 * (G + G.conj().transpose(1, 2)))

    st = kern.davidson_state(p=p, k=k, mmax=mmax, ldg=ld, tol_rel=tol, tol_abs_floor=1e-300,
                             msz=msz, nact=nact, act=act, restart=restart, done=done,
                             n_not_done=n_not_done, theta=theta, rn2=rn2, resmax=resmax,
                             den_floor=den_floor, G=G)
    Xv = BlockView(Xb, cnt_fixed=p)
    HXv = BlockView(HXb, cnt_fixed=p)
    Tv = BlockView(V, cnt_fixed=p, off=msz, cnt=nact)      # free tail of the basis
    Tin = BlockView(V, cnt_fixed=p, off=msz, cnt=nin)      # same window, pre-SVQB count
    HTv = BlockView(W, cnt_fixed=p, off=msz, cnt=nact)
    Vm = BlockView(V, cnt_fixed=mmax, cnt=msz)             # current basis
    Wm = BlockView(W, cnt_fixed=mmax, cnt=msz)
    Vt = BlockView(V, cnt_fixed=mmax, cnt=mtot)            # basis + new block

    it = 0
    converged = False
    lowblock = None
    for it in range(1, maxit + 1):
        live = 1 - done
        nheev = msz * live
        xcnt = (p * live).to(i32)
        kern.heev(G, nheev, theta, Zt)
        if it == 1 and lowblock_ns > 0 and getattr(op, 'supports_lowblock', lambda: False)():
            # two-level preconditioner: exact (H_SS - sigma)^-1 on the lowest-energy basis states,
            # sigma a little below the (already accurate, warm-started) lowest Ritz value
            sigma = theta[:, 0] - 0.05 * theta[:, 0].abs()
            lowblock = op.build_lowblock(sigma, lowblock_ns)
        Xv.cnt = xcnt
        HXv.cnt = xcnt
        kern.block_update(Vm, Zt, Xv, op=0, trans=0)
        kern.block_update(Wm, Zt, HXv, op=0, trans=0)
        kern.residual_norms(Xv, HXv, theta, rn2)
        kern.davidson_decide(st, B)
        left = int(n_not_done.item())
        if stats is not None:
            stats.append((it, left))
            if DETAIL_STATS is not None:      # dev instrumentation (tools/nact_stats.py): costs syncs
                lv = (1 - done)
                DETAIL_STATS.append((it, B, left, int((nact * lv).sum()), int((msz * lv).sum()),
                                     int((((nact + 3) // 4) * lv).sum()), int(restart.sum())))
        if left == 0:
            converged = True
            break
        Xv.cnt = None
        HXv.cnt = None
        kern.restart_copy(Xv, HXv, Vp, Wp, restart)
        kern.precond_residual(Xv, HXv, Tv, theta, act, diag, den_floor)
        if lowblock is not None:
            kern.lowblock_apply(Xv, HXv, Tv, theta, act, lowblock['S'], lowblock['Ainv'], lowblock['unit'],
                                lowblock.get('group', 1))
        for _ in range(2):
            kern.gram(Vm, Tv, S1, impl)
            kern.block_update(Vm, S1, Tv, op=1, trans=1)
            kern.gram(Tv, Tv, S2, impl)
            nin.copy_(nact)
            kern.svqb_coeffs(S2, nact, 1e-12, Cq, msz, mtot)
            kern.block_rightmul(Tin, Cq, nact)
        op.apply(Tv, HTv)
        kern.gram(Vt, HTv, S1, impl)
        kern.davidson_insert(st, S1, B)
    return EigResult(theta[:, :k].contiguous(), Xb[:, :k, :].contiguous(), it, resmax, converged,
                     ritz_block=Xb)


class RitzPairStart:
    """Warm start = union of the Ritz blocks of two solved sweep points per point, gathered from a
    pool ``store`` [S][p][N] straight into the basis buffer (one kernel, no [B][2p][N] staging)."""

    def __init__(self, store, ia, ib):
        self.store, self.ia, self.ib = store, ia, ib
        self.p = store.shape[1]
        self.shape = (ia.shape[0], 2 * self.p, store.shape[2])

    def write_into(self, kern, V):
        p, dev = self.p, V.device
        pool = self.store.reshape(1, -1, self.store.shape[2])
        oa = (self.ia.to(device=dev, dtype=torch.int64) * p).to(torch.int32)
        ob = (self.ib.to(device=dev, dtype=torch.int64) * p).to(torch.int32)
        ones = torch.ones((V.shape[0],), dtype=torch.int32, device=dev)
        kern.restart_copy(BlockView(pool, cnt_fixed=p, off=oa, stride_b=0),
                          BlockView(pool, cnt_fixed=p, off=ob, stride_b=0),
                          BlockView(V, cnt_fixed=p), BlockView(V, cnt_fixed=p, first=p), ones)


DETAIL_STATS = None
# entry points of op.apply: the only ones an instrumented backend may keep out of the CUDA graphs
_APPLY_OPS = frozenset({'hx_fused', 'hx_fused_real', 'hx_generic', 'mode_transform', 'potential_apply'})


def solve_davidson_gen(kern, op, B, N, k, diag, p=None, mmax=None, tol=1e-9, maxit=300, x0=None,
                       stats=None, lowblock_ns=0, drop=1e-8, cold_gate=0.25, expand=None, start_noise=0.03,
                       orth_start=None):
    """Block Davidson WITHOUT orthogonalising the expansion block.

    The basis V is allowed to be non-orthogonal inside a restart cycle; Rayleigh-Ritz is the
    generalised problem GA c = theta GB c with GB = V^H V (``sqc_gen_rr``).  Per iteration the
    basis is streamed three times instead of seven: X = V c, HX = W c (``sqc_ritz_update``) and ONE
    paired Gram pass that produces the new columns of GA and GB together (``sqc_gram_pair``); the
    CGS2 / SVQB passes of ``solve_davidson`` disappear.  Thick restarts (V <- Ritz vectors, which
    are GB-orthonormal) put the basis back to an orthonormal one.  Same convergence and accuracy
    as the orthogonalised iteration on every tested circuit (tools/proto_nonortho.py).

    Tuning knobs (all measured on the 0-pi bench, tools/): ``drop`` relative Cholesky pivot below which a
    basis vector counts as dependent; ``cold_gate`` residual of the lowest pair (relative to the spectral
    scale) below which a cold solve switches the two-level preconditioner on; ``expand`` Ritz pairs that get
    a correction vector per iteration (None: all p -- expanding only the k wanted pairs needs 81 instead of
    66 iterations for the same wall time); ``start_noise`` amplitude of the cold-start perturbation;
    ``orth_start`` orthonormalise a two-block warm start in vector space (None: for complex vectors only).

    Real representation (``op.real``): the vectors are real amplitudes viewed as complex blocks of N =
    ceil(N_real / 2) elements; the Gram entries keep only their real parts, everything else is unchanged."""
    dev = op.device
    real = bool(getattr(op, 'real', False))
    if p is None:
        p = min(16, max(k + 2, (k * 6 + 4) // 5))
    p = min(p, 16, N)
    k = min(k, p)
    if mmax is None:
        mmax = 3 * p + 4        # measured best on the 0-pi bench (36 / 40 / 48: 4031 / 4126 / 3997 diag/s)
    mmax = min(mmax, 64, N)
    ld = mmax
    c128, f64, i32 = torch.complex128, torch.float64, torch.int32
    # basis buffers are never read beyond the per-point counters: no need to zero-fill tens of GB.
    # Rows [0, mmax): the basis; rows [mmax, mmax + p): the Ritz block of an iteration that does not
    # restart.  An iteration that restarts writes its Ritz block straight into rows [0, p) (the
    # update kernel reads every basis row of a position before it writes that position), so there
    # is no restart copy; sqc_davidson_decide plans the restart one iteration ahead (xoff / xlast).
    V = kern.workspace('davidson_V', (B, mmax + p, N), c128)
    W = kern.workspace('davidson_W', (B, mmax + p, N), c128)
    GA = torch.zeros((B, ld, ld), dtype=c128, device=dev)
    GB = torch.zeros((B, ld, ld), dtype=c128, device=dev)
    Zt = torch.zeros((B, ld, ld), dtype=c128, device=dev)
    theta = torch.zeros((B, ld), dtype=f64, device=dev)
    SA = torch.zeros((B, mmax, 16), dtype=c128, device=dev)
    SB = torch.zeros((B, mmax, 16), dtype=c128, device=dev)
    msz = torch.full((B,), p, dtype=i32, device=dev)
    mtot = torch.full((B,), p, dtype=i32, device=dev)
    nact = torch.zeros((B,), dtype=i32, device=dev)
    act = torch.zeros((B, p), dtype=i32, device=dev)
    restart = torch.zeros((B,), dtype=i32, device=dev)
    xoff = torch.full((B,), mmax, dtype=i32, device=dev)
    xlast = torch.full((B,), mmax, dtype=i32, device=dev)
    done = torch.zeros((B,), dtype=i32, device=dev)
    n_not_done = torch.zeros((1,), dtype=i32, device=dev)
    rn2 = torch.zeros((B, p), dtype=f64, device=dev)
    resmax = torch.zeros((B,), dtype=f64, device=dev)
    den_floor = torch.zeros((B,), dtype=f64, device=dev)

    Vp = BlockView(V, cnt_fixed=p)
    Wp = BlockView(W, cnt_fixed=p)
    warm = x0 is not None
    if x0 is None:
        # not orthonormal: goes through the general start
        x0 = cold_start_block(diag, B, p, N, dev, k, start_noise=start_noise, real=real)
    if True:
        m0 = min(x0.shape[1], mmax)
        if hasattr(x0, 'write_into'):
            x0.write_into(kern, V)         # gather straight into the basis (no staging copy)
        else:
            V[:, :m0, :] = x0[:, :m0, :]
        msz.fill_(m0)
        if orth_start is None:
            orth_start = not real
        if orth_start and warm and m0 == 2 * p:
            # The union of two neighbours' Ritz blocks is nearly dependent (along a fine sweep the second block
            # differs from the first by ~1e-3 ... 1e-4): as raw basis vectors they give cond(GB) ~ 1e8, the first
            # Ritz vectors come out orthonormal / H-diagonal only to ~1e-6, and the restart (GB <- I,
            # GA <- diag(theta)) bakes that error in -- seen as a stall at 1e-6 on the gate-charge x flux grid
            # of the transmon - resonator - transmon circuit.  So block 1 is orthogonalised against block 0
            # (itself orthonormal: a converged Ritz block) and inside itself (SVQB), twice, in vector space;
            # directions below 1e-10 of the largest are dropped (per-point basis size p + kept).  Default: on for
            # complex vectors; the real representation (0-pi flux sweeps) has never shown the stall and would pay
            # ~15 ms per 300 ms step for the same 67 iterations (measured).
            Ab = BlockView(V, cnt_fixed=p)
            Bk = BlockView(V, cnt_fixed=p, first=p, cnt=nact)
            nin = torch.empty((B,), dtype=i32, device=dev)
            Bkin = BlockView(V, cnt_fixed=p, first=p, cnt=nin)
            S2 = torch.zeros((B, 16, 16), dtype=c128, device=dev)
            Cq = torch.zeros((B, 16, 16), dtype=c128, device=dev)
            nact.fill_(p)
            msz.fill_(p)
            for _ in range(2):
                kern.gram(Ab, Bk, SA)
                if real:
                    SA.imag.zero_()
                kern.block_update(Ab, SA, Bk, op=1, trans=1)
                kern.gram(Bk, Bk, S2)
                if real:
                    S2.imag.zero_()
                nin.copy_(nact)
                kern.svqb_coeffs(S2, nact, 1e-10, Cq, msz, mtot)
                kern.block_rightmul(Bkin, Cq, nact)
            msz.copy_(mtot)
        Vi = BlockView(V, cnt_fixed=mmax, cnt=msz)
        Wi = BlockView(W, cnt_fixed=mmax, cnt=msz)
        op.apply(Vi, Wi)
        for c0 in range(0, m0, p):
            ccnt = torch.clamp(msz - c0, min=0, max=p).to(i32)
            coff = torch.full((B,), c0, dtype=i32, device=dev)
            kern.gram_pair(Vi, BlockView(V, cnt_fixed=p, off=coff, cnt=ccnt),
                           BlockView(W, cnt_fixed=p, off=coff, cnt=ccnt), SB, SA, real=real)
            GB[:, :, c0:c0 + p] = SB[:, :, :p].real if real else SB[:, :, :p]
            GA[:, :, c0:c0 + p] = SA[:, :, :p].real if real else SA[:, :, :p]
    st = kern.davidson_state(p=p, k=k, mmax=mmax, ldg=ld, tol_rel=tol, tol_abs_floor=1e-300,
                             msz=msz, nact=nact, act=act, restart=restart, done=done,
                             n_not_done=n_not_done, theta=theta, rn2=rn2, resmax=resmax,
                             den_floor=den_floor, G=GA, GB=GB, mtot=mtot,
                             nexp=p if expand is None else expand,
                             xoff=xoff, xlast=xlast, xrow=mmax, real_basis=real)
    Xv = BlockView(V, cnt_fixed=p, off=xoff)         # where this iteration's Ritz block goes
    HXv = BlockView(W, cnt_fixed=p, off=xoff)
    Xl = BlockView(V, cnt_fixed=p, off=xlast)        # where the judged Ritz block is (after decide)
    HXl = BlockView(W, cnt_fixed=p, off=xlast)
    Tv = BlockView(V, cnt_fixed=p, off=msz, cnt=nact)
    HTv = BlockView(W, cnt_fixed=p, off=msz, cnt=nact)
    Vm = BlockView(V, cnt_fixed=mmax, cnt=msz)
    Wm = BlockView(W, cnt_fixed=mmax, cnt=msz)
    Vt = BlockView(V, cnt_fixed=mmax, cnt=mtot)
    it = 0
    converged = False
    lowblock = None
    best_left, best_it = B + 1, 0
    res_hist = []
    on_gpu = dev.type == 'cuda'
    live = torch.empty((B,), dtype=i32, device=dev)
    nrr = torch.empty((B,), dtype=i32, device=dev)
    xcnt = torch.empty((B,), dtype=i32, device=dev)
    Xv.cnt = xcnt
    HXv.cnt = xcnt
    if on_gpu:
        host_count = torch.zeros((1,), dtype=i32).pin_memory()
        count_ready = [torch.cuda.Event(), torch.cuda.Event()]
    want_lowblock = lowblock_ns > 0 and getattr(op, 'supports_lowblock', lambda: False)()
    detail = stats is not None and DETAIL_STATS is not None

    # One iteration = three pieces, all enqueued without host synchronisation and all no-ops for finished
    # points: rayleigh_ritz (generalised RR, Ritz block + residual norms, convergence / restart decisions,
    # count of unfinished points -> pinned host memory, preconditioned residuals), op.apply on the
    # correction block, and project (new Gram columns).  Once the two-level preconditioner exists the
    # pieces are captured into CUDA graphs and replayed: a lock-step iteration is ~15 launches plus a few
    # torch element-wise kernels, and on the small levels of a continuation the Python / launch overhead of
    # enqueueing them one by one exceeds their GPU time.
    def rayleigh_ritz():
        torch.bitwise_xor(done, 1, out=live)
        torch.mul(msz, live, out=nrr)
        torch.mul(live, p, out=xcnt)
        kern.gen_rr(GA, GB, nrr, drop, theta, Zt, real=real, n_out=p)
        kern.ritz_update(Vm, Wm, Zt, Xv, HXv, theta, rn2)       # X = V c, HX = W c, ||HX - theta X||^2
        kern.davidson_decide(st, B)
        if on_gpu:
            host_count.copy_(n_not_done, non_blocking=True)
        kern.precond_residual(Xl, HXl, Tv, theta, act, diag, den_floor, real=real)
        if lowblock is not None and lowblock.get('band') is not None:
            off, nblk, nbmax, perm = lowblock['band']
            kern.lowband_apply(Xl, HXl, Tv, theta, act, lowblock['S'], lowblock['Ainv'], off, nblk, nbmax,
                               lowblock['unit'], lowblock.get('group', 1), perm, real=real)
        elif lowblock is not None:
            kern.lowblock_apply(Xl, HXl, Tv, theta, act, lowblock['S'], lowblock['Ainv'], lowblock['unit'],
                                lowblock.get('group', 1), real=real)

    def project():
        kern.gram_pair(Vt, Tv, HTv, SB, SA, real=real)
        kern.davidson_insert2(st, SA, SB, B)

    def whole():
        rayleigh_ritz()
        op.apply(Tv, HTv)
        project()

    graphs = None          # None: not captured yet; False: run eagerly; else the captured pieces
    eager = frozenset(getattr(kern, 'eager_ops', frozenset)())
    if not (on_gpu and getattr(kern, 'use_graphs', False)) or detail or not eager <= _APPLY_OPS:
        graphs = False
    lag_left = None        # count of the iteration before the last one (graph mode reads one iteration late)
    for it in range(1, maxit + 1):
        if want_lowblock and lowblock is None:
            # The two-level preconditioner needs a shift below the spectrum of H_SS (>= lambda_0): warm
            # starts know lambda_0 to ~1e-6 at once; cold starts wait until every point's lowest Ritz value
            # has settled (its own residual below cold_gate x the spectral scale: then an eigenvalue lies
            # within ||r_0|| of theta_0).  Until then the iteration is run piece by piece: the decision
            # needs this iteration's Ritz values on the host.
            torch.bitwise_xor(done, 1, out=live)
            torch.mul(msz, live, out=nrr)
            kern.gen_rr(GA, GB, nrr, drop, theta, Zt, real=real, n_out=p)
            ready = warm
            if not warm and it > 1:
                r0 = rn2[:, 0].sqrt()
                scale = torch.maximum(theta[:, 0].abs(), theta[:, p - 1] - theta[:, 0])
                ready = bool((r0 < cold_gate * scale).all())
            if ready:
                # margin from the spectral scale, not from |theta_0| alone (a ground energy near zero
                # would leave no margin for the unpivoted complex64 / float32 factorisation)
                spread = torch.maximum(theta[:, 0].abs(), theta[:, min(p, N) - 1] - theta[:, 0])
                sigma = theta[:, 0] - 0.05 * spread
                if not warm:
                    sigma = sigma - rn2[:, 0].sqrt()
                lowblock = op.build_lowblock(sigma, lowblock_ns)
            torch.mul(live, p, out=xcnt)
            kern.ritz_update(Vm, Wm, Zt, Xv, HXv, theta, rn2)
            kern.davidson_decide(st, B)
            if on_gpu:
                host_count.copy_(n_not_done, non_blocking=True)
                count_ready[it & 1].record()
            kern.precond_residual(Xl, HXl, Tv, theta, act, diag, den_floor, real=real)
            if lowblock is not None and lowblock.get('band') is not None:
                off, nblk, nbmax, perm = lowblock['band']
                kern.lowband_apply(Xl, HXl, Tv, theta, act, lowblock['S'], lowblock['Ainv'], off, nblk, nbmax,
                                   lowblock['unit'], lowblock.get('group', 1), perm, real=real)
            elif lowblock is not None:
                kern.lowblock_apply(Xl, HXl, Tv, theta, act, lowblock['S'], lowblock['Ainv'], lowblock['unit'],
                                    lowblock.get('group', 1), real=real)
            op.apply(Tv, HTv)
            project()
            if on_gpu:
                count_ready[it & 1].synchronize()
                left = int(host_count[0])
            else:
                left = int(n_not_done.item())
        elif graphs is False or it == 1:
            # (the first iteration always runs eagerly: lazily grown work buffers must not be born inside a capture)
            rayleigh_ritz()
            if on_gpu:
                count_ready[it & 1].record()      # right after the count copy: the rest overlaps the wait
                op.apply(Tv, HTv)
                project()
                count_ready[it & 1].synchronize()
                left = int(host_count[0])
            else:
                op.apply(Tv, HTv)
                project()
                left = int(n_not_done.item())
            if detail:      # dev instrumentation (tools/nact_stats.py): costs syncs
                lv = (1 - done)
                DETAIL_STATS.append((it, B, left, int((nact * lv).sum()), int((msz * lv).sum()),
                                     int((((nact + 3) // 4) * lv).sum()), int(restart.sum())))
        else:
            if graphs is None:
                # every buffer and table of the iteration exists now (the previous iterations ran eagerly)
                if eager:                         # instrumented H.X: stays outside the graphs
                    graphs = (kern.capture(rayleigh_ritz), kern.capture(project))
                else:
                    graphs = (kern.capture(whole),)
            if len(graphs) == 1:
                kern.replay(graphs[0])
            else:
                kern.replay(graphs[0])
                op.apply(Tv, HTv)
                kern.replay(graphs[1])
            count_ready[it & 1].record()
            # The host reads the count of the PREVIOUS iteration while this one runs (the count never
            # increases, so a value that is one iteration newer than expected is just as good): the GPU
            # never waits for the host, at the price of one no-op iteration at the end.
            if lag_left is None:
                lag_left = True
                left = B                           # unknown yet
            else:
                count_ready[(it - 1) & 1].synchronize()
                left = int(host_count[0])
        if stats is not None:
            stats.append((it, left))
        if left == 0:
            converged = True
            break
        # stagnation guard: no sweep point has finished for 30 iterations (a healthy solve finishes
        # points every iteration once the first ones converge, and a cold start needs ~35)
        if left < best_left:
            best_left, best_it = left, it
        elif it % 10 == 0:
            # ... and the worst residual has not halved over those 30 iterations either (a single
            # hard point may legitimately need more than 60 iterations while converging steadily)
            res_hist.append(float((resmax * (1 - done)).max()))
            if it - best_it >= 30 and it >= 60 and len(res_hist) > 3 and res_hist[-1] > 0.5 * res_hist[-4]:
                break
    if not converged and graphs not in (None, False):
        # the last replay has not been judged on the host yet
        torch.cuda.current_stream(dev).synchronize()
        converged = int(host_count[0]) == 0
    # the Ritz block of every point sits at the row its last judged iteration wrote it to
    rows = xlast.long().unsqueeze(1) + torch.arange(p, device=dev).unsqueeze(0)
    ritz = V[torch.arange(B, device=dev).unsqueeze(1), rows]
    del V, W
    return EigResult(theta[:, :k].contiguous(), ritz[:, :k, :], it, resmax, converged, ritz_block=ritz)
