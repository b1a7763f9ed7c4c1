// Small dense per-sweep-point kernels: batched Hermitian Jacobi eigensolver (Rayleigh-Ritz,
// SVQB, x-basis of the harmonic modes, direct diagonalisation of small Hamiltonians) and the
// block-Davidson bookkeeping.  One CTA per sweep point.
#include <math.h>
#include "common.cuh"
#ifdef RR_PROFILE
#include <cstdio>
#endif

#define JAC_THREADS 512
#define JAC_MAX_SWEEPS 30
#define JAC_BIG 1.0e300

// u-th block (P <= Q) of the upper triangle of a half x half block grid, row by row
__device__ __forceinline__ void tri_index(int u, int half, int& P, int& Q) {
    P = (int)((2 * half + 1 - sqrtf((float)((2 * half + 1) * (2 * half + 1) - 8 * u))) * 0.5f);
    while (P > 0 && P * half - P * (P - 1) / 2 > u) --P;
    while ((P + 1) * half - (P + 1) * P / 2 <= u) ++P;
    Q = P + (u - (P * half - P * (P - 1) / 2));
}

// Parallel cyclic (round-robin) two-sided Jacobi for an n x n Hermitian matrix.
//   A  : np x np complex, leading dim lda (shared memory), np even, rows/cols >= n are padding
//   Zt : np rows of length ldz (shared or global); on exit row j = eigenvector of A[j][j]
//   rot: 4*np/2 doubles of shared scratch, prs: np ints of shared scratch, red: 64 doubles
// All threads of the CTA must call.  Eigenvalues are left on the diagonal of A (unsorted).
__device__ void jacobi_heev(cplx* A, int n, int np, int lda, cplx* Zt, int ldz, double* rot,
                            int* prs, double* red) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int half = np >> 1;
    // padding + identity
    for (int e = tid; e < np * np; e += nt) {
        const int r = e / np, c = e % np;
        if (r >= n || c >= n) A[r * lda + c] = (r == c) ? cmake(JAC_BIG, 0.0) : cmake(0.0, 0.0);
        Zt[r * ldz + c] = (r == c) ? cmake(1.0, 0.0) : cmake(0.0, 0.0);
    }
    __syncthreads();
    // thread teams of the rotation phase: the first na_thr threads rotate the 2x2 blocks of the upper
    // triangle of A (one block each for n <= 62), the others the eigenvector rows -- side by side
    // instead of one after the other, the phase is latency bound
    const int ntri = half * (half + 1) / 2;
    int na_thr = (ntri + 31) & ~31;
    if (na_thr > nt - 128) na_thr = nt / 2;
    int P0 = 0, Q0 = 0;
    if (tid < ntri) tri_index(tid, half, P0, Q0);
#ifdef RR_PROFILE
    long long tpa = 0, tpb = 0, tpc = 0, t0p;
#endif
    bool polish = false;
    for (int sweep = 0; sweep < JAC_MAX_SWEEPS; ++sweep) {
#ifdef RR_PROFILE
        t0p = clock64();
#endif
        // convergence test: off-diagonal vs diagonal Frobenius mass (real rows only)
        double off = 0.0, dg = 0.0;
        for (int e = tid; e < n * n; e += nt) {
            const int r = e / n, c = e % n;
            const cplx v = A[r * lda + c];
            const double m2 = v.x * v.x + v.y * v.y;
            if (r == c) dg += m2; else off += m2;
        }
        off = warp_sum(off);
        dg = warp_sum(dg);
        if ((tid & 31) == 0) {
            red[tid >> 5] = off;
            red[32 + (tid >> 5)] = dg;
        }
        __syncthreads();
        if (tid == 0) {
            double o = 0.0, d = 0.0;
            for (int w = 0; w < (nt >> 5); ++w) {
                o += red[w];
                d += red[32 + w];
            }
            red[0] = o;
            red[32] = d;
        }
        __syncthreads();
        const double o = red[0], d = red[32];
        __syncthreads();
        // quadratic convergence: once the off-diagonal mass is below 1e-11 relative, one more
        // sweep leaves it at rounding level
#ifdef RR_PROFILE
        if ((o == 0.0 || polish) && blockIdx.x == 0 && tid == 0) printf("  jacobi n=%d sweeps %d  cycles: params %lld apply %lld conv-test %lld\n", n, sweep, tpa, tpb, tpc);
#endif
        if (o == 0.0 || polish) break;
#ifdef RR_PROFILE
        tpc += clock64() - t0p;
#endif
        if (o <= 1e-22 * (d + o)) polish = true;
        // single-precision rotation angles only while the off-diagonal mass is large: their 1e-7
        // remnants would break the quadratic end game the polish rule relies on
        const bool fast_angle = o > 1e-8 * (d + o);

        for (int step = 0; step < np - 1; ++step) {
#ifdef RR_PROFILE
            t0p = clock64();
#endif
            // schedule + rotation parameters
            for (int k = tid; k < half; k += nt) {
                int p = (k == 0) ? np - 1 : (step + k) % (np - 1);
                int q = (step - k + (np - 1)) % (np - 1);
                if (p > q) { int tmp = p; p = q; q = tmp; }
                prs[2 * k] = p;
                prs[2 * k + 1] = q;
                const double a = A[p * lda + p].x, dd = A[q * lda + q].x;
                const cplx b = A[p * lda + q];
                // this scalar chain is the critical path of a Jacobi step (every other thread waits
                // for it at the barrier): reciprocal square roots instead of hypot / sqrt / divisions
                const double ab2 = b.x * b.x + b.y * b.y;
                double c = 1.0, s = 0.0, ex = 1.0, ey = 0.0;
                const double lim = 1e-17 * (fabs(a) + fabs(dd));
                if (ab2 > 0.0 && ab2 > lim * lim && p < n && q < n) {
                    // phase of the pivot: exact to double precision (unitarity), off the angle's path
                    const double iab = rsqrt(ab2);
                    ex = b.x * iab;
                    ey = b.y * iab;
                    // tangent of the rotation angle in SINGLE precision: its error only leaves a
                    // 1e-7-relative remnant of the pivot for the next sweep, while (c, s) below are
                    // normalised in double, so the transformation stays unitary to rounding
                    const float ab2f = (float)ab2;
                    double t;
                    if (fast_angle && ab2f > 1e-30f && ab2f < 1e30f) {
                        const float thf = 0.5f * (float)(dd - a) * rsqrtf(ab2f);
                        t = (double)(copysignf(1.0f, thf) / (fabsf(thf) + sqrtf(fmaf(thf, thf, 1.0f))));
                    } else {                    // outside the single-precision range: all in double
                        const double th = 0.5 * (dd - a) * iab;
                        t = copysign(1.0, th) / (fabs(th) + sqrt(th * th + 1.0));
                    }
                    c = rsqrt(fma(t, t, 1.0));
                    s = t * c;
                }
                rot[4 * k] = c;
                rot[4 * k + 1] = s;
                rot[4 * k + 2] = ex;
                rot[4 * k + 3] = ey;
            }
            __syncthreads();
#ifdef RR_PROFILE
            tpa += clock64() - t0p;
            t0p = clock64();
#endif
            // A <- J^H A J on 2x2 blocks (pair P of rows, pair Q of columns)
            // Hermitian: only the blocks P <= Q are rotated (block (Q, P) is written as the mirror).
            // The upper triangle is enumerated densely so that whole warps drop out: FP64 issue
            // slots, not threads, are what this kernel runs out of.
            for (int u = tid; u < ntri && tid < na_thr; u += na_thr) {
                int P = P0, Q = Q0;
                if (u != tid) tri_index(u, half, P, Q);
                const int p = prs[2 * P], q = prs[2 * P + 1], pc = prs[2 * Q], qc = prs[2 * Q + 1];
                const double cP = rot[4 * P], sP = rot[4 * P + 1];
                const cplx eP = cmake(rot[4 * P + 2], rot[4 * P + 3]);
                const double cQ = rot[4 * Q], sQ = rot[4 * Q + 1];
                if (sP == 0.0 && sQ == 0.0) continue;      // both rotations are the identity
                const cplx eQc = cmake(rot[4 * Q + 2], -rot[4 * Q + 3]);
                const cplx a00 = A[p * lda + pc], a01 = A[p * lda + qc];
                const cplx a10 = A[q * lda + pc], a11 = A[q * lda + qc];
                // rows:  r0 = c a0 - s e a1 ;  r1 = s a0 + c e a1
                const cplx ea10 = cmul(eP, a10), ea11 = cmul(eP, a11);
                const cplx r00 = csub(cscale(cP, a00), cscale(sP, ea10));
                const cplx r01 = csub(cscale(cP, a01), cscale(sP, ea11));
                const cplx r10 = cadd(cscale(sP, a00), cscale(cP, ea10));
                const cplx r11 = cadd(cscale(sP, a01), cscale(cP, ea11));
                // cols:  o.0 = c r.0 - s conj(e) r.1 ;  o.1 = s r.0 + c conj(e) r.1
                const cplx er01 = cmul(eQc, r01), er11 = cmul(eQc, r11);
                cplx o00 = csub(cscale(cQ, r00), cscale(sQ, er01));
                cplx o01 = cadd(cscale(sQ, r00), cscale(cQ, er01));
                cplx o10 = csub(cscale(cQ, r10), cscale(sQ, er11));
                cplx o11 = cadd(cscale(sQ, r10), cscale(cQ, er11));
                if (P == Q) {
                    o00.y = 0.0;
                    o11.y = 0.0;
                    if (sP != 0.0) {
                        // exact angle: the pivot is annihilated (set it to a clean zero); single-
                        // precision angle: keep the true remnant, or the similarity would be off by it
                        o01 = fast_angle ? o01 : cmake(0.0, 0.0);
                        o10 = cconj(o01);
                    }
                }
                A[p * lda + pc] = o00;
                A[p * lda + qc] = o01;
                A[q * lda + pc] = o10;
                A[q * lda + qc] = o11;
                if (P != Q) {
                    A[pc * lda + p] = cconj(o00);
                    A[qc * lda + p] = cconj(o01);
                    A[pc * lda + q] = cconj(o10);
                    A[qc * lda + q] = cconj(o11);
                }
            }
            // Zt rows:  z_p <- c z_p - s conj(e) z_q ;  z_q <- s z_p + c conj(e) z_q
            for (int e = tid - na_thr; e < half * n && e >= 0; e += nt - na_thr) {
                const int P = e / n, i = e % n;
                const double sP = rot[4 * P + 1];
                if (sP == 0.0) continue;
                const int p = prs[2 * P], q = prs[2 * P + 1];
                const double cP = rot[4 * P];
                const cplx ec = cmake(rot[4 * P + 2], -rot[4 * P + 3]);
                const cplx zp = Zt[p * ldz + i], zq = cmul(ec, Zt[q * ldz + i]);
                Zt[p * ldz + i] = csub(cscale(cP, zp), cscale(sP, zq));
                Zt[q * ldz + i] = cadd(cscale(sP, zp), cscale(cP, zq));
            }
            __syncthreads();
#ifdef RR_PROFILE
            tpb += clock64() - t0p;
#endif
        }
    }
}

// ---------------------------------------------------------------------------------------
// batched heev
// ---------------------------------------------------------------------------------------
template <bool ZSMEM>
__global__ void __launch_bounds__(JAC_THREADS)
heev_kernel(const cplx* __restrict__ Ain, const int32_t* __restrict__ nArr, int n_fixed, int ld,
            double* __restrict__ w, cplx* __restrict__ Zout, int np_max) {
    extern __shared__ double smem[];
    const int b = blockIdx.x;
    const int n = nArr ? nArr[b] : n_fixed;
    const int np = (n + 1) & ~1;
    const int lda = np_max | 1;   // odd (complex elements): a column walk touches every bank
    cplx* A = reinterpret_cast<cplx*>(smem);
    cplx* Zs = A + (size_t)np_max * lda;
    double* rot = reinterpret_cast<double*>(ZSMEM ? Zs + (size_t)np_max * lda : Zs);
    double* red = rot + 2 * np_max;
    int* prs = reinterpret_cast<int*>(red + 64);
    int* rank = prs + np_max;
    double* wv = reinterpret_cast<double*>(rank + np_max + (np_max & 1));
    cplx* zglob = Zout + (int64_t)b * ld * ld;
    const int tid = threadIdx.x;
    double* wout = w + (int64_t)b * ld;
    if (n <= 0) return;   // inactive sweep point: leave its previous results untouched
    const cplx* a = Ain + (int64_t)b * ld * ld;
    for (int e = tid; e < n * n; e += JAC_THREADS) {
        const int r = e / n, c = e % n;
        // Hermitian average of the two stored triangles
        const cplx u = a[r * ld + c], v = a[c * ld + r];
        A[r * lda + c] = cmake(0.5 * (u.x + v.x), 0.5 * (u.y - v.y));
    }
    __syncthreads();
    cplx* Z = ZSMEM ? Zs : zglob;
    const int ldz = ZSMEM ? lda : ld;
    jacobi_heev(A, n, np, lda, Z, ldz, rot, prs, red);
    // sort ascending
    for (int i = tid; i < n; i += JAC_THREADS) wv[i] = A[i * lda + i].x;
    __syncthreads();
    for (int i = tid; i < n; i += JAC_THREADS) {
        const double wi = wv[i];
        int r = 0;
        for (int j = 0; j < n; ++j) r += (wv[j] < wi) || (wv[j] == wi && j < i);
        rank[i] = r;
    }
    __syncthreads();
    for (int i = tid; i < ld; i += JAC_THREADS) wout[i] = INFINITY;
    __syncthreads();
    for (int i = tid; i < n; i += JAC_THREADS) wout[rank[i]] = wv[i];
    if (ZSMEM) {
        for (int e = tid; e < n * n; e += JAC_THREADS) {
            const int j = e / n, i = e % n;
            zglob[rank[j] * ld + i] = Zs[j * ldz + i];
        }
    } else {
        // in-place row permutation by cycle following (row buffer = A's first row area)
        cplx* tmp = A;  // n entries; A is no longer needed except its diagonal (saved in wv)
        for (int start = 0; start < n; ++start) {
            // process the cycle containing `start` only from its smallest member
            bool smallest = true;
            int j = rank[start];
            int guard = 0;
            while (j != start && guard++ < n) {
                if (j < start) { smallest = false; break; }
                j = rank[j];
            }
            if (!smallest || rank[start] == start) continue;
            // carry row `start` along the cycle: row r moves to rank[r]
            for (int i = tid; i < n; i += JAC_THREADS) tmp[i] = zglob[start * ld + i];
            __syncthreads();
            int src = start;
            do {
                const int dst = rank[src];
                for (int i = tid; i < n; i += JAC_THREADS) {
                    const cplx keep = zglob[dst * ld + i];
                    zglob[dst * ld + i] = tmp[i];
                    tmp[i] = keep;
                }
                __syncthreads();
                src = dst;
            } while (src != start);
        }
    }
}

extern "C" int sqc_heev(const void* A, const int32_t* n, int32_t n_fixed, int32_t ld, double* w,
                        void* Zt, int32_t B, void* stream) {
    if (B <= 0 || ld <= 0 || ld > 112 || n_fixed > ld) return SQC_EINVAL;
    const int np_max = (ld + 1) & ~1;
    const bool zs = np_max <= 64;
    size_t bytes = (size_t)np_max * (np_max | 1) * sizeof(cplx) * (zs ? 2 : 1);
    bytes += (2 * np_max + 64) * sizeof(double) + (2 * np_max + 2) * sizeof(int) + np_max * sizeof(double) + 16;
    cudaError_t e;
    if (zs) {
        e = cudaFuncSetAttribute(heev_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
        heev_kernel<true><<<B, JAC_THREADS, bytes, (cudaStream_t)stream>>>((const cplx*)A, n, n_fixed, ld, w, (cplx*)Zt, np_max);
    } else {
        e = cudaFuncSetAttribute(heev_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
        heev_kernel<false><<<B, JAC_THREADS, bytes, (cudaStream_t)stream>>>((const cplx*)A, n, n_fixed, ld, w, (cplx*)Zt, np_max);
    }
    SQC_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------
// Generalised Rayleigh-Ritz  GA c = theta GB c  for a NON-orthogonal Davidson basis
// (GB = basis Gram matrix).  Cholesky GB = L L^H with dead-column handling (pivot below
// drop * original diagonal: the vector is numerically inside the span of the previous ones and
// is removed from the problem), M = L^-1 GA L^-H, Jacobi on M, c = L^-H y.
// ---------------------------------------------------------------------------------------
#define RR_BIG 1.0e290

__global__ void __launch_bounds__(JAC_THREADS)
gen_rr_kernel(const cplx* __restrict__ GAin, const cplx* __restrict__ GBin,
              const int32_t* __restrict__ nArr, int ld, double drop, double* __restrict__ w,
              cplx* __restrict__ Zout, int np_max) {
    extern __shared__ double smem[];
    const int b = blockIdx.x;
    const int n = nArr[b];
    if (n <= 0) return;
    const int np = (n + 1) & ~1;
    const int lda = np_max | 1;   // odd (complex elements): a column walk touches every bank
    cplx* A = reinterpret_cast<cplx*>(smem);               // GA -> M -> eigenvalues on the diagonal
    cplx* Lm = A + (size_t)np_max * lda;                    // GB -> L -> tmp -> Jacobi vectors
    cplx* Zs = Lm + (size_t)np_max * lda;                   // L^-1
    double* rot = reinterpret_cast<double*>(Zs + (size_t)np_max * lda);
    double* red = rot + 2 * np_max;
    double* gbd = red + 64;                                 // original diagonal of GB
    double* wv = gbd + np_max;
    int* prs = reinterpret_cast<int*>(wv + np_max);
    int* rank = prs + np_max;
    int* dead = rank + np_max;
    const int tid = threadIdx.x;
#ifdef RR_PROFILE
    long long tp[8]; int tpi = 0;
#define RR_MARK() do { __syncthreads(); tp[tpi++] = clock64(); } while (0)
#else
#define RR_MARK() do {} while (0)
#endif
    RR_MARK();
    const cplx* ga = GAin + (int64_t)b * ld * ld;
    const cplx* gb = GBin + (int64_t)b * ld * ld;
    for (int e = tid; e < n * n; e += JAC_THREADS) {
        const int r = e / n, c = e % n;
        const cplx u = ga[r * ld + c], v = ga[c * ld + r];
        A[r * lda + c] = cmake(0.5 * (u.x + v.x), 0.5 * (u.y - v.y));
        const cplx u2 = gb[r * ld + c], v2 = gb[c * ld + r];
        Lm[r * lda + c] = cmake(0.5 * (u2.x + v2.x), 0.5 * (u2.y - v2.y));
    }
    __syncthreads();
    for (int i = tid; i < n; i += JAC_THREADS) {
        gbd[i] = Lm[i * lda + i].x;
        dead[i] = 0;
    }
    __syncthreads();
    // ---- Cholesky (lower), in place ----
    RR_MARK();
    for (int j = 0; j < n; ++j) {
        if (tid == 0) {
            const double d = Lm[j * lda + j].x;
            if (!(d > drop * gbd[j]) || !(gbd[j] > 0.0)) {
                dead[j] = 1;
                Lm[j * lda + j] = cmake(1.0, 0.0);
            } else {
                Lm[j * lda + j] = cmake(sqrt(d), 0.0);
            }
        }
        __syncthreads();
        const bool dj = dead[j] != 0;
        const double inv = 1.0 / Lm[j * lda + j].x;
        for (int i = j + 1 + tid; i < n; i += JAC_THREADS)
            Lm[i * lda + j] = dj ? cmake(0.0, 0.0) : cscale(inv, Lm[i * lda + j]);
        if (dj)
            for (int k = tid; k < j; k += JAC_THREADS) Lm[j * lda + k] = cmake(0.0, 0.0);
        __syncthreads();
        if (!dj) {
            const int rem = n - j - 1;
            for (int e = tid; e < rem * rem; e += JAC_THREADS) {
                const int i = j + 1 + e / rem, k = j + 1 + e % rem;
                if (k <= i) {
                    const cplx li = Lm[i * lda + j], lk = Lm[k * lda + j];
                    cplx v = Lm[i * lda + k];                 // -= li * conj(lk)
                    v.x -= li.x * lk.x + li.y * lk.y;
                    v.y -= li.y * lk.x - li.x * lk.y;
                    Lm[i * lda + k] = v;
                }
            }
        }
        __syncthreads();
    }
    // removed (dependent) vectors contribute nothing
    RR_MARK();
    for (int e = tid; e < n * n; e += JAC_THREADS) {
        const int r = e / n, c = e % n;
        if (dead[r] || dead[c]) A[r * lda + c] = cmake(0.0, 0.0);
    }
    // ---- Zs = L^-1 (lower triangular), one thread per column ----
    for (int c = tid; c < n; c += JAC_THREADS) {
        for (int i = 0; i < c; ++i) Zs[i * lda + c] = cmake(0.0, 0.0);
        for (int i = c; i < n; ++i) {
            cplx acc = (i == c) ? cmake(1.0, 0.0) : cmake(0.0, 0.0);
            for (int k = c; k < i; ++k) {
                const cplx l = Lm[i * lda + k], x = Zs[k * lda + c];
                acc.x -= l.x * x.x - l.y * x.y;
                acc.y -= l.x * x.y + l.y * x.x;
            }
            Zs[i * lda + c] = cscale(1.0 / Lm[i * lda + i].x, acc);
        }
    }
    __syncthreads();
    // ---- Lm = Zs * A ;  A = Lm * Zs^H ----
    RR_MARK();
    for (int e = tid; e < n * n; e += JAC_THREADS) {
        const int r = e / n, c = e % n;
        cplx acc = cmake(0.0, 0.0);
        for (int k = 0; k <= r; ++k) cfma(acc, Zs[r * lda + k], A[k * lda + c]);
        Lm[r * lda + c] = acc;
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += JAC_THREADS) {
        const int r = e / n, c = e % n;
        cplx acc = cmake(0.0, 0.0);
        for (int k = 0; k <= c; ++k) cfma(acc, Lm[r * lda + k], cconj(Zs[c * lda + k]));
        A[r * lda + c] = acc;
    }
    // alive index list (compaction of the removed vectors)
    RR_MARK();
    __shared__ int s_na;
    if (tid == 0) {
        int cnt = 0;
        for (int i = 0; i < n; ++i)
            if (!dead[i]) prs[cnt++] = i;       // prs is free until Jacobi starts
        s_na = cnt;
    }
    __syncthreads();
    const int na = s_na;
    const int npa = (na + 1) & ~1;
    // M restricted to the alive vectors -> Lm (Hermitian average); rank[] keeps the alive list
    for (int i = tid; i < na; i += JAC_THREADS) rank[i] = prs[i];
    __syncthreads();
    for (int e = tid; e < na * na; e += JAC_THREADS) {
        const int r = e / na, c = e % na;
        const cplx u = A[rank[r] * lda + rank[c]], v = A[rank[c] * lda + rank[r]];
        Lm[r * lda + c] = cmake(0.5 * (u.x + v.x), (r == c) ? 0.0 : 0.5 * (u.y - v.y));
    }
    __syncthreads();
    // dead[] is reused as the alive list because rank[] / prs[] are scratch of the Jacobi + sort
    for (int i = tid; i < na; i += JAC_THREADS) dead[i] = rank[i];
    __syncthreads();
    RR_MARK();
    jacobi_heev(Lm, na, npa, lda, A, lda, rot, prs, red);     // eigenvectors: rows of A
    RR_MARK();
    for (int i = tid; i < na; i += JAC_THREADS) wv[i] = Lm[i * lda + i].x;
    __syncthreads();
    for (int i = tid; i < na; i += JAC_THREADS) {
        const double wi = wv[i];
        int r = 0;
        for (int j = 0; j < na; ++j) r += (wv[j] < wi) || (wv[j] == wi && j < i);
        rank[i] = r;
    }
    __syncthreads();
    double* wout = w + (int64_t)b * ld;
    for (int i = tid; i < ld; i += JAC_THREADS) wout[i] = INFINITY;
    cplx* zg = Zout + (int64_t)b * ld * ld;
    for (int e = tid; e < n * ld; e += JAC_THREADS) zg[e] = cmake(0.0, 0.0);
    __syncthreads();
    for (int i = tid; i < na; i += JAC_THREADS) wout[rank[i]] = wv[i];
    // c_j = L^-H y_j :  Zout[rank j][i] = sum_k' conj(Zs[alive k'][i]) * y_j[k'],  y_j = row j of A
    for (int e = tid; e < na * n; e += JAC_THREADS) {
        const int j = e / n, i = e % n;
        cplx acc = cmake(0.0, 0.0);
        for (int kp = 0; kp < na; ++kp) {
            const int kk = dead[kp];
            if (kk >= i) cfma(acc, cconj(Zs[kk * lda + i]), A[j * lda + kp]);
        }
        zg[rank[j] * ld + i] = acc;
    }
#ifdef RR_PROFILE
    RR_MARK();
    if (b == 0 && tid == 0)
        printf("gen_rr n=%d na=%d cycles: load %lld chol %lld inv %lld mm %lld compact %lld jacobi %lld back %lld\n", n, na,
               tp[1] - tp[0], tp[2] - tp[1], tp[3] - tp[2], tp[4] - tp[3], tp[5] - tp[4], tp[6] - tp[5], tp[7] - tp[6]);
#endif
}

extern "C" int sqc_gen_rr(const void* GA, const void* GB, const int32_t* n, int32_t ld, double drop,
                          double* w, void* Zt, int32_t B, void* stream) {
    if (B <= 0 || ld <= 0 || ld > 68 || !n) return SQC_EINVAL;
    const int np_max = (ld + 1) & ~1;
    size_t bytes = 3 * (size_t)np_max * (np_max | 1) * sizeof(cplx) + (4 * np_max + 64) * sizeof(double) +
                   (3 * np_max + 4) * sizeof(int) + 16;
    if (bytes > 227 * 1024) return SQC_ELIMIT;
    cudaError_t e = cudaFuncSetAttribute(gen_rr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    gen_rr_kernel<<<B, JAC_THREADS, bytes, (cudaStream_t)stream>>>((const cplx*)GA, (const cplx*)GB, n, ld, drop, w,
                                                                   (cplx*)Zt, np_max);
    SQC_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------
// x-basis of a harmonic mode: eigen-decomposition of the truncated a + a^dag
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(JAC_THREADS)
xbasis_kernel(int m, double* __restrict__ V, double* __restrict__ lam, cplx* __restrict__ zwork) {
    extern __shared__ double smem[];
    const int np = (m + 1) & ~1;
    cplx* A = reinterpret_cast<cplx*>(smem);
    double* rot = reinterpret_cast<double*>(A + (size_t)np * np);
    double* red = rot + 2 * np;
    int* prs = reinterpret_cast<int*>(red + 64);
    int* rank = prs + np;
    double* wv = reinterpret_cast<double*>(rank + np + (np & 1));
    const int tid = threadIdx.x;
    for (int e = tid; e < m * m; e += JAC_THREADS) {
        const int r = e / m, c = e % m;
        double v = 0.0;
        if (c == r + 1) v = sqrt((double)(r + 1));
        if (r == c + 1) v = sqrt((double)(c + 1));
        A[r * np + c] = cmake(v, 0.0);
    }
    __syncthreads();
    jacobi_heev(A, m, np, np, zwork, np, rot, prs, red);
    for (int i = tid; i < m; i += JAC_THREADS) wv[i] = A[i * np + i].x;
    __syncthreads();
    for (int i = tid; i < m; i += JAC_THREADS) {
        const double wi = wv[i];
        int r = 0;
        for (int j = 0; j < m; ++j) r += (wv[j] < wi) || (wv[j] == wi && j < i);
        rank[i] = r;
        lam[r] = wi;
    }
    __syncthreads();
    // V[nf][i] = component nf of eigenvector i  (sign: first non-negligible component positive)
    for (int j = tid; j < m; j += JAC_THREADS) {
        double sg = 1.0;
        for (int i = 0; i < m; ++i) {
            const double v = zwork[j * np + i].x;
            if (fabs(v) > 1e-8) { sg = v < 0.0 ? -1.0 : 1.0; break; }
        }
        for (int i = 0; i < m; ++i) V[i * m + rank[j]] = sg * zwork[j * np + i].x;
    }
}

extern "C" int sqc_xbasis(int32_t m, double* V, double* lam, void* stream) {
    if (m <= 0 || m > 112) return SQC_ELIMIT;
    const int np = (m + 1) & ~1;
    size_t bytes = (size_t)np * np * sizeof(cplx) + (2 * np + 64) * sizeof(double) +
                   (2 * np + 2) * sizeof(int) + np * sizeof(double) + 16;
    cplx* zwork = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&zwork, (size_t)np * np * sizeof(cplx), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(xbasis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    xbasis_kernel<<<1, JAC_THREADS, bytes, (cudaStream_t)stream>>>(m, V, lam, zwork);
    ++g_sqc_launches;
    e = cudaGetLastError();
    cudaFreeAsync(zwork, (cudaStream_t)stream);
    return (int)e;
}

// ---------------------------------------------------------------------------------------
// SVQB coefficients for orthonormalising a block of <= 16 vectors from its Gram matrix
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
svqb_kernel(const cplx* __restrict__ S, int ld, int32_t* __restrict__ nact, double drop,
            cplx* __restrict__ Cq, const int32_t* __restrict__ msz, int32_t* __restrict__ mtot) {
    __shared__ cplx A[16 * 16];
    __shared__ cplx Z[16 * 16];
    __shared__ double rot[32], red[64], dsc[16], wv[16];
    __shared__ int prs[16], order[16];
    __shared__ int nkeep;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = min(nact[b], min(ld, 16));
    cplx* cq = Cq + (int64_t)b * ld * ld;
    for (int e = tid; e < ld * ld; e += 64) cq[e] = cmake(0.0, 0.0);
    if (n <= 0) {
        if (tid == 0 && mtot && msz) mtot[b] = msz[b];
        return;
    }
    const cplx* s = S + (int64_t)b * ld * ld;
    if (tid < n) {
        const double dd = s[tid * ld + tid].x;
        dsc[tid] = dd > 0.0 ? 1.0 / sqrt(dd) : 0.0;
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += 64) {
        const int r = e / n, c = e % n;
        const cplx u = s[r * ld + c], v = s[c * ld + r];
        const double sc = dsc[r] * dsc[c];
        A[r * 16 + c] = cmake(0.5 * (u.x + v.x) * sc, 0.5 * (u.y - v.y) * sc);
    }
    __syncthreads();
    const int np = (n + 1) & ~1;
    jacobi_heev(A, n, np, 16, Z, 16, rot, prs, red);
    if (tid < n) wv[tid] = A[tid * 16 + tid].x;
    __syncthreads();
    if (tid == 0) {
        double wmax = 0.0;
        for (int i = 0; i < n; ++i) wmax = fmax(wmax, wv[i]);
        // keep directions in DESCENDING sigma order
        int cnt = 0;
        bool used[16];
        for (int i = 0; i < n; ++i) used[i] = false;
        for (int r = 0; r < n; ++r) {
            int best = -1;
            for (int i = 0; i < n; ++i)
                if (!used[i] && (best < 0 || wv[i] > wv[best])) best = i;
            used[best] = true;
            if (wv[best] > drop * wmax && wv[best] > 0.0) order[cnt++] = best;
        }
        nkeep = cnt;
        nact[b] = cnt;
        if (mtot && msz) mtot[b] = msz[b] + cnt;
    }
    __syncthreads();
    // Cq[j][i] = D_i * U_i,src(j) / sqrt(sigma_src(j));  U column src = Z row src
    for (int e = tid; e < nkeep * n; e += 64) {
        const int j = e / n, i = e % n;
        const int src = order[j];
        const double f = dsc[i] / sqrt(wv[src]);
        const cplx u = Z[src * 16 + i];
        cq[j * ld + i] = cmake(u.x * f, u.y * f);
    }
}

extern "C" int sqc_svqb_coeffs(const void* S, int32_t ld, int32_t* nact, double drop, void* Cq,
                               const int32_t* msz, int32_t* mtot, int32_t B, void* stream) {
    if (B <= 0 || ld <= 0 || ld > 16) return SQC_EINVAL;
    svqb_kernel<<<B, 64, 0, (cudaStream_t)stream>>>((const cplx*)S, ld, nact, drop, (cplx*)Cq, msz, mtot);
    SQC_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------
// Davidson bookkeeping
// ---------------------------------------------------------------------------------------
__global__ void zero_counter_kernel(int32_t* c) { *c = 0; }

__global__ void __launch_bounds__(64)
davidson_decide_kernel(sqc_davidson_state_t st) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const int p = st.p, k = st.k, ldg = st.ldg;
    const int nexp = st.nexp > 0 ? st.nexp : p;
    __shared__ int s_restart;
    double* theta = st.theta + (int64_t)b * ldg;
    if (st.done[b]) {
        if (tid == 0) {
            st.nact[b] = 0;
            st.restart[b] = 0;
        }
        return;
    }
    if (tid == 0) {
        const int msz = st.msz[b];
        const int pe = min(p, msz);
        // A wanted Ritz value that is not finite (the generalised Rayleigh-Ritz dropped dependent basis vectors
        // and fewer than k pairs are alive) is NOT converged: without the test its relative residual would
        // be rn / inf = 0 (or NaN, which fmax and > both discard) and the point would be marked done.
        double scale = st.tol_abs_floor;
        bool finite = pe >= min(k, p);
        for (int j = 0; j < min(k, pe); ++j) {
            if (isfinite(theta[j])) scale = fmax(scale, fabs(theta[j]));
            else finite = false;
        }
        double rmax = finite ? 0.0 : 1e300;
        int n = 0;
        for (int j = 0; j < pe; ++j) {
            const double rn2 = st.rn2[(int64_t)b * p + j];
            if (!isfinite(theta[j]) || !isfinite(rn2)) continue;       // no correction vector for a dead pair
            const double rel = sqrt(rn2) / scale;
            if (j < k) rmax = fmax(rmax, rel);
            if (rel > 0.5 * st.tol_rel && j < nexp) st.act[(int64_t)b * p + n++] = j;
        }
        st.resmax[b] = rmax;
        int restart = 0;
        const bool planned = st.xoff != nullptr;      // copy-free restarts, see the header
        if (planned) st.xlast[b] = st.xoff[b];
        if (rmax <= st.tol_rel) {
            st.done[b] = 1;
            n = 0;
        } else {
            atomicAdd(st.n_not_done, 1);
            if (planned) {
                restart = st.restart[b];
                const int mcur = restart ? pe : msz;
                if (mcur + n > st.mmax) n = st.mmax - mcur;
            } else if (msz + n > st.mmax) {
                restart = 1;
            }
        }
        st.nact[b] = n;
        if (!planned) {
            st.restart[b] = restart;
        } else if (!st.done[b]) {
            // next iteration: the basis will hold (restart ? pe : msz) + n vectors; it restarts if
            // another n new ones would not fit
            const int mnext = (restart ? pe : msz) + n;
            const int rnext = mnext + n > st.mmax;
            st.restart[b] = rnext;
            st.xoff[b] = rnext ? 0 : st.xrow;
        }
        int hi = pe - 1;                                   // largest finite Ritz value of the block
        while (hi > 0 && !isfinite(theta[hi])) --hi;
        double fl = isfinite(theta[0]) ? fmax(fabs(theta[0]), fabs(theta[hi])) : 0.0;
        if (isfinite(theta[0])) fl = fmax(fl, theta[hi] - theta[0]);
        st.den_floor[b] = fmax(1e-3 * fl, 1e-300);
        s_restart = restart;
        if (restart) st.msz[b] = pe;
    }
    __syncthreads();
    if (s_restart) {
        cplx* G = reinterpret_cast<cplx*>(st.G) + (int64_t)b * ldg * ldg;
        cplx* GB = st.GB ? reinterpret_cast<cplx*>(st.GB) + (int64_t)b * ldg * ldg : nullptr;
        for (int e = tid; e < ldg * ldg; e += 64) {
            const int r = e / ldg, c = e % ldg;
            G[e] = (r == c && r < p) ? cmake(theta[r], 0.0) : cmake(0.0, 0.0);
            if (GB) GB[e] = (r == c && r < p) ? cmake(1.0, 0.0) : cmake(0.0, 0.0);
        }
    }
    if (tid == 0 && st.mtot) st.mtot[b] = st.msz[b] + st.nact[b];
}

extern "C" int sqc_davidson_decide(const sqc_davidson_state_t* st, int32_t B, void* stream) {
    if (!st || B <= 0 || st->p <= 0 || st->p > 16 || st->k > st->p) return SQC_EINVAL;
    zero_counter_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(st->n_not_done);
    SQC_LAUNCHED();
    davidson_decide_kernel<<<B, 64, 0, (cudaStream_t)stream>>>(*st);
    SQC_LAUNCHED();
    return 0;
}

__global__ void __launch_bounds__(128)
davidson_insert_kernel(sqc_davidson_state_t st, const cplx* __restrict__ S, int ldr, int ldc) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const int ldg = st.ldg;
    const int msz = st.msz[b], na = st.nact[b];
    if (na <= 0) return;
    cplx* G = reinterpret_cast<cplx*>(st.G) + (int64_t)b * ldg * ldg;
    const cplx* s = S + (int64_t)b * ldr * ldc;
    const int rows = msz + na;
    for (int e = tid; e < rows * na; e += 128) {
        const int i = e / na, j = e % na;
        cplx v = s[i * ldc + j];
        if (i >= msz) {  // new diagonal block: Hermitian average
            const cplx u = s[(msz + j) * ldc + (i - msz)];
            v = cmake(0.5 * (v.x + u.x), 0.5 * (v.y - u.y));
        }
        G[i * ldg + msz + j] = v;
        G[(msz + j) * ldg + i] = cconj(v);
    }
    __syncthreads();
    if (tid == 0) st.msz[b] = rows;
}

// Non-orthogonalised variant: new columns of BOTH the projected Hamiltonian (SA = [V T]^H H T)
// and the basis Gram matrix (SB = [V T]^H T).
__global__ void __launch_bounds__(128)
davidson_insert2_kernel(sqc_davidson_state_t st, const cplx* __restrict__ SA, const cplx* __restrict__ SB,
                        int ldr, int ldc) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const int ldg = st.ldg;
    const int msz = st.msz[b], na = st.nact[b];
    if (na <= 0) return;
    cplx* G = reinterpret_cast<cplx*>(st.G) + (int64_t)b * ldg * ldg;
    cplx* GB = reinterpret_cast<cplx*>(st.GB) + (int64_t)b * ldg * ldg;
    const cplx* sa = SA + (int64_t)b * ldr * ldc;
    const cplx* sb = SB + (int64_t)b * ldr * ldc;
    const int rows = msz + na;
    for (int e = tid; e < rows * na; e += 128) {
        const int i = e / na, j = e % na;
        cplx va = sa[i * ldc + j], vb = sb[i * ldc + j];
        if (i >= msz) {
            const cplx ua = sa[(msz + j) * ldc + (i - msz)], ub = sb[(msz + j) * ldc + (i - msz)];
            va = cmake(0.5 * (va.x + ua.x), 0.5 * (va.y - ua.y));
            vb = cmake(0.5 * (vb.x + ub.x), 0.5 * (vb.y - ub.y));
        }
        if (st.real_basis) va.y = vb.y = 0.0;      // real vectors viewed as complex: Im is not an inner product
        G[i * ldg + msz + j] = va;
        G[(msz + j) * ldg + i] = cconj(va);
        GB[i * ldg + msz + j] = vb;
        GB[(msz + j) * ldg + i] = cconj(vb);
    }
    __syncthreads();
    if (tid == 0) st.msz[b] = rows;
}

extern "C" int sqc_davidson_insert2(const sqc_davidson_state_t* st, const void* SA, const void* SB,
                                    int32_t ldr, int32_t ldc, int32_t B, void* stream) {
    if (!st || B <= 0 || !st->GB) return SQC_EINVAL;
    davidson_insert2_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(*st, (const cplx*)SA, (const cplx*)SB, ldr, ldc);
    SQC_LAUNCHED();
    return 0;
}

extern "C" int sqc_davidson_insert(const sqc_davidson_state_t* st, const void* S, int32_t ldr,
                                   int32_t ldc, int32_t B, void* stream) {
    if (!st || B <= 0) return SQC_EINVAL;
    davidson_insert_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(*st, (const cplx*)S, ldr, ldc);
    SQC_LAUNCHED();
    return 0;
}
