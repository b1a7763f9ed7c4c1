// Generalised Rayleigh-Ritz  GA c = theta GB c  for a REAL symmetric pencil (the Davidson basis of the real
// representation, see hx_real.cu): same contract as sqc_gen_rr (small.cu) -- Cholesky GB = L L^T with removal
// of numerically dependent basis vectors, M = L^-1 GA L^-T, cyclic Jacobi on M, c = L^-T y, eigenpairs in
// ascending order -- in real arithmetic.
//
// The complex kernel is latency bound (one 512-thread CTA per sweep point, 79 KB of shared memory: two CTAs
// per SM).  Here a problem takes 39 KB (mmax = 40) and a 384-thread CTA, so five problems are resident per SM
// and the rotations are real (a quarter of the flops).  The time of a call is (waves of resident problems) x
// (latency of one problem), and a level of the continuation only fills ~3 waves: the thread count is chosen
// for the latency of ONE problem (one round of 2x2 blocks per thread, three of eigenvector elements).
#include <math.h>
#include "common.cuh"
#ifdef RRR_PROFILE      // dev build (tools/rrr_profile.sh): phase cycle counts of the first pencil, printed once
#include <cstdio>
__device__ long long g_rrr_t[16];
#define RRR_MARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_rrr_t[i] = clock64(); } while (0)
#else
#define RRR_MARK(i) do {} while (0)
#endif

#define RRR_THREADS 384
#define RRR_WARPS (RRR_THREADS / 32)
#define RRR_BLK_THREADS 128       // threads that rotate the 2x2 blocks of A; the others rotate the rows of Z
#define RRR_MAX_SWEEPS 30
#define RRR_BIG 1.0e300

// eigen-stage of sqc_gen_rr_real: 0 = tridiagonal path when only part of the spectrum is consumed (default),
// 1 = tridiagonal path whenever possible, 2 = always the cyclic Jacobi (dev tool / tests)
__device__ int g_rrr_eig = 0;
extern "C" void sqc_debug_set_rrr_eig(int v) {
    cudaMemcpyToSymbol(g_rrr_eig, &v, sizeof(int));
}
#define RRR_REG_BLOCKS 4          // blocks per thread whose (P, Q) indices are kept in registers

// u-th block (P <= Q) of the upper triangle of a half x half block grid, row by row
__device__ __forceinline__ void rrr_tri_index(int u, int half, int& P, int& Q) {
    P = (int)((2 * half + 1 - sqrtf((float)((2 * half + 1) * (2 * half + 1) - 8 * u))) * 0.5f);
    while (P > 0 && P * half - P * (P - 1) / 2 > u) --P;
    while ((P + 1) * half - (P + 1) * P / 2 <= u) ++P;
    Q = P + (u - (P * half - P * (P - 1) / 2));
}

// Parallel cyclic (round-robin) two-sided Jacobi for an n x n real symmetric matrix in shared memory.
//   A : np x np, leading dim lda, np even, rows / columns >= n are padding
//   Zt: np rows of length ldz; on exit row j = eigenvector of A[j][j]
//   rot: np doubles, prs: np ints, red: 2 RRR_WARPS doubles of shared scratch.  All RRR_THREADS threads must call.
__device__ void jacobi_syev_real(double* A, int n, int np, int lda, double* Zt, int ldz, double* rot, int* prs,
                                 double* red) {
    const int tid = threadIdx.x;
    const int half = np >> 1;
    for (int e = tid; e < np * np; e += RRR_THREADS) {
        const int r = e / np, c = e % np;
        if (r >= n || c >= n) A[r * lda + c] = (r == c) ? RRR_BIG : 0.0;
        Zt[r * ldz + c] = (r == c) ? 1.0 : 0.0;
    }
    const int ntri = half * (half + 1) / 2;
    int pq[RRR_REG_BLOCKS];
#pragma unroll
    for (int r = 0; r < RRR_REG_BLOCKS; ++r) {
        const int u = tid + r * RRR_BLK_THREADS;
        int P = 0, Q = 0;
        if (tid < RRR_BLK_THREADS && u < ntri) rrr_tri_index(u, half, P, Q);
        pq[r] = P | (Q << 8);
    }
    __syncthreads();
    bool polish = false;
    for (int sweep = 0; sweep < RRR_MAX_SWEEPS; ++sweep) {
        // convergence test: off-diagonal against diagonal Frobenius mass
        double off = 0.0, dg = 0.0;
        for (int e = tid; e < n * n; e += RRR_THREADS) {
            const int r = e / n, c = e % n;
            const double v = A[r * lda + c];
            if (r == c) dg = fma(v, v, dg); else off = fma(v, v, off);
        }
        off = warp_sum(off);
        dg = warp_sum(dg);
        if ((tid & 31) == 0) {
            red[tid >> 5] = off;
            red[RRR_WARPS + (tid >> 5)] = dg;
        }
        __syncthreads();
        double o = 0.0, d = 0.0;
#pragma unroll
        for (int w_ = 0; w_ < RRR_WARPS; ++w_) {
            o += red[w_];
            d += red[RRR_WARPS + w_];
        }
        __syncthreads();
        // quadratic convergence: once the off-diagonal mass is below 1e-11 relative, one more sweep leaves
        // it at rounding level
        if (o == 0.0 || polish) break;
        if (o <= 1e-22 * (d + o)) polish = true;
        // single-precision rotation angles only while the off-diagonal mass is large (their 1e-7 remnants
        // would break the quadratic end game the polish rule relies on)
        const bool fast_angle = o > 1e-8 * (d + o);

        for (int step = 0; step < np - 1; ++step) {
            // round-robin schedule + rotation parameters of the half pivots of this step
            if (tid < half) {
                const int k = tid;
                int p = (k == 0) ? np - 1 : (step + k) % (np - 1);
                int q = (step - k + (np - 1)) % (np - 1);
                if (p > q) { const int t_ = p; p = q; q = t_; }
                prs[2 * k] = p;
                prs[2 * k + 1] = q;
                const double a = A[p * lda + p], dd = A[q * lda + q], b = A[p * lda + q];
                double c = 1.0, s = 0.0;
                const double lim = 1e-17 * (fabs(a) + fabs(dd));
                if (fabs(b) > lim && b != 0.0 && p < n && q < n) {
                    double t;
                    const float bf = (float)b;
                    if (fast_angle && fabsf(bf) > 1e-15f && fabsf(bf) < 1e15f) {
                        // tangent in single precision: its error only leaves a 1e-7-relative remnant of the
                        // pivot for the next sweep; (c, s) are normalised in double, the rotation stays
                        // orthogonal to rounding
                        const float thf = 0.5f * (float)(dd - a) / bf;
                        t = (double)(copysignf(1.0f, thf) / (fabsf(thf) + sqrtf(fmaf(thf, thf, 1.0f))));
                    } else {
                        const double th = 0.5 * (dd - a) / b;
                        t = copysign(1.0, th) / (fabs(th) + sqrt(fma(th, th, 1.0)));
                    }
                    c = rsqrt(fma(t, t, 1.0));
                    s = t * c;
                }
                rot[2 * k] = c;
                rot[2 * k + 1] = s;
            }
            __syncthreads();
            if (tid < RRR_BLK_THREADS) {
                // A <- J^T A J on the 2x2 blocks P <= Q of the upper triangle (block (Q, P) is the mirror)
                int r = 0;
                for (int u = tid; u < ntri; u += RRR_BLK_THREADS, ++r) {
                    int P, Q;
                    if (r < RRR_REG_BLOCKS) {
                        P = pq[r] & 255;
                        Q = pq[r] >> 8;
                    } else {
                        rrr_tri_index(u, half, P, Q);
                    }
                    const double cP = rot[2 * P], sP = rot[2 * P + 1], cQ = rot[2 * Q], sQ = rot[2 * Q + 1];
                    if (sP == 0.0 && sQ == 0.0) continue;
                    const int p = prs[2 * P], q = prs[2 * P + 1], pc = prs[2 * Q], qc = prs[2 * Q + 1];
                    const double a00 = A[p * lda + pc], a01 = A[p * lda + qc];
                    const double a10 = A[q * lda + pc], a11 = A[q * lda + qc];
                    // rows:  r0 = c a0 - s a1 ;  r1 = s a0 + c a1
                    const double r00 = cP * a00 - sP * a10, r01 = cP * a01 - sP * a11;
                    const double r10 = sP * a00 + cP * a10, r11 = sP * a01 + cP * a11;
                    // cols:  o.0 = c r.0 - s r.1 ;  o.1 = s r.0 + c r.1
                    double o00 = cQ * r00 - sQ * r01, o01 = sQ * r00 + cQ * r01;
                    double o10 = cQ * r10 - sQ * r11, o11 = sQ * r10 + cQ * r11;
                    if (P == Q && sP != 0.0) {
                        // exact angle: the pivot is annihilated (clean zero); single-precision angle: keep the
                        // true remnant, or the similarity would be off by it
                        o01 = fast_angle ? o01 : 0.0;
                        o10 = o01;
                    }
                    A[p * lda + pc] = o00;
                    A[p * lda + qc] = o01;
                    A[q * lda + pc] = o10;
                    A[q * lda + qc] = o11;
                    if (P != Q) {
                        A[pc * lda + p] = o00;
                        A[qc * lda + p] = o01;
                        A[pc * lda + q] = o10;
                        A[qc * lda + q] = o11;
                    }
                }
            } else {
                // Zt rows:  z_p <- c z_p - s z_q ;  z_q <- s z_p + c z_q
                for (int e = tid - RRR_BLK_THREADS; e < half * n; e += RRR_THREADS - RRR_BLK_THREADS) {
                    const int P = e / n, i = e - P * n;
                    const double sP = rot[2 * P + 1];
                    if (sP == 0.0) continue;
                    const int p = prs[2 * P], q = prs[2 * P + 1];
                    const double cP = rot[2 * P];
                    const double zp = Zt[p * ldz + i], zq = Zt[q * ldz + i];
                    Zt[p * ldz + i] = cP * zp - sP * zq;
                    Zt[q * ldz + i] = sP * zp + cP * zq;
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Lowest `nev` eigenpairs of a real symmetric n x n matrix in shared memory WITHOUT the full Jacobi
// diagonalisation: Householder tridiagonalisation, multisection (32 Sturm counts per round and eigenvalue, one warp
// per eigenvalue), inverse iteration with the LAPACK dstein safeguards for clusters, back-transformation of the nev
// vectors.  The Rayleigh-Ritz of the Davidson iteration consumes 12 of up to 40 pairs, and the cyclic Jacobi spends
// ~250 barrier-separated steps of pure latency on all of them (0.31 ms per pencil); the
// sequential depth here is ~35 reflector steps + ~11 multisection rounds + a tridiagonal solve.
//   Ts : n x n (leading dim lda), destroyed: on exit its strict lower part holds the Householder vectors
//   d, e, tau, vv, pp : scratch doubles [n];  lu : scratch [RRR_WARPS][5 n] doubles
//   wl : [nev] eigenvalues, ascending;  Y : [nev][ldy] eigenvectors as rows;  flag : shared int, 0 on entry
// Returns false (uniformly) when a vector fails its residual / orthogonality check: the caller then runs Jacobi.
// All RRR_THREADS threads must call; n >= 3.
__device__ bool tridiag_lowest(double* Ts, int n, int lda, int nev, double* d, double* e, double* tau, double* vv,
                               double* pp, double* lu, double* wl, double* Y, int ldy, int* flag) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double eps = 2.220446049250313e-16;
    // ---- 1. Householder tridiagonalisation (dsytd2, lower), the trailing block kept as a full symmetric matrix ----
    for (int k = 0; k < n - 1; ++k) {
        const int m = n - k - 1;                       // length of the column below the diagonal
        if (warp == 0) {
            double xn = 0.0;
            for (int i = k + 2 + lane; i < n; i += 32) {
                const double x = Ts[i * lda + k];
                xn = fma(x, x, xn);
            }
            xn = warp_sum(xn);
            const double alpha = Ts[(k + 1) * lda + k];
            double t = 0.0, beta = alpha, scale = 0.0;
            if (m > 1 && xn > 0.0) {
                beta = -copysign(sqrt(fma(alpha, alpha, xn)), alpha);
                t = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            for (int i = k + 1 + lane; i < n; i += 32) {
                const double v = (i == k + 1) ? 1.0 : Ts[i * lda + k] * scale;
                vv[i] = v;
                if (i > k + 1) Ts[i * lda + k] = v;    // kept for the back-transformation
            }
            if (lane == 0) {
                e[k] = beta;
                tau[k] = t;
            }
        }
        __syncthreads();
        const double t = tau[k];
        if (t != 0.0) {                                 // uniform
            // p = tau A22 v, one warp per row
            for (int i = k + 1 + warp; i < n; i += RRR_WARPS) {
                double acc = 0.0;
                for (int j = k + 1 + lane; j < n; j += 32) acc = fma(Ts[i * lda + j], vv[j], acc);
                acc = warp_sum(acc);
                if (lane == 0) pp[i] = t * acc;
            }
            __syncthreads();
            // w = p - (tau / 2) (p . v) v   (every warp forms the dot product itself: no extra barrier)
            double dot = 0.0;
            for (int j = k + 1 + lane; j < n; j += 32) dot = fma(pp[j], vv[j], dot);
            dot = warp_sum(dot);
            const double a2 = -0.5 * t * dot;
            // A22 -= v w^T + w v^T   (warp = rows, lanes = columns: no index arithmetic beyond adds)
            for (int i = k + 1 + warp; i < n; i += RRR_WARPS) {
                const double vi = vv[i], wi = fma(a2, vi, pp[i]);
                for (int j = k + 1 + lane; j < n; j += 32) {
                    const double vj = vv[j], wj = fma(a2, vj, pp[j]);
                    Ts[i * lda + j] -= vi * wj + wi * vj;
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += RRR_THREADS) d[i] = Ts[i * lda + i];
    __syncthreads();
    RRR_MARK(5);
    // ---- 2. multisection for the nev lowest eigenvalues (dstebz-style Sturm counts with pivmin) ----
    double glo = 1e300, ghi = -1e300, emax = 0.0;
    for (int i = lane; i < n; i += 32) {
        const double el = i > 0 ? fabs(e[i - 1]) : 0.0, er = i < n - 1 ? fabs(e[i]) : 0.0;
        glo = fmin(glo, d[i] - el - er);
        ghi = fmax(ghi, d[i] + el + er);
        emax = fmax(emax, er);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        glo = fmin(glo, __shfl_xor_sync(0xffffffffu, glo, o));
        ghi = fmax(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
        emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    }
    const double tnorm = fmax(fabs(glo), fabs(ghi));
    const double pivmin = fmax(2.2250738585072014e-308 * fmax(1.0, emax * emax), 1e-290);
    glo -= 2.0 * tnorm * eps * n + 2.0 * pivmin;
    ghi += 2.0 * tnorm * eps * n + 2.0 * pivmin;
    for (int i = tid; i < n - 1; i += RRR_THREADS) pp[i] = e[i] * e[i];
    __syncthreads();
    for (int j = warp; j < nev; j += RRR_WARPS) {
        double lo = glo, hi = ghi;
        for (int round = 0; round < 16; ++round) {
            const double width = hi - lo;
            if (width <= 2.0 * eps * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin) break;
            const double x = lo + width * (double)(lane + 1) * (1.0 / 33.0);
            int cnt = 0;
            double q = d[0] - x;
            if (fabs(q) < pivmin) q = -pivmin;
            cnt += q < 0.0;
            for (int i = 1; i < n; ++i) {
                // (a correctly rounded reciprocal is ample for a Sturm count and half the latency of a division)
                q = fma(-pp[i - 1], __drcp_rn(q), d[i] - x);
                if (fabs(q) < pivmin) q = -pivmin;
                cnt += q < 0.0;
            }
            // lambda_j < x  <=>  count(x) >= j + 1
            const unsigned bal = __ballot_sync(0xffffffffu, cnt >= j + 1);
            const int s = bal ? __ffs(bal) - 1 : 32;
            const double xs = __shfl_sync(0xffffffffu, x, s < 32 ? s : 31);
            const double xp = __shfl_sync(0xffffffffu, x, s > 0 ? s - 1 : 0);
            if (s < 32) hi = xs;
            if (s > 0) lo = xp;
        }
        if (lane == 0) wl[j] = 0.5 * (lo + hi);
    }
    __syncthreads();
    RRR_MARK(6);
    // ---- 3. inverse iteration (dstein): clusters = runs of eigenvalues closer than 1e-3 ||T||; the r-th member of
    // a cluster is computed in round r, orthogonalised against the members before it ----
    const double ortol = 1e-3 * tnorm, pertol = 10.0 * eps * tnorm;
    int maxr = 0;
    {
        int r = 0;
        for (int j = 1; j < nev; ++j) {
            r = (wl[j] - wl[j - 1] < ortol) ? r + 1 : 0;
            maxr = max(maxr, r);
        }
    }
    double* my = lu + (size_t)warp * 5 * n;            // a | b | c | dd | in  of dlagtf
    for (int rr = 0; rr <= maxr; ++rr) {
        for (int j = warp; j < nev; j += RRR_WARPS) {
            int cs = j;
            while (cs > 0 && wl[cs] - wl[cs - 1] < ortol) --cs;
            if (j - cs != rr) continue;                 // uniform per warp
            // separated shift inside a cluster (the members before j were shifted the same way)
            double lam = wl[cs];
            for (int q = cs + 1; q <= j; ++q) lam = fmax(wl[q], lam + pertol);
            double* a = my, *b = my + n, *c = my + 2 * n, *dd = my + 3 * n, *in = my + 4 * n;
            double* x = Y + (size_t)j * ldy;
            // LU with partial pivoting of T - lam I (dlagtf), sequential: lane 0
            if (lane == 0) {
                for (int i = 0; i < n; ++i) a[i] = d[i] - lam;
                for (int i = 0; i < n - 1; ++i) { b[i] = e[i]; c[i] = e[i]; }
                const double tl = fmax(eps * tnorm, pivmin);
                double scale1 = fabs(a[0]) + (n > 1 ? fabs(b[0]) : 0.0);
                for (int k = 0; k < n - 1; ++k) {
                    double scale2 = fabs(c[k]) + fabs(a[k + 1]);
                    if (k < n - 2) scale2 += fabs(b[k + 1]);
                    // pivot choice of dlagtf, |c| / scale2 <= |a| / scale1, without the divisions
                    if (c[k] == 0.0 || (a[k] != 0.0 && fabs(c[k]) * scale1 <= fabs(a[k]) * scale2)) {
                        in[k] = 0.0;
                        scale1 = scale2;
                        if (c[k] != 0.0) {
                            c[k] = c[k] / a[k];
                            a[k + 1] -= c[k] * b[k];
                        }
                        if (k < n - 2) dd[k] = 0.0;
                    } else {
                        in[k] = 1.0;
                        const double mult = a[k] / c[k];
                        a[k] = c[k];
                        const double temp = a[k + 1];
                        a[k + 1] = b[k] - mult * temp;
                        if (k < n - 2) {
                            dd[k] = b[k + 1];
                            b[k + 1] = -mult * dd[k];
                        }
                        b[k] = temp;
                        c[k] = mult;
                    }
                }
                // zero pivots of U are replaced by +-tl when solving (dlagts job = -1); reciprocals for the solves
                for (int i = 0; i < n; ++i) {
                    double ai = a[i];
                    if (fabs(ai) < tl) ai = copysign(tl, ai == 0.0 ? 1.0 : ai);
                    a[i] = 1.0 / ai;
                }
            }
            // deterministic start vector
            for (int i = lane; i < n; i += 32) {
                unsigned h = (unsigned)(i + 1) * 2654435761u ^ (unsigned)(j + 7) * 40503u ^ (unsigned)(rr + 3) * 69069u;
                h ^= h >> 15;
                x[i] = 1.0 + 0.5 * ((double)(h & 0xffffu) * (1.0 / 65536.0) - 0.5);
            }
            __syncwarp();
            for (int it = 0; it < 3; ++it) {
                // orthogonalise against the earlier members of the cluster (modified Gram-Schmidt)
                for (int q = cs; q < j; ++q) {
                    const double* yq = Y + (size_t)q * ldy;
                    double dot = 0.0;
                    for (int i = lane; i < n; i += 32) dot = fma(yq[i], x[i], dot);
                    dot = warp_sum(dot);
                    for (int i = lane; i < n; i += 32) x[i] = fma(-dot, yq[i], x[i]);
                    __syncwarp();
                }
                if (it == 2) break;                     // last pass: only the re-orthogonalisation + norm below
                // scale so that the solve cannot overflow, then x <- (P L U)^-1 x  (dlagts), lane 0
                double nrm = 0.0;
                for (int i = lane; i < n; i += 32) nrm = fmax(nrm, fabs(x[i]));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) nrm = fmax(nrm, __shfl_xor_sync(0xffffffffu, nrm, o));
                const double sc = nrm > 0.0 ? 1.0 / nrm : 1.0;
                for (int i = lane; i < n; i += 32) x[i] *= sc;
                __syncwarp();
                if (lane == 0) {
                    for (int k = 1; k < n; ++k) {
                        if (in[k - 1] == 0.0) {
                            x[k] -= c[k - 1] * x[k - 1];
                        } else {
                            const double temp = x[k - 1];
                            x[k - 1] = x[k];
                            x[k] = temp - c[k - 1] * x[k];
                        }
                    }
                    for (int k = n - 1; k >= 0; --k) {
                        double temp = x[k];
                        if (k <= n - 2) temp -= b[k] * x[k + 1];
                        if (k <= n - 3) temp -= dd[k] * x[k + 2];
                        x[k] = temp * a[k];
                    }
                }
                __syncwarp();
            }
            double n2 = 0.0;
            for (int i = lane; i < n; i += 32) n2 = fma(x[i], x[i], n2);
            n2 = warp_sum(n2);
            const double inv = n2 > 0.0 ? rsqrt(n2) : 0.0;
            for (int i = lane; i < n; i += 32) x[i] *= inv;
            __syncwarp();
            // residual check on the tridiagonal matrix: || T x - lambda x ||_inf against ||T||
            double res = 0.0;
            for (int i = lane; i < n; i += 32) {
                double r = (d[i] - wl[j]) * x[i];
                if (i > 0) r = fma(e[i - 1], x[i - 1], r);
                if (i < n - 1) r = fma(e[i], x[i + 1], r);
                res = fmax(res, fabs(r));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) res = fmax(res, __shfl_xor_sync(0xffffffffu, res, o));
            if (lane == 0 && !(res <= 1e-10 * tnorm && n2 > 0.0)) atomicExch(flag, 1);
        }
        __syncthreads();
    }
    RRR_MARK(7);
    if (*flag) return false;
    // ---- 4. back-transformation y <- H_0 H_1 ... H_{n-2} y, one warp per vector ----
    for (int j = warp; j < nev; j += RRR_WARPS) {
        double* x = Y + (size_t)j * ldy;
        // lane holds components lane and lane + 32 (n <= 64); the reflector vectors come from the columns of Ts
        double x0 = lane < n ? x[lane] : 0.0, x1 = lane + 32 < n ? x[lane + 32] : 0.0;
        for (int k = n - 2; k >= 0; --k) {
            const double t = tau[k];
            if (t == 0.0) continue;
            const int i0 = lane, i1 = lane + 32;
            const double v0 = i0 == k + 1 ? 1.0 : (i0 > k + 1 && i0 < n ? Ts[i0 * lda + k] : 0.0);
            const double v1 = i1 == k + 1 ? 1.0 : (i1 > k + 1 && i1 < n ? Ts[i1 * lda + k] : 0.0);
            const double dot = t * warp_sum(fma(v0, x0, v1 * x1));
            x0 = fma(-dot, v0, x0);
            x1 = fma(-dot, v1, x1);
        }
        if (lane < n) x[lane] = x0;
        if (lane + 32 < n) x[lane + 32] = x1;
    }
    __syncthreads();
    return true;
}

__global__ void __launch_bounds__(RRR_THREADS, 3)
gen_rr_real_kernel(const cplx* __restrict__ GAin, const cplx* __restrict__ GBin, const int32_t* __restrict__ nArr,
                   int ld, double drop, double* __restrict__ w, cplx* __restrict__ Zout, int np_max, int n_out) {
    extern __shared__ double smem[];
    const int b = blockIdx.x;
    const int n = nArr[b];
    if (n <= 0) return;
    const int lda = np_max | 1;
    double* A = smem;                                    // GA -> M -> Jacobi vectors
    double* Lm = A + (size_t)np_max * lda;                // GB -> L -> tmp -> compacted M -> eigenvalues
    double* Zs = Lm + (size_t)np_max * lda;               // L^-1
    double* rot = Zs + (size_t)np_max * lda;              // [np_max]
    double* red = rot + np_max;                           // [2 RRR_WARPS]
    double* gbd = red + 2 * RRR_WARPS;                                // original diagonal of GB
    double* wv = gbd + np_max;
    int* prs = reinterpret_cast<int*>(wv + np_max);
    int* rank = prs + np_max;
    int* dead = rank + np_max;
    // scratch of the tridiagonal path: d | e | tau | vv | pp [np_max each], wl [16], Y [16][lda], lu [warps][5 np_max]
    double* td = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(dead + np_max + 4) + 7) & ~(uintptr_t)7);
    double* te = td + np_max;
    double* ttau = te + np_max;
    double* tvv = ttau + np_max;
    double* tpp = tvv + np_max;
    double* twl = tpp + np_max;
    double* tY = twl + 16;
    double* tlu = tY + (size_t)16 * lda;
    __shared__ int s_na, s_flag;
    const int tid = threadIdx.x;
    const cplx* ga = GAin + (int64_t)b * ld * ld;
    const cplx* gb = GBin + (int64_t)b * ld * ld;
    for (int e = tid; e < n * n; e += RRR_THREADS) {
        const int r = e / n, c = e % n;
        A[r * lda + c] = 0.5 * (ga[r * ld + c].x + ga[c * ld + r].x);
        Lm[r * lda + c] = 0.5 * (gb[r * ld + c].x + gb[c * ld + r].x);
    }
    __syncthreads();
    for (int i = tid; i < n; i += RRR_THREADS) {
        gbd[i] = Lm[i * lda + i];
        dead[i] = 0;
    }
    __syncthreads();
    RRR_MARK(0);
    // ---- Cholesky (lower), in place; a pivot below drop * (original diagonal) removes the vector ----
    for (int j = 0; j < n; ++j) {
        if (tid == 0) {
            const double d = Lm[j * lda + j];
            if (!(d > drop * gbd[j]) || !(gbd[j] > 0.0)) {
                dead[j] = 1;
                Lm[j * lda + j] = 1.0;
            } else {
                Lm[j * lda + j] = sqrt(d);
            }
        }
        __syncthreads();
        const bool dj = dead[j] != 0;
        const double inv = 1.0 / Lm[j * lda + j];
        for (int i = j + 1 + tid; i < n; i += RRR_THREADS) Lm[i * lda + j] = dj ? 0.0 : inv * Lm[i * lda + j];
        if (dj)
            for (int k = tid; k < j; k += RRR_THREADS) Lm[j * lda + k] = 0.0;
        __syncthreads();
        if (!dj) {
            const int rem = n - j - 1;
            for (int e = tid; e < rem * rem; e += RRR_THREADS) {
                const int i = j + 1 + e / rem, k = j + 1 + e % rem;
                if (k <= i) Lm[i * lda + k] = fma(-Lm[i * lda + j], Lm[k * lda + j], Lm[i * lda + k]);
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < n * n; e += RRR_THREADS) {
        const int r = e / n, c = e % n;
        if (dead[r] || dead[c]) A[r * lda + c] = 0.0;
    }
    RRR_MARK(1);
    // ---- Zs = L^-1 (lower triangular), one thread per column ----
    for (int c = tid; c < n; c += RRR_THREADS) {
        for (int i = 0; i < c; ++i) Zs[i * lda + c] = 0.0;
        for (int i = c; i < n; ++i) {
            double acc = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) acc = fma(-Lm[i * lda + k], Zs[k * lda + c], acc);
            Zs[i * lda + c] = acc / Lm[i * lda + i];
        }
    }
    __syncthreads();
    RRR_MARK(2);
    // ---- Lm = Zs * A ;  A = Lm * Zs^T ----
    for (int e = tid; e < n * n; e += RRR_THREADS) {
        const int r = e / n, c = e % n;
        double acc = 0.0;
        for (int k = 0; k <= r; ++k) acc = fma(Zs[r * lda + k], A[k * lda + c], acc);
        Lm[r * lda + c] = acc;
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += RRR_THREADS) {
        const int r = e / n, c = e % n;
        double acc = 0.0;
        for (int k = 0; k <= c; ++k) acc = fma(Lm[r * lda + k], Zs[c * lda + k], acc);
        A[r * lda + c] = acc;
    }
    RRR_MARK(3);
    // alive index list (compaction of the removed vectors)
    if (tid == 0) {
        int cnt = 0;
        for (int i = 0; i < n; ++i)
            if (!dead[i]) rank[cnt++] = i;
        s_na = cnt;
    }
    __syncthreads();
    const int na = s_na;
    const int npa = (na + 1) & ~1;
    for (int e = tid; e < na * na; e += RRR_THREADS) {
        const int r = e / na, c = e % na;
        Lm[r * lda + c] = 0.5 * (A[rank[r] * lda + rank[c]] + A[rank[c] * lda + rank[r]]);
    }
    __syncthreads();
    // dead[] is reused as the alive list: rank[] / prs[] are scratch of the Jacobi and the sort
    for (int i = tid; i < na; i += RRR_THREADS) dead[i] = rank[i];
    __syncthreads();
    // Only the n_out lowest pairs are consumed.  When that is a proper subset, take them from the tridiagonal form
    // (multisection + inverse iteration); the full cyclic Jacobi is the fall-back (a vector failing its residual
    // check) and the path of small or fully requested pencils.
    const int nev = n_out < na ? n_out : na;
    RRR_MARK(4);
    bool partial = g_rrr_eig != 2 && nev <= 16 && na >= 4 && (nev < na || g_rrr_eig == 1);
    if (partial) {
        if (tid == 0) s_flag = 0;
        __syncthreads();
        partial = tridiag_lowest(Lm, na, lda, nev, td, te, ttau, tvv, tpp, tlu, twl, tY, lda, &s_flag);
        if (!partial) {
            // restore the compacted matrix for the Jacobi fall-back (A still holds M)
            for (int e = tid; e < na * na; e += RRR_THREADS) {
                const int r = e / na, c = e % na;
                Lm[r * lda + c] = 0.5 * (A[dead[r] * lda + dead[c]] + A[dead[c] * lda + dead[r]]);
            }
            __syncthreads();
        }
    }
    const double* Yv = A;                 // eigenvector rows
    int nvec = na;                        // number of computed pairs
    if (partial) {
        for (int i = tid; i < nev; i += RRR_THREADS) {
            wv[i] = twl[i];
            rank[i] = i;
        }
        Yv = tY;
        nvec = nev;
    } else {
        jacobi_syev_real(Lm, na, npa, lda, A, lda, rot, prs, red);        // eigenvectors: rows of A
        for (int i = tid; i < na; i += RRR_THREADS) wv[i] = Lm[i * lda + i];
        __syncthreads();
        for (int i = tid; i < na; i += RRR_THREADS) {
            const double wi = wv[i];
            int r = 0;
            for (int j = 0; j < na; ++j) r += (wv[j] < wi) || (wv[j] == wi && j < i);
            rank[i] = r;
        }
    }
    RRR_MARK(8);
    double* wout = w + (int64_t)b * ld;
    for (int i = tid; i < ld; i += RRR_THREADS) wout[i] = INFINITY;
    cplx* zg = Zout + (int64_t)b * ld * ld;
    const int nrows = n_out < n ? n_out : n;
    for (int e = tid; e < nrows * ld; e += RRR_THREADS) zg[e] = cmake(0.0, 0.0);
    __syncthreads();
    for (int i = tid; i < nvec; i += RRR_THREADS) wout[rank[i]] = wv[i];
    // c_j = L^-T y_j :  Zout[rank j][i] = sum_k' Zs[alive k'][i] * y_j[k'],  y_j = row j of the eigenvector array;
    // only the n_out lowest pairs are consumed (Ritz block / thick restart)
    for (int e = tid; e < nvec * n; e += RRR_THREADS) {
        const int j = e / n, i = e % n;
        const int rj = rank[j];
        if (rj >= nrows) continue;
        double acc = 0.0;
        for (int kp = 0; kp < na; ++kp) {
            const int kk = dead[kp];
            if (kk >= i) acc = fma(Zs[kk * lda + i], Yv[j * lda + kp], acc);
        }
        zg[rj * ld + i] = cmake(acc, 0.0);
    }
#ifdef RRR_PROFILE
    __syncthreads();
    RRR_MARK(9);
    if (blockIdx.x == 0 && tid == 0)
        printf("[rrr n=%d] cholesky %lld  Linv %lld  products %lld  compact %lld  tridiag %lld  multisection %lld  inverse-it %lld  "
               "back+sort %lld  output %lld  total %lld\n", n, g_rrr_t[1] - g_rrr_t[0], g_rrr_t[2] - g_rrr_t[1],
               g_rrr_t[3] - g_rrr_t[2], g_rrr_t[4] - g_rrr_t[3], g_rrr_t[5] - g_rrr_t[4], g_rrr_t[6] - g_rrr_t[5],
               g_rrr_t[7] - g_rrr_t[6], g_rrr_t[8] - g_rrr_t[7], g_rrr_t[9] - g_rrr_t[8], g_rrr_t[9] - g_rrr_t[0]);
#endif
}

extern "C" int sqc_gen_rr_real(const void* GA, const void* GB, const int32_t* n, int32_t ld, double drop,
                               double* w, void* Zt, int32_t n_out, int32_t B, void* stream) {
    if (B <= 0 || ld <= 0 || ld > 68 || !n || n_out <= 0) return SQC_EINVAL;
    const int np_max = (ld + 1) & ~1;
    const size_t bytes = 3 * (size_t)np_max * (np_max | 1) * sizeof(double) + (3 * np_max + 2 * RRR_WARPS) * sizeof(double) +
                         (3 * np_max + 4) * sizeof(int) + 16 +
                         // tridiagonal path: five vectors, 16 eigenvalues, 16 eigenvector rows, LU scratch per warp
                         (5 * (size_t)np_max + 16 + 16 * (size_t)(np_max | 1) + (size_t)RRR_WARPS * 5 * np_max) * sizeof(double) + 16;
    cudaError_t e = cudaFuncSetAttribute(gen_rr_real_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    gen_rr_real_kernel<<<B, RRR_THREADS, bytes, (cudaStream_t)stream>>>((const cplx*)GA, (const cplx*)GB, n, ld, drop,
                                                                         w, (cplx*)Zt, np_max, n_out);
    SQC_LAUNCHED();
    return 0;
}
