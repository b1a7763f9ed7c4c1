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

#define RRR_THREADS 384
#define RRR_WARPS (RRR_THREADS / 32)
#define RRR_BLK_THREADS 128       // threads that rotate the 2x2 blocks of A; the others rotate the rows of Z
#define RRR_MAX_SWEEPS 30
#define RRR_BIG 1.0e300
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

__global__ void __launch_bounds__(RRR_THREADS)
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
    __shared__ int s_na;
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
    jacobi_syev_real(Lm, na, npa, lda, A, lda, rot, prs, red);        // eigenvectors: rows of A
    for (int i = tid; i < na; i += RRR_THREADS) wv[i] = Lm[i * lda + i];
    __syncthreads();
    for (int i = tid; i < na; i += RRR_THREADS) {
        const double wi = wv[i];
        int r = 0;
        for (int j = 0; j < na; ++j) r += (wv[j] < wi) || (wv[j] == wi && j < i);
        rank[i] = r;
    }
    double* wout = w + (int64_t)b * ld;
    for (int i = tid; i < ld; i += RRR_THREADS) wout[i] = INFINITY;
    cplx* zg = Zout + (int64_t)b * ld * ld;
    const int nrows = n_out < n ? n_out : n;
    for (int e = tid; e < nrows * ld; e += RRR_THREADS) zg[e] = cmake(0.0, 0.0);
    __syncthreads();
    for (int i = tid; i < na; i += RRR_THREADS) wout[rank[i]] = wv[i];
    // c_j = L^-T y_j :  Zout[rank j][i] = sum_k' Zs[alive k'][i] * y_j[k'],  y_j = row j of A; only the n_out
    // lowest pairs are consumed (Ritz block / thick restart)
    for (int e = tid; e < na * n; e += RRR_THREADS) {
        const int j = e / n, i = e % n;
        const int rj = rank[j];
        if (rj >= nrows) continue;
        double acc = 0.0;
        for (int kp = 0; kp < na; ++kp) {
            const int kk = dead[kp];
            if (kk >= i) acc = fma(Zs[kk * lda + i], A[j * lda + kp], acc);
        }
        zg[rj * ld + i] = cmake(acc, 0.0);
    }
}

extern "C" int sqc_gen_rr_real(const void* GA, const void* GB, const int32_t* n, int32_t ld, double drop,
                               double* w, void* Zt, int32_t n_out, int32_t B, void* stream) {
    if (B <= 0 || ld <= 0 || ld > 68 || !n || n_out <= 0) return SQC_EINVAL;
    const int np_max = (ld + 1) & ~1;
    const size_t bytes = 3 * (size_t)np_max * (np_max | 1) * sizeof(double) + (3 * np_max + 2 * RRR_WARPS) * sizeof(double) +
                         (3 * np_max + 4) * sizeof(int) + 16;
    cudaError_t e = cudaFuncSetAttribute(gen_rr_real_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    gen_rr_real_kernel<<<B, RRR_THREADS, bytes, (cudaStream_t)stream>>>((const cplx*)GA, (const cplx*)GB, n, ld, drop,
                                                                         w, (cplx*)Zt, np_max, n_out);
    SQC_LAUNCHED();
    return 0;
}

// =====================================================================================
// Batched real symmetric eigensolver for the dense path (sqc_syev): the Hamiltonian of a circuit WITHOUT
// charge modes is real symmetric in the Fock basis (a_t D_t + h.c. = 2 Re(a_t D_t) for the complex-symmetric
// displacement tables D_t = V diag(e^{i s lam}) V^T), so the 1024-point fluxonium sweep of BASELINE configs[1]
// needs a quarter of the flops of sqc_heev and holds BOTH the matrix and the eigenvectors in shared memory
// (2 x 81 KB at n = 100; the complex kernel keeps the eigenvectors in global memory).
// Same parallel cyclic Jacobi as above, with the thread split of the large complex kernel (small.cu): the
// 2 x 2 blocks of the upper triangle and the eigenvector rows are rotated side by side by two teams.
// =====================================================================================
#define SYV_THREADS 1024

__device__ void jacobi_syev_big(double* A, int n, int np, int lda, double* Zt, int ldz, double* rot, int* prs,
                                double* red) {
    const int tid = threadIdx.x, nt = SYV_THREADS;
    const int half = np >> 1;
    for (int e = tid; e < np * np; e += nt) {
        const int r = e / np, c = e % np;
        if (r >= n || c >= n) A[r * lda + c] = (r == c) ? RRR_BIG : 0.0;
        Zt[r * ldz + c] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    const int ntri = half * (half + 1) / 2;
    // a 2 x 2 block costs about four eigenvector elements: split the threads accordingly
    int na_thr = (int)((long long)nt * 4 * ntri / (4LL * ntri + (long long)half * n));
    na_thr = (na_thr + 31) & ~31;
    if (na_thr < 32) na_thr = 32;
    if (na_thr > nt - 32) na_thr = nt - 32;
    int P0 = 0, Q0 = 0;
    if (tid < ntri && tid < na_thr) rrr_tri_index(tid, half, P0, Q0);
    bool polish = false;
    for (int sweep = 0; sweep < RRR_MAX_SWEEPS; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int e = tid; e < n * n; e += nt) {
            const int r = e / n, c = e % n;
            const double v = A[r * lda + c];
            if (r == c) dg = fma(v, v, dg); else off = fma(v, v, off);
        }
        off = warp_sum(off);
        dg = warp_sum(dg);
        if ((tid & 31) == 0) {
            red[tid >> 5] = off;
            red[32 + (tid >> 5)] = dg;
        }
        __syncthreads();
        double o = 0.0, d = 0.0;
        for (int w_ = 0; w_ < nt / 32; ++w_) {
            o += red[w_];
            d += red[32 + w_];
        }
        __syncthreads();
        if (o == 0.0 || polish) break;
        if (o <= 1e-22 * (d + o)) polish = true;
        const bool fast_angle = o > 1e-8 * (d + o);

        for (int step = 0; step < np - 1; ++step) {
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
            if (tid < na_thr) {
                for (int u = tid; u < ntri; u += na_thr) {
                    int P = P0, Q = Q0;
                    if (u != tid) rrr_tri_index(u, half, P, Q);
                    const double cP = rot[2 * P], sP = rot[2 * P + 1], cQ = rot[2 * Q], sQ = rot[2 * Q + 1];
                    if (sP == 0.0 && sQ == 0.0) continue;
                    const int p = prs[2 * P], q = prs[2 * P + 1], pc = prs[2 * Q], qc = prs[2 * Q + 1];
                    const double a00 = A[p * lda + pc], a01 = A[p * lda + qc];
                    const double a10 = A[q * lda + pc], a11 = A[q * lda + qc];
                    const double r00 = cP * a00 - sP * a10, r01 = cP * a01 - sP * a11;
                    const double r10 = sP * a00 + cP * a10, r11 = sP * a01 + cP * a11;
                    double o00 = cQ * r00 - sQ * r01, o01 = sQ * r00 + cQ * r01;
                    double o10 = cQ * r10 - sQ * r11, o11 = sQ * r10 + cQ * r11;
                    if (P == Q && sP != 0.0) {
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
                for (int e = tid - na_thr; e < half * n; e += nt - na_thr) {
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

__global__ void __launch_bounds__(SYV_THREADS, 1)
syev_kernel(const double* __restrict__ Ain, int n, int ld, double* __restrict__ w, double* __restrict__ Zout) {
    extern __shared__ double smem[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int np = (n + 1) & ~1;
    const int lda = np | 1;
    double* A = smem;
    double* Zs = A + (size_t)np * lda;
    double* rot = Zs + (size_t)np * lda;       // [np]
    double* red = rot + np;                    // [64]
    double* wv = red + 64;                     // [np]
    int* prs = reinterpret_cast<int*>(wv + np);
    int* rank = prs + np;
    const double* a = Ain + (int64_t)b * ld * ld;
    for (int e = tid; e < n * n; e += SYV_THREADS) {
        const int r = e / n, c = e % n;
        A[r * lda + c] = 0.5 * (a[r * ld + c] + a[c * ld + r]);
    }
    __syncthreads();
    jacobi_syev_big(A, n, np, lda, Zs, lda, rot, prs, red);
    for (int i = tid; i < n; i += SYV_THREADS) wv[i] = A[i * lda + i];
    __syncthreads();
    for (int i = tid; i < n; i += SYV_THREADS) {
        const double wi = wv[i];
        int r = 0;
        for (int j = 0; j < n; ++j) r += (wv[j] < wi) || (wv[j] == wi && j < i);
        rank[i] = r;
    }
    __syncthreads();
    double* wout = w + (int64_t)b * ld;
    for (int i = tid; i < ld; i += SYV_THREADS) wout[i] = INFINITY;
    __syncthreads();
    for (int i = tid; i < n; i += SYV_THREADS) wout[rank[i]] = wv[i];
    double* zg = Zout + (int64_t)b * ld * ld;
    for (int e = tid; e < n * n; e += SYV_THREADS) {
        const int j = e / n, i = e - j * n;
        zg[rank[j] * ld + i] = Zs[j * lda + i];
    }
}

extern "C" int sqc_syev(const double* A, int32_t n, int32_t ld, double* w, double* Zt, int32_t B, void* stream) {
    if (!A || !w || !Zt || B <= 0 || n <= 0 || ld < n || n > 112) return SQC_EINVAL;
    const int np = (n + 1) & ~1;
    const size_t bytes = ((size_t)2 * np * (np | 1) + 2 * np + 64) * sizeof(double) + 2 * np * sizeof(int) + 16;
    if (bytes > 227 * 1024) return SQC_ELIMIT;
    cudaError_t e = cudaFuncSetAttribute(syev_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    syev_kernel<<<B, SYV_THREADS, bytes, (cudaStream_t)stream>>>(A, n, ld, w, Zt);
    SQC_LAUNCHED();
    return 0;
}
