// Two-level Davidson preconditioner: exact inverse of (H - sigma) restricted to the ns basis
// states with the smallest LC energy (where the diagonal preconditioner is poor), the LC diagonal
// everywhere else.  The block is assembled analytically from the Fock-basis displacement tables
// (no operator applications), inverted in shared memory in complex64 (a preconditioner does not
// need more), and applied as gather -> small mat-mul -> scatter.
// Layout: one harmonic mode (mode 0, dim m) x one charge mode (mode 1, dim inner).
#include "common.cuh"

#define LB_MAXT 4

struct LbGeom {
    int m, inner, ns, n_terms;
    int shift[LB_MAXT];
};

__global__ void __launch_bounds__(256)
lowblock_assemble_kernel(LbGeom gm, const int32_t* __restrict__ S, long long s_stride_b,
                         const cplx* __restrict__ Dt, const cplx* __restrict__ a,
                         const double* __restrict__ g, const double* __restrict__ d,
                         long long d_stride_b, const double* __restrict__ sigma, double inv_unit,
                         float2* __restrict__ out) {
    const int b = blockIdx.y;
    const int e = blockIdx.x * 256 + threadIdx.x;
    const int ns = gm.ns;
    if (e >= ns * ns) return;
    const int r = e / ns, q = e % ns;
    const int32_t* s = S + b * s_stride_b;
    const int ir = s[r], iq = s[q];
    const int n = ir / gm.inner, c = ir % gm.inner;
    const int n2 = iq / gm.inner, c2 = iq % gm.inner;
    double vr = 0.0, vi = 0.0;
    if (r == q) vr += d[b * d_stride_b + ir] - sigma[b];
    if (c == c2) {
        // V diag(g0 + g1 lam) V^T = g0 I + g1 (a + a^dag)
        const double g0 = g[(long long)b * 3], g1 = g[(long long)b * 3 + 1];
        if (n == n2) vr += g0;
        if (n == n2 + 1 || n2 == n + 1) vr += g1 * sqrt((double)max(n, n2));
    }
    for (int t = 0; t < gm.n_terms; ++t) {
        const int sh = gm.shift[t];
        const cplx at = a[(long long)b * gm.n_terms + t];
        if (c2 == c - sh) {            // a_t E_t :  D_t[n][n'] on (c, c - s)
            const cplx w = cmul(at, Dt[((long long)t * gm.m + n) * gm.m + n2]);
            vr += w.x;
            vi += w.y;
        }
        if (c2 == c + sh) {            // h.c. :  conj(a_t) conj(D_t[n'][n]) on (c, c + s)
            const cplx w = cmul(at, Dt[((long long)t * gm.m + n2) * gm.m + n]);
            vr += w.x;
            vi -= w.y;
        }
    }
    out[(long long)b * ns * ns + e] = make_float2((float)(vr * inv_unit), (float)(vi * inv_unit));
}

extern "C" int sqc_lowblock_assemble(const sqc_space_t* sp, int32_t ns, const int32_t* S, int64_t s_stride_b,
                                     const void* Dt, const sqc_potential_t* pot, const double* d,
                                     int64_t d_stride_b, const double* sigma, double unit, void* out,
                                     int32_t B, void* stream) {
    if (!sp || !pot || B <= 0 || ns <= 0) return SQC_EINVAL;
    if (sp->n_modes != 2 || !sp->harmonic[0] || sp->harmonic[1] || pot->n_terms > LB_MAXT) return SQC_ELIMIT;
    LbGeom gm;
    gm.m = sp->dims[0];
    gm.inner = sp->dims[1];
    gm.ns = ns;
    gm.n_terms = pot->n_terms;
    for (int t = 0; t < gm.n_terms; ++t) gm.shift[t] = pot->shift[t][1];
    dim3 grid((ns * ns + 255) / 256, B);
    lowblock_assemble_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        gm, S, s_stride_b, (const cplx*)Dt, (const cplx*)pot->a, pot->g, d, d_stride_b, sigma, 1.0 / unit,
        (float2*)out);
    SQC_LAUNCHED();
    return 0;
}

// In-place Gauss-Jordan inverse (no pivoting: the matrices are Hermitian positive definite),
// whole matrix in shared memory, complex64.
__device__ __forceinline__ float2 cmulf(float2 x, float2 y) {
    return make_float2(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x);
}

__global__ void __launch_bounds__(1024)
hpd_inverse_kernel(float2* __restrict__ Aio, int ns) {
    extern __shared__ float2 sm[];
    const int lda = ns + 1;
    float2* A = sm;                 // [ns][lda]
    float2* rk = A + ns * lda;      // scaled pivot row
    float2* ck = rk + ns;           // pivot column
    float2* g = Aio + (long long)blockIdx.x * ns * ns;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < ns * ns; e += nt) A[(e / ns) * lda + (e % ns)] = g[e];
    __syncthreads();
    for (int k = 0; k < ns; ++k) {
        if (tid < ns) {
            const float2 p = A[k * lda + k];
            const float den = 1.0f / (p.x * p.x + p.y * p.y);
            const float2 dinv = make_float2(p.x * den, -p.y * den);
            float2 v = (tid == k) ? make_float2(1.0f, 0.0f) : A[k * lda + tid];
            rk[tid] = cmulf(v, dinv);
            ck[tid] = A[tid * lda + k];
        }
        __syncthreads();
        for (int e = tid; e < ns * ns; e += nt) {
            const int i = e / ns, j = e % ns;
            float2 v;
            if (i == k) {
                v = rk[j];
            } else {
                const float2 old = (j == k) ? make_float2(0.0f, 0.0f) : A[i * lda + j];
                const float2 w = cmulf(ck[i], rk[j]);
                v = make_float2(old.x - w.x, old.y - w.y);
            }
            A[i * lda + j] = v;
        }
        __syncthreads();
    }
    for (int e = tid; e < ns * ns; e += nt) g[e] = A[(e / ns) * lda + (e % ns)];
}

extern "C" int sqc_hpd_inverse_c64(void* A, int32_t ns, int32_t B, void* stream) {
    if (B <= 0 || ns <= 0) return SQC_EINVAL;
    size_t smem = ((size_t)ns * (ns + 1) + 2 * (size_t)ns) * sizeof(float2);
    if (smem > 227 * 1024) return SQC_ELIMIT;
    cudaError_t e = cudaFuncSetAttribute(hpd_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    hpd_inverse_kernel<<<B, 1024, smem, (cudaStream_t)stream>>>((float2*)A, ns);
    SQC_LAUNCHED();
    return 0;
}

// T[b][jj][S[r]] = (1/unit) * sum_q Ainv[b][r][q] * (HX_a - theta_a X_a)[S[q]],  a = act[b][jj]
__global__ void __launch_bounds__(256)
lowblock_apply_kernel(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* __restrict__ theta,
                      int ld_theta, const int32_t* __restrict__ act, int ld_act,
                      const int32_t* __restrict__ S, long long s_stride_b, const float2* __restrict__ Ainv,
                      int ns, double inv_unit, int group, long long N) {
    extern __shared__ double smd[];
    cplx* rS = reinterpret_cast<cplx*>(smd);        // [na][ns]
    const int b = blockIdx.x;
    const int na = blk_count(T, b);
    if (na <= 0) return;
    const int slot = b / group;
    const int32_t* s = S + slot * s_stride_b;
    const int tid = threadIdx.x;
    for (int e = tid; e < na * ns; e += 256) {
        const int jj = e / ns, q = e % ns;
        const int av = act[(long long)b * ld_act + jj];
        const double th = theta[(long long)b * ld_theta + av];
        const cplx x = blk_vec(X, b, av, N)[s[q]], h = blk_vec(HX, b, av, N)[s[q]];
        rS[e] = cmake(h.x - th * x.x, h.y - th * x.y);
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const float2* ai = Ainv + (long long)slot * ns * ns;
    for (int r = warp; r < ns; r += 8) {
        const float2* row = ai + (long long)r * ns;
        for (int jj = 0; jj < na; ++jj) {
            double sr = 0.0, si = 0.0;
            for (int q = lane; q < ns; q += 32) {
                const float2 m = row[q];
                const cplx v = rS[jj * ns + q];
                sr += (double)m.x * v.x - (double)m.y * v.y;
                si += (double)m.x * v.y + (double)m.y * v.x;
            }
            sr = warp_sum(sr);
            si = warp_sum(si);
            if (lane == 0) blk_vec(T, b, jj, N)[s[r]] = cmake(sr * inv_unit, si * inv_unit);
        }
    }
}

extern "C" int sqc_lowblock_apply(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* theta,
                                  int32_t ld_theta, const int32_t* act, int32_t ld_act, const int32_t* S,
                                  int64_t s_stride_b, const void* Ainv, int32_t ns, double unit, int32_t group,
                                  int64_t N, int32_t B, void* stream) {
    if (B <= 0 || ns <= 0 || T.cnt_fixed <= 0 || group < 1) return SQC_EINVAL;
    size_t smem = (size_t)T.cnt_fixed * ns * sizeof(cplx);
    cudaError_t e = cudaFuncSetAttribute(lowblock_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    lowblock_apply_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(X, HX, T, theta, ld_theta, act, ld_act, S,
                                                                 s_stride_b, (const float2*)Ainv, ns, 1.0 / unit, group, N);
    SQC_LAUNCHED();
    return 0;
}
