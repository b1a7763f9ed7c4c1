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

// Any layout (up to SQC_MAX_MODES harmonic / charge modes): the same block from per-mode displacement tables
//   Dt[b][t][off_i + n m_i + n'] = (V_i diag(etab_t^(i)) V_i^T)[n][n']   (harmonic modes with m_i > 1, concatenated),
// the matrix of junction term t on mode i in the Fock basis.  H_SS[r][q] between the basis states r, q (multi-indices
// decoded from the linear indices in S):  LC diagonal - sigma;  g_0 + sum_i g_i (a + a^dag)_i on states that differ
// in harmonic mode i only;  a_t prod_i Dt_i[n_i][n'_i] where every charge index of q is that of r minus the shift
// of term t, and the Hermitian conjugate.  Used by multi-mode sweeps and by batches of circuits (per-point tables).
struct LbgGeom {
    int M, ns, n_terms;
    int dims[SQC_MAX_MODES], harm[SQC_MAX_MODES], stride[SQC_MAX_MODES], dtoff[SQC_MAX_MODES];
    int shift[SQC_MAX_TERMS][SQC_MAX_MODES];
    int dtlen;
};

// DENSE = 1: the whole Hamiltonian in complex128 (S = all states in natural order, no shift, no scaling) for the
// dense eigensolver path (sqc_hamiltonian_dense); DENSE = 2: its real part as doubles (layouts without charge
// modes: H is real symmetric in the Fock basis).
template <int DENSE>
__global__ void __launch_bounds__(256)
lowblock_assemble_generic_kernel(const __grid_constant__ LbgGeom gm, const int32_t* __restrict__ S, long long s_stride_b,
                                 const cplx* __restrict__ Dt, long long dt_stride_b, const cplx* __restrict__ a,
                                 const double* __restrict__ g, const double* __restrict__ d, long long d_stride_b,
                                 const double* __restrict__ sigma, double inv_unit, void* __restrict__ out_) {
    const int b = blockIdx.y;
    const int e = blockIdx.x * 256 + threadIdx.x;
    const int ns = gm.ns, M = gm.M;
    if (e >= ns * ns) return;
    const int r = e / ns, q = e % ns;
    const int32_t* s = DENSE ? nullptr : S + b * s_stride_b;
    int ir = DENSE ? r : s[r], iq = DENSE ? q : s[q];
    int nr[SQC_MAX_MODES], nq[SQC_MAX_MODES];
    int ndiff = 0, dmode = -1;                       // number of modes in which r and q differ, the last such mode
    {
        int xr = ir, xq = iq;
#pragma unroll
        for (int i = 0; i < SQC_MAX_MODES; ++i) {
            if (i < M) {
                nr[i] = xr / gm.stride[i];
                xr -= nr[i] * gm.stride[i];
                nq[i] = xq / gm.stride[i];
                xq -= nq[i] * gm.stride[i];
                if (nr[i] != nq[i]) {
                    ++ndiff;
                    dmode = i;
                }
            }
        }
    }
    double vr = 0.0, vi = 0.0;
    const double* gb = g + (long long)b * (1 + M);
    if (ndiff == 0) vr += d[b * d_stride_b + ir] - (DENSE ? 0.0 : sigma[b]) + gb[0];
    if (ndiff == 1 && gm.harm[dmode] && abs(nr[dmode] - nq[dmode]) == 1)
        vr += gb[1 + dmode] * sqrt((double)max(nr[dmode], nq[dmode]));
    const cplx* dt = Dt + b * dt_stride_b;
    for (int t = 0; t < gm.n_terms; ++t) {
        bool fw = true, bw = true;
#pragma unroll
        for (int i = 0; i < SQC_MAX_MODES; ++i) {
            if (i < M && !gm.harm[i]) {
                fw = fw && (nq[i] == nr[i] - gm.shift[t][i]);
                bw = bw && (nq[i] == nr[i] + gm.shift[t][i]);
            }
        }
        if (!fw && !bw) continue;
        cplx pf = cmake(1.0, 0.0), pb = cmake(1.0, 0.0);      // prod_i Dt_i[n_r][n_q] and prod_i Dt_i[n_q][n_r]
#pragma unroll
        for (int i = 0; i < SQC_MAX_MODES; ++i) {
            if (i < M && gm.harm[i]) {
                if (gm.dims[i] > 1) {
                    const cplx* di = dt + (long long)t * gm.dtlen + gm.dtoff[i];
                    pf = cmul(pf, di[nr[i] * gm.dims[i] + nq[i]]);
                    pb = cmul(pb, di[nq[i] * gm.dims[i] + nr[i]]);
                }
            }
        }
        const cplx at = a[(long long)b * gm.n_terms + t];
        if (fw) {
            const cplx w = cmul(at, pf);
            vr += w.x;
            vi += w.y;
        }
        if (bw) {
            const cplx w = cmul(at, pb);
            vr += w.x;
            vi -= w.y;
        }
    }
    if (DENSE == 2) reinterpret_cast<double*>(out_)[(long long)b * ns * ns + e] = vr;
    else if (DENSE == 1) reinterpret_cast<cplx*>(out_)[(long long)b * ns * ns + e] = cmake(vr, vi);
    else reinterpret_cast<float2*>(out_)[(long long)b * ns * ns + e] = make_float2((float)(vr * inv_unit), (float)(vi * inv_unit));
}

static int lbg_launch(int dense, const sqc_space_t* sp, int32_t ns, const int32_t* S, int64_t s_stride_b,
                      const void* Dt, int64_t dt_stride_b, const sqc_potential_t* pot,
                      const double* d, int64_t d_stride_b, const double* sigma, double unit,
                      void* out, int32_t B, void* stream) {
    if (!sp || !pot || B <= 0 || ns <= 0 || sp->n_modes < 1 || sp->n_modes > SQC_MAX_MODES) return SQC_EINVAL;
    if (pot->n_terms < 0 || pot->n_terms > SQC_MAX_TERMS) return SQC_EINVAL;
    LbgGeom gm = {};
    gm.M = sp->n_modes;
    gm.ns = ns;
    gm.n_terms = pot->n_terms;
    long long N = 1;
    for (int i = sp->n_modes - 1; i >= 0; --i) {
        gm.dims[i] = sp->dims[i];
        gm.harm[i] = sp->harmonic[i];
        gm.stride[i] = (int)N;
        N *= sp->dims[i];
        if (N > 0x7fffffffLL) return SQC_ELIMIT;
    }
    int off = 0;
    for (int i = 0; i < sp->n_modes; ++i) {
        gm.dtoff[i] = off;
        if (sp->harmonic[i] && sp->dims[i] > 1) off += sp->dims[i] * sp->dims[i];
    }
    gm.dtlen = off;
    for (int t = 0; t < pot->n_terms; ++t)
        for (int i = 0; i < sp->n_modes; ++i) gm.shift[t][i] = sp->harmonic[i] ? 0 : pot->shift[t][i];
    if (dense && (long long)ns != N) return SQC_EINVAL;
    dim3 grid((ns * ns + 255) / 256, B);
    if (dense == 2)
        lowblock_assemble_generic_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(
            gm, S, s_stride_b, (const cplx*)Dt, dt_stride_b, (const cplx*)pot->a, pot->g, d, d_stride_b, sigma, 1.0, out);
    else if (dense == 1)
        lowblock_assemble_generic_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(
            gm, S, s_stride_b, (const cplx*)Dt, dt_stride_b, (const cplx*)pot->a, pot->g, d, d_stride_b, sigma, 1.0, out);
    else
        lowblock_assemble_generic_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(
            gm, S, s_stride_b, (const cplx*)Dt, dt_stride_b, (const cplx*)pot->a, pot->g, d, d_stride_b, sigma, 1.0 / unit,
            out);
    SQC_LAUNCHED();
    return 0;
}

extern "C" int sqc_lowblock_assemble_generic(const sqc_space_t* sp, int32_t ns, const int32_t* S, int64_t s_stride_b,
                                             const void* Dt, int64_t dt_stride_b, const sqc_potential_t* pot,
                                             const double* d, int64_t d_stride_b, const double* sigma, double unit,
                                             void* out, int32_t B, void* stream) {
    if (!S || !sigma) return SQC_EINVAL;
    return lbg_launch(0, sp, ns, S, s_stride_b, Dt, dt_stride_b, pot, d, d_stride_b, sigma, unit, out, B, stream);
}

extern "C" int sqc_hamiltonian_dense(const sqc_space_t* sp, const void* Dt, int64_t dt_stride_b,
                                     const sqc_potential_t* pot, const double* d, int64_t d_stride_b, void* out,
                                     int32_t real_out, int32_t B, void* stream) {
    if (!sp) return SQC_EINVAL;
    if (real_out)       // only layouts without charge modes have a real symmetric H in this basis
        for (int i = 0; i < sp->n_modes; ++i)
            if (!sp->harmonic[i] && sp->dims[i] > 1) return SQC_EINVAL;
    long long N = 1;
    for (int i = 0; i < sp->n_modes; ++i) N *= sp->dims[i];
    if (N > 4096) return SQC_ELIMIT;
    return lbg_launch(real_out ? 2 : 1, sp, (int32_t)N, nullptr, 0, Dt, dt_stride_b, pot, d, d_stride_b, nullptr, 1.0, out, B, stream);
}

// Real (time-reversal adapted) representation, see hx_real.cu: the index set holds REAL basis states
// R = i * inner + j (Fock i; j = 0: charge 0, 1 <= j <= nc: cos-type (|n> + |-n>)/sqrt2, j > nc: sin-type
// i(|n> - |-n>)/sqrt2 with n = j resp. j - nc).  H_r[R][R'] = sum_{s,s'} conj(u_s) u'_s' h(i, s n; i', s' n')
// is real; it is stored as complex64 with a zero imaginary part so that the factorisation and
// substitution kernels below serve both representations.
__device__ __forceinline__ cplx lb_h_elem(const LbGeom& gm, const cplx* __restrict__ Dt, const cplx* __restrict__ a,
                                          double g0, double g1, int b, int n, int q, int n2, int q2) {
    // element between (Fock n, charge q) and (Fock n2, charge q2), without the LC diagonal
    double vr = 0.0, vi = 0.0;
    if (q == q2) {
        if (n == n2) vr += g0;
        if (n == n2 + 1 || n2 == n + 1) vr += g1 * sqrt((double)max(n, n2));
    }
    for (int t = 0; t < gm.n_terms; ++t) {
        const int sh = gm.shift[t];
        const cplx at = a[(long long)b * gm.n_terms + t];
        if (q2 == q - sh) {
            const cplx w = cmul(at, Dt[((long long)t * gm.m + n) * gm.m + n2]);
            vr += w.x;
            vi += w.y;
        }
        if (q2 == q + sh) {
            const cplx w = cmul(at, Dt[((long long)t * gm.m + n2) * gm.m + n]);
            vr += w.x;
            vi -= w.y;
        }
    }
    return cmake(vr, vi);
}

__global__ void __launch_bounds__(256)
lowblock_assemble_real_kernel(LbGeom gm, const int32_t* __restrict__ S, long long s_stride_b,
                              const cplx* __restrict__ Dt, const cplx* __restrict__ a,
                              const double* __restrict__ g, const double* __restrict__ d,
                              long long d_stride_b, const double* __restrict__ sigma, double inv_unit,
                              float2* __restrict__ out, int real_storage) {
    const int b = blockIdx.y;
    const int e = blockIdx.x * 256 + threadIdx.x;
    const int ns = gm.ns;
    if (e >= ns * ns) return;
    const int r = e / ns, q = e % ns;
    const int32_t* s = S + b * s_stride_b;
    const int ir = s[r], iq = s[q];
    const int nc = (gm.inner - 1) / 2;
    const int n = ir / gm.inner, j = ir % gm.inner;
    const int n2 = iq / gm.inner, j2 = iq % gm.inner;
    const int c = j > nc ? j - nc : j, c2 = j2 > nc ? j2 - nc : j2;      // |charge|
    const double rs2 = 0.70710678118654752440;
    // u_+ , u_- of both states (only u_+ for charge 0)
    const cplx up = j == 0 ? cmake(1.0, 0.0) : (j <= nc ? cmake(rs2, 0.0) : cmake(0.0, rs2));
    const cplx um = j == 0 ? cmake(0.0, 0.0) : (j <= nc ? cmake(rs2, 0.0) : cmake(0.0, -rs2));
    const cplx vp = j2 == 0 ? cmake(1.0, 0.0) : (j2 <= nc ? cmake(rs2, 0.0) : cmake(0.0, rs2));
    const cplx vm = j2 == 0 ? cmake(0.0, 0.0) : (j2 <= nc ? cmake(rs2, 0.0) : cmake(0.0, -rs2));
    const double g0 = g[(long long)b * 3], g1 = g[(long long)b * 3 + 1];
    double acc = 0.0;
    for (int sa = 0; sa < (j == 0 ? 1 : 2); ++sa) {
        const cplx ua = sa ? um : up;
        const int qa = sa ? -c : c;
        for (int sb = 0; sb < (j2 == 0 ? 1 : 2); ++sb) {
            const cplx ub = sb ? vm : vp;
            const int qb = sb ? -c2 : c2;
            if (abs(qa - qb) > 1) continue;
            const cplx h = lb_h_elem(gm, Dt, a, g0, g1, b, n, qa, n2, qb);
            const cplx w = cmul(cmulc(ua, ub), h);        // conj(u_a) u_b h
            acc += w.x;
        }
    }
    if (r == q) acc += d[b * d_stride_b + ir] - sigma[b];
    if (real_storage) reinterpret_cast<float*>(out)[(long long)b * ns * ns + e] = (float)(acc * inv_unit);
    else out[(long long)b * ns * ns + e] = make_float2((float)(acc * inv_unit), 0.0f);
}

extern "C" int sqc_lowblock_assemble_real(const sqc_space_t* sp, int32_t ns, const int32_t* S, int64_t s_stride_b,
                                          const void* Dt, const sqc_potential_t* pot, const double* d,
                                          int64_t d_stride_b, const double* sigma, double unit, void* out,
                                          int32_t real_storage, int32_t B, void* stream) {
    if (!sp || !pot || B <= 0 || ns <= 0) return SQC_EINVAL;
    if (sp->n_modes != 2 || !sp->harmonic[0] || sp->harmonic[1] || pot->n_terms > LB_MAXT) return SQC_ELIMIT;
    if (!(sp->dims[1] & 1)) return SQC_ELIMIT;
    LbGeom gm;
    gm.m = sp->dims[0];
    gm.inner = sp->dims[1];
    gm.ns = ns;
    gm.n_terms = pot->n_terms;
    for (int t = 0; t < gm.n_terms; ++t) gm.shift[t] = pot->shift[t][1];
    dim3 grid((ns * ns + 255) / 256, B);
    lowblock_assemble_real_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        gm, S, s_stride_b, (const cplx*)Dt, (const cplx*)pot->a, pot->g, d, d_stride_b, sigma, 1.0 / unit,
        (float2*)out, real_storage);
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
// REAL: the vectors are real arrays of 2 N doubles (real representation) and S holds real indices
template <bool REAL>
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
        if (REAL) {
            const double x = reinterpret_cast<const double*>(blk_vec(X, b, av, N))[s[q]];
            const double h = reinterpret_cast<const double*>(blk_vec(HX, b, av, N))[s[q]];
            rS[e] = cmake(h - th * x, 0.0);
        } else {
            const cplx x = blk_vec(X, b, av, N)[s[q]], h = blk_vec(HX, b, av, N)[s[q]];
            rS[e] = cmake(h.x - th * x.x, h.y - th * x.y);
        }
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
            if (lane == 0) {
                if (REAL) reinterpret_cast<double*>(blk_vec(T, b, jj, N))[s[r]] = sr * inv_unit;
                else blk_vec(T, b, jj, N)[s[r]] = cmake(sr * inv_unit, si * inv_unit);
            }
        }
    }
}

template <bool REAL>
static int lowblock_apply_launch(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* theta,
                                 int32_t ld_theta, const int32_t* act, int32_t ld_act, const int32_t* S,
                                 int64_t s_stride_b, const void* Ainv, int32_t ns, double unit, int32_t group,
                                 int64_t N, int32_t B, void* stream) {
    if (B <= 0 || ns <= 0 || T.cnt_fixed <= 0 || group < 1) return SQC_EINVAL;
    size_t smem = (size_t)T.cnt_fixed * ns * sizeof(cplx);
    cudaError_t e = cudaFuncSetAttribute(lowblock_apply_kernel<REAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    lowblock_apply_kernel<REAL><<<B, 256, smem, (cudaStream_t)stream>>>(X, HX, T, theta, ld_theta, act, ld_act, S,
                                                                       s_stride_b, (const float2*)Ainv, ns, 1.0 / unit, group, N);
    SQC_LAUNCHED();
    return 0;
}

extern "C" int sqc_lowblock_apply(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* theta,
                                  int32_t ld_theta, const int32_t* act, int32_t ld_act, const int32_t* S,
                                  int64_t s_stride_b, const void* Ainv, int32_t ns, double unit, int32_t group,
                                  int64_t N, int32_t B, void* stream) {
    return lowblock_apply_launch<false>(X, HX, T, theta, ld_theta, act, ld_act, S, s_stride_b, Ainv, ns, unit, group, N,
                                        B, stream);
}

extern "C" int sqc_lowblock_apply_real(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* theta,
                                       int32_t ld_theta, const int32_t* act, int32_t ld_act, const int32_t* S,
                                       int64_t s_stride_b, const void* Ainv, int32_t ns, double unit, int32_t group,
                                       int64_t N, int32_t B, void* stream) {
    return lowblock_apply_launch<true>(X, HX, T, theta, ld_theta, act, ld_act, S, s_stride_b, Ainv, ns, unit, group, N,
                                       B, stream);
}

// =====================================================================================
// Banded variant for LARGE low-energy blocks (ns up to 1024).
//
// With the index set ordered charge-major, H_SS is block tridiagonal: the junction terms shift the
// charge index by at most one, everything else is diagonal in it.  Block b_q holds the Fock states
// kept at charge q (a few dozen), so instead of a dense ns x ns inverse (ns^2 per right-hand side,
// ns^3 to build) the block LDL^H factorisation is used:
//     D'_0 = D_0,   D'_q = D_q - E_{q-1}^H P_{q-1} E_{q-1},   P_q = D'_q^-1
//     forward   y_q = r_q - E_{q-1}^H z_{q-1},   z_q = P_q y_q
//     backward  x_q = z_q - P_q (E_q x_{q+1})
// (D_q: diagonal blocks, E_q: block (q, q+1)).  Cost per right-hand side ~ 3 sum_q n_q^2 -- for the
// 512 lowest states of the 0-pi bench about the same as the dense 160 x 160 product, while the
// Davidson iteration count drops by a fifth.
// sqc_lowband_factor overwrites the diagonal blocks of the assembled matrix (sqc_lowblock_assemble
// with the charge-major index set) by P_q, in place; the E_q stay where they are.
// =====================================================================================
#define LBN_THREADS 256

// Gauss-Jordan inverse of the n x n complex64 matrix M (leading dim ld) in shared memory, in place.
// rk, ck: n entries of scratch each.  All threads of the CTA call.
__device__ void gj_inverse_smem(float2* M, int n, int ld, float2* rk, float2* ck) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = 0; k < n; ++k) {
        if (tid < n) {
            const float2 p = M[k * ld + k];
            const float den = 1.0f / (p.x * p.x + p.y * p.y);
            const float2 dinv = make_float2(p.x * den, -p.y * den);
            const float2 v = (tid == k) ? make_float2(1.0f, 0.0f) : M[k * ld + tid];
            rk[tid] = cmulf(v, dinv);
            ck[tid] = M[tid * ld + k];
        }
        __syncthreads();
        for (int e = tid; e < n * n; e += nt) {
            const int i = e / n, j = e % n;
            float2 v;
            if (i == k) {
                v = rk[j];
            } else {
                const float2 old = (j == k) ? make_float2(0.0f, 0.0f) : M[i * ld + j];
                const float2 w = cmulf(ck[i], rk[j]);
                v = make_float2(old.x - w.x, old.y - w.y);
            }
            M[i * ld + j] = v;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(LBN_THREADS)
lowband_factor_kernel(float2* __restrict__ Aio, int ns, const int32_t* __restrict__ blk_off, int ld_off,
                      const int32_t* __restrict__ nblk, int nbmax) {
    extern __shared__ float2 smf[];
    const int ld = nbmax | 1;
    float2* M = smf;                    // current Schur complement -> P_q
    float2* Pp = M + nbmax * ld;        // P_{q-1}
    float2* E = Pp + nbmax * ld;        // E_{q-1}  [n_{q-1}][n_q]
    float2* Tm = E + nbmax * ld;        // P_{q-1} E_{q-1}
    float2* rk = Tm + nbmax * ld;
    float2* ck = rk + nbmax;
    const int slot = blockIdx.x, tid = threadIdx.x;
    float2* A = Aio + (long long)slot * ns * ns;
    const int32_t* off = blk_off + (long long)slot * ld_off;
    const int nb = nblk[slot];
    int nprev = 0, oprev = 0;
    for (int q = 0; q < nb; ++q) {
        const int o = off[q], n = off[q + 1] - o;
        for (int e = tid; e < n * n; e += LBN_THREADS) M[(e / n) * ld + (e % n)] = A[(long long)(o + e / n) * ns + o + e % n];
        if (q > 0)
            for (int e = tid; e < nprev * n; e += LBN_THREADS)
                E[(e / n) * ld + (e % n)] = A[(long long)(oprev + e / n) * ns + o + e % n];
        __syncthreads();
        if (q > 0) {
            // Tm = Pp E   (nprev x n)
            for (int e = tid; e < nprev * n; e += LBN_THREADS) {
                const int i = e / n, j = e % n;
                float2 acc = make_float2(0.0f, 0.0f);
                for (int k = 0; k < nprev; ++k) {
                    const float2 w = cmulf(Pp[i * ld + k], E[k * ld + j]);
                    acc.x += w.x;
                    acc.y += w.y;
                }
                Tm[i * ld + j] = acc;
            }
            __syncthreads();
            // M -= E^H Tm   (n x n)
            for (int e = tid; e < n * n; e += LBN_THREADS) {
                const int i = e / n, j = e % n;
                float2 acc = make_float2(0.0f, 0.0f);
                for (int k = 0; k < nprev; ++k) {
                    const float2 ec = E[k * ld + i], t = Tm[k * ld + j];
                    acc.x += ec.x * t.x + ec.y * t.y;          // conj(ec) * t
                    acc.y += ec.x * t.y - ec.y * t.x;
                }
                M[i * ld + j].x -= acc.x;
                M[i * ld + j].y -= acc.y;
            }
            __syncthreads();
        }
        gj_inverse_smem(M, n, ld, rk, ck);
        for (int e = tid; e < n * n; e += LBN_THREADS) {
            const float2 v = M[(e / n) * ld + (e % n)];
            A[(long long)(o + e / n) * ns + o + e % n] = v;
            Pp[(e / n) * ld + (e % n)] = v;
        }
        __syncthreads();
        nprev = n;
        oprev = o;
    }
}

extern "C" int sqc_lowband_factor(void* A, int32_t ns, const int32_t* blk_off, int32_t ld_off,
                                  const int32_t* nblk, int32_t nbmax, int32_t B, void* stream) {
    if (B <= 0 || ns <= 0 || !A || !blk_off || !nblk || nbmax <= 0) return SQC_EINVAL;
    if (nbmax > 96) return SQC_ELIMIT;
    const size_t smem = ((size_t)4 * nbmax * (nbmax | 1) + 2 * (size_t)nbmax) * sizeof(float2);
    cudaError_t e = cudaFuncSetAttribute(lowband_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    lowband_factor_kernel<<<B, LBN_THREADS, smem, (cudaStream_t)stream>>>((float2*)A, ns, blk_off, ld_off, nblk, nbmax);
    SQC_LAUNCHED();
    return 0;
}

// Apply: T[b][jj][S[i]] = ((H_SS - sigma)^-1 (HX_a - theta_a X_a)[S])_i,  a = act[b][jj], through the block
// factorisation above.  One CTA per sweep point; a warp owns LBC right-hand sides (stored
// interleaved, [row][LBC]) and a lane one row of the current block, so a factor entry is fetched from
// shared memory once per four multiply-adds and the four vector entries arrive as two broadcast
// 16-byte loads -- the kernel is shared-memory-bandwidth bound (ncu: one RHS per warp issued
// 3 wavefronts per multiply-add).  The factor blocks of a step are staged once for all warps,
// one step ahead (cp.async, two buffers).
#ifndef LBC
#define LBC 2
#endif

struct f2x4 { float2 c[LBC]; };

// acc[c] (+/-)= sum_k conj(M[k][i]) x[k][c]    (Mcol = &M[0][i];  x: [n][LBC])
template <bool SUB>
__device__ __forceinline__ void dot4_conj_col(f2x4& acc, const float2* __restrict__ Mcol, int ld,
                                              const float2* __restrict__ x, int n) {
    for (int k = 0; k < n; ++k) {
        const float2 m = Mcol[k * ld];
        float2 xs[LBC];
#pragma unroll
        for (int h = 0; h < LBC / 2; ++h) {
            const float4 xa = *reinterpret_cast<const float4*>(x + k * LBC + 2 * h);
            xs[2 * h] = make_float2(xa.x, xa.y);
            xs[2 * h + 1] = make_float2(xa.z, xa.w);
        }
#pragma unroll
        for (int c = 0; c < LBC; ++c) {
            const float re = m.x * xs[c].x + m.y * xs[c].y, im = m.x * xs[c].y - m.y * xs[c].x;
            acc.c[c].x += SUB ? -re : re;
            acc.c[c].y += SUB ? -im : im;
        }
    }
}
// acc[c] = sum_k M[i][k] x[k][c]    (Mrow = &M[i][0])
__device__ __forceinline__ void dot4_row(f2x4& acc, const float2* __restrict__ Mrow, const float2* __restrict__ x, int n) {
    for (int k = 0; k < n; ++k) {
        const float2 m = Mrow[k];
        float2 xs[LBC];
#pragma unroll
        for (int h = 0; h < LBC / 2; ++h) {
            const float4 xa = *reinterpret_cast<const float4*>(x + k * LBC + 2 * h);
            xs[2 * h] = make_float2(xa.x, xa.y);
            xs[2 * h + 1] = make_float2(xa.z, xa.w);
        }
#pragma unroll
        for (int c = 0; c < LBC; ++c) {
            acc.c[c].x += m.x * xs[c].x - m.y * xs[c].y;
            acc.c[c].y += m.x * xs[c].y + m.y * xs[c].x;
        }
    }
}
__device__ __forceinline__ void st4(float2* dst, const f2x4& v) {
#pragma unroll
    for (int h = 0; h < LBC / 2; ++h)
        *reinterpret_cast<float4*>(dst + 2 * h) = make_float4(v.c[2 * h].x, v.c[2 * h].y, v.c[2 * h + 1].x, v.c[2 * h + 1].y);
}
__device__ __forceinline__ f2x4 ld4(const float2* src) {
    f2x4 v;
#pragma unroll
    for (int h = 0; h < LBC / 2; ++h) {
        const float4 a = *reinterpret_cast<const float4*>(src + 2 * h);
        v.c[2 * h] = make_float2(a.x, a.y);
        v.c[2 * h + 1] = make_float2(a.z, a.w);
    }
    return v;
}

__global__ void __launch_bounds__(32 * (16 / LBC))
lowband_apply_kernel(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* __restrict__ theta,
                     int ld_theta, const int32_t* __restrict__ act, int ld_act,
                     const int32_t* __restrict__ S, long long s_stride_b, const float2* __restrict__ Afac,
                     int ns, const int32_t* __restrict__ blk_off, int ld_off, const int32_t* __restrict__ nblk,
                     int nbmax, double inv_unit, int group, long long N, const int32_t* __restrict__ perm) {
    extern __shared__ float2 smf[];
    const int b = blockIdx.x;
    const int na = blk_count(T, b);
    if (na <= 0) return;
    const int ld = nbmax | 1;
    const int nwarp = blockDim.x >> 5;
    const int bufsz = 2 * nbmax * ld;          // P block + E block
    float2* buf = smf;                         // two buffers of (P, E)
    float2* vec = buf + 2 * bufsz + (2 * bufsz & 1);      // 16-byte aligned: [nwarp][ns][LBC]  r -> z -> x
    float2* tmp = vec + (size_t)nwarp * ns * LBC;         // [nwarp][nbmax][LBC]
    const int slot = b / group;
    const int32_t* s = S + slot * s_stride_b;
    const int32_t* pm = perm ? perm + slot * s_stride_b : nullptr;
    const float2* A = Afac + (long long)slot * ns * ns;
    const int32_t* off = blk_off + (long long)slot * ld_off;
    const int nb = nblk[slot];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c0 = warp * LBC;                 // first right-hand side of this warp
    const bool mine = c0 < na;
    float2* v = vec + (size_t)warp * ns * LBC;
    float2* u = tmp + (size_t)warp * nbmax * LBC;
    // stage step `st` (0 .. 2 nb - 2: forward steps q = st, then backward steps q = 2 nb - 2 - st)
    // into buffer st & 1:  P_q, and E_{q-1} (forward) or E_q (backward).  8-byte cp.async.
    auto stage = [&](int st) {
        if (st <= 2 * nb - 2) {
            float2* P = buf + (st & 1) * bufsz;
            float2* E = P + nbmax * ld;
            const bool fwd = st < nb;
            const int q = fwd ? st : 2 * nb - 2 - st;
            const int o = off[q], n = off[q + 1] - o;
            for (int e = tid; e < n * n; e += blockDim.x)
                cp_async8(P + (e / n) * ld + (e % n), A + (long long)(o + e / n) * ns + o + e % n);
            if (fwd && q > 0) {
                const int op = off[q - 1], np_ = o - op;
                for (int e = tid; e < np_ * n; e += blockDim.x)
                    cp_async8(E + (e / n) * ld + (e % n), A + (long long)(op + e / n) * ns + o + e % n);
            } else if (!fwd) {
                const int on = off[q + 1], nn = off[q + 2] - on;
                for (int e = tid; e < n * nn; e += blockDim.x)
                    cp_async8(E + (e / nn) * ld + (e % nn), A + (long long)(o + e / nn) * ns + on + e % nn);
            }
        }
        cp_async_commit();
    };
    stage(0);
    if (mine) {
        // gather in ascending-address order (perm sorts the index set by address)
#pragma unroll
        for (int c = 0; c < LBC; ++c) {
            if (c0 + c < na) {
                const int av = act[(long long)b * ld_act + c0 + c];
                const double th = theta[(long long)b * ld_theta + av];
                const cplx* x = blk_vec(X, b, av, N);
                const cplx* h = blk_vec(HX, b, av, N);
                for (int j = lane; j < ns; j += 32) {
                    const int i = pm ? pm[j] : j;
                    const int idx = s[i];
                    const cplx xv = x[idx], hv = h[idx];
                    v[i * LBC + c] = make_float2((float)((hv.x - th * xv.x) * inv_unit),
                                                 (float)((hv.y - th * xv.y) * inv_unit));
                }
            } else {
                for (int j = lane; j < ns; j += 32) v[j * LBC + c] = make_float2(0.0f, 0.0f);
            }
        }
    }
    for (int st = 0; st <= 2 * nb - 2; ++st) {
        __syncthreads();                       // everyone is done with buffer (st + 1) & 1 (step st - 1)
        stage(st + 1);
        cp_async_wait<1>();                    // this thread's copies of step st have landed
        __syncthreads();                       // ... and everybody else's
        const float2* P = buf + (st & 1) * bufsz;
        const float2* E = P + nbmax * ld;
        if (!mine) continue;
        if (st < nb) {
            // ---- forward:  y = r_q - E_{q-1}^H z_{q-1},  z_q = P_q y ----
            const int q = st, o = off[q], n = off[q + 1] - o;
            const int op = q > 0 ? off[q - 1] : 0, np_ = q > 0 ? o - op : 0;
            for (int i = lane; i < n; i += 32) {
                f2x4 acc = ld4(v + (o + i) * LBC);
                dot4_conj_col<true>(acc, E + i, ld, v + op * LBC, np_);
                st4(u + i * LBC, acc);
            }
            __syncwarp();
            // P Hermitian: P[i][k] = conj(P[k][i]) -> lanes read consecutive addresses
            for (int i = lane; i < n; i += 32) {
                f2x4 acc;
#pragma unroll
                for (int c = 0; c < LBC; ++c) acc.c[c] = make_float2(0.0f, 0.0f);
                dot4_conj_col<false>(acc, P + i, ld, u, n);
                st4(v + (o + i) * LBC, acc);
            }
            __syncwarp();
        } else {
            // ---- backward:  x_q = z_q - P_q (E_q x_{q+1}) ----
            const int q = 2 * nb - 2 - st, o = off[q], n = off[q + 1] - o;
            const int on = off[q + 1], nn = off[q + 2] - on;
            for (int i = lane; i < n; i += 32) {
                f2x4 acc;
#pragma unroll
                for (int c = 0; c < LBC; ++c) acc.c[c] = make_float2(0.0f, 0.0f);
                dot4_row(acc, E + i * ld, v + on * LBC, nn);
                st4(u + i * LBC, acc);
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) {
                f2x4 acc = ld4(v + (o + i) * LBC);
                dot4_conj_col<true>(acc, P + i, ld, u, n);
                st4(v + (o + i) * LBC, acc);
            }
            __syncwarp();
        }
    }
    cp_async_wait<0>();
    if (mine) {
#pragma unroll
        for (int c = 0; c < LBC; ++c) {
            if (c0 + c < na) {
                cplx* t = blk_vec(T, b, c0 + c, N);
                for (int j = lane; j < ns; j += 32) {
                    const int i = pm ? pm[j] : j;
                    const float2 w = v[i * LBC + c];
                    t[s[i]] = cmake((double)w.x, (double)w.y);
                }
            }
        }
    }
}

static int lowband_apply_launch(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* theta,
                                int32_t ld_theta, const int32_t* act, int32_t ld_act, const int32_t* S,
                                int64_t s_stride_b, const void* Afac, int32_t ns, const int32_t* blk_off,
                                int32_t ld_off, const int32_t* nblk, int32_t nbmax, double unit, int32_t group,
                                const int32_t* perm, int64_t N, int32_t B, void* stream) {
    if (B <= 0 || ns <= 0 || T.cnt_fixed <= 0 || T.cnt_fixed > 16 || group < 1 || nbmax <= 0) return SQC_EINVAL;
    const int nwarp = (T.cnt_fixed + LBC - 1) / LBC;
    const size_t smem = ((size_t)4 * nbmax * (nbmax | 1) + 1 + (size_t)nwarp * ns * LBC + (size_t)nwarp * nbmax * LBC) *
                        sizeof(float2);
    if (smem > 227 * 1024) return SQC_ELIMIT;
    cudaError_t e = cudaFuncSetAttribute(lowband_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    lowband_apply_kernel<<<B, 32 * nwarp, smem, (cudaStream_t)stream>>>(
        X, HX, T, theta, ld_theta, act, ld_act, S, s_stride_b, (const float2*)Afac, ns, blk_off, ld_off, nblk, nbmax,
        1.0 / unit, group, N, perm);
    SQC_LAUNCHED();
    return 0;
}

extern "C" int sqc_lowband_apply(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* theta,
                                 int32_t ld_theta, const int32_t* act, int32_t ld_act, const int32_t* S,
                                 int64_t s_stride_b, const void* Afac, int32_t ns, const int32_t* blk_off,
                                 int32_t ld_off, const int32_t* nblk, int32_t nbmax, double unit, int32_t group,
                                 const int32_t* perm, int64_t N, int32_t B, void* stream) {
    return lowband_apply_launch(X, HX, T, theta, ld_theta, act, ld_act, S, s_stride_b, Afac, ns, blk_off, ld_off,
                                       nblk, nbmax, unit, group, perm, N, B, stream);
}



// =====================================================================================
// Real-storage variants of the banded preconditioner (real representation, see hx_real.cu): the low block
// is real symmetric, so the factor is kept as float [ns][ns] (half the shared-memory traffic of the
// complex64 kernels, a quarter of the multiply-adds) and a warp carries FOUR right-hand sides as one float4
// per row.  Same block LDL^T algebra as above.
// =====================================================================================
__device__ void gj_inverse_smem_real(float* M, int n, int ld, float* rk, float* ck) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = 0; k < n; ++k) {
        if (tid < n) {
            const float dinv = 1.0f / M[k * ld + k];
            rk[tid] = ((tid == k) ? 1.0f : M[k * ld + tid]) * dinv;
            ck[tid] = M[tid * ld + k];
        }
        __syncthreads();
        for (int e = tid; e < n * n; e += nt) {
            const int i = e / n, j = e % n;
            float v;
            if (i == k) v = rk[j];
            else v = ((j == k) ? 0.0f : M[i * ld + j]) - ck[i] * rk[j];
            M[i * ld + j] = v;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(LBN_THREADS)
lowband_factor_real_kernel(float* __restrict__ Aio, int ns, const int32_t* __restrict__ blk_off, int ld_off,
                           const int32_t* __restrict__ nblk, int nbmax) {
    extern __shared__ float smr[];
    const int ld = nbmax | 1;
    float* M = smr;                    // current Schur complement -> P_q
    float* Pp = M + nbmax * ld;        // P_{q-1}
    float* E = Pp + nbmax * ld;        // E_{q-1}  [n_{q-1}][n_q]
    float* Tm = E + nbmax * ld;        // P_{q-1} E_{q-1}
    float* rk = Tm + nbmax * ld;
    float* ck = rk + nbmax;
    const int slot = blockIdx.x, tid = threadIdx.x;
    float* A = Aio + (long long)slot * ns * ns;
    const int32_t* off = blk_off + (long long)slot * ld_off;
    const int nb = nblk[slot];
    int nprev = 0, oprev = 0;
    for (int q = 0; q < nb; ++q) {
        const int o = off[q], n = off[q + 1] - o;
        for (int e = tid; e < n * n; e += LBN_THREADS) M[(e / n) * ld + (e % n)] = A[(long long)(o + e / n) * ns + o + e % n];
        if (q > 0)
            for (int e = tid; e < nprev * n; e += LBN_THREADS)
                E[(e / n) * ld + (e % n)] = A[(long long)(oprev + e / n) * ns + o + e % n];
        __syncthreads();
        if (q > 0) {
            for (int e = tid; e < nprev * n; e += LBN_THREADS) {          // Tm = Pp E
                const int i = e / n, j = e % n;
                float acc = 0.0f;
                for (int k = 0; k < nprev; ++k) acc = fmaf(Pp[i * ld + k], E[k * ld + j], acc);
                Tm[i * ld + j] = acc;
            }
            __syncthreads();
            for (int e = tid; e < n * n; e += LBN_THREADS) {              // M -= E^T Tm
                const int i = e / n, j = e % n;
                float acc = 0.0f;
                for (int k = 0; k < nprev; ++k) acc = fmaf(E[k * ld + i], Tm[k * ld + j], acc);
                M[i * ld + j] -= acc;
            }
            __syncthreads();
        }
        gj_inverse_smem_real(M, n, ld, rk, ck);
        for (int e = tid; e < n * n; e += LBN_THREADS) {
            const float v = M[(e / n) * ld + (e % n)];
            A[(long long)(o + e / n) * ns + o + e % n] = v;
            Pp[(e / n) * ld + (e % n)] = v;
        }
        __syncthreads();
        nprev = n;
        oprev = o;
    }
}

extern "C" int sqc_lowband_factor_real(void* A, int32_t ns, const int32_t* blk_off, int32_t ld_off,
                                       const int32_t* nblk, int32_t nbmax, int32_t B, void* stream) {
    if (B <= 0 || ns <= 0 || !A || !blk_off || !nblk || nbmax <= 0) return SQC_EINVAL;
    if (nbmax > 128) return SQC_ELIMIT;
    const size_t smem = ((size_t)4 * nbmax * (nbmax | 1) + 2 * (size_t)nbmax) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(lowband_factor_real_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    lowband_factor_real_kernel<<<B, LBN_THREADS, smem, (cudaStream_t)stream>>>((float*)A, ns, blk_off, ld_off, nblk, nbmax);
    SQC_LAUNCHED();
    return 0;
}

#define LBR 4          // right-hand sides per warp (one float4 per row)

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem_src));
}
__device__ __forceinline__ void fma4(float4& acc, float m, const float4& x) {
    acc.x = fmaf(m, x.x, acc.x);
    acc.y = fmaf(m, x.y, acc.y);
    acc.z = fmaf(m, x.z, acc.z);
    acc.w = fmaf(m, x.w, acc.w);
}

__global__ void __launch_bounds__(32 * (16 / LBR))
lowband_apply_real_kernel(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* __restrict__ theta,
                          int ld_theta, const int32_t* __restrict__ act, int ld_act,
                          const int32_t* __restrict__ S, long long s_stride_b, const float* __restrict__ Afac,
                          int ns, const int32_t* __restrict__ blk_off, int ld_off, const int32_t* __restrict__ nblk,
                          int nbmax, double inv_unit, int group, long long N, const int32_t* __restrict__ perm) {
    extern __shared__ float4 smr4[];
    const int b = blockIdx.x;
    const int na = blk_count(T, b);
    if (na <= 0) return;
    const int ld = nbmax | 1;
    const int nwarp = blockDim.x >> 5;
    const int bufsz = 2 * nbmax * ld;          // P block + E block (floats)
    float4* vec = smr4;                        // [nwarp][ns]  r -> z -> x, one float4 (LBR right-hand sides) per row
    float4* tmp = vec + (size_t)nwarp * ns;    // [nwarp][nbmax]
    float* buf = reinterpret_cast<float*>(tmp + (size_t)nwarp * nbmax);      // two buffers of (P, E)
    const int slot = b / group;
    const int32_t* s = S + slot * s_stride_b;
    const int32_t* pm = perm ? perm + slot * s_stride_b : nullptr;
    const float* A = Afac + (long long)slot * ns * ns;
    const int32_t* off = blk_off + (long long)slot * ld_off;
    const int nb = nblk[slot];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c0 = warp * LBR;
    const bool mine = c0 < na;
    float4* v = vec + (size_t)warp * ns;
    float4* u = tmp + (size_t)warp * nbmax;
    auto stage = [&](int st) {
        if (st <= 2 * nb - 2) {
            float* P = buf + (st & 1) * bufsz;
            float* E = P + nbmax * ld;
            const bool fwd = st < nb;
            const int q = fwd ? st : 2 * nb - 2 - st;
            const int o = off[q], n = off[q + 1] - o;
            for (int e = tid; e < n * n; e += blockDim.x)
                cp_async4(P + (e / n) * ld + (e % n), A + (long long)(o + e / n) * ns + o + e % n);
            if (fwd && q > 0) {
                const int op = off[q - 1], np_ = o - op;
                for (int e = tid; e < np_ * n; e += blockDim.x)
                    cp_async4(E + (e / n) * ld + (e % n), A + (long long)(op + e / n) * ns + o + e % n);
            } else if (!fwd) {
                const int on = off[q + 1], nn = off[q + 2] - on;
                for (int e = tid; e < n * nn; e += blockDim.x)
                    cp_async4(E + (e / nn) * ld + (e % nn), A + (long long)(o + e / nn) * ns + on + e % nn);
            }
        }
        cp_async_commit();
    };
    stage(0);
    if (mine) {
        // gather the residual entries of the index set (ascending-address order through perm)
        const double* xp[LBR];
        const double* hp[LBR];
        double th[LBR];
#pragma unroll
        for (int c = 0; c < LBR; ++c) {
            const bool on = c0 + c < na;
            const int av = on ? act[(long long)b * ld_act + c0 + c] : 0;
            th[c] = on ? theta[(long long)b * ld_theta + av] : 0.0;
            xp[c] = on ? reinterpret_cast<const double*>(blk_vec(X, b, av, N)) : nullptr;
            hp[c] = on ? reinterpret_cast<const double*>(blk_vec(HX, b, av, N)) : nullptr;
        }
        for (int j = lane; j < ns; j += 32) {
            const int i = pm ? pm[j] : j;
            const int idx = s[i];
            float r[LBR];
#pragma unroll
            for (int c = 0; c < LBR; ++c)
                r[c] = xp[c] ? (float)((hp[c][idx] - th[c] * xp[c][idx]) * inv_unit) : 0.0f;
            v[i] = make_float4(r[0], r[1], r[2], r[3]);
        }
    }
    for (int st = 0; st <= 2 * nb - 2; ++st) {
        __syncthreads();                       // everyone is done with buffer (st + 1) & 1 (step st - 1)
        stage(st + 1);
        cp_async_wait<1>();
        __syncthreads();
        const float* P = buf + (st & 1) * bufsz;
        const float* E = P + nbmax * ld;
        if (!mine) continue;
        if (st < nb) {
            // ---- forward:  y = r_q - E_{q-1}^T z_{q-1},  z_q = P_q y ----
            const int q = st, o = off[q], n = off[q + 1] - o;
            const int op = q > 0 ? off[q - 1] : 0, np_ = q > 0 ? o - op : 0;
            for (int i = lane; i < n; i += 32) {
                float4 acc = v[o + i];
                for (int k = 0; k < np_; ++k) fma4(acc, -E[k * ld + i], v[op + k]);
                u[i] = acc;
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) {              // P symmetric: column walk = consecutive lanes
                float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                for (int k = 0; k < n; ++k) fma4(acc, P[k * ld + i], u[k]);
                v[o + i] = acc;
            }
            __syncwarp();
        } else {
            // ---- backward:  x_q = z_q - P_q (E_q x_{q+1}) ----
            const int q = 2 * nb - 2 - st, o = off[q], n = off[q + 1] - o;
            const int on = off[q + 1], nn = off[q + 2] - on;
            for (int i = lane; i < n; i += 32) {
                float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                for (int k = 0; k < nn; ++k) fma4(acc, E[i * ld + k], v[on + k]);
                u[i] = acc;
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) {
                float4 acc = v[o + i];
                for (int k = 0; k < n; ++k) fma4(acc, -P[k * ld + i], u[k]);
                v[o + i] = acc;
            }
            __syncwarp();
        }
    }
    cp_async_wait<0>();
    if (mine) {
        double* tp[LBR];
#pragma unroll
        for (int c = 0; c < LBR; ++c)
            tp[c] = c0 + c < na ? reinterpret_cast<double*>(blk_vec(T, b, c0 + c, N)) : nullptr;
        for (int j = lane; j < ns; j += 32) {
            const int i = pm ? pm[j] : j;
            const int idx = s[i];
            const float4 w = v[i];
            if (tp[0]) tp[0][idx] = (double)w.x;
            if (tp[1]) tp[1][idx] = (double)w.y;
            if (tp[2]) tp[2][idx] = (double)w.z;
            if (tp[3]) tp[3][idx] = (double)w.w;
        }
    }
}

extern "C" int sqc_lowband_apply_real(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* theta,
                                        int32_t ld_theta, const int32_t* act, int32_t ld_act, const int32_t* S,
                                        int64_t s_stride_b, const void* Afac, int32_t ns, const int32_t* blk_off,
                                        int32_t ld_off, const int32_t* nblk, int32_t nbmax, double unit,
                                        int32_t group, const int32_t* perm, int64_t N, int32_t B, void* stream) {
    if (B <= 0 || ns <= 0 || T.cnt_fixed <= 0 || T.cnt_fixed > 16 || group < 1 || nbmax <= 0) return SQC_EINVAL;
    const int nwarp = (T.cnt_fixed + LBR - 1) / LBR;
    const size_t smem = ((size_t)nwarp * ns + (size_t)nwarp * nbmax) * sizeof(float4) +
                        (size_t)4 * nbmax * (nbmax | 1) * sizeof(float);
    if (smem > 227 * 1024) return SQC_ELIMIT;
    cudaError_t e = cudaFuncSetAttribute(lowband_apply_real_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    lowband_apply_real_kernel<<<B, 32 * nwarp, smem, (cudaStream_t)stream>>>(
        X, HX, T, theta, ld_theta, act, ld_act, S, s_stride_b, (const float*)Afac, ns, blk_off, ld_off, nblk, nbmax,
        1.0 / unit, group, N, perm);
    SQC_LAUNCHED();
    return 0;
}
