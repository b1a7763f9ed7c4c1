// Operator construction and matrix-free operator application.
//   H X = d_LC .* X + (x_m V_m) [ P ( (x_m V_m^T) X ) ]
// with V_m the x-basis of harmonic mode m, P the x/charge-basis potential (junction cosines:
// diagonal phases times integer shifts on the charge modes; external-flux inductor terms:
// diagonal).  Reference: Circuit._get_LC_hamil / _get_inductive_hamil / hamiltonian
// (circuit.py:1731-1829, 2147-2164), which builds the same operator as a scipy CSR matrix.
#include <math.h>
#include "common.cuh"

struct SpaceDev {
    int n_modes;
    int dims[SQC_MAX_MODES];
    int stride[SQC_MAX_MODES];   // element stride of each mode (last mode fastest)
    int harm[SQC_MAX_MODES];
    int toff[SQC_MAX_MODES];     // offset of the mode in concatenated per-mode tables
    int tlen;
    long long N;
};

static int make_space(const sqc_space_t* sp, SpaceDev* d) {
    if (!sp || sp->n_modes <= 0 || sp->n_modes > SQC_MAX_MODES) return SQC_EINVAL;
    d->n_modes = sp->n_modes;
    long long N = 1;
    for (int m = sp->n_modes - 1; m >= 0; --m) {
        if (sp->dims[m] <= 0) return SQC_EINVAL;
        d->dims[m] = sp->dims[m];
        d->harm[m] = sp->harmonic[m];
        d->stride[m] = (int)N;
        N *= sp->dims[m];
        if (N > 0x7fffffffLL) return SQC_ELIMIT;
    }
    int off = 0;
    for (int m = 0; m < sp->n_modes; ++m) {
        d->toff[m] = off;
        off += sp->dims[m];
    }
    d->tlen = off;
    d->N = N;
    return 0;
}

// ---------------------------------------------------------------------------------------
// (1) operator construction
// ---------------------------------------------------------------------------------------
__global__ void phase_tables_kernel(SpaceDev sp, int n_terms, const double* __restrict__ lam,
                                    const double* __restrict__ s, long long s_stride_b,
                                    cplx* __restrict__ etab, long long etab_stride_b) {
    const int b = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_terms * sp.tlen) return;
    const int t = e / sp.tlen, pos = e % sp.tlen;
    int m = 0;
    while (m + 1 < sp.n_modes && pos >= sp.toff[m + 1]) ++m;
    cplx v = cmake(1.0, 0.0);
    if (sp.harm[m]) {
        const double ang = s[b * s_stride_b + t * sp.n_modes + m] * lam[pos];
        double sn, cs;
        sincos(ang, &sn, &cs);
        v = cmake(cs, sn);
    }
    etab[b * etab_stride_b + (long long)t * sp.tlen + pos] = v;
}

extern "C" int sqc_phase_tables(const sqc_space_t* sp, int32_t n_terms, const double* lam,
                                const double* s, int64_t s_stride_b, void* etab,
                                int64_t etab_stride_b, int32_t B, void* stream) {
    SpaceDev d;
    int rc = make_space(sp, &d);
    if (rc) return rc;
    if (n_terms <= 0 || n_terms > SQC_MAX_TERMS || B <= 0) return SQC_EINVAL;
    dim3 grid((n_terms * d.tlen + 127) / 128, B);
    phase_tables_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(d, n_terms, lam, s, s_stride_b,
                                                              (cplx*)etab, etab_stride_b);
    SQC_LAUNCHED();
    return 0;
}

__global__ void lc_diagonal_kernel(SpaceDev sp, const double* __restrict__ omega,
                                   const double* __restrict__ q, const double* __restrict__ ng,
                                   long long pstride, double* __restrict__ out) {
    const int b = blockIdx.y;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= sp.N) return;
    const int M = sp.n_modes;
    double val[SQC_MAX_MODES];
    long long r = idx;
#pragma unroll
    for (int m = 0; m < SQC_MAX_MODES; ++m) {
        if (m < M) {
            const int im = (int)(r / sp.stride[m]);
            r -= (long long)im * sp.stride[m];
            // charge modes: integer charge n - ng with n = -(dim-1)/2 .. (dim-1)/2
            val[m] = sp.harm[m] ? (double)im
                                : (double)(im - (sp.dims[m] - 1) / 2) - ng[b * pstride * M + m];
        }
    }
    double acc = 0.0;
    for (int m = 0; m < M; ++m) {
        if (sp.harm[m]) {
            acc += omega[b * pstride * M + m] * val[m];
        } else {
            for (int m2 = m; m2 < M; ++m2)
                if (!sp.harm[m2]) acc += q[b * pstride * M * M + m * M + m2] * val[m] * val[m2];
        }
    }
    out[b * sp.N + idx] = acc;
}

extern "C" int sqc_lc_diagonal(const sqc_space_t* sp, const double* omega, const double* q,
                               const double* ng, int64_t param_stride_b, double* out,
                               int32_t B_out, void* stream) {
    SpaceDev d;
    int rc = make_space(sp, &d);
    if (rc) return rc;
    if (B_out <= 0) return SQC_EINVAL;
    dim3 grid((unsigned)((d.N + 255) / 256), B_out);
    lc_diagonal_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d, omega, q, ng, param_stride_b, out);
    SQC_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------
// (2a) mode-wise contraction, plain FP64 SIMT version
//   tile = all m rows x TC columns, where a "column" is one (vector, outer, inner) position
// ---------------------------------------------------------------------------------------
#define MT_THREADS 256
#define MT_TC 32
#define MT_RI 16   // rows per thread (supports m <= 8 * 16 = 128)

__global__ void __launch_bounds__(MT_THREADS)
mode_transform_simt_kernel(int m, long long inner, long long outer, const double* __restrict__ Mat,
                           long long mat_stride_b, int transpose, sqc_block_t X, sqc_block_t Z,
                           const double* __restrict__ ediag, long long ediag_stride_b,
                           sqc_block_t EX, long long N, int vmax, long long tiles_per_vec) {
    extern __shared__ double smem[];
    double* sM = smem;                                    // [k][i]  (m x m)
    cplx* sX = reinterpret_cast<cplx*>(smem + m * m + (m * m & 1));  // [k][MT_TC]
    const long long vec = blockIdx.x / tiles_per_vec;
    const long long tile = blockIdx.x % tiles_per_vec;
    const int b = (int)(vec / vmax), v = (int)(vec % vmax);
    if (v >= blk_count(X, b)) return;
    const int tid = threadIdx.x;
    const double* mat = Mat + b * mat_stride_b;
    for (int e = tid; e < m * m; e += MT_THREADS) {
        const int k = e / m, i = e % m;
        sM[e] = transpose ? mat[k * m + i] : mat[i * m + k];
    }
    // columns of this tile: c in [c0, c0 + MT_TC) over (outer, inner) flattened
    const long long c0 = tile * MT_TC;
    const long long ncols = outer * inner;
    const cplx* xv = blk_vec(X, b, v, N);
    for (int e = tid; e < m * MT_TC; e += MT_THREADS) {
        const int k = e / MT_TC, c = e % MT_TC;
        const long long col = c0 + c;
        cplx val = cmake(0.0, 0.0);
        if (col < ncols) {
            const long long o = col / inner, r = col % inner;
            val = xv[(o * m + k) * inner + r];
        }
        sX[e] = val;
    }
    __syncthreads();
    const int c = tid % MT_TC, rg = tid / MT_TC;   // 8 row groups
    const int ri = (m + 7) / 8;
    const int i0 = rg * ri;
    cplx acc[MT_RI];
#pragma unroll
    for (int ii = 0; ii < MT_RI; ++ii) acc[ii] = cmake(0.0, 0.0);
    for (int k = 0; k < m; ++k) {
        const cplx x = sX[k * MT_TC + c];
        const double* mr = sM + k * m + i0;
#pragma unroll
        for (int ii = 0; ii < MT_RI; ++ii) {
            if (ii < ri && i0 + ii < m) {
                const double w = mr[ii];
                acc[ii].x = fma(w, x.x, acc[ii].x);
                acc[ii].y = fma(w, x.y, acc[ii].y);
            }
        }
    }
    const long long col = c0 + c;
    if (col >= ncols) return;
    const long long o = col / inner, r = col % inner;
    cplx* zv = blk_vec(Z, b, v, N);
    const cplx* ex = ediag ? blk_vec(EX, b, v, N) : nullptr;
#pragma unroll
    for (int ii = 0; ii < MT_RI; ++ii) {
        if (ii < ri && i0 + ii < m) {
            const long long idx = (o * m + (i0 + ii)) * inner + r;
            cplx out = acc[ii];
            if (ediag) {
                const double dv = ediag[b * ediag_stride_b + idx];
                const cplx e = ex[idx];
                out.x = fma(dv, e.x, out.x);
                out.y = fma(dv, e.y, out.y);
            }
            zv[idx] = out;
        }
    }
}

// ---------------------------------------------------------------------------------------
// (2a') mode-wise contraction on the FP64 tensor cores (DMMA m8n8k4).
// The complex tile is a real matrix with twice the columns (Mat is real), so
//   Z_r[m x 2TC] = Mat[m x m] . X_r[m x 2TC]
// Warp w owns row tiles {w, w+8}; every warp sweeps all column tiles.  The matrix and the
// staged panel are read as fragments from padded shared memory (row stride == 4 mod 16
// doubles keeps both fragment patterns bank-conflict free).
// ---------------------------------------------------------------------------------------
#define MD_THREADS 256
#define MD_TC 32                         // complex columns per panel -> 64 real columns
#define MD_XLD (2 * MD_TC + 4)           // 68

__global__ void __launch_bounds__(MD_THREADS)
mode_transform_dmma_kernel(int m, int mp, int mld, long long inner, long long outer,
                           const double* __restrict__ Mat, long long mat_stride_b, int transpose,
                           sqc_block_t X, sqc_block_t Z, const double* __restrict__ ediag,
                           long long ediag_stride_b, sqc_block_t EX, long long N, int vmax,
                           long long tiles_per_vec, long long total_items) {
    extern __shared__ double smem[];
    double* sM = smem;                 // [i][k]  mp x mld   (A operand, row-major)
    double* sX = smem + mp * mld;      // [k][2*MD_TC] mp x MD_XLD  (B operand: K x N)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const long long ncols = outer * inner;
    const int nrt = mp / 8;            // row tiles (<= 16)
    int loaded_b = -1;
    // persistent CTA: the matrix is staged once (per sweep point when it is point dependent)
    for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {
        const long long vec = item / tiles_per_vec;
        const long long tile = item % tiles_per_vec;
        const int b = (int)(vec / vmax), v = (int)(vec % vmax);
        if (v >= blk_count(X, b)) continue;   // uniform for the CTA
        if (loaded_b < 0 || (mat_stride_b != 0 && b != loaded_b)) {
            const double* mat = Mat + b * mat_stride_b;
            for (int e = tid; e < mp * mld; e += MD_THREADS) {
                const int i = e / mld, k = e % mld;
                double val = 0.0;
                if (i < m && k < m) val = transpose ? mat[k * m + i] : mat[i * m + k];
                sM[e] = val;
            }
            loaded_b = b;
        }
        const long long c0 = tile * MD_TC;
        const cplx* xv = blk_vec(X, b, v, N);
        for (int e = tid; e < mp * MD_TC; e += MD_THREADS) {
            const int k = e / MD_TC, c = e % MD_TC;
            const long long col = c0 + c;
            cplx val = cmake(0.0, 0.0);
            if (col < ncols && k < m) {
                const long long o = col / inner, r = col % inner;
                val = xv[(o * m + k) * inner + r];
            }
            sX[k * MD_XLD + 2 * c] = val.x;
            sX[k * MD_XLD + 2 * c + 1] = val.y;
        }
        __syncthreads();
        cplx* zv = blk_vec(Z, b, v, N);
        const cplx* ex = ediag ? blk_vec(EX, b, v, N) : nullptr;
        for (int rt = warp; rt < nrt; rt += 8) {
            double acc[2 * MD_TC / 8][2];
#pragma unroll
            for (int ct = 0; ct < 2 * MD_TC / 8; ++ct) acc[ct][0] = acc[ct][1] = 0.0;
            for (int k0 = 0; k0 < mp; k0 += 4) {
                const double af = sM[(8 * rt + g) * mld + k0 + t];
#pragma unroll
                for (int ct = 0; ct < 2 * MD_TC / 8; ++ct) {
                    const double bf = sX[(k0 + t) * MD_XLD + 8 * ct + g];
                    dmma884(acc[ct][0], acc[ct][1], af, bf);
                }
            }
            // C fragment: row 8rt+g, real columns 8ct+2t, 8ct+2t+1  == complex column 4ct+t
            const int i = 8 * rt + g;
            if (i < m) {
#pragma unroll
                for (int ct = 0; ct < 2 * MD_TC / 8; ++ct) {
                    const long long col = c0 + 4 * ct + t;
                    if (col < ncols) {
                        const long long o = col / inner, r = col % inner;
                        const long long idx = (o * m + i) * inner + r;
                        cplx out = cmake(acc[ct][0], acc[ct][1]);
                        if (ediag) {
                            const double dv = ediag[b * ediag_stride_b + idx];
                            const cplx e = ex[idx];
                            out.x = fma(dv, e.x, out.x);
                            out.y = fma(dv, e.y, out.y);
                        }
                        zv[idx] = out;
                    }
                }
            }
        }
        __syncthreads();   // panel is re-staged by the next item
    }
}

extern "C" int sqc_mode_transform(const sqc_space_t* sp, int32_t mode, const double* Mat,
                                  int64_t mat_stride_b, int32_t transpose, sqc_block_t X,
                                  sqc_block_t Z, const double* ediag, int64_t ediag_stride_b,
                                  sqc_block_t EX, int32_t B, int32_t impl, void* stream) {
    SpaceDev d;
    int rc = make_space(sp, &d);
    if (rc) return rc;
    if (mode < 0 || mode >= d.n_modes || B <= 0 || X.cnt_fixed <= 0) return SQC_EINVAL;
    const int m = d.dims[mode];
    if (m > 128) return SQC_ELIMIT;
    const long long inner = d.stride[mode];
    const long long outer = d.N / (inner * m);
    const long long ncols = outer * inner;
    cudaStream_t st = (cudaStream_t)stream;
    if (impl == 1) {
        const long long tiles = (ncols + MT_TC - 1) / MT_TC;
        const long long grid = (long long)B * X.cnt_fixed * tiles;
        if (grid > 0x7fffffffLL) return SQC_ELIMIT;
        size_t smem = ((size_t)m * m + ((m * m) & 1)) * sizeof(double) + (size_t)m * MT_TC * sizeof(cplx);
        cudaError_t e = cudaFuncSetAttribute(mode_transform_simt_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        mode_transform_simt_kernel<<<(unsigned)grid, MT_THREADS, smem, st>>>(
            m, inner, outer, Mat, mat_stride_b, transpose, X, Z, ediag, ediag_stride_b, EX, d.N,
            X.cnt_fixed, tiles);
        SQC_LAUNCHED();
        return 0;
    }
    const int mp = (m + 7) & ~7;
    int mld = mp;
    while ((mld & 15) != 4) ++mld;     // row stride == 4 (mod 16) doubles
    const long long tiles = (ncols + MD_TC - 1) / MD_TC;
    const long long total = (long long)B * X.cnt_fixed * tiles;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    size_t smem = ((size_t)mp * mld + (size_t)mp * MD_XLD) * sizeof(double);
    const int per_sm = smem <= 72 * 1024 ? 3 : (smem <= 110 * 1024 ? 2 : 1);
    long long grid = (long long)nsm * per_sm;
    if (grid > total) grid = total;
    cudaError_t e = cudaFuncSetAttribute(mode_transform_dmma_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    mode_transform_dmma_kernel<<<(unsigned)grid, MD_THREADS, smem, st>>>(
        m, mp, mld, inner, outer, Mat, mat_stride_b, transpose, X, Z, ediag, ediag_stride_b, EX, d.N,
        X.cnt_fixed, tiles, total);
    SQC_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------
// (2b) x/charge-basis potential
// ---------------------------------------------------------------------------------------
struct PotDev {
    int n_terms;
    int shift[SQC_MAX_TERMS][SQC_MAX_MODES];
    int lin_shift[SQC_MAX_TERMS];     // linear index offset of the shift
    const double* lam;
    const cplx* etab;
    long long etab_stride_b;
    const double* g;
    const cplx* a;
    const double* diag;
    long long diag_stride_b;
};

__global__ void __launch_bounds__(256)
potential_kernel(SpaceDev sp, PotDev pt, sqc_block_t X, sqc_block_t Z, int vmax, int nchunk) {
    const long long vec = blockIdx.x / nchunk;
    const int b = (int)(vec / vmax), v = (int)(vec % vmax);
    if (v >= blk_count(X, b)) return;
    const long long idx = (long long)(blockIdx.x % nchunk) * 256 + threadIdx.x;
    if (idx >= sp.N) return;
    const int M = sp.n_modes;
    int im[SQC_MAX_MODES];
    int r = (int)idx;                      // N < 2^31 (make_space): 32-bit divisions, not the emulated 64-bit ones
#pragma unroll
    for (int m = 0; m < SQC_MAX_MODES; ++m) {
        if (m < M) {
            im[m] = r / sp.stride[m];
            r -= im[m] * sp.stride[m];
        }
    }
    const cplx* x = blk_vec(X, b, v, sp.N);
    const double* gb = pt.g + (long long)b * (1 + M);
    double lin = gb[0];
    for (int m = 0; m < M; ++m)
        if (sp.harm[m]) lin = fma(gb[1 + m], pt.lam[sp.toff[m] + im[m]], lin);
    if (pt.diag) lin += pt.diag[b * pt.diag_stride_b + idx];
    const cplx x0 = x[idx];
    cplx acc = cmake(lin * x0.x, lin * x0.y);
    const cplx* et = pt.etab + b * pt.etab_stride_b;
    for (int t = 0; t < pt.n_terms; ++t) {
        cplx ph = pt.a[(long long)b * pt.n_terms + t];
        bool lo = true, hi = true;   // idx - shift / idx + shift inside the truncated space
        for (int m = 0; m < M; ++m) {
            if (sp.harm[m]) {
                ph = cmul(ph, et[(long long)t * sp.tlen + sp.toff[m] + im[m]]);
            } else {
                const int w = pt.shift[t][m];
                lo = lo && (im[m] - w >= 0) && (im[m] - w < sp.dims[m]);
                hi = hi && (im[m] + w >= 0) && (im[m] + w < sp.dims[m]);
            }
        }
        if (lo) cfma(acc, ph, x[idx - pt.lin_shift[t]]);
        if (hi) cfma(acc, cconj(ph), x[idx + pt.lin_shift[t]]);
    }
    blk_vec(Z, b, v, sp.N)[idx] = acc;
}

extern "C" int sqc_potential_apply(const sqc_space_t* sp, const sqc_potential_t* pot, sqc_block_t X,
                                   sqc_block_t Z, int32_t B, void* stream) {
    SpaceDev d;
    int rc = make_space(sp, &d);
    if (rc) return rc;
    if (!pot || B <= 0 || X.cnt_fixed <= 0 || pot->n_terms < 0 || pot->n_terms > SQC_MAX_TERMS)
        return SQC_EINVAL;
    if (X.ptr == Z.ptr) return SQC_EINVAL;   // reads neighbours: not in-place safe
    PotDev pd;
    pd.n_terms = pot->n_terms;
    for (int t = 0; t < pot->n_terms; ++t) {
        int ls = 0;
        for (int m = 0; m < d.n_modes; ++m) {
            pd.shift[t][m] = d.harm[m] ? 0 : pot->shift[t][m];
            ls += pd.shift[t][m] * d.stride[m];
        }
        pd.lin_shift[t] = ls;
    }
    pd.lam = pot->lam;
    pd.etab = (const cplx*)pot->etab;
    pd.etab_stride_b = pot->etab_stride_b;
    pd.g = pot->g;
    pd.a = (const cplx*)pot->a;
    pd.diag = pot->diag;
    pd.diag_stride_b = pot->diag_stride_b;
    const int nchunk = (int)((d.N + 255) / 256);
    const long long grid = (long long)B * X.cnt_fixed * nchunk;
    if (grid > 0x7fffffffLL) return SQC_ELIMIT;
    potential_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(d, pd, X, Z, X.cnt_fixed, nchunk);
    SQC_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------
// (2b') matrix elements of a potential-type operator between a few x-basis vectors per point
// ---------------------------------------------------------------------------------------
// out[b][i][j] = sum_idx conj(z_i[idx]) (P z_j)[idx] for the nv <= 4 vectors of every point, P as in
// sqc_potential_apply.  One pass over the vectors, nothing written but nv^2 numbers per point: the decoherence
// channels of a T1 / Tphi query (circuit.py:2529-2647) each need the 2 x 2 matrix of one such operator between the
// same pair of eigenstates, rotated to the x basis once (F^T P F between a and b = P between F a and F b).
// Block = POTEL_CHUNK consecutive indices of one point; per-chunk partial sums in `work`, then a fixed-order
// reduction over the chunks (deterministic).
#define POTEL_THREADS 256
#define POTEL_PER_THREAD 4
#define POTEL_CHUNK (POTEL_THREADS * POTEL_PER_THREAD)
#define POTEL_MAXV 4

template <int NV>
__global__ void __launch_bounds__(POTEL_THREADS)
potential_elements_kernel(SpaceDev sp, PotDev pt, sqc_block_t Z, int nchunk, cplx* __restrict__ work) {
    const int b = blockIdx.x / nchunk, ch = blockIdx.x - b * nchunk;
    const int M = sp.n_modes;
    const cplx* z[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) z[v] = blk_vec(Z, b, v, sp.N);
    const double* gb = pt.g + (long long)b * (1 + M);
    const cplx* et = pt.etab + b * pt.etab_stride_b;
    cplx acc[NV][NV];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < NV; ++j) acc[i][j] = cmake(0.0, 0.0);
    for (int q = 0; q < POTEL_PER_THREAD; ++q) {
        const long long idx = (long long)ch * POTEL_CHUNK + q * POTEL_THREADS + threadIdx.x;
        if (idx >= sp.N) break;
        int im[SQC_MAX_MODES];
        int r = (int)idx;                  // N < 2^31 (make_space): 32-bit divisions, not the emulated 64-bit ones
#pragma unroll
        for (int m = 0; m < SQC_MAX_MODES; ++m) {
            if (m < M) {
                im[m] = r / sp.stride[m];
                r -= im[m] * sp.stride[m];
            }
        }
        double lin = gb[0];
        for (int m = 0; m < M; ++m)
            if (sp.harm[m]) lin = fma(gb[1 + m], pt.lam[sp.toff[m] + im[m]], lin);
        if (pt.diag) lin += pt.diag[b * pt.diag_stride_b + idx];
        cplx pz[NV], zc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            zc[v] = z[v][idx];
            pz[v] = cmake(lin * zc[v].x, lin * zc[v].y);
        }
        for (int t = 0; t < pt.n_terms; ++t) {
            cplx ph = pt.a[(long long)b * pt.n_terms + t];
            bool lo = true, hi = true;
            for (int m = 0; m < M; ++m) {
                if (sp.harm[m]) {
                    ph = cmul(ph, et[(long long)t * sp.tlen + sp.toff[m] + im[m]]);
                } else {
                    const int w = pt.shift[t][m];
                    lo = lo && (im[m] - w >= 0) && (im[m] - w < sp.dims[m]);
                    hi = hi && (im[m] + w >= 0) && (im[m] + w < sp.dims[m]);
                }
            }
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                if (lo) cfma(pz[v], ph, z[v][idx - pt.lin_shift[t]]);
                if (hi) cfma(pz[v], cconj(ph), z[v][idx + pt.lin_shift[t]]);
            }
        }
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < NV; ++j) cfmac(acc[i][j], zc[i], pz[j]);
    }
    // block reduction: warp shuffles, then one partial per warp through shared memory, fixed order
    __shared__ double red[POTEL_THREADS / 32][2 * NV * NV];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const double re = warp_sum(acc[i][j].x), imv = warp_sum(acc[i][j].y);
            if (lane == 0) {
                red[wp][2 * (i * NV + j)] = re;
                red[wp][2 * (i * NV + j) + 1] = imv;
            }
        }
    __syncthreads();
    if (threadIdx.x < 2 * NV * NV) {
        double sum = 0.0;
#pragma unroll
        for (int w_ = 0; w_ < POTEL_THREADS / 32; ++w_) sum += red[w_][threadIdx.x];
        reinterpret_cast<double*>(work)[((long long)blockIdx.x) * 2 * NV * NV + threadIdx.x] = sum;
    }
}

__global__ void potential_elements_reduce_kernel(const cplx* __restrict__ work, int nchunk, int nn, cplx* __restrict__ out) {
    const int b = blockIdx.x, e = threadIdx.x;
    if (e >= nn) return;
    double re = 0.0, imv = 0.0;
    for (int c = 0; c < nchunk; ++c) {
        const cplx v = work[((long long)b * nchunk + c) * nn + e];
        re += v.x;
        imv += v.y;
    }
    out[(long long)b * nn + e] = cmake(re, imv);
}

extern "C" int64_t sqc_potential_elements_work_bytes(int64_t N, int32_t nv, int32_t B) {
    if (N <= 0 || nv <= 0 || nv > POTEL_MAXV || B <= 0) return 0;
    const int64_t nchunk = (N + POTEL_CHUNK - 1) / POTEL_CHUNK;
    return (int64_t)B * nchunk * nv * nv * (int64_t)sizeof(cplx);
}

extern "C" int sqc_potential_elements(const sqc_space_t* sp, const sqc_potential_t* pot, sqc_block_t Z, int32_t nv,
                                      void* out, void* work, int32_t B, void* stream) {
    SpaceDev d;
    int rc = make_space(sp, &d);
    if (rc) return rc;
    if (!pot || !out || !work || B <= 0 || nv <= 0 || nv > POTEL_MAXV || Z.cnt_fixed < nv || Z.cnt ||
        pot->n_terms < 0 || pot->n_terms > SQC_MAX_TERMS)
        return SQC_EINVAL;
    PotDev pd;
    pd.n_terms = pot->n_terms;
    for (int t = 0; t < pot->n_terms; ++t) {
        int ls = 0;
        for (int m = 0; m < d.n_modes; ++m) {
            pd.shift[t][m] = d.harm[m] ? 0 : pot->shift[t][m];
            ls += pd.shift[t][m] * d.stride[m];
        }
        pd.lin_shift[t] = ls;
    }
    pd.lam = pot->lam;
    pd.etab = (const cplx*)pot->etab;
    pd.etab_stride_b = pot->etab_stride_b;
    pd.g = pot->g;
    pd.a = (const cplx*)pot->a;
    pd.diag = pot->diag;
    pd.diag_stride_b = pot->diag_stride_b;
    const int nchunk = (int)((d.N + POTEL_CHUNK - 1) / POTEL_CHUNK);
    const long long grid = (long long)B * nchunk;
    if (grid > 0x7fffffffLL) return SQC_ELIMIT;
    cudaStream_t st = (cudaStream_t)stream;
    switch (nv) {
        case 1: potential_elements_kernel<1><<<(unsigned)grid, POTEL_THREADS, 0, st>>>(d, pd, Z, nchunk, (cplx*)work); break;
        case 2: potential_elements_kernel<2><<<(unsigned)grid, POTEL_THREADS, 0, st>>>(d, pd, Z, nchunk, (cplx*)work); break;
        case 3: potential_elements_kernel<3><<<(unsigned)grid, POTEL_THREADS, 0, st>>>(d, pd, Z, nchunk, (cplx*)work); break;
        default: potential_elements_kernel<4><<<(unsigned)grid, POTEL_THREADS, 0, st>>>(d, pd, Z, nchunk, (cplx*)work); break;
    }
    SQC_LAUNCHED();
    potential_elements_reduce_kernel<<<B, 32, 0, st>>>((const cplx*)work, nchunk, nv * nv, (cplx*)out);
    SQC_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------
// (2c) ladder operators of a harmonic mode
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ladder_kernel(SpaceDev sp, int mode, const cplx* __restrict__ c_dn, const cplx* __restrict__ c_up,
              sqc_block_t X, sqc_block_t Y, int vmax, int nchunk, int accumulate) {
    const long long vec = blockIdx.x / nchunk;
    const int b = (int)(vec / vmax), v = (int)(vec % vmax);
    if (v >= blk_count(X, b)) return;
    const long long idx = (long long)(blockIdx.x % nchunk) * 256 + threadIdx.x;
    if (idx >= sp.N) return;
    const int st = sp.stride[mode], m = sp.dims[mode];
    const int i = (int)((idx / st) % m);
    const cplx* x = blk_vec(X, b, v, sp.N);
    cplx acc = cmake(0.0, 0.0);
    if (i + 1 < m) {   // (a x)[i] = sqrt(i+1) x[i+1]
        const cplx xv = x[idx + st];
        cfma(acc, cscale(sqrt((double)(i + 1)), c_dn[b]), xv);
    }
    if (i > 0) {       // (a^dag x)[i] = sqrt(i) x[i-1]
        const cplx xv = x[idx - st];
        cfma(acc, cscale(sqrt((double)i), c_up[b]), xv);
    }
    cplx* y = blk_vec(Y, b, v, sp.N) + idx;
    if (accumulate) acc = cadd(acc, *y);
    *y = acc;
}

extern "C" int sqc_ladder_apply(const sqc_space_t* sp, int32_t mode, const void* c_dn,
                                const void* c_up, sqc_block_t X, sqc_block_t Y, int32_t B,
                                int32_t accumulate, void* stream) {
    SpaceDev d;
    int rc = make_space(sp, &d);
    if (rc) return rc;
    if (mode < 0 || mode >= d.n_modes || B <= 0 || X.cnt_fixed <= 0 || X.ptr == Y.ptr) return SQC_EINVAL;
    const int nchunk = (int)((d.N + 255) / 256);
    const long long grid = (long long)B * X.cnt_fixed * nchunk;
    if (grid > 0x7fffffffLL) return SQC_ELIMIT;
    ladder_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        d, mode, (const cplx*)c_dn, (const cplx*)c_up, X, Y, X.cnt_fixed, nchunk, accumulate);
    SQC_LAUNCHED();
    return 0;
}
