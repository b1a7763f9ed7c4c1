// Real (time-reversal adapted) representation of harmonic x charge circuits at zero gate charge
// (see hx_real.cu): conversions to / from the complex Fock x charge basis of the reference
// (circuit.py:1938-1986 returns eigenvectors in that basis) and the element-wise Davidson kernels
// that need to know that a "complex" element of a block is really two independent real amplitudes.
//
// A real vector has N = m * inner doubles y[i][j] (planar layout: j = 0 charge 0, 1 <= j <= nc
// sqrt2 Re z(i, j), nc < j <= 2 nc sqrt2 Im z(i, j - nc)); it is stored in a complex128 block with
// Nc = ceil(N / 2) elements per vector, the padding double of an odd N stays zero.
#include "common.cuh"

// Zc[b][v][i * inner + nc +- n] from Xr[b][v]  (one thread per (i, n >= 0))
__global__ void __launch_bounds__(256)
real_to_complex_kernel(sqc_block_t Xr, sqc_block_t Zc, int m, int inner, int vmax, int nchunk) {
    const int vec = blockIdx.x / nchunk;
    const int b = vec / vmax, v = vec % vmax;
    if (v >= blk_count(Xr, b)) return;
    const int nc = (inner - 1) / 2;
    const int e = (blockIdx.x % nchunk) * 256 + threadIdx.x;
    if (e >= m * (nc + 1)) return;
    const int i = e / (nc + 1), n = e % (nc + 1);
    const long long N = (long long)m * inner, Ncx = (N + 1) / 2;
    const double* y = reinterpret_cast<const double*>(blk_vec(Xr, b, v, Ncx)) + (long long)i * inner;
    cplx* z = blk_vec(Zc, b, v, N) + (long long)i * inner + nc;
    if (n == 0) {
        z[0] = cmake(y[0], 0.0);
    } else {
        const double rs2 = 0.70710678118654752440;
        const double re = rs2 * y[n], im = rs2 * y[nc + n];
        z[n] = cmake(re, im);
        z[-n] = cmake(re, -im);
    }
}

// Xr[b][v] = T-symmetric part of Zc[b][v] in real coordinates:  y = sqrt2 * (z(n) + conj z(-n)) / 2
__global__ void __launch_bounds__(256)
complex_to_real_kernel(sqc_block_t Zc, sqc_block_t Xr, int m, int inner, int vmax, int nchunk) {
    const int vec = blockIdx.x / nchunk;
    const int b = vec / vmax, v = vec % vmax;
    if (v >= blk_count(Zc, b)) return;
    const int nc = (inner - 1) / 2;
    const int e = (blockIdx.x % nchunk) * 256 + threadIdx.x;
    const long long N = (long long)m * inner, Ncx = (N + 1) / 2;
    double* yv = reinterpret_cast<double*>(blk_vec(Xr, b, v, Ncx));
    if (e == 0 && (N & 1)) yv[N] = 0.0;
    if (e >= m * (nc + 1)) return;
    const int i = e / (nc + 1), n = e % (nc + 1);
    double* y = yv + (long long)i * inner;
    const cplx* z = blk_vec(Zc, b, v, N) + (long long)i * inner + nc;
    if (n == 0) {
        y[0] = z[0].x;
    } else {
        const double rs2 = 0.70710678118654752440;
        const cplx p = z[n], q = z[-n];
        y[n] = rs2 * (p.x + q.x);
        y[nc + n] = rs2 * (p.y - q.y);
    }
}

static int conv_args(const sqc_space_t* sp, int32_t B, const sqc_block_t& src, int* m, int* inner, int* nchunk,
                     int64_t* grid) {
    if (!sp || B <= 0 || src.cnt_fixed <= 0) return SQC_EINVAL;
    if (sp->n_modes != 2 || !sp->harmonic[0] || sp->harmonic[1] || !(sp->dims[1] & 1)) return SQC_ELIMIT;
    *m = sp->dims[0];
    *inner = sp->dims[1];
    *nchunk = (*m * ((*inner + 1) / 2) + 255) / 256;
    *grid = (int64_t)B * src.cnt_fixed * *nchunk;
    return *grid > 0x7fffffffLL ? SQC_ELIMIT : 0;
}

extern "C" int sqc_real_to_complex(const sqc_space_t* sp, sqc_block_t Xr, sqc_block_t Zc, int32_t B, void* stream) {
    int m, inner, nchunk;
    int64_t grid;
    const int rc = conv_args(sp, B, Xr, &m, &inner, &nchunk, &grid);
    if (rc) return rc;
    real_to_complex_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(Xr, Zc, m, inner, Xr.cnt_fixed, nchunk);
    SQC_LAUNCHED();
    return 0;
}

extern "C" int sqc_complex_to_real(const sqc_space_t* sp, sqc_block_t Zc, sqc_block_t Xr, int32_t B, void* stream) {
    int m, inner, nchunk;
    int64_t grid;
    const int rc = conv_args(sp, B, Zc, &m, &inner, &nchunk, &grid);
    if (rc) return rc;
    complex_to_real_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(Zc, Xr, m, inner, Zc.cnt_fixed, nchunk);
    SQC_LAUNCHED();
    return 0;
}

// T[b][jj] = (HX_a - theta_a X_a) / clamp(d - theta_a) on REAL amplitudes: d has 2 N entries per sweep point
// (one per double of the block element), N = complex elements per vector.
__global__ void __launch_bounds__(256)
precond_residual_real_kernel(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* __restrict__ theta,
                             int ld_theta, const int32_t* __restrict__ act, int ld_act,
                             const double* __restrict__ d, int64_t d_stride_b,
                             const double* __restrict__ den_floor, int64_t N, int pmax, int nchunk) {
    const int vec = blockIdx.x / nchunk;
    const int b = vec / pmax, jj = vec % pmax;
    if (jj >= blk_count(T, b)) return;
    const int64_t n = (int64_t)(blockIdx.x % nchunk) * 256 + threadIdx.x;
    if (n >= N) return;
    const int a = act[(int64_t)b * ld_act + jj];
    const double th = theta[(int64_t)b * ld_theta + a];
    const double fl = den_floor[b];
    const cplx xv = blk_vec(X, b, a, N)[n], hv = blk_vec(HX, b, a, N)[n];
    const double2 dd = *reinterpret_cast<const double2*>(d + (int64_t)b * d_stride_b + 2 * n);
    double den0 = dd.x - th, den1 = dd.y - th;
    if (fabs(den0) < fl) den0 = (den0 < 0.0) ? -fl : fl;
    if (fabs(den1) < fl) den1 = (den1 < 0.0) ? -fl : fl;
    blk_vec(T, b, jj, N)[n] = cmake((hv.x - th * xv.x) / den0, (hv.y - th * xv.y) / den1);
}

extern "C" int sqc_precond_residual_real(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* theta,
                                         int32_t ld_theta, const int32_t* act, int32_t ld_act,
                                         const double* d, int64_t d_stride_b, const double* den_floor,
                                         int64_t N, int32_t B, void* stream) {
    if (B <= 0 || N <= 0 || T.cnt_fixed <= 0 || (d_stride_b & 1)) return SQC_EINVAL;
    const int nchunk = (int)((N + 255) / 256);
    const int64_t grid = (int64_t)B * T.cnt_fixed * nchunk;
    if (grid > 0x7fffffffLL) return SQC_ELIMIT;
    precond_residual_real_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        X, HX, T, theta, ld_theta, act, ld_act, d, d_stride_b, den_floor, N, T.cnt_fixed, nchunk);
    SQC_LAUNCHED();
    return 0;
}
