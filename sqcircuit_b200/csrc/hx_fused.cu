// Fused matrix-free H.X for circuits with one harmonic mode (mode 0) and one charge mode
// (mode 1), e.g. the 0-pi qubit:   Y = d .* X + V [ P ( V^T X ) ]   in ONE kernel.
//
// A CTA owns a panel of the state vector: all m Fock rows x PC charge columns (TC useful
// columns plus one halo column on each side for the junction's +-1 charge shift).  Per panel:
//   1. cp.async the panel into shared memory (even Fock rows first, odd rows second)
//   2. forward transform on the FP64 tensor cores using the parity of the x basis:
//        E = Ve^T X_even, O = Vo^T X_odd,  Z[i] = E+O, Z[m-1-i] = E-O     (half the flops)
//      each warp owns one 8-wide column tile, so the result is written back in place
//   3. potential in place: one warp per mirrored row pair, Z' = lin*Z + sum_t ph_t Z(c-s) + cc
//      and immediately S = Z'[i] + Z'[m-1-i], D = Z'[i] - Z'[m-1-i]
//   4. backward transform  Y_even = Ve S, Y_odd = Vo D, epilogue  + d .* X, store.
// HBM traffic: X is read once (+2 halo columns per panel), Y written once.
// Replaces the CSR mat-vec on the Hamiltonian of Circuit.hamiltonian() (circuit.py:2147-2164).
#include "common.cuh"

#define HX_MAXRT 8       // row tiles per parity half  -> m <= 128
#define HX_MAXT 4        // junction terms

struct HxGeom {
    int m, he, ho;       // Fock dim, #even rows, #odd rows
    int hep;             // he rounded up to 8 (rows of each half in shared memory)
    int kp;              // he rounded up to 4 (contraction length)
    int nrt;             // hep / 8
    int inner;           // charge-mode dimension (columns)
    int np, TC, PC, CT;  // panels per vector, useful cols per panel, cols incl. halo, col tiles
    int vld, pld;        // leading dims (doubles) of the half matrices and of the panel
    int n_terms;
    int shift[HX_MAXT];
    long long N;
};

template <int NRT>
__global__ void __launch_bounds__(288, 2)
hx_fused_kernel(HxGeom gm, const double* __restrict__ Ve, const double* __restrict__ Vo,
                const double* __restrict__ lam, const cplx* __restrict__ etab, long long etab_stride_b,
                const double* __restrict__ g, const cplx* __restrict__ a,
                const double* __restrict__ fdiag, long long fdiag_stride_b,
                sqc_block_t X, sqc_block_t Y, int vmax, long long total_items) {
    extern __shared__ double smem[];
    // rows: kp (= he rounded up to the k-step) per half; row tiles that stick out past kp are
    // masked when fragments are read / results written
    double* sVe = smem;                                   // [kp][vld]   Ve[a][i] = V[2a][i]
    double* sVo = sVe + gm.kp * gm.vld;                   // [kp][vld]   Vo[a][i] = V[2a+1][i]
    double* sP = sVo + gm.kp * gm.vld;                    // [2*kp][pld] panel: S/even half, D/odd half
    double* sLin = sP + 2 * gm.kp * gm.pld;               // [m]
    cplx* sPh = reinterpret_cast<cplx*>(sLin + ((gm.m + 1) & ~1));   // [n_terms][m]

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int nwarp = nthr >> 5;
    const int m = gm.m, he = gm.he, ho = gm.ho, hep = gm.kp, vld = gm.vld, pld = gm.pld;

    // half matrices (zero padded), staged once per CTA
    for (int e = tid; e < hep * vld; e += nthr) {
        const int r = e / vld, c = e % vld;
        sVe[e] = (r < he && c < he) ? Ve[r * he + c] : 0.0;
        sVo[e] = (r < ho && c < ho) ? Vo[r * ho + c] : 0.0;
    }
    int loaded_b = -1;
    for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {
        const long long vec = item / gm.np;
        const int panel = (int)(item % gm.np);
        const int b = (int)(vec / vmax), v = (int)(vec % vmax);
        if (v >= blk_count(X, b)) continue;
        if (b != loaded_b) {
            // per-sweep-point tables: lin(i) = g0 + g1 lam_i ;  ph_t(i) = a_t etab_t[i]
            const double g0 = g[(long long)b * 3], g1 = g[(long long)b * 3 + 1];
            for (int i = tid; i < m; i += nthr) sLin[i] = fma(g1, lam[i], g0);
            for (int e = tid; e < gm.n_terms * m; e += nthr) {
                const int t = e / m, i = e % m;
                sPh[e] = cmul(a[(long long)b * gm.n_terms + t],
                              etab[b * etab_stride_b + (long long)t * (m + gm.inner) + i]);
            }
            loaded_b = b;
        }
        const cplx* xv = blk_vec(X, b, v, gm.N);
        const int c_first = panel * gm.TC - 1;          // global column of panel column 0
        // ---- 1. stage the panel: Fock row n -> (n even ? n/2 : hep + n/2) ----
        for (int e = tid; e < 2 * hep * gm.PC; e += nthr) {
            const int r = e / gm.PC, pc = e % gm.PC;
            const int n = (r < hep) ? 2 * r : 2 * (r - hep) + 1;
            const int rr = (r < hep) ? r : r - hep;
            const bool rowok = (r < hep) ? (rr < he) : (rr < ho);
            const int c = c_first + pc;
            double* dst = sP + r * pld + 2 * pc;
            if (rowok && c >= 0 && c < gm.inner) {
                cp_async16(dst, xv + (long long)n * gm.inner + c);
            } else {
                dst[0] = 0.0;
                dst[1] = 0.0;
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();

        // ---- 2. forward transform (warp = column tile), in place ----
        if (warp < gm.CT) {
            double aE[NRT][2], aO[NRT][2];
#pragma unroll
            for (int r = 0; r < NRT; ++r) aE[r][0] = aE[r][1] = aO[r][0] = aO[r][1] = 0.0;
            const int cb = 8 * warp + gq;
            for (int k0 = 0; k0 < gm.kp; k0 += 4) {
                const double bE = sP[(k0 + tq) * pld + cb];
                const double bO = sP[(hep + k0 + tq) * pld + cb];
                const double* ve = sVe + (k0 + tq) * vld + gq;     // A[i][a] = Ve[a][i]
                const double* vo = sVo + (k0 + tq) * vld + gq;
#pragma unroll
                for (int r = 0; r < NRT; ++r) {
                    const bool in = 8 * r + gq < hep;
                    dmma884(aE[r][0], aE[r][1], in ? ve[8 * r] : 0.0, bE);
                    dmma884(aO[r][0], aO[r][1], in ? vo[8 * r] : 0.0, bO);
                }
            }
            __syncwarp();
            const int cc = 8 * warp + 2 * tq;
#pragma unroll
            for (int r = 0; r < NRT; ++r) {
                const int i = 8 * r + gq;
                if (i < hep) {
                    double2 zs = make_double2(aE[r][0] + aO[r][0], aE[r][1] + aO[r][1]);   // Z[i]
                    double2 zd = make_double2(aE[r][0] - aO[r][0], aE[r][1] - aO[r][1]);   // Z[m-1-i]
                    if (i >= ho) zd = make_double2(0.0, 0.0);
                    *reinterpret_cast<double2*>(sP + i * pld + cc) = zs;
                    *reinterpret_cast<double2*>(sP + (hep + i) * pld + cc) = zd;
                }
            }
        }
        __syncthreads();

        // ---- 3. potential + mirror combination, in place (warp = mirrored row pair) ----
        for (int i = warp; i < he; i += nwarp) {
            const bool pair = i < ho;                 // centre row of an odd m has no mirror
            const int im = m - 1 - i;
            const double lin_i = sLin[i], lin_m = pair ? sLin[im] : 0.0;
            cplx zi[2], zm[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int pc = lane + 32 * q;
                zi[q] = cmake(0.0, 0.0);
                zm[q] = cmake(0.0, 0.0);
                if (pc < gm.PC) {
                    const double* ri = sP + i * pld;
                    const double* rm = sP + (hep + i) * pld;
                    const cplx c0 = *reinterpret_cast<const cplx*>(ri + 2 * pc);
                    const cplx m0 = *reinterpret_cast<const cplx*>(rm + 2 * pc);
                    cplx oi = cmake(lin_i * c0.x, lin_i * c0.y);
                    cplx om = cmake(lin_m * m0.x, lin_m * m0.y);
                    for (int t = 0; t < gm.n_terms; ++t) {
                        const int s = gm.shift[t];
                        const cplx pi_ = sPh[t * m + i];
                        const cplx pm_ = pair ? sPh[t * m + im] : cmake(0.0, 0.0);
                        const int lo = pc - s, hi = pc + s;
                        if (lo >= 0 && lo < gm.PC) {
                            cfma(oi, pi_, *reinterpret_cast<const cplx*>(ri + 2 * lo));
                            cfma(om, pm_, *reinterpret_cast<const cplx*>(rm + 2 * lo));
                        }
                        if (hi >= 0 && hi < gm.PC) {
                            cfma(oi, cconj(pi_), *reinterpret_cast<const cplx*>(ri + 2 * hi));
                            cfma(om, cconj(pm_), *reinterpret_cast<const cplx*>(rm + 2 * hi));
                        }
                    }
                    zi[q] = oi;
                    zm[q] = om;
                }
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int pc = lane + 32 * q;
                if (pc < gm.PC) {
                    *reinterpret_cast<cplx*>(sP + i * pld + 2 * pc) = cadd(zi[q], zm[q]);          // S_i
                    *reinterpret_cast<cplx*>(sP + (hep + i) * pld + 2 * pc) =
                        pair ? csub(zi[q], zm[q]) : cmake(0.0, 0.0);                               // D_i
                }
            }
        }
        __syncthreads();

        // ---- 4. backward transform + LC diagonal epilogue ----
        if (warp < gm.CT) {
            double yE[NRT][2], yO[NRT][2];
#pragma unroll
            for (int r = 0; r < NRT; ++r) yE[r][0] = yE[r][1] = yO[r][0] = yO[r][1] = 0.0;
            const int cb = 8 * warp + gq;
            for (int k0 = 0; k0 < gm.kp; k0 += 4) {
                const double bS = sP[(k0 + tq) * pld + cb];
                const double bD = sP[(hep + k0 + tq) * pld + cb];
                const double* ve = sVe + gq * vld + k0 + tq;       // A[a][i] = Ve[a][i]
                const double* vo = sVo + gq * vld + k0 + tq;
#pragma unroll
                for (int r = 0; r < NRT; ++r) {
                    const bool in = 8 * r + gq < hep;
                    dmma884(yE[r][0], yE[r][1], in ? ve[8 * r * vld] : 0.0, bS);
                    dmma884(yO[r][0], yO[r][1], in ? vo[8 * r * vld] : 0.0, bD);
                }
            }
            const int pcc = 4 * warp + tq;              // complex panel column of this lane
            const int c = c_first + pcc;
            if (pcc >= 1 && pcc <= gm.TC && c < gm.inner) {
                cplx* yv = blk_vec(Y, b, v, gm.N);
                const double* fd = fdiag ? fdiag + b * fdiag_stride_b : nullptr;
#pragma unroll
                for (int r = 0; r < NRT; ++r) {
                    {
                        const int aidx = 8 * r + gq;
                        if (aidx < he) {
                            const long long idx = (long long)(2 * aidx) * gm.inner + c;
                            cplx o = cmake(yE[r][0], yE[r][1]);
                            if (fd) {
                                const double dv = fd[idx];
                                const cplx x0 = xv[idx];
                                o.x = fma(dv, x0.x, o.x);
                                o.y = fma(dv, x0.y, o.y);
                            }
                            yv[idx] = o;
                        }
                        if (aidx < ho) {
                            const long long idx = (long long)(2 * aidx + 1) * gm.inner + c;
                            cplx o = cmake(yO[r][0], yO[r][1]);
                            if (fd) {
                                const double dv = fd[idx];
                                const cplx x0 = xv[idx];
                                o.x = fma(dv, x0.x, o.x);
                                o.y = fma(dv, x0.y, o.y);
                            }
                            yv[idx] = o;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

extern "C" int sqc_hx_fused(const sqc_space_t* sp, const double* Ve, const double* Vo,
                            const sqc_potential_t* pot, const double* fdiag, int64_t fdiag_stride_b,
                            sqc_block_t X, sqc_block_t Y, int32_t B, void* stream) {
    if (!sp || !pot || B <= 0 || X.cnt_fixed <= 0 || X.ptr == Y.ptr) return SQC_EINVAL;
    if (sp->n_modes != 2 || !sp->harmonic[0] || sp->harmonic[1]) return SQC_ELIMIT;
    if (pot->n_terms < 0 || pot->n_terms > HX_MAXT || pot->diag) return SQC_ELIMIT;
    HxGeom gm;
    gm.m = sp->dims[0];
    gm.inner = sp->dims[1];
    if (gm.m < 2 || gm.m > 128) return SQC_ELIMIT;
    gm.he = (gm.m + 1) / 2;
    gm.ho = gm.m / 2;
    gm.hep = (gm.he + 7) & ~7;
    gm.kp = (gm.he + 3) & ~3;
    gm.nrt = gm.hep / 8;
    gm.N = (long long)gm.m * gm.inner;
    gm.n_terms = pot->n_terms;
    for (int t = 0; t < gm.n_terms; ++t) {
        gm.shift[t] = pot->shift[t][1];
        if (gm.shift[t] < -1 || gm.shift[t] > 1) return SQC_ELIMIT;
    }
    // panels: PC = TC + 2 columns, PC a multiple of 4 (8 real columns per tile), PC <= 36 (9 warps)
    int best_np = 0, best_pc = 0;
    for (int np = 1; np <= gm.inner; ++np) {
        int tc = (gm.inner + np - 1) / np;
        int pc = ((tc + 2) + 3) & ~3;
        if (pc <= 36) { best_np = np; best_pc = pc; break; }
    }
    gm.np = best_np;
    gm.PC = best_pc;
    gm.TC = (gm.inner + gm.np - 1) / gm.np;
    gm.CT = gm.PC / 4;
    // leading dims == 4 (mod 8) doubles keep both DMMA fragment patterns bank-conflict free
    gm.vld = gm.kp;
    while ((gm.vld & 7) != 4) ++gm.vld;
    gm.pld = 2 * gm.PC;
    while ((gm.pld & 7) != 4) ++gm.pld;
    int nwarp = gm.CT < 4 ? 4 : gm.CT;
    size_t smem = ((size_t)2 * gm.kp * gm.vld + (size_t)2 * gm.kp * gm.pld + ((gm.m + 1) & ~1)) * sizeof(double) +
                  (size_t)(gm.n_terms > 0 ? gm.n_terms : 1) * gm.m * sizeof(cplx);
    if (smem > 227 * 1024) return SQC_ELIMIT;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int per_sm = (smem <= 112 * 1024) ? 2 : 1;
    const long long total = (long long)B * X.cnt_fixed * gm.np;
    long long grid = (long long)nsm * per_sm;
    if (grid > total) grid = total;
#define HX_CASE(R)                                                                                     \
    case R: {                                                                                          \
        cudaError_t e = cudaFuncSetAttribute(hx_fused_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem);                                               \
        if (e != cudaSuccess) return (int)e;                                                           \
        hx_fused_kernel<R><<<(unsigned)grid, 32 * nwarp, smem, (cudaStream_t)stream>>>(                \
            gm, Ve, Vo, pot->lam, (const cplx*)pot->etab, pot->etab_stride_b, pot->g, (const cplx*)pot->a, \
            fdiag, fdiag_stride_b, X, Y, X.cnt_fixed, total);                                          \
        break;                                                                                         \
    }
    switch (gm.nrt) {
        HX_CASE(1) HX_CASE(2) HX_CASE(3) HX_CASE(4) HX_CASE(5) HX_CASE(6) HX_CASE(7) HX_CASE(8)
        default: return SQC_ELIMIT;
    }
#undef HX_CASE
    SQC_LAUNCHED();
    return 0;
}
