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

#define HX_MAXT 4        // junction terms
#ifndef HX_TEAMS
#define HX_TEAMS 1      // 1: a single 8-warp team (30 + 2 column panels); 2: two independent 4-warp teams per
                        // CTA (14 + 2 column panels) -- measured slower (7.5 vs 6.7 ms at 2048 x 12 vectors): the extra
                        // halo work costs more than the additional stream overlaps
#endif
#ifndef HX_KUNROLL
#define HX_KUNROLL 2
#endif
#ifndef HX_EB
#define HX_EB 2
#endif

struct HxGeom {
    int m, he, ho;       // Fock dim, #even rows, #odd rows
    int kp;              // he rounded up to 4 (rows of each half / contraction length)
    int inner;           // charge-mode dimension (columns)
    int np, TC, PC, CT;  // panels per vector, useful cols per panel, cols incl. halo, col tiles
    int CT_last;         // column tiles of the last (narrower) panel of a vector
    int vld, pld;        // leading dims (doubles) of the half matrices and of the panel
    int pg;              // row pairs per warp in the potential pass
    int rotate;          // panel index rotated by the round (needs grid x teams to be a multiple of np)
    int ng;              // independent warp teams per CTA (each works on its own panel; they share Ve / Vo)
    int team_doubles;    // shared-memory doubles of one team (panel + per-point tables)
    int n_terms;
    int shpack;          // junction charge shifts (-1, 0, +1), 2 bits each, biased by one
    long long N;
};

// KS = contraction steps (kp / 4) fixes every shared-memory extent of the half matrices at compile
// time, so all A-fragment offsets in both transforms are immediates.
template <int KS>
__global__ void __launch_bounds__(256, 2)
hx_fused_kernel(HxGeom gm, const double* __restrict__ Ve, const double* __restrict__ Vo,
                const double* __restrict__ lam, const cplx* __restrict__ etab, long long etab_stride_b,
                const double* __restrict__ g, const cplx* __restrict__ a,
                const double* __restrict__ fdiag, long long fdiag_stride_b,
                sqc_block_t X, sqc_block_t Y, int vmax, long long total_items) {
    constexpr int KP = 4 * KS;                            // rows of each parity half (he rounded to 4)
    constexpr int NRT = (KS + 1) / 2;                     // 8-row tiles per half
    constexpr int VLD = (KP % 8 == 4) ? KP : KP + 4;      // == 4 (mod 8): conflict-free fragments
    constexpr int KU = HX_KUNROLL;
    extern __shared__ double smem[];
    // row tiles that stick out past KP read finite neighbouring data and are never written back
    double* sVe = smem;                                   // [KP][VLD]   Ve[a][i] = V[2a][i]
    double* sVo = sVe + KP * VLD;                         // [KP][VLD]   Vo[a][i] = V[2a+1][i]
    // Teams: the warps of a CTA are split into gm.ng independent teams.  A team streams its own panels
    // (own panel buffer, own per-point tables, own named barrier); the teams only share the half
    // matrices.  More independent panel streams per SM = better overlap of the tensor-core phases
    // of one stream with the staging / potential / epilogue phases of the others.
    const int tid_cta = threadIdx.x;
    const int nwarp = (blockDim.x >> 5) / gm.ng;          // warps per team
    const int team = (tid_cta >> 5) / nwarp;
    const int tid = tid_cta - team * nwarp * 32, nthr = nwarp * 32;      // thread index / count inside the team
    const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int m = gm.m, he = gm.he, ho = gm.ho, pld = gm.pld;
    double* sP = sVo + KP * VLD + team * gm.team_doubles; // [2*KP][pld] panel: S/even half, D/odd half
    double* sLin = sP + 2 * KP * pld;                     // [m]
    cplx* sPh = reinterpret_cast<cplx*>(sLin + ((m + 1) & ~1));   // [n_terms][m]
    auto team_sync = [&]() {
        if (gm.ng == 1) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(nthr) : "memory");
    };

    // half matrices (zero padded), staged once per CTA
    for (int e = tid_cta; e < KP * VLD; e += blockDim.x) {
        const int r = e / VLD, c = e % VLD;
        sVe[e] = (r < he && c < he) ? Ve[r * he + c] : 0.0;
        sVo[e] = (r < ho && c < ho) ? Vo[r * ho + c] : 0.0;
    }
    for (int e = tid; e < 2 * KP * pld; e += nthr) sP[e] = 0.0;
    __syncthreads();
    // DMMA warps: the lane's complex column inside the warp's 4-column tile
    const int pcc = 4 * warp + tq;
    int loaded_b = -1;
    // grid x teams is a multiple of np, so all panels of a vector are dealt in the same round; the panel
    // index is rotated by the round so that every CTA sees wide and narrow panels alike
    int round = 0;
    for (long long item = (long long)blockIdx.x * gm.ng + team; item < total_items;
         item += (long long)gridDim.x * gm.ng, ++round) {
        const long long vec = item / gm.np;
        const int panel = (int)((item % gm.np + (gm.rotate ? round : 0)) % gm.np);
        const int b = (int)(vec / vmax), v = (int)(vec % vmax);
        if (v >= blk_count(X, b)) continue;
        if (b != loaded_b) {
            // per-sweep-point tables: lin(i) = g0 + g1 lam_i ;  ph_t(i) = a_t etab_t[i]
            // (only the potential pass reads them; it sits between CTA barriers)
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
        // the last panel of a vector is narrower: fewer column tiles (still a multiple of four,
        // so that the tensor-core work stays evenly spread over the four SM sub-partitions)
        const int ct = (panel == gm.np - 1) ? gm.CT_last : gm.CT;
        const int pcn = 4 * ct;
        const int c = c_first + pcc;                    // this lane's global column (DMMA warps)
        if (warp < ct) {
            // ---- 1. warp-private staging of the warp's own 4-column tile (64 B per row) ----
            // Fock row n -> (n even ? n/2 : KP + n/2); padding rows stay zero from the start
            {
                double* dcol = sP + 2 * pcc;
                if (c >= 0 && c < gm.inner) {
                    const cplx* src = xv + c;
                    for (int a2 = gq; a2 < he; a2 += 8)
                        cp_async16(dcol + a2 * pld, src + (long long)(2 * a2) * gm.inner);
                    for (int a2 = gq; a2 < ho; a2 += 8)
                        cp_async16(dcol + (KP + a2) * pld, src + (long long)(2 * a2 + 1) * gm.inner);
                } else {
                    for (int a2 = gq; a2 < he; a2 += 8)
                        *reinterpret_cast<double2*>(dcol + a2 * pld) = make_double2(0.0, 0.0);
                    for (int a2 = gq; a2 < ho; a2 += 8)
                        *reinterpret_cast<double2*>(dcol + (KP + a2) * pld) = make_double2(0.0, 0.0);
                }
                cp_async_commit();
                cp_async_wait<0>();
                __syncwarp();
            }
            // ---- 2. forward transform of the tile, in place ----
            double aE[NRT][2], aO[NRT][2];
#pragma unroll
            for (int r = 0; r < NRT; ++r) aE[r][0] = aE[r][1] = aO[r][0] = aO[r][1] = 0.0;
            const double* bp = sP + tq * pld + 8 * warp + gq;
            const double* ve = sVe + tq * VLD + gq;               // A[i][a] = Ve[a][i]
            const double* vo = sVo + tq * VLD + gq;
#pragma unroll KU
            for (int k = 0; k < KS; ++k) {
                const double bE = bp[(4 * k) * pld];
                const double bO = bp[(KP + 4 * k) * pld];
#pragma unroll
                for (int r = 0; r < NRT; ++r) {
                    dmma884(aE[r][0], aE[r][1], ve[4 * k * VLD + 8 * r], bE);
                    dmma884(aO[r][0], aO[r][1], vo[4 * k * VLD + 8 * r], bO);
                }
            }
            __syncwarp();
            double* zc = sP + 8 * warp + 2 * tq;
#pragma unroll
            for (int r = 0; r < NRT; ++r) {
                const int i = 8 * r + gq;
                if (i < KP) {
                    double2 zs = make_double2(aE[r][0] + aO[r][0], aE[r][1] + aO[r][1]);   // Z[i]
                    double2 zd = make_double2(aE[r][0] - aO[r][0], aE[r][1] - aO[r][1]);   // Z[m-1-i]
                    if (i >= ho) zd = make_double2(0.0, 0.0);
                    *reinterpret_cast<double2*>(zc + i * pld) = zs;
                    *reinterpret_cast<double2*>(zc + (KP + i) * pld) = zd;
                }
            }
        }
        team_sync();

        // ---- 3. potential + mirror combination, in place ----
        // a warp owns pg whole row pairs; the (at most 32) columns of a row are done by its lanes:
        // read, warp sync, write back -- nothing waits on the CTA inside the pass.
        for (int i_base = warp * gm.pg; i_base < he; i_base += nwarp * gm.pg) {
            const int i_end = min(i_base + gm.pg, he);
            auto eval = [&](int i, int pc, cplx& S, cplx& D) {
                const bool pair = i < ho;                 // centre row of an odd m has no mirror
                const int im = m - 1 - i;
                const double* ri = sP + i * pld;
                const double* rm = sP + (KP + i) * pld;
                const double lin_i = sLin[i], lin_m = pair ? sLin[im] : 0.0;
                const cplx c0 = *reinterpret_cast<const cplx*>(ri + 2 * pc);
                const cplx m0 = *reinterpret_cast<const cplx*>(rm + 2 * pc);
                cplx oi = cmake(lin_i * c0.x, lin_i * c0.y);
                cplx om = cmake(lin_m * m0.x, lin_m * m0.y);
                for (int t = 0; t < gm.n_terms; ++t) {
                    const int sh = ((gm.shpack >> (2 * t)) & 3) - 1;
                    const cplx pi_ = sPh[t * m + i];
                    const cplx pm_ = pair ? sPh[t * m + im] : cmake(0.0, 0.0);
                    const int lo = pc - sh, hi = pc + sh;
                    if (lo >= 0 && lo < pcn) {
                        cfma(oi, pi_, *reinterpret_cast<const cplx*>(ri + 2 * lo));
                        cfma(om, pm_, *reinterpret_cast<const cplx*>(rm + 2 * lo));
                    }
                    if (hi >= 0 && hi < pcn) {
                        cfma(oi, cconj(pi_), *reinterpret_cast<const cplx*>(ri + 2 * hi));
                        cfma(om, cconj(pm_), *reinterpret_cast<const cplx*>(rm + 2 * hi));
                    }
                }
                S = cadd(oi, om);                                        // S_i
                D = pair ? csub(oi, om) : cmake(0.0, 0.0);              // D_i
            };
            // two evaluations per trip: their loads and FP64 chains are independent (the pass is
            // latency bound); with <= 16 panel columns the lanes of one evaluation cover two rows
            const int cw = pcn <= 16 ? 16 : 32, rpl = 32 / cw;
            const int pcol = lane % cw, prow = lane / cw;
            for (int i = i_base; i < i_end; i += 2 * rpl) {
                const int r0 = i + prow, r1 = i + rpl + prow;
                const bool ok0 = pcol < pcn && r0 < i_end, ok1 = pcol < pcn && r1 < i_end;
                cplx s0, d0, s1, d1;
                if (ok0) eval(r0, pcol, s0, d0);
                if (ok1) eval(r1, pcol, s1, d1);
                __syncwarp();
                if (ok0) {
                    *reinterpret_cast<cplx*>(sP + r0 * pld + 2 * pcol) = s0;
                    *reinterpret_cast<cplx*>(sP + (KP + r0) * pld + 2 * pcol) = d0;
                }
                if (ok1) {
                    *reinterpret_cast<cplx*>(sP + r1 * pld + 2 * pcol) = s1;
                    *reinterpret_cast<cplx*>(sP + (KP + r1) * pld + 2 * pcol) = d1;
                }
            }
        }
        team_sync();

        // ---- 4. backward transform of the warp's tile + LC diagonal epilogue (warp-private) ----
        if (warp < ct) {
            double yE[NRT][2], yO[NRT][2];
#pragma unroll
            for (int r = 0; r < NRT; ++r) yE[r][0] = yE[r][1] = yO[r][0] = yO[r][1] = 0.0;
            const double* bp = sP + tq * pld + 8 * warp + gq;
            const double* ve = sVe + gq * VLD + tq;               // A[a][i] = Ve[a][i]
            const double* vo = sVo + gq * VLD + tq;
#pragma unroll KU
            for (int k = 0; k < KS; ++k) {
                const double bS = bp[(4 * k) * pld];
                const double bD = bp[(KP + 4 * k) * pld];
#pragma unroll
                for (int r = 0; r < NRT; ++r) {
                    dmma884(yE[r][0], yE[r][1], ve[8 * r * VLD + 4 * k], bS);
                    dmma884(yO[r][0], yO[r][1], vo[8 * r * VLD + 4 * k], bD);
                }
            }
            // the lane holds complex column pcc of Fock rows 2(8r+gq) [yE] and 2(8r+gq)+1 [yO];
            // all loads of a batch are issued before its first store (X, Y and fdiag may alias
            // as far as the compiler knows)
            if (pcc >= 1 && pcc <= gm.TC && c < gm.inner) {
                cplx* yv = blk_vec(Y, b, v, gm.N);
                const double* fd = fdiag ? fdiag + b * fdiag_stride_b : nullptr;
                constexpr int EB = HX_EB;                       // rows per batch
#pragma unroll
                for (int r0 = 0; r0 < NRT; r0 += EB) {
#pragma unroll
                    for (int par = 0; par < 2; ++par) {
                        const int hcnt = par ? ho : he;
                        double dd[EB];
                        cplx xx[EB];
#pragma unroll
                        for (int j = 0; j < EB; ++j) {
                            const int aidx = 8 * (r0 + j) + gq;
                            dd[j] = 0.0;
                            xx[j] = cmake(0.0, 0.0);
                            if (fd && r0 + j < NRT && aidx < hcnt) {
                                const long long idx = (long long)(2 * aidx + par) * gm.inner + c;
                                dd[j] = fd[idx];
                                xx[j] = xv[idx];
                            }
                        }
#pragma unroll
                        for (int j = 0; j < EB; ++j) {
                            const int aidx = 8 * (r0 + j) + gq;
                            if (r0 + j < NRT && aidx < hcnt) {
                                const long long idx = (long long)(2 * aidx + par) * gm.inner + c;
                                const int rr = (r0 + j < NRT) ? r0 + j : 0;
                                cplx o = par ? cmake(yO[rr][0], yO[rr][1]) : cmake(yE[rr][0], yE[rr][1]);
                                o.x = fma(dd[j], xx[j].x, o.x);
                                o.y = fma(dd[j], xx[j].y, o.y);
                                yv[idx] = o;
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
        // no CTA barrier here: from the backward transform to the next forward transform a warp
        // only touches its own column tile
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
    gm.kp = (gm.he + 3) & ~3;
    gm.N = (long long)gm.m * gm.inner;
    gm.n_terms = pot->n_terms;
    gm.shpack = 0;
    for (int t = 0; t < gm.n_terms; ++t) {
        const int sh = pot->shift[t][1];
        if (sh < -1 || sh > 1) return SQC_ELIMIT;
        gm.shpack |= (sh + 1) << (2 * t);
    }
    // panels: PC = TC + 2 columns (one halo column each side), 4 complex columns per warp tile.
    // The tile count is kept a multiple of four wherever possible: a CTA's warps are dealt
    // round-robin to the four SM sub-partitions, each with its own FP64 tensor pipe, and e.g. nine
    // equal tiles put 3+3 warps of two resident CTAs on one sub-partition (measured: 6.2 instead of
    // 4.0 cycles per DMMA).  Wide modes use 30 + 2 column panels; the last panel takes the remainder.
    // HX_TEAMS teams of 4 warps (14 + 2 column panels) or one team of 8 warps (30 + 2 column panels)
    const int teams = HX_TEAMS;
    const int tcw = teams == 2 ? 14 : 30, ctw = teams == 2 ? 4 : 8;
    if (gm.inner + 2 <= 4 * ctw) {
        gm.np = 1;
        gm.TC = gm.inner;
        gm.PC = (gm.inner + 2 + 3) & ~3;
        gm.CT = gm.CT_last = gm.PC / 4;
    } else {
        gm.TC = tcw;
        gm.PC = 4 * ctw;
        gm.CT = ctw;
        gm.np = (gm.inner + gm.TC - 1) / gm.TC;
        const int tl = gm.inner - gm.TC * (gm.np - 1);
        gm.CT_last = ((tl + 2 + 3) / 4 + 3) & ~3;
        if (gm.CT_last > ctw) gm.CT_last = ctw;
    }
    gm.ng = teams;
    // leading dims == 4 (mod 8) doubles keep both DMMA fragment patterns bank-conflict free
    gm.vld = (gm.kp % 8 == 4) ? gm.kp : gm.kp + 4;        // must match VLD in the kernel
    gm.pld = 2 * gm.PC;
    while ((gm.pld & 7) != 4) ++gm.pld;
    int nwarp = gm.CT < 4 ? 4 : gm.CT;       // warps per team
    gm.pg = (gm.he + nwarp - 1) / nwarp;      // row pairs per warp in the potential pass
    gm.team_doubles = 2 * gm.kp * gm.pld + ((gm.m + 1) & ~1) + 2 * (gm.n_terms > 0 ? gm.n_terms : 1) * gm.m;
    size_t smem = ((size_t)2 * gm.kp * gm.vld + (size_t)gm.ng * gm.team_doubles) * sizeof(double);
    if (smem > 227 * 1024) return SQC_ELIMIT;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int per_sm = (smem <= 112 * 1024) ? 2 : 1;
    const long long total = (long long)B * X.cnt_fixed * gm.np;
    long long grid = (long long)nsm * per_sm;
    if (grid * gm.ng > total) grid = (total + gm.ng - 1) / gm.ng;
    // panel rotation (see the kernel) needs grid x teams to be a multiple of np; tiny launches go without
    long long gr = grid;
    while ((gr * gm.ng) % gm.np != 0 && gr > 1) --gr;
    gm.rotate = (gr * gm.ng) % gm.np == 0;
    if (gm.rotate) grid = gr;
#define HX_CASE(R)                                                                                     \
    case R: {                                                                                          \
        cudaError_t e = cudaFuncSetAttribute(hx_fused_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem);                                               \
        if (e != cudaSuccess) return (int)e;                                                           \
        hx_fused_kernel<R><<<(unsigned)grid, 32 * nwarp * gm.ng, smem, (cudaStream_t)stream>>>(                \
            gm, Ve, Vo, pot->lam, (const cplx*)pot->etab, pot->etab_stride_b, pot->g, (const cplx*)pot->a, \
            fdiag, fdiag_stride_b, X, Y, X.cnt_fixed, total);                                          \
        break;                                                                                         \
    }
    switch (gm.kp / 4) {
        HX_CASE(1) HX_CASE(2) HX_CASE(3) HX_CASE(4) HX_CASE(5) HX_CASE(6) HX_CASE(7) HX_CASE(8)
        HX_CASE(9) HX_CASE(10) HX_CASE(11) HX_CASE(12) HX_CASE(13) HX_CASE(14) HX_CASE(15) HX_CASE(16)
        default: return SQC_ELIMIT;
    }
#undef HX_CASE
    SQC_LAUNCHED();
    return 0;
}
