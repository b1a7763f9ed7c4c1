// Fused matrix-free H.X in the REAL (time-reversal adapted) representation, for circuits with one
// harmonic mode (mode 0) and one charge mode (mode 1) at zero gate charge -- the 0-pi flux sweep.
//
// With n_g = 0 the Hamiltonian of Circuit.hamiltonian() (circuit.py:2147-2164) commutes with
// T = (complex conjugation in the Fock x charge basis) x (n -> -n):  every junction term
// e^{i phi_ext} D(alpha) (x) d^{s} is mapped onto its own Hermitian conjugate.  T-invariant vectors have
// z(i, -n) = conj z(i, n), so only the charges n >= 0 are stored, as N = m * inner REAL numbers
//     y[i][0] = z(i, 0),   y[i][n] = sqrt2 Re z(i, n),   y[i][nc + n] = sqrt2 Im z(i, n)   (1 <= n <= nc)
// ("planar" layout, inner = 2 nc + 1), an orthonormal real basis in which H is real symmetric.  Half the
// bytes and half the transform flops of the complex kernel (hx_fused.cu), and no halo columns: the charge
// -1 neighbour of column 0 is the conjugate of column 1.
//
// One CTA per SM, two independent teams of one warp per 8-column tile (13 warps for inner = 101: every
// warp carries the same tensor work, nothing waits at the team barriers for a warp with two tiles); a team
// owns one whole vector at a time:
//   1. TMA bulk copies (cp.async.bulk + mbarrier, one per pair of Fock rows -- the row pitch of 8*inner
//      bytes is an odd multiple of 8, so a 2-D tensor map cannot describe it) land the vector in shared
//      memory in its natural row order; a pair of rows is padded so that DMMA fragment loads are
//      conflict free.
//   2. forward transform on the FP64 tensor cores with the parity split of the x basis
//      (E = Ve^T X_even -> rows 2i, O = Vo^T X_odd -> rows 2i+1), in place, one parity half at a time;
//   3. Z_i = E_i + O_i, Z_{m-1-i} = E_i - O_i, potential and mirror sums in place (warp = row pair,
//      lanes = charges);
//   4. backward transform (Y_even = Ve S, Y_odd = Vo D) + LC diagonal (d and X come from L2 -- the vector was
//      staged through it a few microseconds earlier -- software-pipelined two k-steps ahead of the DMMAs
//      that hide their latency), written in place in natural row order, then TMA bulk stores.
// While one team waits for its store / load, the other one keeps the tensor pipe busy.
#include "common.cuh"
#ifdef HXR_PROFILE      // dev build (tools/hxr_profile.sh): phase cycle counts of one team, printed once
#include <cstdio>
#define HXR_MARK(i) do { if (prof) tp[i] = clock64(); } while (0)
#else
#define HXR_MARK(i) do {} while (0)
#endif

#ifndef HXR_LAG_SEP        // (tuning knobs of the epilogue pipeline, tools/hxr_variants.sh; measured at the bench
#define HXR_LAG_SEP 2      //  geometry, 12288 vectors: 4 -> 1.994 ms, 3 -> 1.963, 2 -> 1.907, 1 -> 1.930, no pipeline 2.007)
#endif
#ifndef HXR_PIPE
#define HXR_PIPE 1
#endif
#define HXR_MAXT 4
#define HXR_MAX_WARPS 28    // warps per CTA (registers: 65536 / (28 * 32) = 73 per thread)
#define HXR_MAX_TEAMS 4

struct HxrGeom {
    int m, he, ho;       // Fock dim, #even rows, #odd rows
    int inner, nc;       // real columns per Fock row (2 nc + 1), charge cut-off
    int pp;              // doubles per padded pair of rows in shared memory
    int nct;             // 8-wide column tiles = warps per team
    int teams;           // teams per CTA (as many vectors as fit in shared memory and in HXR_MAX_WARPS)
    int buf_doubles;     // shared-memory doubles of one team's vector buffer
    int n_terms;
    int shpack;          // junction charge shifts (-1, 0, +1), 2 bits each, biased by one
    long long N;         // reals per vector (m * inner)
    long long vstride;   // reals between consecutive vectors of a block (2 * Nc)
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// TMA bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_store(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One half (par = 0: even Fock rows / Ve, par = 1: odd rows / Vo) of the parity-split transform of one column
// tile (8 real columns) of the team's buffer, all row tiles, in place:
//   FWD : row(2i + par) <- sum_a V[a][i] row(2a + par)      (A[i][a] = V[a][i];  E_i resp. O_i)
//   !FWD: row(2a + par) <- sum_i V[a][i] row(2i + par) + d .* X   (A[a][i] = V[a][i];  Y_even resp. Y_odd)
// The warp owns these columns of these rows, so a __syncwarp separates its reads from its writes.
// Epilogue of the backward pass: X comes from global memory (L2: the vector was staged through it a few
// microseconds earlier) LAG k-steps ahead of its use, the DMMAs in between hide the latency.  d is either
// the full LC diagonal fd (same indexing as X) or, when it is separable (SEP), only its Fock part dh[row]
// from shared memory -- the charge part was added in the x basis by the potential phase.
template <int KS, bool FWD, bool SEP>
__device__ __forceinline__ void hxr_pass(const double* __restrict__ sV, double* buf, const HxrGeom& gm, int par,
                                         int col0, int gq, int tq, const double* __restrict__ xv,
                                         const double* __restrict__ fd, const double* __restrict__ sDh) {
    constexpr int KP = 4 * KS, NRT = (KS + 1) / 2;
    constexpr int VLD = (KP % 8 == 4) ? KP : KP + 4;
    constexpr int LAG = SEP ? HXR_LAG_SEP : 2;      // row tiles in flight (registers: 2 resp. 4 doubles per tile)
    constexpr bool PIPE = HXR_PIPE && KS >= NRT + LAG;
    const int pp = gm.pp, inner = gm.inner;
    const int hcnt = par ? gm.ho : gm.he;
    double acc[NRT][2];
#pragma unroll
    for (int r = 0; r < NRT; ++r) acc[r][0] = acc[r][1] = 0.0;
    const double* va = FWD ? sV + tq * VLD + gq : sV + gq * VLD + tq;
    double* rows = buf + par * inner;
    const double* bp = rows + tq * pp + col0 + gq;
    const int c = col0 + 2 * tq;
    const bool ok0 = c < inner, ok1 = c + 1 < inner;
    const bool epi = !FWD && xv != nullptr;
    double px[LAG][2], pd[SEP ? 1 : LAG][2];
    auto epi_load = [&](int r) {
        const int i = 8 * r + gq;
        const long long idx = (long long)(2 * i + par) * inner + c;
        const bool on = epi && i < hcnt;
        px[r % LAG][0] = (on && ok0) ? xv[idx] : 0.0;
        px[r % LAG][1] = (on && ok1) ? xv[idx + 1] : 0.0;
        if (!SEP) {
            pd[r % LAG][0] = (on && ok0) ? fd[idx] : 0.0;
            pd[r % LAG][1] = (on && ok1) ? fd[idx + 1] : 0.0;
        }
    };
    auto epi_use = [&](int r) {
        if (SEP) {
            const int i = 8 * r + gq;
            const double dh = i < hcnt ? sDh[2 * i + par] : 0.0;
            acc[r][0] = fma(dh, px[r % LAG][0], acc[r][0]);
            acc[r][1] = fma(dh, px[r % LAG][1], acc[r][1]);
        } else {
            acc[r][0] = fma(pd[r % LAG][0], px[r % LAG][0], acc[r][0]);
            acc[r][1] = fma(pd[r % LAG][1], px[r % LAG][1], acc[r][1]);
        }
    };
    if (!FWD && PIPE) {
#pragma unroll
        for (int r = 0; r < LAG - 1 && r < NRT; ++r) epi_load(r);
    }
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        if (!FWD && PIPE && k + LAG - 1 < NRT) epi_load(k + LAG - 1);
        const double bf = bp[(4 * k) * pp];
#pragma unroll
        for (int r = 0; r < NRT; ++r) {
            const double af = FWD ? va[4 * k * VLD + 8 * r] : va[8 * r * VLD + 4 * k];
            dmma884(acc[r][0], acc[r][1], af, bf);
        }
        if (!FWD && PIPE && k < NRT) epi_use(k);       // loaded LAG - 1 k-steps ago
    }
    if (!FWD && !PIPE) {
        // small transforms: too few k-steps to pipeline
#pragma unroll
        for (int r = 0; r < NRT; ++r) {
            epi_load(r);
            epi_use(r);
        }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < NRT; ++r) {
        const int i = 8 * r + gq;
        if (i < hcnt) {
            double* z = rows + i * pp + c;
            if (ok0) z[0] = acc[r][0];
            if (ok1) z[1] = acc[r][1];
        }
    }
}

template <int KS, bool FWD, bool SEP>
__device__ __forceinline__ void hxr_transform(const double* sVe, const double* sVo, double* buf, const HxrGeom& gm,
                                              int col0, int gq, int tq, const double* __restrict__ xv,
                                              const double* __restrict__ fd, const double* __restrict__ sDh) {
    hxr_pass<KS, FWD, SEP>(sVe, buf, gm, 0, col0, gq, tq, xv, fd, sDh);
    hxr_pass<KS, FWD, SEP>(sVo, buf, gm, 1, col0, gq, tq, xv, fd, sDh);
}

template <int KS, bool SEP>
__global__ void __launch_bounds__(32 * HXR_MAX_WARPS, 1)
hx_real_kernel(HxrGeom gm, const double* __restrict__ Ve, const double* __restrict__ Vo,
               const double* __restrict__ lam, const cplx* __restrict__ etab, long long etab_stride_b,
               const double* __restrict__ g, const cplx* __restrict__ a,
               const double* __restrict__ fdiag, long long fdiag_stride_b, const double* __restrict__ dsep,
               sqc_block_t X, sqc_block_t Y, int vmax, long long total_items) {
    constexpr int KP = 4 * KS;
    constexpr int VLD = (KP % 8 == 4) ? KP : KP + 4;
    extern __shared__ double smem[];
    double* sVe = smem;                                   // [KP][VLD]   Ve[a][i] = V[2a][i]
    double* sVo = sVe + KP * VLD;                         // [KP][VLD]   Vo[a][i] = V[2a+1][i]
    const int tid_cta = threadIdx.x;
    const int team_warps = gm.nct, team_threads = 32 * gm.nct;
    const int team = tid_cta / team_threads;
    const int tid = tid_cta - team * team_threads;
    const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int m = gm.m, he = gm.he, ho = gm.ho, inner = gm.inner, nc = gm.nc, pp = gm.pp;
    const int team_doubles = gm.buf_doubles + ((m + 1) & ~1) + 2 * HXR_MAXT * m + 8;
    double* buf = sVo + KP * VLD + team * team_doubles;   // [KP pairs][pp]: rows 2a (offset 0), 2a+1 (offset inner)
    double* sLin = buf + gm.buf_doubles;                  // [m]
    cplx* sPh = reinterpret_cast<cplx*>(sLin + ((m + 1) & ~1));   // [n_terms][m]
    // descriptor of the team's current / next item, written by lane 0 of warp 0: {live b, v, next item}
    long long* desc = reinterpret_cast<long long*>(sLin + ((m + 1) & ~1) + 2 * HXR_MAXT * m);
    // separable LC diagonal d(i, j) = dh[i] + dc[|charge of column j|], shared by the teams
    double* sDh = sVo + KP * VLD + gm.teams * team_doubles;              // [m]
    double* sDc = sDh + ((m + 1) & ~1);                                  // [nc + 1]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sDc + ((nc + 2) & ~1));
    unsigned long long* bar = bars + team;
    auto team_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(team_threads) : "memory"); };

    for (int e = tid_cta; e < KP * VLD; e += blockDim.x) {
        const int r = e / VLD, c = e % VLD;
        sVe[e] = (r < he && c < he) ? Ve[r * he + c] : 0.0;
        sVo[e] = (r < ho && c < ho) ? Vo[r * ho + c] : 0.0;
    }
    for (int e = tid_cta; e < m + nc + 1; e += blockDim.x) {
        if (e < m) sDh[e] = SEP ? dsep[e] : 0.0;
        else sDc[e - m] = SEP ? dsep[e] : 0.0;          // real-layout columns 0 .. nc: charges 0 .. nc
    }
    if (team < gm.teams) {
        for (int e = tid; e < gm.buf_doubles; e += team_threads) buf[e] = 0.0;
        if (tid == 0) mbar_init(bar, 1);
    }
    fence_async_smem();
    __syncthreads();
    if (team >= gm.teams) return;

    const int col0 = 8 * warp;                            // this warp's column tile
    const double rs2 = 0.70710678118654752440, s2 = 1.41421356237309504880;
    const unsigned pair_bytes = 16u * (unsigned)inner;
    // the last pair of an odd m carries one row plus the zero padding element of the vector
    const unsigned last_bytes = (m & 1) ? 8u * (unsigned)(inner + 1) : pair_bytes;
    const unsigned vec_bytes = (unsigned)(he - 1) * pair_bytes + last_bytes;
    int loaded_b = -1;
    unsigned phase = 0;
    const long long stride = (long long)gridDim.x * gm.teams;
    // first live item at or after `it` (a sweep point may hold fewer than vmax vectors)
    auto next_live = [&](long long it) {
        while (it < total_items) {
            const int b = (int)(it / vmax), v = (int)(it % vmax);
            if (v < blk_count(X, b)) break;
            it += stride;
        }
        return it;
    };
    auto issue_load = [&](long long it) {          // warp 0 of the team
        const int b = (int)(it / vmax), v = (int)(it % vmax);
        const double* src = reinterpret_cast<const double*>(blk_vec(X, b, v, gm.vstride / 2));
        if (lane == 0) mbar_expect_tx(bar, vec_bytes);
        __syncwarp();
        for (int p = lane; p < he; p += 32)
            bulk_load(buf + p * pp, src + (long long)p * 2 * inner, p == he - 1 ? last_bytes : pair_bytes, bar);
    };
    if (warp == 0) {
        const long long first = next_live((long long)blockIdx.x * gm.teams + team);
        if (first < total_items) issue_load(first);
        if (lane == 0) desc[0] = first;
    }
    team_sync();
    long long item = desc[0];

#ifdef HXR_PROFILE
    long long tp[10];
    int nitem = 0;
#endif
    while (item < total_items) {
#ifdef HXR_PROFILE
        const bool prof = blockIdx.x == 1 && tid == 32 && ++nitem == 4;
#endif
        HXR_MARK(0);
        const int b = (int)(item / vmax), v = (int)(item % vmax);
        if (warp == 0 && lane == 0) desc[1] = next_live(item + stride);     // overlaps the wait for the load
        if (b != loaded_b) {
            // per-sweep-point tables: lin(i) = g0 + g1 lam_i ;  ph_t(i) = a_t etab_t[i]
            const double g0 = g[(long long)b * 3], g1 = g[(long long)b * 3 + 1];
            for (int i = tid; i < m; i += team_threads) sLin[i] = fma(g1, lam[i], g0);
            for (int e = tid; e < gm.n_terms * m; e += team_threads) {
                const int t = e / m, i = e % m;
                sPh[e] = cmul(a[(long long)b * gm.n_terms + t], etab[b * etab_stride_b + (long long)t * (m + inner) + i]);
            }
            loaded_b = b;
        }
        const double* xv = reinterpret_cast<const double*>(blk_vec(X, b, v, gm.vstride / 2));
        double* yv = reinterpret_cast<double*>(blk_vec(Y, b, v, gm.vstride / 2));
        mbar_wait(bar, phase);     // (letting warp 0 alone poll, the others sleeping in a team barrier: 2 % slower)
        phase ^= 1;
        HXR_MARK(1);

        // ---- forward transform, in place: row 2i <- E_i, row 2i+1 <- O_i ----
        hxr_transform<KS, true, SEP>(sVe, sVo, buf, gm, col0, gq, tq, nullptr, nullptr, nullptr);
        HXR_MARK(2);
        team_sync();
        HXR_MARK(3);

        // ---- potential + mirror combination, in place: a warp owns whole row pairs ----
        // Z_i = E_i + O_i (x point i), Z_{m-1-i} = E_i - O_i (its mirror);  S = P Z_i + P Z_mirror -> row 2i,
        // D = P Z_i - P Z_mirror -> row 2i+1.  Lane = charge n (two passes); y(n) and its neighbours "y(n-1)",
        // "y(n+1)" of a stored row are  (c_re row[i_re], c_im row[i_im])  with per-lane indices and
        // coefficients (no divergent code): the charge -1 neighbour of column 0 is conj y(1) / sqrt2,
        // columns 0 and 1 are linked by sqrt2, y(nc + 1) = 0.
        for (int i = warp; i < he; i += team_warps) {
            const bool pair = i < ho;
            const int im = pair ? m - 1 - i : i;
            double* re = buf + i * pp;
            double* ro = re + inner;
            const double lin_i = sLin[i], lin_m = pair ? sLin[im] : 0.0;
            const double pairf = pair ? 1.0 : 0.0;
            cplx S[2], D[2];
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int n0 = lane + 32 * pass;
                const int n = n0 <= nc ? n0 : nc;               // idle lanes repeat the last charge, do not store
                const int il_re = n >= 2 ? n - 1 : (n == 1 ? 0 : (nc >= 1 ? 1 : 0));
                const int il_im = n >= 2 ? nc + n - 1 : (n == 1 ? 0 : (nc >= 1 ? nc + 1 : 0));
                const double cl_re = n >= 2 ? 1.0 : (n == 1 ? s2 : (nc >= 1 ? rs2 : 0.0));
                const double cl_im = n >= 2 ? 1.0 : (n == 1 ? 0.0 : (nc >= 1 ? -rs2 : 0.0));
                const int ir_re = n >= nc ? 0 : n + 1, ir_im = n >= nc ? 0 : nc + n + 1;
                const double cr = n >= nc ? 0.0 : (n == 0 ? rs2 : 1.0);
                const int ic_im = n >= 1 ? nc + n : 0;
                const double cc_im = n >= 1 ? 1.0 : 0.0;
                const cplx ec = cmake(re[n], cc_im * re[ic_im]), oc = cmake(ro[n], cc_im * ro[ic_im]);
                const cplx el = cmake(cl_re * re[il_re], cl_im * re[il_im]);
                const cplx ol = cmake(cl_re * ro[il_re], cl_im * ro[il_im]);
                const cplx er = cmake(cr * re[ir_re], cr * re[ir_im]);
                const cplx orr = cmake(cr * ro[ir_re], cr * ro[ir_im]);
                const cplx ci = cadd(ec, oc), cm = csub(ec, oc);
                const cplx li = cadd(el, ol), lm_ = csub(el, ol), rgi = cadd(er, orr), rgm = csub(er, orr);
                const double dcn = SEP ? sDc[n] : 0.0;         // charge part of the LC diagonal (x basis: exact)
                const double di = lin_i + dcn, dm = pairf * (lin_m + dcn);
                cplx oi = cmake(di * ci.x, di * ci.y);
                cplx om = cmake(dm * cm.x, dm * cm.y);
                for (int t = 0; t < gm.n_terms; ++t) {
                    const int sh = ((gm.shpack >> (2 * t)) & 3) - 1;
                    const cplx pi_ = sPh[t * m + i];
                    const cplx pm_ = cscale(pairf, sPh[t * m + im]);
                    // ph * y(n - sh) + conj(ph) * y(n + sh)
                    const cplx ai = sh > 0 ? li : (sh < 0 ? rgi : ci), bi = sh > 0 ? rgi : (sh < 0 ? li : ci);
                    const cplx am = sh > 0 ? lm_ : (sh < 0 ? rgm : cm), bm = sh > 0 ? rgm : (sh < 0 ? lm_ : cm);
                    cfma(oi, pi_, ai);
                    cfma(oi, cconj(pi_), bi);
                    cfma(om, pm_, am);
                    cfma(om, cconj(pm_), bm);
                }
                S[pass] = cadd(oi, om);
                D[pass] = cscale(pairf, csub(oi, om));
            }
            __syncwarp();
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int n = lane + 32 * pass;
                if (n <= nc) {
                    re[n] = S[pass].x;
                    ro[n] = D[pass].x;
                    if (n >= 1) {
                        re[nc + n] = S[pass].y;
                        ro[nc + n] = D[pass].y;
                    }
                }
            }
        }
        HXR_MARK(4);
        team_sync();
        HXR_MARK(5);

        // ---- backward transform + LC diagonal, in place ----
        const double* fd = fdiag ? fdiag + b * fdiag_stride_b : nullptr;
        hxr_transform<KS, false, SEP>(sVe, sVo, buf, gm, col0, gq, tq, (SEP || fd) ? xv : nullptr, fd, sDh);
        HXR_MARK(6);
        fence_async_smem();          // generic-proxy writes -> visible to the bulk store
        team_sync();
        HXR_MARK(7);

        // ---- TMA bulk stores, then the next vector's loads into the same buffer ----
        const long long next = desc[1];
        if (warp == 0) {
            for (int p = lane; p < he; p += 32)
                bulk_store(yv + (long long)p * 2 * inner, buf + p * pp, p == he - 1 ? last_bytes : pair_bytes);
            bulk_commit();
            bulk_wait_read0();       // the stores have read the buffer
            __syncwarp();
            // the padding slot of an odd m received the vector's pad element; rows >= m stay zero because
            // nothing ever writes them
            if (next < total_items) issue_load(next);
        }
        HXR_MARK(8);
#ifdef HXR_PROFILE
        if (prof)
            printf("hxr cycles (team %d warp %d): wait-load %lld  fwd %lld  sync %lld  potential %lld  sync %lld  bwd %lld  "
                   "sync %lld  store+issue(warp0 only) %lld  total %lld\n", team, warp, tp[1] - tp[0], tp[2] - tp[1],
                   tp[3] - tp[2], tp[4] - tp[3], tp[5] - tp[4], tp[6] - tp[5], tp[7] - tp[6], tp[8] - tp[7], tp[8] - tp[0]);
#endif
        item = next;
    }
    // the last bulk stores must be complete before the CTA's shared memory goes away
    if (warp == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

extern "C" int sqc_hx_fused_real(const sqc_space_t* sp, const double* Ve, const double* Vo,
                                 const sqc_potential_t* pot, const double* fdiag, int64_t fdiag_stride_b,
                                 const double* fdiag_sep, sqc_block_t X, sqc_block_t Y, int32_t B, void* stream) {
    if (!sp || !pot || B <= 0 || X.cnt_fixed <= 0 || X.ptr == Y.ptr) return SQC_EINVAL;
    if (sp->n_modes != 2 || !sp->harmonic[0] || sp->harmonic[1]) return SQC_ELIMIT;
    if (pot->n_terms < 0 || pot->n_terms > HXR_MAXT || pot->diag) return SQC_ELIMIT;
    HxrGeom gm;
    gm.m = sp->dims[0];
    gm.inner = sp->dims[1];
    if (gm.m < 2 || gm.m > 128 || !(gm.inner & 1)) return SQC_ELIMIT;
    gm.nc = (gm.inner - 1) / 2;
    gm.he = (gm.m + 1) / 2;
    gm.ho = gm.m / 2;
    const int kp = (gm.he + 3) & ~3;
    gm.N = (long long)gm.m * gm.inner;
    gm.vstride = 2 * ((gm.N + 1) / 2);
    gm.n_terms = pot->n_terms;
    gm.shpack = 0;
    for (int t = 0; t < gm.n_terms; ++t) {
        const int sh = pot->shift[t][1];
        if (sh < -1 || sh > 1) return SQC_ELIMIT;
        gm.shpack |= (sh + 1) << (2 * t);
    }
    gm.nct = (gm.inner + 7) / 8;                 // one warp per 8-column tile
    if (gm.nct > 16 || gm.nc + 1 > 64) return SQC_ELIMIT;
    // pair pitch: even (16-byte aligned bulk copies) and == 4 or 12 (mod 16) doubles, so that the four
    // k-rows of a DMMA B fragment (consecutive pairs) fall on disjoint banks
    gm.pp = 2 * gm.inner;
    while ((gm.pp & 15) != 4 && (gm.pp & 15) != 12) gm.pp += 2;
    // + one pitch of slack: the last column tile reads a few doubles past the end of a row
    gm.buf_doubles = kp * gm.pp + 16;
    const int vld = (kp % 8 == 4) ? kp : kp + 4;
    const int team_doubles = gm.buf_doubles + ((gm.m + 1) & ~1) + 2 * HXR_MAXT * gm.m + 8;
    size_t smem = 0;
    gm.teams = HXR_MAX_WARPS / gm.nct;
    if (gm.teams > HXR_MAX_TEAMS) gm.teams = HXR_MAX_TEAMS;
    for (; gm.teams >= 1; --gm.teams) {
        smem = ((size_t)2 * kp * vld + (size_t)gm.teams * team_doubles + ((gm.m + 1) & ~1) + ((gm.nc + 2) & ~1) +
                HXR_MAX_TEAMS) * sizeof(double);
        if (smem <= 227 * 1024) break;
    }
    if (gm.teams < 1) return SQC_ELIMIT;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const long long total = (long long)B * X.cnt_fixed;
    long long grid = nsm;
    if (grid * gm.teams > total) grid = (total + gm.teams - 1) / gm.teams;
    const unsigned threads = 32u * gm.nct * gm.teams;
#define HXR_LAUNCH(R, SEP)                                                                             \
    {                                                                                                  \
        cudaError_t e = cudaFuncSetAttribute(hx_real_kernel<R, SEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem);                                               \
        if (e != cudaSuccess) return (int)e;                                                           \
        hx_real_kernel<R, SEP><<<(unsigned)grid, threads, smem, (cudaStream_t)stream>>>(               \
            gm, Ve, Vo, pot->lam, (const cplx*)pot->etab, pot->etab_stride_b, pot->g, (const cplx*)pot->a, \
            fdiag, fdiag_stride_b, fdiag_sep, X, Y, X.cnt_fixed, total);                               \
    }
#define HXR_CASE(R)                                                                                    \
    case R:                                                                                            \
        if (fdiag_sep) HXR_LAUNCH(R, true) else HXR_LAUNCH(R, false)                                   \
        break;
    switch (kp / 4) {
        HXR_CASE(1) HXR_CASE(2) HXR_CASE(3) HXR_CASE(4) HXR_CASE(5) HXR_CASE(6) HXR_CASE(7) HXR_CASE(8)
        HXR_CASE(9) HXR_CASE(10) HXR_CASE(11) HXR_CASE(12) HXR_CASE(13) HXR_CASE(14) HXR_CASE(15) HXR_CASE(16)
        default: return SQC_ELIMIT;
    }
#undef HXR_CASE
#undef HXR_LAUNCH
    SQC_LAUNCHED();
    return 0;
}
