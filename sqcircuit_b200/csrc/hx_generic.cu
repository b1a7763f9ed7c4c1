// sqc_hx_generic: H.X in ONE kernel for any layout whose harmonic modes lead the Kronecker order
// (up to four harmonic modes of up to 32 levels each, up to eight charge modes):
//
//     Y = d .* X  +  (x_i V_i) [ P ( (x_i V_i^T) X ) ]
//
// A state vector is a matrix [H][C]: H = product of the harmonic dimensions (rows), C = product of the charge
// dimensions (columns, contiguous in memory).  The mode transforms act on the rows, the junction terms shift
// along the columns.  One CTA (512 threads) owns a tile of ALL rows x a window of columns:
//   1. TMA bulk copies (cp.async.bulk + mbarrier), one per row: window + halo -> shared memory
//   2. forward transforms, one harmonic mode after the other, in place: a thread owns whole fibres (the m
//      amplitudes of one mode at fixed other indices), holds them in registers and writes the m outputs back
//      -- plain FP64 FMAs: B200 issues DFMA at the DMMA rate, and m ~ 10 would waste half of an 8x8x4 tile
//   3. potential: linear terms and junction cosines (phase table per row x unit shifts along the columns,
//      masks per column from the charge indices); every thread first gathers its results in registers, the
//      CTA synchronises, then they overwrite the tile -- no second buffer
//   4. backward transforms (window columns only), epilogue + d .* X (X re-read through L2), TMA bulk stores.
// When the whole vector fits in shared memory (e.g. a 4-mode circuit of dimension 8100) the window is the
// vector and there is no halo; otherwise the halo is the largest column offset of a junction shift.
// HBM traffic: X read once (+ halo), Y written once -- instead of one read + one write per harmonic mode and
// direction plus one for the potential (sqc_mode_transform / sqc_potential_apply).
// Replaces the CSR mat-vec on the Hamiltonian of Circuit.hamiltonian() (circuit.py:2147-2164).
#include "common.cuh"

#define HXG_MAXH 4          // harmonic modes

static int g_hxg_threads = 0;      // 0: choose; 512: force the 512-thread build (dev tool, tools/hxg_bench.py)
extern "C" void sqc_debug_set_hxg_threads(int n) { g_hxg_threads = n; }

struct HxgGeom {
    int nh;                          // harmonic modes with more than one level
    int hm[HXG_MAXH];                // their dimensions
    int hs[HXG_MAXH];                // row stride of the mode inside [H]
    int htoff[HXG_MAXH];             // offset of the mode inside lam / etab rows
    int hvoff[HXG_MAXH];             // offset of V_i inside the concatenated matrices
    int haoff[HXG_MAXH];             // offset of the padded (V, V^T) pair of the mode in shared memory
    int hg[HXG_MAXH];                // index of the mode's linear coefficient in g[b][:]
    int H, C;
    int nc;                          // charge modes with more than one state
    int cd[SQC_MAX_MODES];           // their dimensions
    int cs[SQC_MAX_MODES];           // their column strides
    int n_terms;
    int tsh[SQC_MAX_TERMS][SQC_MAX_MODES];   // shift of term t on charge mode j
    int tdelta[SQC_MAX_TERMS];       // the same shift as a column offset
    unsigned tpos[SQC_MAX_MODES];    // bit t: term t shifts charge mode j by +1
    unsigned tneg[SQC_MAX_MODES];    // bit t: term t shifts charge mode j by -1
    int halo, ncore, ntile, ld;      // tile geometry: window columns, tiles per vector, row pitch (complex)
    int rp;                          // rows per pass of the potential phase = spare rows of the tile
    int tlen, gcount, vlen, alen;
    long long N;
};

__device__ __forceinline__ unsigned hxg_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hxg_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hxg_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void hxg_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hxg_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hxg_mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "HXG_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra HXG_DONE;\n"
        "bra HXG_WAIT;\n"
        "HXG_DONE:\n"
        "}\n" ::"r"(hxg_smem_u32(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void hxg_bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(hxg_smem_u32(dst)), "l"(src), "r"(bytes), "r"(hxg_smem_u32(bar)) : "memory");
}
// TMA bulk copy shared -> global
__device__ __forceinline__ void hxg_bulk_store(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(hxg_smem_u32(src)), "r"(bytes) : "memory");
}

// Thread layout of a phase that walks a [rows] x [n columns] piece of the tile: TC = n rounded up to whole warps
// (at most the CTA), TR = 512 / TC row lanes; thread -> (tr, tc).  No division inside the loops.
struct HxgLanes {
    int TC, TR, tr, tc;
    __device__ __forceinline__ HxgLanes(int n, int tid, int nthreads) {
        TC = min(nthreads, (n + 31) & ~31);
        TR = nthreads / TC;
        tr = tid / TC;
        tc = tid - tr * TC;
        if (tr >= TR) tr = -1;                 // idle thread of this phase
    }
};

// One harmonic mode, all fibres of the tile columns [c_lo, c_hi), in place: out[j] = sum_k A[k][j] in[k] with A the
// padded matrix [m][mp] in shared memory (A = V: to the x basis; A = V^T: back).  A thread owns whole fibres: the m
// inputs sit in registers, two outputs are formed at a time so that one 16-byte load of A feeds four FMAs.
template <int MAXM>
__device__ __forceinline__ void hxg_transform(cplx* __restrict__ T, const double* __restrict__ sA, int m, int mp, int s,
                                              int H, int ld, int c_lo, int c_hi, int tid, int nthreads) {
    const HxgLanes ln(c_hi - c_lo, tid, nthreads);
    if (ln.tr < 0) return;
    const long long step = (long long)s * ld;
    const int nf = H / m;
    if (MAXM <= 12) {
        for (int f = ln.tr; f < nf; f += ln.TR) {
            const int o = f / s, r = f - o * s;
            for (int c = c_lo + ln.tc; c < c_hi; c += ln.TC) {
                cplx* base = T + (long long)(o * m * s + r) * ld + c;
                cplx in[MAXM];
#pragma unroll
                for (int k = 0; k < MAXM; ++k)
                    if (k < m) in[k] = base[k * step];
                for (int j = 0; j < m; j += 2) {
                    const double2* ap = reinterpret_cast<const double2*>(sA + j);
                    const int as = mp >> 1;
                    double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
#pragma unroll
                    for (int k = 0; k < MAXM; ++k) {
                        if (k < m) {
                            const double2 v = ap[k * as];
                            ax = fma(v.x, in[k].x, ax);
                            ay = fma(v.x, in[k].y, ay);
                            bx = fma(v.y, in[k].x, bx);
                            by = fma(v.y, in[k].y, by);
                        }
                    }
                    base[j * step] = cmake(ax, ay);
                    if (j + 1 < m) base[(j + 1) * step] = cmake(bx, by);
                }
            }
        }
    } else {
        // long fibres: real and imaginary parts are separate fibres (half the registers)
        double* Td = reinterpret_cast<double*>(T);
        for (int f = ln.tr; f < nf; f += ln.TR) {
            const int o = f / s, r = f - o * s;
            for (int c = c_lo + ln.tc; c < c_hi; c += ln.TC) {
#pragma unroll 1
                for (int comp = 0; comp < 2; ++comp) {
                    double* base = Td + 2 * ((long long)(o * m * s + r) * ld + c) + comp;
                    double in[MAXM];
#pragma unroll
                    for (int k = 0; k < MAXM; ++k)
                        if (k < m) in[k] = base[2 * k * step];
                    for (int j = 0; j < m; j += 2) {
                        const double2* ap = reinterpret_cast<const double2*>(sA + j);
                        const int as = mp >> 1;
                        double a0 = 0.0, a1 = 0.0;
#pragma unroll
                        for (int k = 0; k < MAXM; ++k) {
                            if (k < m) {
                                const double2 v = ap[k * as];
                                a0 = fma(v.x, in[k], a0);
                                a1 = fma(v.y, in[k], a1);
                            }
                        }
                        base[2 * j * step] = a0;
                        if (j + 1 < m) base[2 * (j + 1) * step] = a1;
                    }
                }
            }
        }
    }
}

template <int MAXDIM>
__device__ __forceinline__ void hxg_transform_any(cplx* T, const double* sA, int m, int mp, int s, int H, int ld,
                                                  int c_lo, int c_hi, int tid, int nthreads) {
    if (m <= 4) hxg_transform<4>(T, sA, m, mp, s, H, ld, c_lo, c_hi, tid, nthreads);
    else if (m <= 6) hxg_transform<6>(T, sA, m, mp, s, H, ld, c_lo, c_hi, tid, nthreads);
    else if (m <= 8) hxg_transform<8>(T, sA, m, mp, s, H, ld, c_lo, c_hi, tid, nthreads);
    else if (m <= 10) hxg_transform<10>(T, sA, m, mp, s, H, ld, c_lo, c_hi, tid, nthreads);
    else if (MAXDIM <= 10) return;
    else if (m <= 12) hxg_transform<12>(T, sA, m, mp, s, H, ld, c_lo, c_hi, tid, nthreads);
    else if (m <= 16) hxg_transform<16>(T, sA, m, mp, s, H, ld, c_lo, c_hi, tid, nthreads);
    else if (m <= 20) hxg_transform<20>(T, sA, m, mp, s, H, ld, c_lo, c_hi, tid, nthreads);
    else if (m <= 24) hxg_transform<24>(T, sA, m, mp, s, H, ld, c_lo, c_hi, tid, nthreads);
    else hxg_transform<32>(T, sA, m, mp, s, H, ld, c_lo, c_hi, tid, nthreads);
}

template <int NTHREADS, int MAXDIM>
__global__ void __launch_bounds__(NTHREADS, 1)
hx_generic_kernel(const __grid_constant__ HxgGeom gm, const double* __restrict__ Vcat, const double* __restrict__ lam,
                  const cplx* __restrict__ etab, long long etab_stride_b, const double* __restrict__ g,
                  const cplx* __restrict__ a, const double* __restrict__ fdiag, long long fdiag_stride_b,
                  sqc_block_t X, sqc_block_t Y, int vmax, long long nvec, long long nitems) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int H = gm.H, C = gm.C, ld = gm.ld, T_ = gm.n_terms;
    // shared memory: tile [H][ld] complex | phase table [T][H] complex | padded V and V^T per mode | lin[H] |
    // column masks | mbarrier
    cplx* tile = reinterpret_cast<cplx*>(smem_raw);           // [H + rp][ld]: final position of the vector
    cplx* tload = tile + (long long)gm.rp * ld;               // where it is loaded and transformed forward
    cplx* sPh = tile + (long long)(H + gm.rp) * ld;
    double* sA = reinterpret_cast<double*>(sPh + (long long)max(T_, 1) * H);
    double* sLin = sA + gm.alen;
    unsigned* sMask = reinterpret_cast<unsigned*>(sLin + H);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(
        (reinterpret_cast<uintptr_t>(sMask + ld) + 7) & ~(uintptr_t)7);

    for (int i = 0; i < gm.nh; ++i) {
        const int m = gm.hm[i], mp = m + (m & 1);
        double* A = sA + gm.haoff[i];
        const double* V = Vcat + gm.hvoff[i];
        for (int e = tid; e < m * mp; e += NTHREADS) {
            const int k = e / mp, j = e - k * mp;
            A[e] = j < m ? V[k * m + j] : 0.0;                    // forward:  out[j] = sum_k V[k][j] in[k]
            A[m * mp + e] = j < m ? V[j * m + k] : 0.0;           // backward: out[j] = sum_k V[j][k] in[k]
        }
    }
    if (tid == 0) hxg_mbar_init(bar, 1);
    __syncthreads();

    int last_b = -1, last_tile = -1;
    unsigned phase = 0;
    // items are tile-major: a CTA's consecutive items share the column window (masks), the point changes
    for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int ti = (int)(item / nvec);
        const long long vec = item - (long long)ti * nvec;
        const int b = (int)(vec / vmax), v = (int)(vec - (long long)b * vmax);
        if (v >= blk_count(X, b)) continue;          // uniform over the CTA
        const cplx* x = blk_vec(X, b, v, gm.N);
        cplx* y = blk_vec(Y, b, v, gm.N);
        // window [c0, c1) of this tile, loaded columns [l0, l1) (window + halo inside the vector)
        const int c0 = ti * gm.ncore, c1 = min(C, c0 + gm.ncore);
        const int l0 = max(0, c0 - gm.halo), l1 = min(C, c1 + gm.halo);
        const int nld = l1 - l0;

        // ---- 1. TMA loads: one bulk copy per row ----
        if (tid < 32) {
            if (tid == 0) hxg_mbar_expect_tx(bar, (unsigned)(H * nld * 16));
            __syncwarp();
            for (int h = tid; h < H; h += 32)
                hxg_bulk_load(tload + (long long)h * ld, x + (long long)h * C + l0, (unsigned)(nld * 16), bar);
        }
        // ---- tables (while the copies fly): per sweep point the phases a_t prod_i etab_t[i][h_i] and the linear
        // term of every row; per tile the in-range masks of every column ----
        if (b != last_b) {
            const cplx* et = etab + b * etab_stride_b;
            const double* gb = g + (long long)b * gm.gcount;
            for (int h = tid; h < H; h += NTHREADS) {
                double lin = gb[0];
                int hi[HXG_MAXH];
#pragma unroll
                for (int i = 0; i < HXG_MAXH; ++i) {
                    if (i < gm.nh) {
                        hi[i] = (h / gm.hs[i]) % gm.hm[i];
                        lin = fma(gb[gm.hg[i]], lam[gm.htoff[i] + hi[i]], lin);
                    }
                }
                sLin[h] = lin;
                for (int t = 0; t < T_; ++t) {
                    cplx ph = a[(long long)b * T_ + t];
#pragma unroll
                    for (int i = 0; i < HXG_MAXH; ++i)
                        if (i < gm.nh) ph = cmul(ph, et[(long long)t * gm.tlen + gm.htoff[i] + hi[i]]);
                    sPh[t * H + h] = ph;
                }
            }
            last_b = b;
        }
        if (ti != last_tile) {
            for (int cl = tid; cl < nld; cl += NTHREADS) {
                int rem = l0 + cl;
                unsigned lo = 0xffffffffu, hi = 0xffffffffu;      // bit t: the shifted index stays inside the space
                for (int j = 0; j < gm.nc; ++j) {
                    const int ij = rem / gm.cs[j];
                    rem -= ij * gm.cs[j];
                    if (ij == 0) { lo &= ~gm.tpos[j]; hi &= ~gm.tneg[j]; }              // ij - 1 / ij + (-1) < 0
                    if (ij == gm.cd[j] - 1) { lo &= ~gm.tneg[j]; hi &= ~gm.tpos[j]; }  // ij + 1 >= dim
                }
                unsigned mk = 0;
                for (int t = 0; t < T_; ++t) mk |= (lo >> t & 1u) << (2 * t) | (hi >> t & 1u) << (2 * t + 1);
                sMask[cl] = mk;
            }
            last_tile = ti;
        }
        hxg_mbar_wait(bar, phase);
        phase ^= 1;
        __syncthreads();

        // ---- 2. forward transforms (all loaded columns: the halo is needed in the x basis too) ----
        for (int i = 0; i < gm.nh; ++i) {
            const int m = gm.hm[i], mp = m + (m & 1);
            hxg_transform_any<MAXDIM>(tload, sA + gm.haoff[i], m, mp, gm.hs[i], H, ld, 0, nld, tid, NTHREADS);
            __syncthreads();
        }

        // ---- 3. potential on the window columns, out of place with ONE spare group of rows: the vector was loaded
        // gm.rp rows below its final position; pass p reads rows [p rp, (p+1) rp) at the loaded position and writes
        // them rp rows higher -- into the spare rows first, then into the rows the previous pass has consumed (the
        // potential couples the columns of one row only).  One barrier per pass, plain loops, no staging registers.
        const int w0 = c0 - l0, nw = c1 - c0;            // window inside the loaded columns
        const HxgLanes ln(nw, tid, NTHREADS);
        for (int h0 = 0; h0 < H; h0 += gm.rp) {
            const int h1 = min(H, h0 + gm.rp);
            if (ln.tr >= 0) {
                for (int h = h0 + ln.tr; h < h1; h += ln.TR) {
                    const cplx* row = tload + (long long)h * ld + w0;
                    cplx* dst = tile + (long long)h * ld + w0;
                    const double lin = sLin[h];
                    // terms in blocks of four: phases and offsets of the row in registers, invalid neighbours (outside
                    // the truncated charge range) enter with coefficient zero through a valid address -- no branches
                    for (int t0 = 0; t0 == 0 || t0 < T_; t0 += 4) {
                        cplx ph[4];
                        int dl[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const bool on = t0 + u < T_;
                            ph[u] = on ? sPh[(t0 + u) * H + h] : cmake(0.0, 0.0);
                            dl[u] = on ? gm.tdelta[t0 + u] : 0;
                        }
                        for (int cc = ln.tc; cc < nw; cc += ln.TC) {
                            cplx acc;
                            if (t0 == 0) {
                                const cplx z = row[cc];
                                acc = cmake(lin * z.x, lin * z.y);
                            } else {
                                acc = dst[cc];
                            }
                            const unsigned mk = sMask[w0 + cc] >> (2 * t0);
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const bool lo = mk >> (2 * u) & 1u, hi = mk >> (2 * u + 1) & 1u;
                                const cplx zl = row[lo ? cc - dl[u] : cc], zh = row[hi ? cc + dl[u] : cc];
                                const cplx pl = lo ? ph[u] : cmake(0.0, 0.0);
                                const cplx pc = hi ? cconj(ph[u]) : cmake(0.0, 0.0);
                                cfma(acc, pl, zl);
                                cfma(acc, pc, zh);
                            }
                            dst[cc] = acc;
                        }
                    }
                }
            }
            __syncthreads();
        }

        // ---- 4. backward transforms on the window, epilogue, TMA stores ----
        for (int i = gm.nh - 1; i >= 0; --i) {
            const int m = gm.hm[i], mp = m + (m & 1);
            hxg_transform_any<MAXDIM>(tile, sA + gm.haoff[i] + m * mp, m, mp, gm.hs[i], H, ld, w0, w0 + nw, tid, NTHREADS);
            __syncthreads();
        }
        if (fdiag && ln.tr >= 0) {
            const double* fd = fdiag + b * fdiag_stride_b;
            for (int h = ln.tr; h < H; h += ln.TR) {
                const long long g0 = (long long)h * C + c0;
                cplx* trow = tile + (long long)h * ld + w0;
#pragma unroll 4
                for (int cc = ln.tc; cc < nw; cc += ln.TC) {
                    const double d = fd[g0 + cc];
                    const cplx xv = x[g0 + cc];
                    cplx val = trow[cc];
                    val.x = fma(d, xv.x, val.x);
                    val.y = fma(d, xv.y, val.y);
                    trow[cc] = val;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid < 32) {
            for (int h = tid; h < H; h += 32)
                hxg_bulk_store(y + (long long)h * C + c0, tile + (long long)h * ld + w0, (unsigned)(nw * 16));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the stores have read the tile
        }
        __syncthreads();
    }
    if (tid < 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

extern "C" int sqc_hx_generic(const sqc_space_t* sp, const double* Vcat, const sqc_potential_t* pot,
                              const double* fdiag, int64_t fdiag_stride_b, sqc_block_t X, sqc_block_t Y,
                              int32_t B, void* stream) {
    if (!sp || !pot || B <= 0 || X.cnt_fixed <= 0 || sp->n_modes < 1 || sp->n_modes > SQC_MAX_MODES) return SQC_EINVAL;
    if (pot->diag || pot->n_terms < 0 || pot->n_terms > SQC_MAX_TERMS || X.ptr == Y.ptr) return SQC_EINVAL;
    HxgGeom gm = {};
    const int M = sp->n_modes;
    // harmonic modes (more than one level) must lead the Kronecker order; modes of dimension one are transparent
    int toff[SQC_MAX_MODES], off = 0;
    bool seen_charge = false;
    long long Hh = 1, Cc = 1;
    for (int i = 0; i < M; ++i) {
        toff[i] = off;
        off += sp->dims[i];
        if (sp->dims[i] <= 1) continue;
        if (sp->harmonic[i]) {
            if (seen_charge || gm.nh >= HXG_MAXH || sp->dims[i] > 32) return SQC_ELIMIT;
            gm.hm[gm.nh] = sp->dims[i];
            gm.htoff[gm.nh] = toff[i];
            gm.hg[gm.nh] = 1 + i;
            ++gm.nh;
            Hh *= sp->dims[i];
        } else {
            seen_charge = true;
            gm.cd[gm.nc] = sp->dims[i];
            for (int t = 0; t < pot->n_terms; ++t) {
                const int w = pot->shift[t][i];
                if (w < -1 || w > 1) return SQC_ELIMIT;
                gm.tsh[t][gm.nc] = w;
            }
            ++gm.nc;
            Cc *= sp->dims[i];
        }
    }
    if (gm.nh == 0 || Hh > 4096 || Cc > (1 << 24)) return SQC_ELIMIT;
    gm.tlen = off;
    gm.gcount = 1 + M;
    gm.H = (int)Hh;
    gm.C = (int)Cc;
    gm.N = Hh * Cc;
    int s = (int)Hh, vo = 0;
    for (int i = 0; i < gm.nh; ++i) {
        s /= gm.hm[i];
        gm.hs[i] = s;
        gm.hvoff[i] = vo;
        vo += gm.hm[i] * gm.hm[i];
        gm.haoff[i] = gm.alen;
        gm.alen += 2 * gm.hm[i] * (gm.hm[i] + (gm.hm[i] & 1));
    }
    gm.vlen = vo;
    int cs = (int)Cc;
    for (int j = 0; j < gm.nc; ++j) {
        cs /= gm.cd[j];
        gm.cs[j] = cs;
    }
    gm.n_terms = pot->n_terms;
    int halo = 0;
    for (int t = 0; t < pot->n_terms; ++t) {
        int d = 0;
        for (int j = 0; j < gm.nc; ++j) {
            d += gm.tsh[t][j] * gm.cs[j];
            if (gm.tsh[t][j] > 0) gm.tpos[j] |= 1u << t;
            if (gm.tsh[t][j] < 0) gm.tneg[j] |= 1u << t;
        }
        gm.tdelta[t] = d;
        halo = max(halo, abs(d));
    }
    // tile geometry: as many window columns as shared memory allows.  Rows of the tile: H + rp (spare rows of the
    // out-of-place potential phase); rp from the thread layout of a window of w columns: whole row lanes, about two
    // elements per thread and pass.
    const long long fixed = (long long)max(gm.n_terms, 1) * gm.H * 16 + (long long)gm.alen * 8 + (long long)gm.H * 8 + 64;
    const long long budget = 227 * 1024 - fixed;
    int maxdim = 0;
    for (int i = 0; i < gm.nh; ++i) maxdim = max(maxdim, gm.hm[i]);
    // Fibres of at most ten levels fit the 64-register budget of a 1024-thread CTA: twice the warps to hide the
    // shared-memory and FP64 latencies of the in-place phases (measured: 2.23 -> 2.03 ms on 6144 vectors of a
    // 10 x 10 x 9 x 9 space; no gain on 10 x 99 x 99, whose ten rows leave one fibre per column).
    const int nthreads = (maxdim <= 10 && Hh >= 64 && g_hxg_threads != 512) ? 1024 : 512;
    auto rows_per_pass = [&](long long w) {
        const long long tc = w >= nthreads ? nthreads : ((w + 31) & ~31LL);
        const long long tr = nthreads / tc, cs = (w + tc - 1) / tc;
        const long long rp = tr * (cs >= 2 ? 1 : 2);
        return (int)(rp < gm.H ? rp : gm.H);
    };
    auto bytes_for = [&](long long ncols, long long w) { return ncols * ((long long)(gm.H + rows_per_pass(w)) * 16 + 4); };
    int ncore, ntile;
    if (bytes_for(gm.C, gm.C) <= budget) {
        ncore = gm.C;
        ntile = 1;
        gm.halo = 0;
    } else {
        // largest window w whose balanced tiling (ncore columns per tile, its own rows per pass) fits the budget
        long long w = budget / ((long long)(gm.H + 1) * 16 + 4) - 2 * halo;
        ncore = ntile = 0;
        for (; w >= halo && w >= 1; --w) {
            const int nt = (int)((gm.C + w - 1) / w);
            const int nc = (gm.C + nt - 1) / nt;
            if (bytes_for(nc + 2 * halo, nc) <= budget) {
                ncore = nc;
                ntile = (gm.C + nc - 1) / nc;
                break;
            }
        }
        if (ncore == 0) return SQC_ELIMIT;       // halo would dominate
        gm.halo = halo;
    }
    gm.ncore = ncore;
    gm.ntile = ntile;
    gm.ld = ntile == 1 ? gm.C : min(gm.C, ncore + 2 * gm.halo);
    gm.rp = rows_per_pass(ncore);
    const size_t smem = (size_t)((long long)(gm.H + gm.rp) * gm.ld * 16 + fixed + (long long)gm.ld * 4 + 16);
    if (smem > 227 * 1024) return SQC_ELIMIT;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long nvec = (long long)B * X.cnt_fixed;
    const long long nitems = nvec * ntile;
    const int grid = (int)(nitems < sms ? nitems : sms);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(hx_generic_kernel<1024, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(hx_generic_kernel<512, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    if (nthreads == 1024)
        hx_generic_kernel<1024, 10><<<grid, 1024, smem, (cudaStream_t)stream>>>(
            gm, Vcat, pot->lam, (const cplx*)pot->etab, pot->etab_stride_b, pot->g, (const cplx*)pot->a, fdiag,
            fdiag_stride_b, X, Y, X.cnt_fixed, nvec, nitems);
    else
        hx_generic_kernel<512, 32><<<grid, 512, smem, (cudaStream_t)stream>>>(
            gm, Vcat, pot->lam, (const cplx*)pot->etab, pot->etab_stride_b, pot->g, (const cplx*)pot->a, fdiag,
            fdiag_stride_b, X, Y, X.cnt_fixed, nvec, nitems);
    SQC_LAUNCHED();
    return 0;
}
