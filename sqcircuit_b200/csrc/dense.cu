// Tall-skinny batched linear algebra of the block-Davidson solver:
//   Gram / Rayleigh-Ritz projections (FP64 tensor cores), basis updates,
//   residual norms, preconditioned residuals, restart copies.
// Replaces the ARPACK call in Circuit._diag_np (reference circuit.py:1962-1969).
#include "common.cuh"

long long g_sqc_launches = 0;

// =====================================================================================
// Gram:  C[b][i][j] = sum_n conj(A[b][i][n]) B[b][j][n]
//
// Real formulation on the interleaved data (K index t runs over the 2N doubles of a vector):
//   Re C = A_r . B_r^T           Im C = A_r . B'_r^T,   B'[2n] = Im b_n, B'[2n+1] = -Re b_n
// so one DMMA B-fragment plus a lane-xor shuffle feeds both the real and the imaginary tile.
// CTA = 4 warps, 112 KB of stages so that two CTAs are always resident per SM, >= 3 cp.async stages sized by the rows the sweep point really
// has (one block sync per chunk: the refill of the slot consumed one iteration earlier is issued
// before the tensor-core work of the current chunk); the warps split the K range of every staged
// chunk and each warp keeps all (NAT x NCT) 8x8 accumulator tiles; a final shared-memory
// reduction combines them.
// =====================================================================================
#ifndef GRAM_THREADS
#define GRAM_THREADS 128
#endif
#define GRAM_WARPS (GRAM_THREADS / 32)
#define GRAM_CH 32                  // complex elements per staged chunk
#define GRAM_LD (2 * GRAM_CH + 4)   // doubles per smem row: 68 == 4 (mod 8) -> conflict-free frags

// Tensor-core work of one staged chunk for exactly NATR populated row tiles.  The row-tile count
// is a template parameter (dispatched by a warp-uniform switch) because a run-time guard is
// compiled to PREDICATED DMMAs, and predicated-off DMMAs still occupy the FP64 tensor pipe.
// B fragment of column tile j: row (brow + bstep * j) of the staged right-hand block, element k0 + bcol.
template <int NATR, int NCTR, int NAT, int NCT, int KW>
__device__ __forceinline__ void gram_chunk(double (&acc)[NAT][NCT][2], const double* __restrict__ sAb,
                                           const double* __restrict__ sBb, int warp, int g, int t,
                                           int bcol, double bsign, int brow, int bstep) {
#pragma unroll
    for (int ks = 0; ks < (2 * GRAM_CH / 4) / KW; ++ks) {
        const int k0 = 4 * (warp + KW * ks);
        double bf[NCTR];
#pragma unroll
        for (int j = 0; j < NCTR; ++j) bf[j] = bsign * sBb[(bstep * j + brow) * GRAM_LD + k0 + bcol];
#pragma unroll
        for (int i = 0; i < NATR; ++i) {
            const double af = sAb[(8 * i + g) * GRAM_LD + k0 + t];
#pragma unroll
            for (int j = 0; j < NCTR; ++j) dmma884(acc[i][j][0], acc[i][j][1], af, bf[j]);
        }
    }
}

// column tiles actually populated (the active right-hand vectors of the sweep point), then row tiles
template <int NATR, int NAT, int NCT, int KW>
__device__ __forceinline__ void gram_chunk_cols(double (&acc)[NAT][NCT][2], const double* __restrict__ sAb,
                                                const double* __restrict__ sBb, int warp, int g, int t,
                                                int bcol, double bsign, int brow, int bstep, int nctr) {
    switch (nctr) {
        case 1: gram_chunk<NATR, 1, NAT, NCT, KW>(acc, sAb, sBb, warp, g, t, bcol, bsign, brow, bstep); break;
        case 2: if constexpr (NCT >= 2) gram_chunk<NATR, 2, NAT, NCT, KW>(acc, sAb, sBb, warp, g, t, bcol, bsign, brow, bstep); break;
        case 3: if constexpr (NCT >= 3) gram_chunk<NATR, 3, NAT, NCT, KW>(acc, sAb, sBb, warp, g, t, bcol, bsign, brow, bstep); break;
        default: if constexpr (NCT >= 4) gram_chunk<NATR, 4, NAT, NCT, KW>(acc, sAb, sBb, warp, g, t, bcol, bsign, brow, bstep); break;
    }
}

// NAT: max row tiles (8 vectors of A each); NCT: column tiles over the 2*nb packed real
// columns (column 2j = Re C[.][j], column 2j+1 = Im C[.][j]).
// PAIR: two B blocks (Bm -> C, Bm2 -> C2) against the same A in ONE pass over A: warps 0,1 take
// Bm, warps 2,3 take Bm2, each pair splitting K two ways.
// REAL: the vectors are real arrays viewed as complex blocks (real representation, hx_real.cu): only the real
// inner products exist, a column tile holds EIGHT right-hand vectors (12 vectors = 2 tiles instead of 3, half
// the DMMAs of the paired kernel) and the results are stored with zero imaginary parts.
template <int NAT, int NCT, bool PAIR, bool REAL>
__global__ void __launch_bounds__(GRAM_THREADS)
gram_dmma_kernel(sqc_block_t A, sqc_block_t Bm, sqc_block_t Bm2, int64_t N, int nsplit,
                 cplx* __restrict__ C, cplx* __restrict__ C2, int ldr, int ldc, cplx* __restrict__ work,
                 int smem_doubles) {
    constexpr int NA = NAT * 8, NB = NCT * (REAL ? 8 : 4);
    constexpr int NBLK = PAIR ? 2 : 1;          // B blocks
    constexpr int KW = GRAM_WARPS / NBLK;       // warps that split K for one B block
    extern __shared__ double smem[];

    const int b = blockIdx.x / nsplit, split = blockIdx.x % nsplit;
    const int na = min(blk_count(A, b), NA), nb = min(blk_count(Bm, b), NB);
    const int nb2 = PAIR ? min(blk_count(Bm2, b), NB) : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int grp = PAIR ? (warp / KW) : 0, kw = warp % KW;

    cplx* outs[2];
    outs[0] = (nsplit == 1) ? C + (int64_t)b * ldr * ldc
                            : work + (((int64_t)b * nsplit + split) * NBLK + 0) * (NA * NB);
    outs[1] = !PAIR ? nullptr
              : (nsplit == 1) ? C2 + (int64_t)b * ldr * ldc
                              : work + (((int64_t)b * nsplit + split) * NBLK + 1) * (NA * NB);
    const int out_ld = (nsplit == 1) ? ldc : NB;
    const int out_rows = (nsplit == 1) ? min(A.cnt_fixed, ldr) : NA;
    const int out_cols = (nsplit == 1) ? min(Bm.cnt_fixed, ldc) : NB;

    if (na == 0 || (nb == 0 && nb2 == 0)) {
        for (int o = 0; o < NBLK; ++o)
            for (int e = tid; e < out_rows * out_cols; e += GRAM_THREADS)
                outs[o][(e / out_cols) * out_ld + (e % out_cols)] = cmake(0.0, 0.0);
        return;
    }
    const int nat = (na + 7) >> 3;          // row tiles actually populated for this sweep point
    // column tiles (4 vectors each) populated in this warp's right-hand block
    const int nctr = REAL ? max(1, ((PAIR && grp == 1 ? nb2 : nb) + 7) >> 3)
                          : max(1, ((PAIR && grp == 1 ? nb2 : nb) + 3) >> 2);
    // stages are sized by the rows this point really has, so small bases get a deeper pipeline
    const int stage_doubles = (nat * 8 + NBLK * NB) * GRAM_LD;
    int nstage = smem_doubles / stage_doubles;     // >= 2 by construction of the launch
    nstage = nstage > 8 ? 8 : nstage;
    const cplx* a0 = blk_vec(A, b, 0, N);
    const cplx* b0 = blk_vec(Bm, b, 0, N);
    const cplx* b1 = PAIR ? blk_vec(Bm2, b, 0, N) : b0;

    // chunks are dealt round-robin to the nsplit CTAs of a sweep point: CTAs with adjacent
    // blockIdx run concurrently and stream ADJACENT 512 B pieces of every vector, so DRAM sees
    // multi-KB bursts per vector instead of isolated 512 B requests
    const int64_t nchunks = (N + GRAM_CH - 1) / GRAM_CH;
    const int64_t c_lo = 0, c_hi = (nchunks - split + nsplit - 1) / nsplit;   // local chunk indices

    double acc[NAT][NCT][2];
#pragma unroll
    for (int i = 0; i < NAT; ++i)
#pragma unroll
        for (int j = 0; j < NCT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // one commit group per call, even when the chunk is past the end (keeps the group count uniform).
    // Every thread owns one column of the chunk and walks down the rows with plain pointer
    // increments (the first version recomputed row/column/validity per 16-byte copy: three times
    // the instructions of the tensor-core work).  Rows that do not exist for this sweep point are
    // zeroed once (below) and never touched again.
    const int scol = tid % GRAM_CH;            // column of this thread inside a chunk
    const int srow0 = tid / GRAM_CH;           // first row, step GRAM_THREADS / GRAM_CH
    constexpr int SROWSTEP = GRAM_THREADS / GRAM_CH;
    auto stage = [&](int64_t chunk) {
        if (chunk < c_hi) {
            double* base = smem + (int)((chunk - c_lo) % nstage) * stage_doubles + 2 * scol;
            const int64_t n = (split + chunk * nsplit) * GRAM_CH + scol;
            if (n < N) {
                const cplx* src = a0 + n;
                for (int r = srow0; r < na; r += SROWSTEP) cp_async16(base + r * GRAM_LD, src + (int64_t)r * N);
                double* bb = base + nat * 8 * GRAM_LD;
                src = b0 + n;
                for (int r = srow0; r < nb; r += SROWSTEP) cp_async16(bb + r * GRAM_LD, src + (int64_t)r * N);
                if (PAIR) {
                    bb += NB * GRAM_LD;
                    src = b1 + n;
                    for (int r = srow0; r < nb2; r += SROWSTEP) cp_async16(bb + r * GRAM_LD, src + (int64_t)r * N);
                }
            } else {
                // tail of the last chunk: this column is past the end of the vectors
                const int rows = nat * 8 + NBLK * NB;
                for (int r = srow0; r < rows; r += SROWSTEP) {
                    base[r * GRAM_LD] = 0.0;
                    base[r * GRAM_LD + 1] = 0.0;
                }
            }
        }
        cp_async_commit();
    };
    auto wait_oldest = [&]() {      // all but the newest (nstage - 2) groups complete
        switch (nstage) {
            case 2: cp_async_wait<0>(); break;
            case 3: cp_async_wait<1>(); break;
            case 4: cp_async_wait<2>(); break;
            case 5: cp_async_wait<3>(); break;
            case 6: cp_async_wait<4>(); break;
            case 7: cp_async_wait<5>(); break;
            default: cp_async_wait<6>(); break;
        }
    };

    // B fragment of packed column n = 8ct + g:  vector j = 4ct + (g>>1), part = g&1
    //   part 0: B[j][k]        part 1: (k even) ? B[j][k+1] : -B[j][k-1]      (k = k0 + t)
    // REAL: packed column n = 8ct + g is vector j = 8ct + g itself
    const int bpart = REAL ? 0 : (g & 1);
    const int bcol = t ^ bpart;
    const double bsign = (bpart && (t & 1)) ? -1.0 : 1.0;
    const int brow = REAL ? g : (g >> 1), bstep = REAL ? 8 : 4;

    for (int e = tid; e < nstage * stage_doubles; e += GRAM_THREADS) smem[e] = 0.0;
    __syncthreads();
    for (int pre = 0; pre < nstage - 1; ++pre) stage(c_lo + pre);
    for (int64_t ch = c_lo; ch < c_hi; ++ch) {
        wait_oldest();
        __syncthreads();            // chunk ch visible; everyone is done with chunk ch - 1
        stage(ch + nstage - 1);     // refill the slot of chunk ch - 1 while computing chunk ch
        const double* sAb = smem + (int)((ch - c_lo) % nstage) * stage_doubles;
        const double* sBb = sAb + (nat * 8 + grp * NB) * GRAM_LD;
        switch (nat) {
            case 1: gram_chunk_cols<1, NAT, NCT, KW>(acc, sAb, sBb, kw, g, t, bcol, bsign, brow, bstep, nctr); break;
            case 2: if constexpr (NAT >= 2) gram_chunk_cols<2, NAT, NCT, KW>(acc, sAb, sBb, kw, g, t, bcol, bsign, brow, bstep, nctr); break;
            case 3: if constexpr (NAT >= 3) gram_chunk_cols<3, NAT, NCT, KW>(acc, sAb, sBb, kw, g, t, bcol, bsign, brow, bstep, nctr); break;
            case 4: if constexpr (NAT >= 4) gram_chunk_cols<4, NAT, NCT, KW>(acc, sAb, sBb, kw, g, t, bcol, bsign, brow, bstep, nctr); break;
            case 5: if constexpr (NAT >= 5) gram_chunk_cols<5, NAT, NCT, KW>(acc, sAb, sBb, kw, g, t, bcol, bsign, brow, bstep, nctr); break;
            case 6: if constexpr (NAT >= 6) gram_chunk_cols<6, NAT, NCT, KW>(acc, sAb, sBb, kw, g, t, bcol, bsign, brow, bstep, nctr); break;
            case 7: if constexpr (NAT >= 7) gram_chunk_cols<7, NAT, NCT, KW>(acc, sAb, sBb, kw, g, t, bcol, bsign, brow, bstep, nctr); break;
            default: if constexpr (NAT >= 8) gram_chunk_cols<8, NAT, NCT, KW>(acc, sAb, sBb, kw, g, t, bcol, bsign, brow, bstep, nctr); break;
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    // cross-warp reduction through shared memory: red[warp][row][NB] complex
    cplx* red = reinterpret_cast<cplx*>(smem);
#pragma unroll
    for (int i = 0; i < NAT; ++i)
#pragma unroll
        for (int j = 0; j < NCT; ++j) {
            if (REAL) {
                red[(warp * NA + 8 * i + g) * NB + 8 * j + 2 * t] = cmake(acc[i][j][0], 0.0);
                red[(warp * NA + 8 * i + g) * NB + 8 * j + 2 * t + 1] = cmake(acc[i][j][1], 0.0);
            } else {
                red[(warp * NA + 8 * i + g) * NB + 4 * j + t] = cmake(acc[i][j][0], acc[i][j][1]);
            }
        }
    __syncthreads();
    for (int o = 0; o < NBLK; ++o) {
        for (int e = tid; e < out_rows * out_cols; e += GRAM_THREADS) {
            const int r = e / out_cols, c = e % out_cols;
            double re = 0.0, im = 0.0;
            if (r < NA && c < NB) {
#pragma unroll
                for (int w = 0; w < KW; ++w) {
                    const cplx v = red[((o * KW + w) * NA + r) * NB + c];
                    re += v.x;
                    im += v.y;
                }
            }
            outs[o][r * out_ld + c] = cmake(re, im);
        }
    }
}

__global__ void gram_reduce_kernel(const cplx* __restrict__ work, int nsplit, int NA, int NB, int nblk,
                                   int blk, cplx* __restrict__ C, int ldr, int ldc, int rows, int cols) {
    const int b = blockIdx.x;
    for (int e = threadIdx.x; e < rows * cols; e += blockDim.x) {
        const int r = e / cols, c = e % cols;
        double re = 0.0, im = 0.0;
        if (r < NA && c < NB)
            for (int s = 0; s < nsplit; ++s) {
                const cplx v = work[(((int64_t)b * nsplit + s) * nblk + blk) * (NA * NB) + r * NB + c];
                re += v.x;
                im += v.y;
            }
        C[(int64_t)b * ldr * ldc + r * ldc + c] = cmake(re, im);
    }
}

// plain SIMT reference (one warp per output element); used by tests to cross-check the DMMA path
__global__ void gram_ref_kernel(sqc_block_t A, sqc_block_t Bm, int64_t N, cplx* __restrict__ C,
                                int ldr, int ldc) {
    const int rows = min(A.cnt_fixed, ldr), cols = min(Bm.cnt_fixed, ldc);
    const int b = blockIdx.x / (rows * cols);
    const int e = blockIdx.x % (rows * cols);
    const int i = e / cols, j = e % cols;
    const int na = blk_count(A, b), nb = blk_count(Bm, b);
    double re = 0.0, im = 0.0;
    if (i < na && j < nb) {
        const cplx* a = blk_vec(A, b, i, N);
        const cplx* c = blk_vec(Bm, b, j, N);
        cplx acc = cmake(0.0, 0.0);
        for (int64_t n = threadIdx.x; n < N; n += blockDim.x) cfmac(acc, a[n], c[n]);
        re = acc.x;
        im = acc.y;
    }
    __shared__ double sr[32], si[32];
    re = warp_sum(re);
    im = warp_sum(im);
    if ((threadIdx.x & 31) == 0) {
        sr[threadIdx.x >> 5] = re;
        si[threadIdx.x >> 5] = im;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0, m = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) {
            r += sr[w];
            m += si[w];
        }
        C[(int64_t)b * ldr * ldc + i * ldc + j] = cmake(r, m);
    }
}

// Small Gram (at most 4 x 4, e.g. the 2 x 2 transition matrix elements of the decoherence rates):
// the tensor-core pipeline above has nothing to amortise its staging over, so this is a plain
// streaming kernel -- one CTA per sweep point, every vector read exactly once.
template <int RA, int RB>
__global__ void __launch_bounds__(256)
gram_small_kernel(sqc_block_t A, sqc_block_t Bm, int64_t N, cplx* __restrict__ C, int ldr, int ldc) {
    const int b = blockIdx.x;
    const int na = min(blk_count(A, b), RA), nb = min(blk_count(Bm, b), RB);
    const cplx* a0 = blk_vec(A, b, 0, N);
    const cplx* b0 = blk_vec(Bm, b, 0, N);
    cplx acc[RA][RB];
#pragma unroll
    for (int i = 0; i < RA; ++i)
#pragma unroll
        for (int j = 0; j < RB; ++j) acc[i][j] = cmake(0.0, 0.0);
    for (int64_t n = threadIdx.x; n < N; n += blockDim.x) {
        cplx av[RA], bv[RB];
#pragma unroll
        for (int i = 0; i < RA; ++i) av[i] = i < na ? a0[(int64_t)i * N + n] : cmake(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < RB; ++j) bv[j] = j < nb ? b0[(int64_t)j * N + n] : cmake(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < RA; ++i)
#pragma unroll
            for (int j = 0; j < RB; ++j) cfmac(acc[i][j], av[i], bv[j]);
    }
    __shared__ double red[8][RA * RB * 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < RA; ++i)
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const double re = warp_sum(acc[i][j].x), im = warp_sum(acc[i][j].y);
            if (lane == 0) {
                red[warp][2 * (i * RB + j)] = re;
                red[warp][2 * (i * RB + j) + 1] = im;
            }
        }
    __syncthreads();
    const int rows = min(A.cnt_fixed, ldr), cols = min(Bm.cnt_fixed, ldc);
    for (int e = threadIdx.x; e < rows * cols; e += blockDim.x) {
        const int i = e / cols, j = e % cols;
        double re = 0.0, im = 0.0;
        if (i < RA && j < RB)
            for (int w = 0; w < 8; ++w) {
                re += red[w][2 * (i * RB + j)];
                im += red[w][2 * (i * RB + j) + 1];
            }
        C[(int64_t)b * ldr * ldc + i * ldc + j] = cmake(re, im);
    }
}

static int g_gram_split_min = 1;
static int gram_nsplit(int B, int64_t N, bool pair = false) {
    // enough CTAs to fill 148 SMs x 2 resident CTAs, at least g_gram_split_min interleaved CTAs per
    // sweep point (measured: the paired kernel gains ~5% from 8 interleaved CTAs per point, the
    // single-block kernel nothing), but never below 8 chunks per CTA
    int want = (148 * 2 + B - 1) / B;
    if (want < g_gram_split_min) want = g_gram_split_min;
    if (pair && want < 8) want = 8;
    int64_t nchunks = (N + GRAM_CH - 1) / GRAM_CH;
    int64_t maxs = nchunks / 8 > 0 ? nchunks / 8 : 1;
    if (want > maxs) want = (int)maxs;
    if (want < 1) want = 1;
    if (want > 64) want = 64;
    return want;
}
extern "C" void sqc_debug_set_gram_split(int v) { g_gram_split_min = v; }
static void gram_tiles(int na_max, int nb_max, int* nat, int* nct) {
    *nat = (na_max + 7) / 8;
    if (*nat < 1) *nat = 1;
    if (*nat > 2 && (*nat & 1)) ++*nat;  // instantiated: 1,2,4,6,8
    if (*nat == 3) *nat = 4;
    *nct = (2 * nb_max + 7) / 8;         // 1..4
    if (*nct < 1) *nct = 1;
}

extern "C" int64_t sqc_gram_work_bytes(int32_t B, int32_t na_max, int32_t nb_max, int64_t N) {
    int nat, nct;
    gram_tiles(na_max, nb_max, &nat, &nct);
    int ns = gram_nsplit(B, N, true);   // upper bound over both kernel flavours
    if (ns == 1) return 0;
    int nbv = nct * 4;                                  // complex kernels: 4 vectors per column tile ...
    if (nbv < ((nb_max + 7) / 8) * 8) nbv = ((nb_max + 7) / 8) * 8;      // ... the real one: 8
    return 2 * (int64_t)B * ns * (nat * 8) * nbv * (int64_t)sizeof(cplx);
}

template <int NAT, int NCT, bool PAIR, bool REAL = false>
static int launch_gram(sqc_block_t A, sqc_block_t Bm, sqc_block_t Bm2, int64_t N, int B, cplx* C, cplx* C2,
                       int ldr, int ldc, cplx* work, cudaStream_t st) {
    constexpr int NBV = NCT * (REAL ? 8 : 4);        // right-hand vectors the column tiles can hold
    const int ns = gram_nsplit(B, N, PAIR);
    if (ns > 1 && !work) return SQC_EINVAL;
    // always 112 KB so that TWO CTAs (8 warps) are resident per SM: one CTA alone cannot keep the
    // FP64 tensor pipe busy.  The kernel cuts it into stages sized by the rows the sweep point
    // really has (>= 2 stages even for a full 64-row basis, >= 3 up to 40 rows + a pair of blocks).
    size_t stage_bytes = (size_t)(NAT * 8 + (PAIR ? 2 : 1) * NBV) * GRAM_LD * sizeof(double);
    size_t smem = 112 * 1024;
    if (smem < 2 * stage_bytes) smem = 2 * stage_bytes;
    size_t red_bytes = GRAM_WARPS * (size_t)(NAT * 8) * NBV * sizeof(cplx);
    if (smem < red_bytes) smem = red_bytes;
    if (smem > 227 * 1024) return SQC_ELIMIT;
    auto kern = gram_dmma_kernel<NAT, NCT, PAIR, REAL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<B * ns, GRAM_THREADS, smem, st>>>(A, Bm, Bm2, N, ns, C, C2, ldr, ldc, work, (int)(smem / sizeof(double)));
    SQC_LAUNCHED();
    if (ns > 1) {
        int rows = A.cnt_fixed < ldr ? A.cnt_fixed : ldr, cols = Bm.cnt_fixed < ldc ? Bm.cnt_fixed : ldc;
        const int nblk = PAIR ? 2 : 1;
        gram_reduce_kernel<<<B, 128, 0, st>>>(work, ns, NAT * 8, NBV, nblk, 0, C, ldr, ldc, rows, cols);
        SQC_LAUNCHED();
        if (PAIR) {
            gram_reduce_kernel<<<B, 128, 0, st>>>(work, ns, NAT * 8, NBV, nblk, 1, C2, ldr, ldc, rows, cols);
            SQC_LAUNCHED();
        }
    }
    return 0;
}

extern "C" int sqc_gram(sqc_block_t A, sqc_block_t Bm, int64_t N, int32_t B, void* C, int32_t ldr,
                        int32_t ldc, void* work, int32_t impl, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || N <= 0 || A.cnt_fixed <= 0 || Bm.cnt_fixed <= 0) return SQC_EINVAL;
    if (impl == 1) {
        int rows = A.cnt_fixed < ldr ? A.cnt_fixed : ldr, cols = Bm.cnt_fixed < ldc ? Bm.cnt_fixed : ldc;
        gram_ref_kernel<<<B * rows * cols, 128, 0, st>>>(A, Bm, N, (cplx*)C, ldr, ldc);
        SQC_LAUNCHED();
        return 0;
    }
    if (A.cnt_fixed <= 4 && Bm.cnt_fixed <= 4) {
        const int ra = A.cnt_fixed <= 1 ? 1 : A.cnt_fixed <= 2 ? 2 : 4, rb = Bm.cnt_fixed <= 1 ? 1 : Bm.cnt_fixed <= 2 ? 2 : 4;
#define GRAMS_CASE(a, b_) \
    if (ra == a && rb == b_) gram_small_kernel<a, b_><<<B, 256, 0, st>>>(A, Bm, N, (cplx*)C, ldr, ldc)
        GRAMS_CASE(1, 1); GRAMS_CASE(1, 2); GRAMS_CASE(1, 4); GRAMS_CASE(2, 1); GRAMS_CASE(2, 2); GRAMS_CASE(2, 4);
        GRAMS_CASE(4, 1); GRAMS_CASE(4, 2); GRAMS_CASE(4, 4);
#undef GRAMS_CASE
        SQC_LAUNCHED();
        return 0;
    }
    if (A.cnt_fixed > 64 || Bm.cnt_fixed > 16) return SQC_ELIMIT;
    int nat, nct;
    gram_tiles(A.cnt_fixed, Bm.cnt_fixed, &nat, &nct);
#define GRAM_CASE(a, b_) \
    if (nat == a && nct == b_) return launch_gram<a, b_, false>(A, Bm, Bm, N, B, (cplx*)C, nullptr, ldr, ldc, (cplx*)work, st)
    GRAM_CASE(1, 1); GRAM_CASE(2, 1); GRAM_CASE(4, 1); GRAM_CASE(6, 1); GRAM_CASE(8, 1);
    GRAM_CASE(1, 2); GRAM_CASE(2, 2); GRAM_CASE(4, 2); GRAM_CASE(6, 2); GRAM_CASE(8, 2);
    GRAM_CASE(1, 3); GRAM_CASE(2, 3); GRAM_CASE(4, 3); GRAM_CASE(6, 3); GRAM_CASE(8, 3);
    GRAM_CASE(1, 4); GRAM_CASE(2, 4); GRAM_CASE(4, 4); GRAM_CASE(6, 4); GRAM_CASE(8, 4);
#undef GRAM_CASE
    return SQC_ELIMIT;
}

// One pass over A for two right-hand blocks: C = A^H Bm, C2 = A^H Bm2 (the new columns of the
// basis Gram matrix and of the projected Hamiltonian in the non-orthogonalised Davidson).
extern "C" int sqc_gram_pair(sqc_block_t A, sqc_block_t Bm, sqc_block_t Bm2, int64_t N, int32_t B, void* C,
                             void* C2, int32_t ldr, int32_t ldc, void* work, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || N <= 0 || A.cnt_fixed <= 0 || Bm.cnt_fixed <= 0 || Bm.cnt_fixed != Bm2.cnt_fixed) return SQC_EINVAL;
    if (A.cnt_fixed > 64 || Bm.cnt_fixed > 16) return SQC_ELIMIT;
    int nat, nct;
    gram_tiles(A.cnt_fixed, Bm.cnt_fixed, &nat, &nct);
#define GRAMP_CASE(a, b_) \
    if (nat == a && nct == b_) return launch_gram<a, b_, true>(A, Bm, Bm2, N, B, (cplx*)C, (cplx*)C2, ldr, ldc, (cplx*)work, st)
    GRAMP_CASE(1, 1); GRAMP_CASE(2, 1); GRAMP_CASE(4, 1); GRAMP_CASE(6, 1); GRAMP_CASE(8, 1);
    GRAMP_CASE(1, 2); GRAMP_CASE(2, 2); GRAMP_CASE(4, 2); GRAMP_CASE(6, 2); GRAMP_CASE(8, 2);
    GRAMP_CASE(1, 3); GRAMP_CASE(2, 3); GRAMP_CASE(4, 3); GRAMP_CASE(6, 3); GRAMP_CASE(8, 3);
    GRAMP_CASE(1, 4); GRAMP_CASE(2, 4); GRAMP_CASE(4, 4); GRAMP_CASE(6, 4); GRAMP_CASE(8, 4);
#undef GRAMP_CASE
    return SQC_ELIMIT;
}

// sqc_gram_pair for blocks of the real representation (real arrays viewed as complex blocks, N = complex
// elements per vector): C = Re(A^H Bm), C2 = Re(A^H Bm2) with zero imaginary parts.
extern "C" int sqc_gram_pair_real(sqc_block_t A, sqc_block_t Bm, sqc_block_t Bm2, int64_t N, int32_t B, void* C,
                                  void* C2, int32_t ldr, int32_t ldc, void* work, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || N <= 0 || A.cnt_fixed <= 0 || Bm.cnt_fixed <= 0 || Bm.cnt_fixed != Bm2.cnt_fixed) return SQC_EINVAL;
    if (A.cnt_fixed > 64 || Bm.cnt_fixed > 16) return SQC_ELIMIT;
    int nat, nct;
    gram_tiles(A.cnt_fixed, Bm.cnt_fixed, &nat, &nct);
    nct = (Bm.cnt_fixed + 7) / 8;        // 8 real right-hand vectors per column tile
#define GRAMR_CASE(a, b_) \
    if (nat == a && nct == b_) return launch_gram<a, b_, true, true>(A, Bm, Bm2, N, B, (cplx*)C, (cplx*)C2, ldr, ldc, (cplx*)work, st)
    GRAMR_CASE(1, 1); GRAMR_CASE(2, 1); GRAMR_CASE(4, 1); GRAMR_CASE(6, 1); GRAMR_CASE(8, 1);
    GRAMR_CASE(1, 2); GRAMR_CASE(2, 2); GRAMR_CASE(4, 2); GRAMR_CASE(6, 2); GRAMR_CASE(8, 2);
#undef GRAMR_CASE
    return SQC_ELIMIT;
}

// =====================================================================================
// Block update:  Out[b][j][n] (op)= sum_i Cf[b][j][i] A[b][i][n]
// One thread per n; NB accumulators in registers; coefficients broadcast from shared memory.
// =====================================================================================
#define UPD_THREADS 128

template <int NB>
__global__ void __launch_bounds__(UPD_THREADS)
block_update_kernel(sqc_block_t A, const cplx* __restrict__ Cf, int ldr, int ldc, int trans,
                    sqc_block_t Out, int64_t N, int nchunk, int op) {
    extern __shared__ double smem[];
    cplx* sC = reinterpret_cast<cplx*>(smem);  // [na][NB]  (i-major so that j is contiguous)
    const int b = blockIdx.x / nchunk;
    const int64_t n = (int64_t)(blockIdx.x % nchunk) * UPD_THREADS + threadIdx.x;
    const int na = blk_count(A, b), nb = min(blk_count(Out, b), NB);
    if (nb == 0) return;
    if (na == 0 && op == 1) return;
    const cplx* cf = Cf + (int64_t)b * ldr * ldc;
    for (int e = threadIdx.x; e < na * NB; e += UPD_THREADS) {
        const int i = e / NB, j = e % NB;
        sC[e] = (j < nb) ? (trans ? cf[i * ldc + j] : cf[j * ldc + i]) : cmake(0.0, 0.0);
    }
    __syncthreads();
    if (n >= N) return;
    cplx acc[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[j] = cmake(0.0, 0.0);
    const cplx* a = blk_vec(A, b, 0, N) + n;
#pragma unroll 2
    for (int i = 0; i < na; ++i) {
        const cplx av = a[(int64_t)i * N];
#pragma unroll
        for (int j = 0; j < NB; ++j) cfma(acc[j], sC[i * NB + j], av);
    }
    cplx* o = blk_vec(Out, b, 0, N) + n;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        if (j < nb) {
            if (op == 1) {
                cplx v = o[(int64_t)j * N];
                o[(int64_t)j * N] = csub(v, acc[j]);
            } else {
                o[(int64_t)j * N] = acc[j];
            }
        }
    }
}

// Tensor-core version.  Real formulation (M = positions n, K = (vector i, re/im), N = packed
// (output j, re/im)):   [Re Out, Im Out][n][j] = sum_i [Re A, Im A][i][n] . [[cre, cim], [-cim, cre]]
// A warp owns 32 consecutive positions (4 M-tiles) and all outputs; A fragments are read straight
// from global memory (per k-step two vectors x 512 contiguous bytes), the packed coefficient
// matrix sits in shared memory.
#define UPD2_THREADS 256
#define UPD2_MT 4
#define UPD2_DEPTH 10            // ring slots per warp (k-steps in flight = DEPTH - 1)
#define UPD2_VS 72                // doubles per staged vector piece: 32 positions x (re, im) + 8 of padding, so the two
                                  // vectors of a k-step sit 16 banks apart (64 would put them on the same banks)
#define UPD2_SLOT (2 * UPD2_VS)   // doubles per slot: 2 vectors

template <int NCT>
__global__ void __launch_bounds__(UPD2_THREADS, 2)
block_update_dmma_kernel(sqc_block_t A, const cplx* __restrict__ Cf, int ldr, int ldc, int trans,
                         sqc_block_t Out, int64_t N, int nchunk, int op, int ldm, int kk_max) {
    extern __shared__ double smem[];     // Mr[kk_max][ldm] | ring[warps][DEPTH][SLOT]
    const int b = blockIdx.x / nchunk;
    const int na = blk_count(A, b), nb = min(blk_count(Out, b), NCT * 4);
    if (nb == 0) return;
    if (na == 0 && op == 1) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const cplx* cf = Cf + (int64_t)b * ldr * ldc;
    const int kk = (2 * na + 3) & ~3;
    for (int e = tid; e < kk * (NCT * 8); e += UPD2_THREADS) {
        const int k = e / (NCT * 8), n = e % (NCT * 8);
        const int i = k >> 1, pin = k & 1, j = n >> 1, pout = n & 1;
        double v = 0.0;
        if (i < na && j < nb) {
            const cplx c = trans ? cf[i * ldc + j] : cf[j * ldc + i];
            v = (pin == pout) ? c.x : (pin == 0 ? c.y : -c.y);
        }
        smem[k * ldm + n] = v;
    }
    __syncthreads();
    const int64_t nbase = ((int64_t)(blockIdx.x % nchunk) * (UPD2_THREADS / 32) + warp) * (8 * UPD2_MT);
    if (nbase >= N) return;
    double* ring = smem + kk_max * ldm + warp * (UPD2_DEPTH * UPD2_SLOT);
    const cplx* a = blk_vec(A, b, 0, N);
    const int nks = kk / 4;                 // k-steps; k-step s uses vectors 2s, 2s+1
    // slot layout: [vector (2)][position (32)][re, im]; every lane copies its position of both vectors
    auto issue = [&](int s) {
        if (s < nks) {
            double* dst = ring + (s % UPD2_DEPTH) * UPD2_SLOT + 2 * lane;
            const int64_t n = nbase + lane;
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const int i = 2 * s + v;
                if (i < na && n < N) {
                    cp_async16(dst + v * UPD2_VS, a + (int64_t)i * N + n);
                } else {
                    dst[v * UPD2_VS] = 0.0;
                    dst[v * UPD2_VS + 1] = 0.0;
                }
            }
        }
        cp_async_commit();
    };
    double acc[UPD2_MT][NCT][2];
#pragma unroll
    for (int mt = 0; mt < UPD2_MT; ++mt)
#pragma unroll
        for (int j = 0; j < NCT; ++j) acc[mt][j][0] = acc[mt][j][1] = 0.0;
#pragma unroll
    for (int s = 0; s < UPD2_DEPTH - 1; ++s) issue(s);
    for (int s = 0; s < nks; ++s) {
        cp_async_wait<UPD2_DEPTH - 2>();    // the group of k-step s has landed (for this lane)
        __syncwarp();
        const double* slot = ring + (s % UPD2_DEPTH) * UPD2_SLOT;
        // A fragment: row n = 8 mt + g, k = (vector t>>1, part t&1)
        double af[UPD2_MT];
#pragma unroll
        for (int mt = 0; mt < UPD2_MT; ++mt) af[mt] = slot[(t >> 1) * UPD2_VS + 2 * (8 * mt + g) + (t & 1)];
#pragma unroll
        for (int j = 0; j < NCT; ++j) {
            const double bf = smem[(4 * s + t) * ldm + 8 * j + g];
#pragma unroll
            for (int mt = 0; mt < UPD2_MT; ++mt) dmma884(acc[mt][j][0], acc[mt][j][1], af[mt], bf);
        }
        __syncwarp();                       // slot may be overwritten by the next issue
        issue(s + UPD2_DEPTH - 1);
    }
    cp_async_wait<0>();
    // C fragment: row n = nbase + 8 mt + g, packed cols 8j + 2t, +1  -> Out[4j + t][n] (re, im)
    cplx* o = blk_vec(Out, b, 0, N);
#pragma unroll
    for (int mt = 0; mt < UPD2_MT; ++mt) {
        const int64_t n = nbase + 8 * mt + g;
        if (n < N) {
#pragma unroll
            for (int j = 0; j < NCT; ++j) {
                const int jj = 4 * j + t;
                if (jj < nb) {
                    cplx* p = o + (int64_t)jj * N + n;
                    cplx val = cmake(acc[mt][j][0], acc[mt][j][1]);
                    if (op == 1) val = csub(*p, val);
                    *p = val;
                }
            }
        }
    }
}

extern "C" int sqc_block_update(sqc_block_t A, const void* Cf, int32_t ldr, int32_t ldc, int32_t trans,
                                sqc_block_t Out, int64_t N, int32_t B, int32_t op, int32_t impl,
                                void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || N <= 0 || Out.cnt_fixed <= 0 || Out.cnt_fixed > 16 || A.cnt_fixed > 128) return SQC_EINVAL;
    const int nbm = Out.cnt_fixed;
    if (impl == 0 && A.cnt_fixed <= 64) {
        const int per_cta = (UPD2_THREADS / 32) * 8 * UPD2_MT;
        const int nchunk = (int)((N + per_cta - 1) / per_cta);
        const int64_t grid = (int64_t)B * nchunk;
        if (grid > 0x7fffffffLL) return SQC_ELIMIT;
        const int nct = (2 * nbm + 7) / 8;
        int ldm = nct * 8;
        while ((ldm & 15) != 4) ++ldm;
        const int kk = (2 * (A.cnt_fixed > 0 ? A.cnt_fixed : 1) + 3) & ~3;
        size_t smem = ((size_t)kk * ldm + (size_t)(UPD2_THREADS / 32) * UPD2_DEPTH * UPD2_SLOT) * sizeof(double);
#define UPD2_CASE(NCTv)                                                                               \
    if (nct == NCTv) {                                                                                \
        auto kern = block_update_dmma_kernel<NCTv>;                                                   \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return (int)e;                                                          \
        kern<<<(unsigned)grid, UPD2_THREADS, smem, st>>>(A, (const cplx*)Cf, ldr, ldc, trans, Out, N, nchunk, op, ldm, kk); \
        SQC_LAUNCHED();                                                                               \
        return 0;                                                                                     \
    }
        UPD2_CASE(1) UPD2_CASE(2) UPD2_CASE(3) UPD2_CASE(4)
#undef UPD2_CASE
        return SQC_ELIMIT;
    }
    const int nchunk = (int)((N + UPD_THREADS - 1) / UPD_THREADS);
    const int64_t grid = (int64_t)B * nchunk;
    if (grid > 0x7fffffffLL) return SQC_ELIMIT;
#define UPD_CASE(NBv)                                                                              \
    if (nbm <= NBv) {                                                                              \
        size_t smem = (size_t)(A.cnt_fixed > 0 ? A.cnt_fixed : 1) * NBv * sizeof(cplx);             \
        auto kern = block_update_kernel<NBv>;                                                      \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return (int)e;                                                       \
        kern<<<(unsigned)grid, UPD_THREADS, smem, st>>>(A, (const cplx*)Cf, ldr, ldc, trans, Out, N, nchunk, op); \
        SQC_LAUNCHED();                                                                            \
        return 0;                                                                                  \
    }
    UPD_CASE(4) UPD_CASE(8) UPD_CASE(12) UPD_CASE(16)
#undef UPD_CASE
    return SQC_ELIMIT;
}

// ---------------------------------------------------------------------------------------
// Ritz block in ONE pass over the basis and its image:  X = Cf . V,  HX = Cf . W  and the squared
// residual norms rn2[j] = || HX_j - theta_j X_j ||^2 (replaces two sqc_block_update calls and
// sqc_residual_norms: the Ritz vectors are not read back from HBM for their residuals).
// Same real DMMA formulation as block_update_dmma_kernel; a warp owns 32 positions and streams its
// piece of the basis first and of the image second, keeping both accumulator sets in registers.
// Residuals: per-CTA partial sums -> work[b][chunk][16], summed by a second tiny kernel in a fixed
// order (no atomics: the convergence decisions stay reproducible run to run).
// ---------------------------------------------------------------------------------------
#define UPD3_THREADS 256
#define UPD3_MT 4
#define UPD3_DEPTH 9             // 9 x 8 warps x 1.15 KB + coefficients stays under 113 KB: two CTAs per SM
#define UPD3_SLOT (2 * UPD2_VS)   // doubles per slot: 2 vectors x (32 positions x (re, im) + padding)

template <int NCT>
__global__ void __launch_bounds__(UPD3_THREADS, 2)
ritz_update_dmma_kernel(sqc_block_t A, sqc_block_t A2, const cplx* __restrict__ Cf, int ldr, int ldc,
                        sqc_block_t Out, sqc_block_t Out2, const double* __restrict__ theta, int ld_theta,
                        double* __restrict__ work, int64_t N, int nchunk, int ldm, int kk_max) {
    extern __shared__ double smem[];     // Mr[kk_max][ldm] | ring[warps][DEPTH][SLOT] | red[warps][16]
    const int b = blockIdx.x / nchunk, chunk = blockIdx.x % nchunk;
    const int na = blk_count(A, b), nb = min(blk_count(Out, b), NCT * 4);
    if (nb == 0) return;                 // inactive sweep point: nothing is written, not even partials
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const cplx* cf = Cf + (int64_t)b * ldr * ldc;
    const int kk = (2 * na + 3) & ~3;
    for (int e = tid; e < kk * (NCT * 8); e += UPD3_THREADS) {
        const int k = e / (NCT * 8), n = e % (NCT * 8);
        const int i = k >> 1, pin = k & 1, j = n >> 1, pout = n & 1;
        double v = 0.0;
        if (i < na && j < nb) {
            const cplx c = cf[j * ldc + i];
            v = (pin == pout) ? c.x : (pin == 0 ? c.y : -c.y);
        }
        smem[k * ldm + n] = v;
    }
    double* red = smem + kk_max * ldm + (UPD3_THREADS / 32) * (UPD3_DEPTH * UPD3_SLOT);
    __syncthreads();
    const int64_t nbase = ((int64_t)chunk * (UPD3_THREADS / 32) + warp) * (8 * UPD3_MT);
    double part[NCT];
#pragma unroll
    for (int j = 0; j < NCT; ++j) part[j] = 0.0;
    if (nbase < N) {
        double* ring = smem + kk_max * ldm + warp * (UPD3_DEPTH * UPD3_SLOT);
        // the warp streams its 32 positions of the basis first, then of the image: one long
        // cp.async ring over 2 * nks k-steps (k-step s of a matrix uses its vectors 2s, 2s+1)
        const cplx* a0 = blk_vec(A, b, 0, N);
        const cplx* a1 = blk_vec(A2, b, 0, N);
        const int nks = kk / 4;
        auto issue = [&](int s) {
            if (s < 2 * nks) {
                const cplx* a = s < nks ? a0 : a1;
                const int sl = s < nks ? s : s - nks;
                double* dst = ring + (s % UPD3_DEPTH) * UPD3_SLOT + 2 * lane;
                const int64_t n = nbase + lane;
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const int i = 2 * sl + v;
                    if (i < na && n < N) {
                        cp_async16(dst + v * UPD2_VS, a + (int64_t)i * N + n);
                    } else {
                        dst[v * UPD2_VS] = 0.0;
                        dst[v * UPD2_VS + 1] = 0.0;
                    }
                }
            }
            cp_async_commit();
        };
        double acc[2][UPD3_MT][NCT][2];
#pragma unroll
        for (int w = 0; w < 2; ++w)
#pragma unroll
            for (int mt = 0; mt < UPD3_MT; ++mt)
#pragma unroll
                for (int j = 0; j < NCT; ++j) acc[w][mt][j][0] = acc[w][mt][j][1] = 0.0;
#pragma unroll
        for (int s = 0; s < UPD3_DEPTH - 1; ++s) issue(s);
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            for (int sl = 0; sl < nks; ++sl) {
                const int s = w * nks + sl;
                cp_async_wait<UPD3_DEPTH - 2>();
                __syncwarp();
                const double* slot = ring + (s % UPD3_DEPTH) * UPD3_SLOT;
                // A fragment: row n = 8 mt + g, k = (vector t>>1, part t&1)
                double af[UPD3_MT];
#pragma unroll
                for (int mt = 0; mt < UPD3_MT; ++mt) af[mt] = slot[(t >> 1) * UPD2_VS + 2 * (8 * mt + g) + (t & 1)];
#pragma unroll
                for (int j = 0; j < NCT; ++j) {
                    const double bf = smem[(4 * sl + t) * ldm + 8 * j + g];
#pragma unroll
                    for (int mt = 0; mt < UPD3_MT; ++mt) dmma884(acc[w][mt][j][0], acc[w][mt][j][1], af[mt], bf);
                }
                __syncwarp();
                issue(s + UPD3_DEPTH - 1);
            }
        }
        cp_async_wait<0>();
        // C fragment: row n = nbase + 8 mt + g, packed cols 8j + 2t, +1  -> vector 4j + t (re, im)
        cplx* o = blk_vec(Out, b, 0, N);
        cplx* o2 = blk_vec(Out2, b, 0, N);
#pragma unroll
        for (int j = 0; j < NCT; ++j) {
            const int jj = 4 * j + t;
            if (jj < nb) {
                const double th = theta[(int64_t)b * ld_theta + jj];
#pragma unroll
                for (int mt = 0; mt < UPD3_MT; ++mt) {
                    const int64_t n = nbase + 8 * mt + g;
                    if (n < N) {
                        const cplx x = cmake(acc[0][mt][j][0], acc[0][mt][j][1]);
                        const cplx h = cmake(acc[1][mt][j][0], acc[1][mt][j][1]);
                        o[(int64_t)jj * N + n] = x;
                        o2[(int64_t)jj * N + n] = h;
                        const double rr = h.x - th * x.x, ri = h.y - th * x.y;
                        part[j] = fma(rr, rr, part[j]);
                        part[j] = fma(ri, ri, part[j]);
                    }
                }
            }
        }
    }
    // sum over the 8 row groups g of the warp (lanes with equal t), then over the warps of the CTA
#pragma unroll
    for (int j = 0; j < NCT; ++j) {
        part[j] += __shfl_xor_sync(0xffffffffu, part[j], 4);
        part[j] += __shfl_xor_sync(0xffffffffu, part[j], 8);
        part[j] += __shfl_xor_sync(0xffffffffu, part[j], 16);
        if (g == 0) red[warp * 16 + 4 * j + t] = part[j];
    }
    __syncthreads();
    if (tid < NCT * 4) {
        double sacc = 0.0;
#pragma unroll
        for (int w = 0; w < UPD3_THREADS / 32; ++w) sacc += red[w * 16 + tid];
        work[((int64_t)b * nchunk + chunk) * 16 + tid] = sacc;
    }
}

__global__ void ritz_rn2_reduce_kernel(const double* __restrict__ work, int nchunk, sqc_block_t Out, int p,
                                       double* __restrict__ rn2) {
    const int b = blockIdx.x, j = threadIdx.x;
    if (j >= p) return;
    const int nb = blk_count(Out, b);
    if (nb == 0) return;                 // inactive point: rn2 keeps its last values
    double sacc = 0.0;
    if (j < nb)
        for (int c = 0; c < nchunk; ++c) sacc += work[((int64_t)b * nchunk + c) * 16 + j];
    rn2[(int64_t)b * p + j] = sacc;
}

static int ritz_nchunk(int64_t N) {
    const int per_cta = (UPD3_THREADS / 32) * 8 * UPD3_MT;
    return (int)((N + per_cta - 1) / per_cta);
}

extern "C" int64_t sqc_ritz_update_work_bytes(int32_t B, int64_t N) {
    return (int64_t)B * ritz_nchunk(N) * 16 * (int64_t)sizeof(double);
}

extern "C" int sqc_ritz_update(sqc_block_t A, sqc_block_t A2, const void* Cf, int32_t ldr, int32_t ldc,
                               sqc_block_t Out, sqc_block_t Out2, const double* theta, int32_t ld_theta,
                               double* rn2, void* work, int64_t N, int32_t B, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || N <= 0 || Out.cnt_fixed <= 0 || Out.cnt_fixed > 16 || A.cnt_fixed > 64 || !work || !rn2 || !theta)
        return SQC_EINVAL;
    if (A.cnt_fixed != A2.cnt_fixed || Out.cnt_fixed != Out2.cnt_fixed) return SQC_EINVAL;
    const int nchunk = ritz_nchunk(N);
    const int64_t grid = (int64_t)B * nchunk;
    if (grid > 0x7fffffffLL) return SQC_ELIMIT;
    const int nct = (2 * Out.cnt_fixed + 7) / 8;
    int ldm = nct * 8;
    while ((ldm & 15) != 4) ++ldm;
    const int kk = (2 * (A.cnt_fixed > 0 ? A.cnt_fixed : 1) + 3) & ~3;
    size_t smem = ((size_t)kk * ldm + (size_t)(UPD3_THREADS / 32) * UPD3_DEPTH * UPD3_SLOT + (UPD3_THREADS / 32) * 16) *
                  sizeof(double);
#define UPD3_CASE(NCTv)                                                                               \
    if (nct == NCTv) {                                                                                \
        auto kern = ritz_update_dmma_kernel<NCTv>;                                                    \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return (int)e;                                                          \
        kern<<<(unsigned)grid, UPD3_THREADS, smem, st>>>(A, A2, (const cplx*)Cf, ldr, ldc, Out, Out2, theta, ld_theta, \
                                                         (double*)work, N, nchunk, ldm, kk);          \
        SQC_LAUNCHED();                                                                               \
        ritz_rn2_reduce_kernel<<<B, 32, 0, st>>>((const double*)work, nchunk, Out, Out.cnt_fixed, rn2); \
        SQC_LAUNCHED();                                                                               \
        return 0;                                                                                     \
    }
    UPD3_CASE(1) UPD3_CASE(2) UPD3_CASE(3) UPD3_CASE(4)
#undef UPD3_CASE
    return SQC_ELIMIT;
}

// In-place right multiplication of a (<=16 vector) block by a small matrix.
__global__ void __launch_bounds__(UPD_THREADS)
block_rightmul_kernel(sqc_block_t T, const cplx* __restrict__ Cq, int ldr, int ldc,
                      const int32_t* __restrict__ nout, int64_t N, int nchunk) {
    __shared__ cplx sC[16 * 16];  // [j][i]
    const int b = blockIdx.x / nchunk;
    const int64_t n = (int64_t)(blockIdx.x % nchunk) * UPD_THREADS + threadIdx.x;
    const int nin = min(blk_count(T, b), 16), no = min(nout[b], 16);
    if (nin == 0) return;
    const cplx* cq = Cq + (int64_t)b * ldr * ldc;
    for (int e = threadIdx.x; e < 256; e += UPD_THREADS) {
        const int j = e >> 4, i = e & 15;
        sC[e] = (j < no && i < nin) ? cq[j * ldc + i] : cmake(0.0, 0.0);
    }
    __syncthreads();
    if (n >= N) return;
    cplx* tp = blk_vec(T, b, 0, N) + n;
    cplx tin[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) tin[i] = (i < nin) ? tp[(int64_t)i * N] : cmake(0.0, 0.0);
    for (int j = 0; j < nin; ++j) {   // vectors j >= nout are zeroed (dropped directions)
        cplx acc = cmake(0.0, 0.0);
        if (j < no) {
#pragma unroll
            for (int i = 0; i < 16; ++i) cfma(acc, sC[j * 16 + i], tin[i]);
        }
        tp[(int64_t)j * N] = acc;
    }
}

extern "C" int sqc_block_rightmul(sqc_block_t T, const void* Cq, int32_t ldr, int32_t ldc,
                                  const int32_t* nout, int64_t N, int32_t B, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || N <= 0 || T.cnt_fixed <= 0 || T.cnt_fixed > 16 || !nout) return SQC_EINVAL;
    const int nchunk = (int)((N + UPD_THREADS - 1) / UPD_THREADS);
    const int64_t grid = (int64_t)B * nchunk;
    if (grid > 0x7fffffffLL) return SQC_ELIMIT;
    block_rightmul_kernel<<<(unsigned)grid, UPD_THREADS, 0, st>>>(T, (const cplx*)Cq, ldr, ldc, nout, N, nchunk);
    SQC_LAUNCHED();
    return 0;
}

// =====================================================================================
// residual norms, preconditioned residuals, restart copies, diag multiply
// =====================================================================================
__global__ void __launch_bounds__(256)
residual_norms_kernel(sqc_block_t X, sqc_block_t HX, const double* __restrict__ theta, int ld_theta,
                      int64_t N, int p, double* __restrict__ rn2) {
    const int b = blockIdx.x / p, j = blockIdx.x % p;
    double s = 0.0;
    if (j < blk_count(X, b)) {
        const double th = theta[(int64_t)b * ld_theta + j];
        const cplx* x = blk_vec(X, b, j, N);
        const cplx* hx = blk_vec(HX, b, j, N);
        for (int64_t n = threadIdx.x; n < N; n += 256) {
            const cplx a = x[n], h = hx[n];
            const double rr = h.x - th * a.x, ri = h.y - th * a.y;
            s = fma(rr, rr, s);
            s = fma(ri, ri, s);
        }
    }
    __shared__ double sh[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += sh[w];
        rn2[(int64_t)b * p + j] = tot;
    }
}

extern "C" int sqc_residual_norms(sqc_block_t X, sqc_block_t HX, const double* theta,
                                  int32_t ld_theta, int64_t N, int32_t B, double* rn2, void* stream) {
    if (B <= 0 || N <= 0 || X.cnt_fixed <= 0) return SQC_EINVAL;
    const int p = X.cnt_fixed;
    residual_norms_kernel<<<B * p, 256, 0, (cudaStream_t)stream>>>(X, HX, theta, ld_theta, N, p, rn2);
    SQC_LAUNCHED();
    return 0;
}

__global__ void __launch_bounds__(256)
precond_residual_kernel(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* __restrict__ theta,
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
    double den = d[(int64_t)b * d_stride_b + n] - th;
    if (fabs(den) < fl) den = (den < 0.0) ? -fl : fl;
    const double inv = 1.0 / den;
    blk_vec(T, b, jj, N)[n] = cmake((hv.x - th * xv.x) * inv, (hv.y - th * xv.y) * inv);
}

extern "C" int sqc_precond_residual(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* theta,
                                    int32_t ld_theta, const int32_t* act, int32_t ld_act,
                                    const double* d, int64_t d_stride_b, const double* den_floor,
                                    int64_t N, int32_t B, void* stream) {
    if (B <= 0 || N <= 0 || T.cnt_fixed <= 0) return SQC_EINVAL;
    const int nchunk = (int)((N + 255) / 256);
    const int64_t grid = (int64_t)B * T.cnt_fixed * nchunk;
    if (grid > 0x7fffffffLL) return SQC_ELIMIT;
    precond_residual_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        X, HX, T, theta, ld_theta, act, ld_act, d, d_stride_b, den_floor, N, T.cnt_fixed, nchunk);
    SQC_LAUNCHED();
    return 0;
}

__global__ void __launch_bounds__(256)
restart_copy_kernel(sqc_block_t X, sqc_block_t HX, sqc_block_t V, sqc_block_t W,
                    const int32_t* __restrict__ restart, int64_t N, int p, int nchunk) {
    const int vec = blockIdx.x / nchunk;
    const int b = vec / p, j = vec % p;
    if (!restart[b]) return;
    const int64_t n = (int64_t)(blockIdx.x % nchunk) * 256 + threadIdx.x;
    if (n >= N) return;
    blk_vec(V, b, j, N)[n] = blk_vec(X, b, j, N)[n];
    blk_vec(W, b, j, N)[n] = blk_vec(HX, b, j, N)[n];
}

extern "C" int sqc_restart_copy(sqc_block_t X, sqc_block_t HX, sqc_block_t V, sqc_block_t W,
                                const int32_t* restart, int64_t N, int32_t B, void* stream) {
    if (B <= 0 || N <= 0 || X.cnt_fixed <= 0) return SQC_EINVAL;
    const int p = X.cnt_fixed;
    const int nchunk = (int)((N + 255) / 256);
    const int64_t grid = (int64_t)B * p * nchunk;
    if (grid > 0x7fffffffLL) return SQC_ELIMIT;
    restart_copy_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(X, HX, V, W, restart, N, p, nchunk);
    SQC_LAUNCHED();
    return 0;
}

__global__ void __launch_bounds__(256)
diag_mul_kernel(sqc_block_t X, sqc_block_t Y, const double* __restrict__ diag, int64_t diag_stride_b,
                int64_t N, int vmax, int nchunk, int accumulate) {
    const int vec = blockIdx.x / nchunk;
    const int b = vec / vmax, v = vec % vmax;
    if (v >= blk_count(X, b)) return;
    const int64_t n = (int64_t)(blockIdx.x % nchunk) * 256 + threadIdx.x;
    if (n >= N) return;
    const double dv = diag[(int64_t)b * diag_stride_b + n];
    const cplx x = blk_vec(X, b, v, N)[n];
    cplx* y = blk_vec(Y, b, v, N) + n;
    cplx r = cmake(dv * x.x, dv * x.y);
    if (accumulate) r = cadd(r, *y);
    *y = r;
}

extern "C" int sqc_diag_mul(sqc_block_t X, sqc_block_t Y, const double* diag, int64_t diag_stride_b,
                            int64_t N, int32_t B, int32_t accumulate, void* stream) {
    if (B <= 0 || N <= 0 || X.cnt_fixed <= 0) return SQC_EINVAL;
    const int nchunk = (int)((N + 255) / 256);
    const int64_t grid = (int64_t)B * X.cnt_fixed * nchunk;
    if (grid > 0x7fffffffLL) return SQC_ELIMIT;
    diag_mul_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(X, Y, diag, diag_stride_b, N,
                                                                     X.cnt_fixed, nchunk, accumulate);
    SQC_LAUNCHED();
    return 0;
}

extern "C" int sqc_abi_version(void) { return 4; }
extern "C" int64_t sqc_launch_count(void) { return g_sqc_launches; }
extern "C" const char* sqc_error_string(int code) {
    if (code == 0) return "ok";
    if (code == SQC_EINVAL) return "sqcircuit_b200: invalid argument";
    if (code == SQC_ELIMIT) return "sqcircuit_b200: size outside the supported range";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "sqcircuit_b200: unknown error";
}
