// Shared device helpers for the sqcircuit_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sqcircuit_b200.h"

typedef double2 cplx;  // x = re, y = im (torch.complex128 layout)

extern long long g_sqc_launches;  // host-side counter of kernel launches

#define SQC_LAUNCHED()                                     \
    do {                                                   \
        ++g_sqc_launches;                                  \
        cudaError_t e__ = cudaGetLastError();              \
        if (e__ != cudaSuccess) return (int)e__;           \
    } while (0)

__device__ __forceinline__ cplx cmake(double a, double b) { return make_double2(a, b); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// conj(a) * b
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {
    return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cplx cscale(double s, cplx a) { return make_double2(s * a.x, s * a.y); }
// acc += a * b
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a) * b
__device__ __forceinline__ void cfmac(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
}

// ---- block addressing ---------------------------------------------------
__device__ __forceinline__ int blk_count(const sqc_block_t& k, int b) {
    return k.cnt ? k.cnt[b] : k.cnt_fixed;
}
__device__ __forceinline__ cplx* blk_vec(const sqc_block_t& k, int b, int v, int64_t N) {
    int64_t o = k.off ? (int64_t)k.off[b] : 0;
    return reinterpret_cast<cplx*>(k.ptr) + (int64_t)b * k.stride_b + (o + v) * N;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).
// lane = 4*g + t :  a = A[g][t],  b = B[t][g],  c0,c1 = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
