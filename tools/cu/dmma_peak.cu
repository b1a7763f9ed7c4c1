// Dev microbenchmark: issue-rate of the FP64 mma.sync shapes on sm_100a (register-resident operands).
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int ILP>
__global__ void k(double* out, int iters) {
    double c[ILP][4];
    for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 4; ++i) b[i] = 0.5 + threadIdx.x * 1e-4 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (SHAPE == 0)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
            else if (SHAPE == 1)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            else if (SHAPE == 2)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0;
    for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int ILP>
void run(const char* name, double flops_per, int warps_per_sm) {
    double* out; cudaMalloc(&out, 148 * 1024 * sizeof(double) * 4);
    const int iters = 20000;
    int threads = 32 * (warps_per_sm > 16 ? 16 : warps_per_sm);
    int blocks = 148 * (warps_per_sm > 16 ? warps_per_sm / 16 : 1);
    k<SHAPE, ILP><<<blocks, threads>>>(out, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<SHAPE, ILP><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)blocks * (threads / 32) * iters * ILP;
    printf("%-10s ILP %d warps/SM %2d : %.2f TFLOP/s  (%.2f cycles/instr/SM @1.965GHz)  err=%d\n", name, ILP, warps_per_sm,
           n * flops_per / ms / 1e9, ms * 1e-3 * 1.965e9 / (n / 148), (int)cudaGetLastError());
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0, 8>("m8n8k4", 512, w);
        run<1, 8>("m16n8k4", 1024, w);
        run<2, 8>("m16n8k8", 2048, w);
        run<3, 8>("m16n8k16", 4096, w);
    }
    run<0, 2>("m8n8k4", 512, 8); run<0, 4>("m8n8k4", 512, 8); run<0, 16>("m8n8k4", 512, 8);
    run<3, 2>("m16n8k16", 4096, 8); run<3, 4>("m16n8k16", 4096, 8);
    return 0;
}
