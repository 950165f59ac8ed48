// Micro-benchmark: FP64 DFMA rate vs FP64 tensor-core (mma.sync m8n8k4 / m16n8k4... f64) rate on the current GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dfma_k(double* out, int iters) {
    double a[8];
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x + j;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fma(a[j], m, c);
    double s = 0;
    for (int j = 0; j < 8; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dmma884_k(double* out, int iters) {
    double c[8][2];
    for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    double s = 0;
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dmma16816_k(double* out, int iters) {
    double c[4][4];
    for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) c[j][q] = 0.0;
    double a[8], b[4];
    for (int q = 0; q < 8; ++q) a[q] = 1.0 + (threadIdx.x + q) * 1e-9;
    for (int q = 0; q < 4; ++q) b[q] = 1.0 - (threadIdx.x + q) * 1e-9;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    double s = 0;
    for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// even warps run the DFMA loop, odd warps the DMMA loop: do the two paths share one execution unit?
__global__ void __launch_bounds__(256) mixed_k(double* out, int iters) {
    double s = 0;
    if ((threadIdx.x >> 5) & 1) {
        double c[8][2];
        for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = 0.0;
        double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
        for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    } else {
        double a[8];
        for (int j = 0; j < 8; ++j) a[j] = threadIdx.x + j;
        const double m = 1.0000001, c = 1e-9;
        for (int i = 0; i < 8 * iters; ++i)      // 8 DFMA = the flops of one DMMA per thread
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fma(a[j], m, c);
        for (int j = 0; j < 8; ++j) s += a[j];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
double timeit(K k, int grid, int iters, double flops_per_thread_iter, double* d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        k<<<grid, 256>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    return flops_per_thread_iter * iters * (double)grid * 256 / (best * 1e-3) / 1e12;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d; cudaMalloc(&d, sizeof(double) * sms * 8 * 256);
    int grid = sms * 8, iters = 1 << 13;
    printf("SMs %d\n", sms);
    printf("DFMA            : %.2f TFLOP/s\n", timeit(dfma_k, grid, iters, 2.0 * 8, d));
    // m8n8k4: 8*8*4 FMA per warp -> 2*256/32 flops per thread per mma
    printf("DMMA m8n8k4     : %.2f TFLOP/s\n", timeit(dmma884_k, grid, iters, 8 * 2.0 * 256 / 32, d));
    printf("DMMA m16n8k16   : %.2f TFLOP/s\n", timeit(dmma16816_k, grid, iters, 4 * 2.0 * 2048 / 32, d));
    // every warp does `iters * 8 * 16` flops per thread in both branches
    printf("mixed DFMA+DMMA : %.2f TFLOP/s (sum of both halves; ~37 => one shared unit, ~74 => separate units)\n",
           timeit(mixed_k, grid, iters, 8 * 2.0 * 256 / 32, d));
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
