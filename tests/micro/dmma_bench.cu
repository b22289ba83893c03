// microbenchmark: mma.sync.m8n8k4.f64 latency / throughput vs DFMA on this GPU (development aid)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dmma_chain(double *o, int iters) {      // dependent chain: latency
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4, c0 = 0, c1 = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    long long t1 = clock64();
    o[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("dmma dependent: %.1f cycles each\n", (double)(t1 - t0) / iters);
}
__global__ void dmma_tput(double *o, int iters) {       // 8 independent accumulators per warp
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[16] = {0};
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[2 * k]), "+d"(c[2 * k + 1]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < 16; k++) s += c[k];
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("dmma 8 independent/warp, %d warps/CTA: %.1f cycles per mma per warp\n", blockDim.x / 32, (double)(t1 - t0) / (iters * 8));
}
__global__ void dfma_tput(double *o, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[8] = {0};
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] = fma(a, b, c[k]);
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < 8; k++) s += c[k];
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("dfma 8 independent/warp, %d warps/CTA: %.2f cycles per dfma per warp\n", blockDim.x / 32, (double)(t1 - t0) / (iters * 8));
}
int main() {
    double *o; cudaMalloc(&o, 148 * 1024 * 8);
    dmma_chain<<<1, 32>>>(o, 2000); cudaDeviceSynchronize();
    dmma_tput<<<148, 32>>>(o, 2000); cudaDeviceSynchronize();
    dmma_tput<<<148, 256>>>(o, 2000); cudaDeviceSynchronize();
    dmma_tput<<<148, 1024>>>(o, 2000); cudaDeviceSynchronize();
    dfma_tput<<<148, 32>>>(o, 2000); cudaDeviceSynchronize();
    dfma_tput<<<148, 256>>>(o, 2000); cudaDeviceSynchronize();
    dfma_tput<<<148, 1024>>>(o, 2000); cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
