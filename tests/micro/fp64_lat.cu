// Latency microbenchmarks that set the critical path of the panel Cholesky: dependent DFMA, rsqrt(double), SHFL, LDS.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double seed) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed + i;
    __syncthreads();
    double a = seed + threadIdx.x, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    // 1. dependent DFMA chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) a = a * b + c;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 1024;
    // 2. dependent rsqrt chain
    double r = fabs(a) + 2.0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) r = rsqrt(r) + 1.5;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = (t1 - t0) / 256;
    // 3. dependent shfl chain (64-bit = 2 SHFL)
    double s = r;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) s = __shfl_sync(0xffffffffu, s, (i * 7) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = (t1 - t0) / 256;
    // 4. dependent LDS chain
    int idx = threadIdx.x;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) idx = ((int)sm[idx & 1023] + i) & 1023;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = (t1 - t0) / 256;
    // 5. independent DFMA throughput, one warp: 8 chains
    double x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = a + j;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 128; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = x[j] * b + c;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = (t1 - t0);   // 1024 DFMA warp-instructions
    // 6. float rsqrt seed + one cubic step
    double q = fabs(x[0]) + 2.0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
        const double y0 = (double)rsqrtf((float)q);
        const double e = fma(-q * y0, y0, 1.0);
        const double p = fma(0.375, e, 0.5);
        q = fma(y0 * e, p, y0) + 1.5;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = (t1 - t0) / 256;
    // 7. DMUL + FSEL + SHFL + DFMA (one pivot step without rsqrt)
    double u = q, w = s;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
        u = u * b;
        const double l = __shfl_sync(0xffffffffu, u, i & 31);
        w = w - u * l;
        u = w;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = (t1 - t0) / 256;
    double acc = a + r + s + idx + q + w;
    for (int j = 0; j < 8; j++) acc += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 64);
    for (int threads : {32, 128, 512}) {
        k<<<1, threads>>>(out, cyc, 1.25);
        long long h[8];
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads=%d: DFMA dep %lld | rsqrt+add dep %lld | shfl64 dep %lld | LDS dep(+cvt) %lld | 1024 indep DFMA (8 chains) %lld cyc | custom rsqrt+add %lld | mul+shfl+fma step %lld\n",
               threads, h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    }
    return 0;
}
