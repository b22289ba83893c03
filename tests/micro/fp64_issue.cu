// FP64 issue cost with realistic operand patterns (distinct registers, operands from shared memory).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double seed) {
    __shared__ double sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = seed + 1e-3 * i;
    __syncthreads();
    long long t0, t1;
    double x[8], y[8], z[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { x[j] = seed + j; y[j] = 1.0 + 1e-7 * (j + threadIdx.x); z[j] = 1e-9 * (j + 1); }
    // A: x[j] = x[j]*y[j] + z[j]  (3 distinct operands, 8 chains)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 128; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = x[j] * y[j] + z[j];
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // B: rank-1 style: x[j] -= a * l[j], a and l from shared memory each iteration (TRSM inner loop)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 128; i++) {
        const double a = sm[(i * 32 + threadIdx.x) & 2047];
        const double2 *lt = reinterpret_cast<const double2 *>(sm + ((i * 8) & 1023));
#pragma unroll
        for (int q = 0; q < 4; q++) { const double2 l = lt[q]; x[2 * q] -= a * l.x; x[2 * q + 1] -= a * l.y; }
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // C: 16 accumulators, 4x4 outer product per step with operands from smem (rank32_update inner loop)
    double acc[16];
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = 0.0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { a[u] = sm[(i * 4 + u * 34 + threadIdx.x) & 2047]; b[u] = sm[(i * 4 + u * 8 + 1024) & 2047]; }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int v = 0; v < 4; v++) acc[u * 4 + v] += a[u] * b[v];
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // D: DMMA m8n8k4, 4 independent accumulators
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    double fa = x[0], fb = y[0];
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(fa), "d"(fb));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(fb), "d"(fa));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(fa), "d"(fa));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(fb), "d"(fb));
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    double s = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
    for (int j = 0; j < 8; j++) s += x[j];
    for (int j = 0; j < 16; j++) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 64);
    for (int threads : {32, 64, 128, 256, 512}) {
        k<<<1, threads>>>(out, cyc, 1.25);
        long long h[8];
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads=%3d: A distinct-operand DFMA %.2f cyc/instr | B smem rank-1 (8 DFMA + 5 LDS) %.1f cyc/iter | C 4x4 outer (16 DFMA + 8 LDS) %.1f cyc/iter | D DMMA %.2f cyc/instr (per warp)\n",
               threads, h[0] / 1024.0, h[1] / 128.0, h[2] / 64.0, h[3] / 1024.0);
    }
    return 0;
}
