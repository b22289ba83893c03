// microbenchmark: FP64 RED throughput, scattered vs sector-coalesced vs shared-memory RMW (development aid)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void red_scatter(double *buf, size_t nblk, int iters) {   // lane -> own block, element e per instruction
    const size_t gt = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; it++) {
        const size_t blk = (gt * 2654435761u + it * 97u) % nblk;
#pragma unroll
        for (int e = 0; e < 36; e++) atomicAdd(buf + blk * 36 + e, 1.0);
    }
}
__global__ void red_coalesced(double *buf, size_t nblk, int iters) { // warp -> 32 blocks in turn, lane -> element
    const size_t gt = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; it++) {
        for (int p = 0; p < 32; p++) {
            const size_t blk = (((gt - lane) + p) * 2654435761u + it * 97u) % nblk;
            atomicAdd(buf + blk * 36 + lane, 1.0);
            if (lane < 4) atomicAdd(buf + blk * 36 + 32 + lane, 1.0);
        }
    }
}
__global__ void smem_rmw(double *out, int iters) {                  // 36 RMW per lane per iteration into shared memory
    extern __shared__ double sm[];
    const int nblk = 400;
    for (int i = threadIdx.x; i < nblk * 36; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        const int blk = (threadIdx.x * 7 + it * 13) % nblk;          // conflicts ignored: throughput only
#pragma unroll
        for (int e = 0; e < 36; e++) sm[blk * 36 + e] += 1.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = sm[5];
}
__global__ void smem_atomic(double *out, int iters) {               // 36 atomicAdd(double) per lane per iteration into shared memory
    extern __shared__ double sm[];
    const int nblk = 400;
    for (int i = threadIdx.x; i < nblk * 36; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        const int blk = (threadIdx.x * 7 + it * 13) % nblk;
#pragma unroll
        for (int e = 0; e < 36; e++) atomicAdd(&sm[blk * 36 + e], 1.0);
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = sm[5];
}
int main() {
    const size_t nblk = 14000;   // ~4 MB of blocks, L2 resident like the band-stored system
    double *buf; cudaMalloc(&buf, nblk * 36 * 8); cudaMemset(buf, 0, nblk * 36 * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = 148 * 2, threads = 256, iters = 200;
    const double total = (double)grid * threads * iters * 36;
    for (int rep = 0; rep < 2; rep++) {
        float ms;
        cudaEventRecord(a); red_scatter<<<grid, threads>>>(buf, nblk, iters); cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b); printf("scatter   : %.3f ms  %.1f G RED/s\n", ms, total / ms / 1e6);
        cudaEventRecord(a); red_coalesced<<<grid, threads>>>(buf, nblk, iters); cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b); printf("coalesced : %.3f ms  %.1f G RED/s\n", ms, total / ms / 1e6);
        cudaFuncSetAttribute(smem_rmw, cudaFuncAttributeMaxDynamicSharedMemorySize, 400 * 36 * 8);
        cudaEventRecord(a); smem_rmw<<<148, 512, 400 * 36 * 8>>>(buf, iters * 4); cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b); printf("smem rmw  : %.3f ms  %.1f G RMW/s\n", ms, 148.0 * 512 * iters * 4 * 36 / ms / 1e6);
        cudaFuncSetAttribute(smem_atomic, cudaFuncAttributeMaxDynamicSharedMemorySize, 400 * 36 * 8);
        cudaEventRecord(a); smem_atomic<<<148, 512, 400 * 36 * 8>>>(buf, iters * 4); cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b); printf("smem atom : %.3f ms  %.1f G atomicAdd/s\n", ms, 148.0 * 512 * iters * 4 * 36 / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
