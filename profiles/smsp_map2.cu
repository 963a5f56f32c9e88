// Two CTAs of 8 warps on one SM: which sub-partition does warp b of the SECOND CTA share with warp 0 of the first?
// Every CTA draws a ticket per SM; the CTA with ticket 0 runs its warp 0, the CTA with ticket 1 its warp b, both the same loop
// of independent DMMAs (16 cycles per DMMA = pipe to itself, 32 = shared).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smsp_map2 smsp_map2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 2) pair(double* out, long long* cyc, int* tickets, int* ready, int wb, int iters) {
    extern __shared__ unsigned char pad[];  // 100 KB: exactly two CTAs per SM
    __shared__ int s_ticket;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) {
        s_ticket = atomicAdd(tickets + smid, 1);
        atomicAdd(ready + smid, 1);
        while (atomicAdd(ready + smid, 0) < 2) {}  // both CTAs of this SM are resident
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int mine = s_ticket == 0 ? 0 : wb;
    if (s_ticket > 1 || warp != mine) return;
    double c[4][2];
    for (int k = 0; k < 4; ++k) c[k][0] = c[k][1] = threadIdx.x * 1e-9 + k;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) dmma(c[k], 1.0000001, 1e-9);
    const long long t1 = clock64();
    double s = 0;
    for (int k = 0; k < 4; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + pad[0];
    unsigned wid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    if (smid == 0 && (threadIdx.x & 31) == 0) cyc[s_ticket == 0 ? 0 : 2] = t1 - t0, cyc[s_ticket == 0 ? 1 : 3] = wid;
}

int main() {
    double* out;
    long long* cyc;
    int *tickets, *ready;
    cudaMalloc(&out, 8 * 1024 * 1024);
    cudaMallocManaged(&cyc, 64);
    cudaMalloc(&tickets, 1024 * 4);
    cudaMalloc(&ready, 1024 * 4);
    cudaFuncSetAttribute(pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 4096;
    printf("cycles per DMMA of warp 0 of the first CTA of SM 0 while warp b of the second CTA runs the same loop\n");
    for (int b = 0; b < 8; ++b) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(tickets, 0, 1024 * 4);
            cudaMemset(ready, 0, 1024 * 4);
            pair<<<296, 256, 100 * 1024>>>(out, cyc, tickets, ready, b, iters);
            cudaDeviceSynchronize();
        }
        printf("  ticket-0 CTA warp 0 (warpid %lld): %6.2f cycles per DMMA;  ticket-1 CTA warp %d (warpid %lld): %6.2f\n", cyc[1], (double)cyc[0] / (iters * 4.0), b, cyc[3], (double)cyc[2] / (iters * 4.0));
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
