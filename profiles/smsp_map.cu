// Which SM sub-partition does warp w of a CTA run on?  Two warps (a, b) of a 16-warp CTA issue independent DMMAs, the other
// warps exit: if a and b share a sub-partition they share its FP64 pipe and the loop takes twice as long.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smsp_map smsp_map.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__global__ void pair(double* out, long long* cyc, int wa, int wb, int iters) {
    const int warp = threadIdx.x >> 5;
    if (warp != wa && warp != wb) return;
    double c[4][2];
    for (int k = 0; k < 4; ++k) c[k][0] = c[k][1] = threadIdx.x * 1e-9 + k;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) dmma(c[k], 1.0000001, 1e-9);
    const long long t1 = clock64();
    double s = 0;
    for (int k = 0; k < 4; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && warp == wa && (threadIdx.x & 31) == 0) cyc[0] = t1 - t0;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 8 * 1024 * 1024);
    cudaMallocManaged(&cyc, 8);
    const int iters = 4096;
    printf("cycles per DMMA of warp 0 while warp b runs the same loop (16 cycles = pipe to itself, 32 = shared)\n");
    for (int b = 1; b < 16; ++b) {
        pair<<<148, 512>>>(out, cyc, 0, b, iters);
        pair<<<148, 512>>>(out, cyc, 0, b, iters);
        cudaDeviceSynchronize();
        printf("  warps 0 and %2d: %6.2f\n", b, (double)cyc[0] / (iters * 4.0));
    }
    // two CTAs of 8 warps per SM: does warp w of the second CTA land on the same sub-partition as warp w of the first?
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
