// Micro-benchmark: FP64 peaks of this B200 that the roofline fractions in DESIGN.md refer to.
//   DFMA : register-resident chains of fused multiply-adds on the FP64 pipe
//   DMMA : mma.sync.aligned.m8n8k4.row.col.f64 (FP64 tensor path; tcgen05 has no f64 kind)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/_fp64_peaks profiles/fp64_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int TILES>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
    double c[TILES][2];
#pragma unroll
    for (int i = 0; i < TILES; ++i) c[i][0] = threadIdx.x * 1e-3, c[i][1] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < TILES; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < TILES; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        const int threads = warps * 32, blocks = sms;
        float ms = time_ms([&] { dfma_kernel<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double flops = 2.0 * 16 * (double)iters * threads * blocks;
        printf("DFMA  warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", warps, ms, flops / ms / 1e9);
    }
    for (int warps : {4, 8, 16, 32}) {
        const int threads = warps * 32, blocks = sms;
        float ms = time_ms([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double flops = 2.0 * 8 * 8 * 4 * 8 * (double)iters * warps * blocks;
        printf("DMMA  warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", warps, ms, flops / ms / 1e9);
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
