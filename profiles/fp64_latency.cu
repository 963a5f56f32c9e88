// FP64 pipe latencies on sm_100a, measured with clock64 inside one warp (and with 2 / 4 warps per SM sub-partition):
//   dependent DFMA chain, dependent DMMA chain (same accumulator), K independent DMMA accumulators issued round-robin,
//   LDS.128 -> use latency.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int K>
__global__ void dmma_indep(double* out, long long* cyc, int iters, double a, double b) {
    double c[K][2];
#pragma unroll
    for (int k = 0; k < K; ++k) c[k][0] = c[k][1] = threadIdx.x * 1e-9 + k;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) dmma(c[k], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int K>
__global__ void dfma_indep(double* out, long long* cyc, int iters, double a, double b) {
    double c[K];
#pragma unroll
    for (int k = 0; k < K; ++k) c[k] = threadIdx.x * 1e-9 + k;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) c[k] = fma(c[k], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// LDS.128 -> dependent address chain: latency of a shared-memory load
__global__ void lds_chain(double* out, long long* cyc, int iters) {
    __shared__ int4 buf[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) buf[i] = make_int4((i + 33) & 255, 0, 0, 0);
    __syncthreads();
    int idx = threadIdx.x & 255;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) idx = buf[idx].x;
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = idx;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// DMMA -> DFMA -> DMMA dependent (accumulator feeds an FMA feeds the next A operand): mixed chain
__global__ void mixed_chain(double* out, long long* cyc, int iters, double a, double b) {
    double c[2] = {threadIdx.x * 1e-9, 1.0};
    double t = 1.0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        dmma(c, t, b);
        t = fma(c[0], a, b);
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = c[0] + c[1] + t;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// KM independent DMMAs and KF independent DFMAs per iteration: do the two share a pipe (time = sum) or not (time = max)?
template <int KM, int KF>
__global__ void mixed_indep(double* out, long long* cyc, int iters, double a, double b) {
    double c[KM][2], f[KF];
#pragma unroll
    for (int k = 0; k < KM; ++k) c[k][0] = c[k][1] = threadIdx.x * 1e-9 + k;
#pragma unroll
    for (int k = 0; k < KF; ++k) f[k] = threadIdx.x * 1e-9 + k;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < KM; ++k) {
            dmma(c[k], a, b);
#pragma unroll
            for (int j = 0; j < KF / KM; ++j) f[k * (KF / KM) + j] = fma(f[k * (KF / KM) + j], a, b);
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < KM; ++k) s += c[k][0] + c[k][1];
#pragma unroll
    for (int k = 0; k < KF; ++k) s += f[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <class F>
static void run(const char* name, int ops_per_iter, int iters, F launch) {
    long long* cyc;
    cudaMallocManaged(&cyc, 8);
    launch(cyc);
    launch(cyc);
    cudaDeviceSynchronize();
    printf("%-44s %8.2f cycles per op  (%s)\n", name, (double)cyc[0] / ((double)iters * ops_per_iter), cudaGetErrorString(cudaGetLastError()));
    cudaFree(cyc);
}

int main() {
    double* out;
    cudaMalloc(&out, 8 * 1024 * 1024);
    const int iters = 4096;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int warps : {1, 4, 8, 16}) {  // warps per SM (1 = one warp on one sub-partition; 4 = one per sub-partition; 8 = two; 16 = four)
        const int th = warps * 32;
        printf("-- %d warp(s) per SM, %d blocks\n", warps, sms);
        char nm[96];
        snprintf(nm, sizeof nm, "DFMA dependent chain (K=1)"); run(nm, 1, iters, [&](long long* c) { dfma_indep<1><<<sms, th>>>(out, c, iters, 1.0000001, 1e-9); });
        snprintf(nm, sizeof nm, "DFMA 4 independent chains"); run(nm, 4, iters, [&](long long* c) { dfma_indep<4><<<sms, th>>>(out, c, iters, 1.0000001, 1e-9); });
        snprintf(nm, sizeof nm, "DFMA 16 independent chains"); run(nm, 16, iters, [&](long long* c) { dfma_indep<16><<<sms, th>>>(out, c, iters, 1.0000001, 1e-9); });
        snprintf(nm, sizeof nm, "DMMA dependent chain (K=1)"); run(nm, 1, iters, [&](long long* c) { dmma_indep<1><<<sms, th>>>(out, c, iters, 1.0000001, 1e-9); });
        snprintf(nm, sizeof nm, "DMMA 2 independent accumulators"); run(nm, 2, iters, [&](long long* c) { dmma_indep<2><<<sms, th>>>(out, c, iters, 1.0000001, 1e-9); });
        snprintf(nm, sizeof nm, "DMMA 4 independent accumulators"); run(nm, 4, iters, [&](long long* c) { dmma_indep<4><<<sms, th>>>(out, c, iters, 1.0000001, 1e-9); });
        snprintf(nm, sizeof nm, "DMMA 8 independent accumulators"); run(nm, 8, iters, [&](long long* c) { dmma_indep<8><<<sms, th>>>(out, c, iters, 1.0000001, 1e-9); });
        snprintf(nm, sizeof nm, "DMMA -> DFMA -> DMMA dependent (per pair)"); run(nm, 1, iters, [&](long long* c) { mixed_chain<<<sms, th>>>(out, c, iters, 1.0000001, 1e-9); });
        snprintf(nm, sizeof nm, "4 DMMA + 8 DFMA independent (per iteration)"); run(nm, 1, iters, [&](long long* c) { mixed_indep<4, 8><<<sms, th>>>(out, c, iters, 1.0000001, 1e-9); });
        snprintf(nm, sizeof nm, "4 DMMA + 16 DFMA independent (per iteration)"); run(nm, 1, iters, [&](long long* c) { mixed_indep<4, 16><<<sms, th>>>(out, c, iters, 1.0000001, 1e-9); });
        snprintf(nm, sizeof nm, "4 DMMA + 32 DFMA independent (per iteration)"); run(nm, 1, iters, [&](long long* c) { mixed_indep<4, 32><<<sms, th>>>(out, c, iters, 1.0000001, 1e-9); });
        snprintf(nm, sizeof nm, "LDS.128 dependent chain"); run(nm, 1, iters, [&](long long* c) { lds_chain<<<sms, th>>>(out, c, iters); });
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
