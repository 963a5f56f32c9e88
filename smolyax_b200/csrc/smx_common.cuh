// Shared declarations of the CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdint>
#include <string>
#include <vector>

#include "smolyax_b200.h"
#include "smx_plan.h"

namespace smx {

// ---- error reporting: thread-local message, status codes of smolyax_b200.h -------------------------------------
void set_error(const std::string& msg);
int fail(int status, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
extern std::atomic<int64_t> g_launches;
extern thread_local char t_last_kernel[96];  // name and template arguments of the last kernel this thread launched

#define SMX_CUDA(call)                                        \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return smx::cuda_fail(e__, #call); \
    } while (0)

// after every kernel launch: count it, remember which kernel (and which instantiation) it was - smx_last_kernel() - and
// turn a launch error into a status.  Arguments: a printf format and its values.
#define SMX_LAUNCH_CHECK(...)                                                        \
    do {                                                                             \
        smx::g_launches.fetch_add(1, std::memory_order_relaxed);                     \
        std::snprintf(smx::t_last_kernel, sizeof(smx::t_last_kernel), __VA_ARGS__);  \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) return smx::cuda_fail(e__, smx::t_last_kernel);      \
    } while (0)

constexpr int kSeamMaxN = 8;  // active dimensions per summand supported by the per-summand kernels

// Device-side view of one reference-layout group (all pointers are device pointers).
struct SeamGroup {
    int n;
    long long nn;
    int shape[kSeamMaxN];        // tau_j + 1
    long long fstride[kSeamMaxN];  // element stride of axis j inside one (summand, output) block of F
    long long fsize;             // prod shape
    int tw;                      // taumax + 1
    int ent_off[kSeamMaxN];      // prefix sums of shape: slot j's basis occupies rows [ent_off[j], ent_off[j+1])
    int ent_total;
    const double* F;
    const double* nodes;
    const double* weights;
    const long long* dims;
    const long long* degs;
    const long long* zetas;
    const double* quad;
};

int make_seam_group(const smx_group_desc* g, SeamGroup& out);

// per-summand kernels on the reference layout (smx_seam.cu)
int seam_eval(const double* x, int64_t N, int64_t ldx, const SeamGroup& g, int64_t d_out, double* y, cudaStream_t st);
int seam_gradient(const double* x, int64_t N, int64_t ldx, int64_t d_in, const SeamGroup& g, int64_t d_out, double* J,
                  cudaStream_t st);
int seam_integral(const SeamGroup& g, int64_t d_out, double* q, double* workspace, int64_t workspace_doubles,
                  cudaStream_t st);
int64_t seam_integral_workspace(const SeamGroup& g, int64_t d_out);
int fill_rows(double* y, int64_t N, int64_t d_out, const double* row, cudaStream_t st);  // y[p,:] = row (or 0)
int device_compute_weights(const double* nodes, int64_t m, double* w, cudaStream_t st);
int device_basis(const double* x, int64_t N, const double* xi, const double* w, int64_t m, int64_t nu, int derivative,
                 double* out, cudaStream_t st);

// fast path (smx_fast.cu)
struct FastDevice {
    int64_t d_in = 0, d_out = 0;
    int32_t n_tab = 1, n_hot = 0, n_hot_rows = 0, n_levels = 1, n_chunks = 0, hot_dims = 0, n_pairs = 0;
    int32_t warp_off[17] = {0};  // per-warp item lists (balanced at upload for the chosen CTA shape)
    int32_t level_off[kMaxLevels + 2] = {0};
    double* eta = nullptr;
    int32_t* tab_pairs = nullptr;
    int32_t* tab_factors = nullptr;
    int32_t n_flat = 0;
    int32_t* hot_off = nullptr;
    int32_t* hot_pos = nullptr;
    int32_t* chunk_dir = nullptr;
    int32_t* chunk_meta = nullptr;
    double* coef = nullptr;
    double* c0 = nullptr;
    int32_t* nan_off = nullptr;     // per dimension: the nodes at which the reference returns NaN gradients
    double* nan_nodes = nullptr;
    // GEMM-regime form (large d_out): dense term matrix in DMMA fragment order, see smx_plan.h
    bool has_sparse = false, has_dense = false;
    int32_t dense_k4 = 0;
    int32_t* dense_meta = nullptr;
    double* dense_eta0 = nullptr;
    double* dense_coef = nullptr;
    int32_t* dense_tickets = nullptr;  // one counter per SM: co-resident CTAs of the staged dense kernel rotate their block assignment (smx_dense_kernel.cu)
    bool eta0_zero = true;          // every cold block has zero first centres (pi_{j,1} = x_j): kernel variant without the subtraction
    bool has_cold = false;          // some leading entries live on cold columns (their derivatives are block-sparse row sums)
    int64_t bytes = 0;
    int sm_count = 148;
    int warps = 12;  // warps per CTA of the evaluation kernel (12 or 8: one CTA per SM; 4: two CTAs per SM)
    bool flat_ok = false;  // every row has a factor list: the kernel variant without product rows can run
    bool flat = false;     // .. and is the one chosen (8 warps, two CTAs per SM)
    bool deep_ok = false;  // hot parts of five to eight pairs: the plan carries a second factor list per row slot ..
    bool deep = false;     // .. and the eight-factor lean kernel is the one chosen (values only)
    int multi = 0;         // > 0: values run the multi-set kernel with that many coefficient sets per pass (2 <= d_out < 32)
    int rest_warps = 0;    // > 0: outputs beyond a multiple of three run the lean kernel with that many warps (item lists: pipe_dir), the others fast_multi_kernel
    int pipe_warps = 0;    // > 0: single-output values run the pipelined kernel (smx_fast_pipe.cu) with that many worker warps ..
    int32_t pipe_warp_off[17] = {0};  // .. on their own per-warp item lists
    int32_t* pipe_dir = nullptr;
};
// Gradient jobs on the device (smx_grad_kernel.cu; built by grad_upload from FastPlan::grad)
struct GradDevice {
    bool present = false;
    double* records = nullptr;   // per (item, output): metadata + coefficients packed as DMMA B fragments
    int32_t* dir = nullptr;      // 4 ints per item, in the order of the warps' job lists
    int32_t* jobs = nullptr;     // 4 ints per job
    double* job_c0 = nullptr;    // [job][d_out]
    double* job_nodes = nullptr; // 32 doubles per cold job
    int32_t* zero_cols = nullptr;
    int32_t n_items = 0, n_jobs = 0, n_zero = 0;
    int warps = 0;
    int32_t warp_off[17] = {0};
    int64_t bytes = 0;
};
int fast_upload(const FastPlan& plan, FastDevice& dev);
int grad_upload(const FastPlan& plan, const FastDevice& dev, GradDevice& g);
void grad_free(GradDevice& g);
int grad_kernel_warps(const GradDevice& g, const FastDevice& d, int smem_sm);
int grad_kernel_launch(const FastDevice& d, const GradDevice& g, const double* x, int64_t N, int64_t ldx, double* J, bool nan_at_nodes,
                       cudaStream_t st);
void fast_free(FastDevice& dev);
int fast_eval(const FastDevice& dev, const double* x, int64_t N, int64_t ldx, double* y, cudaStream_t st);
int dense_eval(const FastDevice& dev, const double* x, int64_t N, int64_t ldx, double* y, cudaStream_t st);
int fast_gradient(const FastDevice& dev, const GradDevice& g, const double* x, int64_t N, int64_t ldx, double* J, bool nan_at_nodes, cudaStream_t st);

}  // namespace smx
