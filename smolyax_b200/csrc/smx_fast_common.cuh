// Definitions shared by the host side (smx_fast.cu: upload, launch) and the kernel (smx_fast_kernel.cu) of the fast path.
#pragma once
#include <cuda.h>

#include <cstddef>

#include "smx_common.cuh"

namespace smx {

constexpr int kTile = 32;  // points per tile
static_assert(kBlockWidth == 16, "lane mapping assumes 16 entries per block");
static_assert(kChunkRows == 16, "the metadata record holds 16 row indices");
constexpr int kKStepDoubles = 4 * kBlockWidth;  // one DMMA k-step (4 rows x 16 entries) of packed coefficients
constexpr int kMaxWarps = 16;

struct FastArgs {
    const double* eta;
    const int2* tab_pairs;   // value-table rows of level >= 2: (parent row, hot row)
    const int4* tab_factors; // rows of level 2..4 as products of up to four hot rows
    int n_flat;              // number of those rows (they come first among the product rows)
    const int32_t* hot_off;
    const int32_t* hot_pos;  // entry index of hot pair k (dimension-major numbering, = index of its centre in eta)
    const int4* chunk_dir;   // per work item: offset of its first record (128-byte units), rows, flags, first column of x
    const double* coef;      // records: per (item, set) the metadata followed by the coefficients packed in DMMA
                             // B-fragment order (pack_coefficients(), fast_upload())
    const double* c0;
    long long N, ldx, d_out, num_tiles;
    // gradient mode: y is J (N, d_out, d_in); pass q of output o is the function itself (q = 0: stores the row sums of the
    // cold blocks, which are the derivatives w.r.t. their columns) or the derivative set of hot dimension grad_dims[q - 1]
    int gradient;            // 0: values, 1: gradient
    int n_gd;                // number of hot dimensions with a derivative set
    const int32_t* grad_dims;
    long long d_in;
    int n_hot, n_hot_rows, n_tab, n_chunks, n_levels, hot_dims, n_pairs;
    int flat;                // 1: the value table holds the hot rows only, product rows are multiplied on the fly
    int ablate;              // tuning builds only (SMX_TUNING): timing experiments, 0 in the product library
    int nwk;                 // pipelined kernel: number of worker warps (warp nwk is the service warp)
    int o_begin, o_end;      // outputs [o_begin, o_end) of this launch (a call may split its outputs over two kernels: smx_fast_kernel.cu)
    unsigned long long* dbg; // tuning builds only: per-tile time stamps of CTA 0 (service warp and worker 0), else nullptr
    int level_off[kMaxLevels + 2];
    int warp_off[kMaxWarps + 1];  // work items of warp w are [warp_off[w], warp_off[w + 1]) of the (re-ordered) directory
};

// One work item as the kernel sees it in shared memory: metadata record (smx_plan.h, kMetaInts) + packed coefficients.
// Device copy: every value-table row index (tab, ridx, fac) is pre-multiplied by kTabPitch (offset in doubles).
struct alignas(16) ItemBuffer {
    int tab[16];      // value-table row of each entry (hot blocks)
    int deg[16];      // degree of each entry (0: dummy)
    int etaoff[16];   // offset of the entry's centres in `eta`
    int ridx[16];     // value-table rows, transposed: ridx[4 * k + s] = row 4 * s + k (0 beyond the item's rows)
    double eta0[16];  // first centre of each entry
    int4 fac[16];     // row slot i (k-step i >> 2, A-fragment column i & 3) as a product of four hot rows (0 = ones row)
    double coef[kChunkRows * kBlockWidth];
};
static_assert(offsetof(ItemBuffer, coef) == kMetaInts * 4, "metadata record layout");

int fast_kernel_prepare(FastDevice& d);
// tensor map of x: (N rows) x (d_in columns) fp64, row pitch ldx * 8 bytes (16-byte aligned rows); box = 16 columns x 32 rows
int make_x_tensor_map(CUtensorMap* map, const double* x, int64_t d_in, int64_t N, int64_t ldx);
// few outputs (2 <= d_out < 32): several coefficient sets per pass (smx_fast_multi.cu)
bool multi_kernel_shape(const FastDevice& d, int smem_optin, int* sets, int* warps);
int multi_kernel_launch(const CUtensorMap& map, const FastDevice& d, const FastArgs& a, const double* x, double* y, cudaStream_t st);
int fast_kernel_launch(const FastDevice& d, const FastArgs& a, const double* x, double* y, cudaStream_t st);
// single output, FLAT plans: warp-specialised persistent kernel without CTA-wide barriers (smx_fast_pipe.cu)
int pipe_kernel_workers(const FastDevice& d, int smem_optin);  // worker warps that fit (0: the kernel cannot run this plan)
int pipe_kernel_launch(const CUtensorMap& map, const FastDevice& d, const FastArgs& a, const double* x, double* y, cudaStream_t st);

}  // namespace smx
