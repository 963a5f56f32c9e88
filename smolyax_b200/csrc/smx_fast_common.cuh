// Definitions shared by the two evaluation kernels of the fast path (smx_fast.cu: cp.async staging, any alignment;
// smx_fast_tma.cu: TMA staging, needs 16-byte aligned rows of x).
#pragma once
#include <cstddef>

#include "smx_common.cuh"

namespace smx {

constexpr int kTile = 32;  // points per tile
static_assert(kBlockWidth == 16, "lane mapping below assumes 16 entries per block");
static_assert(kChunkRows == 16, "the metadata record holds 16 row indices");

struct FastArgs {
    const int32_t* ent_dim;  // only read for the (rare) cold blocks whose columns are not contiguous
    const double* eta;
    const int2* tab_pairs;   // value-table rows of level >= 2: (parent row, hot row)
    const int32_t* hot_off;
    const int4* chunk_dir;   // per work item: first row slot, rows, flags | block << 4, first column of x
    const int32_t* chunk_meta;
    const double* coef;
    const double* c0;
    long long N, ldx, d_out, num_tiles;
    int n_hot, n_tab, n_chunks, n_levels, hot_dims, n_pairs;
    int x_vec_ok;  // x is 16-byte aligned and ldx is even: contiguous blocks may be staged with 16-byte copies
    int level_off[kMaxLevels + 2];
};

// One work item as the kernel sees it in shared memory: metadata record (smx_plan.h, kMetaInts) + coefficient rows.
struct alignas(16) ItemBuffer {
    int tab[16];      // value-table row of each entry (hot blocks)
    int deg[16];      // degree of each entry (0: dummy)
    int etaoff[16];   // offset of the entry's centres in `eta`
    int ridx[16];     // value-table row of each coefficient row
    double eta0[16];  // first centre of each entry
    double coef[kChunkRows * kBlockWidth];
};
static_assert(offsetof(ItemBuffer, coef) == kMetaInts * 4, "metadata record layout");

// Per-warp staging area, filled with cp.async one item ahead of its use: the x tile of the item (32 points x 16
// entries; consumed into registers at the start of the item, so one buffer is enough) and two item buffers.
struct alignas(16) WarpStage {
    double xs[kTile * kBlockWidth];
    ItemBuffer item[2];
};


int fast_eval_tma(const FastDevice& d, const double* x, int64_t N, int64_t ldx, double* y, cudaStream_t st);
int fast_tma_prepare(FastDevice& d);
void fill_fast_args(const FastDevice& d, const double* x, int64_t N, int64_t ldx, FastArgs& a);

}  // namespace smx
