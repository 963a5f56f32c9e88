// Definitions shared by the host side (smx_fast.cu) and the GEMM-regime kernel (smx_dense_kernel.cu).
#pragma once
#include "smx_common.cuh"

namespace smx {

constexpr int kDenseTile = 32;  // points per CTA
static_assert(kDenseStageK4 == 16 && kDensePadK4 == 64, "stages of 8 or 16 k-steps; look-ahead of up to 3 x 16 k-steps (meta: stage s + 3)");

struct DenseArgs {
    const double* eta;
    const int2* tab_pairs;   // value-table rows of level >= 2: (parent row, hot row)
    const int32_t* hot_off;
    const int32_t* hot_pos;
    const int2* meta;        // per term: (table row of the hot part, table row of the leading entry or -1 - column of x)
    const double* eta0;      // (d_in) first centre of every dimension
    const double* coef;      // [ceil(d_out / 8)][k4][32] DMMA B fragments
    const double* c0;        // (ncol) constant term per column
    const int32_t* colmap;   // NULL: column c of the product is column c of y;  else its position inside a row of y (gradient)
    long long N, ldx;
    long long ncol;          // columns of the product: d_out (values) or d_out * n_gd (derivative sets)
    long long ldy;           // row pitch of y in elements: d_out (values) or d_out * d_in (gradient)
    int k4;                  // k-steps (4 terms each), a multiple of kDenseStageK4; arrays carry kDensePadK4 more (zeros)
    int nblk;                // ceil(ncol / 8)
    int n_tab, n_hot_rows, n_levels, hot_dims;
    int32_t* tickets;        // per-SM CTA counters (256 ints, never reset), see dense_eval_kernel
    int skew;                // staged kernel: half of the warps assemble A before their DMMAs (set by dense_kernel_launch)
    int level_off[kMaxLevels + 2];
};

bool dense_kernel_fits(int n_tab, int smem_optin);
int dense_kernel_launch(const DenseArgs& a, const double* x, double* y, cudaStream_t st);

}  // namespace smx
