// Device-side helpers shared by the block-sparse kernels (smx_fast_kernel.cu, smx_fast_multi.cu): TMA / bulk-copy /
// mbarrier wrappers, the FP64 tensor instruction, the x-tile buffer.
#pragma once
#include <cuda.h>

#include "smx_fast_common.cuh"

namespace smx {
namespace {

constexpr int kXTileBytes = kTile * kBlockWidth * 8;  // 4096
constexpr int kHotRegs = 8;                           // hot coordinates per lane kept in registers for the next tile

// Per-warp staging: the x tile (TMA destination, 128-byte swizzle => 1024-byte alignment; consumed into registers at
// the start of the item, so one buffer is enough), two item buffers, and one mbarrier per item buffer (the x tile
// travels with its item).
struct alignas(1024) XTile {
    double v[kTile * kBlockWidth];
};
// tile point t = gid + 8 * i  ->  position inside a value-table row: points (gid, gid + 8) and (gid + 16, gid + 24)
// are adjacent pairs, the second pair 16 doubles after the first (two LDS.128 fetch a lane's four A-fragment values)
__device__ __forceinline__ int t_slot(int t) { return ((t >> 4) & 1) * 16 + (t & 7) * 2 + ((t >> 3) & 1); }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_copy(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// D (8x8, fp64) += A (8x4, row) * B (4x8, col): one value of A and B per lane, two of D
template <int ABL = 0>
__device__ __forceinline__ void dmma_(double (&c)[2], double a, double b) {
    if (ABL == 1) {  // keep the data dependence, stay off the FP64 pipe
        c[0] = __longlong_as_double(__double_as_longlong(c[0]) ^ __double_as_longlong(a));
        c[1] = __longlong_as_double(__double_as_longlong(c[1]) ^ __double_as_longlong(b));
        return;
    }
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// D = A * B (first k-step of an item: the accumulators need no clearing)
template <int ABL = 0>
__device__ __forceinline__ void dmma_first(double (&c)[2], double a, double b) {
    if (ABL == 1) {
        c[0] = a, c[1] = b;
        return;
    }
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %4};" : "=d"(c[0]), "=d"(c[1]) : "d"(a), "d"(b), "d"(0.0));
}


// ---- staging of one work item by one lane (lean / pipelined kernels) -------------------------------------------------
// The shared-memory directory of these kernels holds, per item: the global address of its first record (x, y), then
// flags | nf << 8 | record size in 128-byte units << 16 | k-steps << 24 (z; bit 7 = last item of its warp), first column (w).
constexpr int kDirLast = 0x80;
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// Item buffer + its mbarrier; the two of a warp sit back to back, so ONE signed stride flips every pointer into them.
// DEEP (hot parts of five to eight pairs): the record ends with a second factor list, factors 5..8 of every row slot,
// which lands right behind the item's coefficients (at most at `fac2`).
template <bool DEEP>
struct alignas(16) LeanStage {
    ItemBuffer item;
    unsigned long long bar, pad;
};
template <>
struct alignas(16) LeanStage<true> {
    ItemBuffer item;
    int4 fac2[16];
    unsigned long long bar, pad;
};
template <bool DEEP>
__host__ __device__ constexpr size_t lean_stage_bytes() { return 2 * sizeof(LeanStage<DEEP>); }  // per warp
__device__ __forceinline__ void stage_lean(const CUtensorMap* xmap, const void* item, const void* bar, double* xs, const int4 dir, int o, int p0) {
    const unsigned bytes = ((unsigned)dir.z >> 9) & 0x7f80u;  // record = metadata + coefficients
    const bool cold = !(dir.z & kChunkHot);
    unsigned long long* b = const_cast<unsigned long long*>(static_cast<const unsigned long long*>(bar));
    mbar_expect_tx(b, bytes + (cold ? kXTileBytes : 0));
    const unsigned long long src = ((unsigned long long)(unsigned)dir.y << 32 | (unsigned)dir.x) + (unsigned long long)((unsigned)o * bytes);
    bulk_copy(const_cast<void*>(item), reinterpret_cast<const void*>(src), bytes, b);
    if (cold) tma_load_2d(xs, xmap, dir.w, p0, b);
}

}  // namespace
}  // namespace smx
