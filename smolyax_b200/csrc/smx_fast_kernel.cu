// Fast evaluation path (K1): fused 1-D basis + block-sparse Kronecker contraction on the hierarchical layout of
// smx_plan.h.  Replaces the whole of reference interpolation.py:281-302 (python loop over groups and summand batches
// around jit(vmap(barycentric.evaluate_tensor_product_interpolant))) by ONE persistent kernel:
//
//   I(x_p) = c_0 + sum_{e=(j,a)} pi_e(x_pj) * sum_r C[r][e] * m_r(x_p)
//
// One CTA works on one tile of 32 evaluation points at a time (16, 12 or 8 warps and one CTA per SM, or 4 warps and two).
//   * prologue : the CTA fills the value table in shared memory (one row of 32 points per index, pitch kTabPitch):
//                row 0 = 1, then the 1-D basis values pi_e(x_p) of the hot entries (lane = point, running product over
//                the Newton centres; the hot coordinates of the NEXT tile are already in registers, loaded during the
//                previous main loop), then level by level the products of two and more hot pairs.
//   * main     : warp w takes the work items (block of 16 leading entries x <= 16 rows) w, w + NW, ..  One item ahead
//                of its use, one lane stages the item with the TMA: the 32-point x 16-column tile of x of a cold block
//                is ONE cp.async.bulk.tensor.2d (128-byte swizzle; rows beyond N and columns beyond d_in are
//                zero-filled by the hardware; every coordinate of x is read from HBM exactly once), the item's
//                metadata record and its packed coefficients are two cp.async.bulk copies; all complete on an mbarrier.
//                The contraction  acc[p][e] = sum_r m_r(x_p) C[r][e]  runs on the FP64 tensor path:
//                mma.sync.m8n8k4.f64 with A = value-table rows (points x rows), B = coefficients (rows x entries),
//                32 points x 16 entries x 4 rows per k-step = 8 DMMA; accumulators stay in registers.
//                Then tot[p] += sum_e pi_e(x_p) * acc[p][e] with pi_e formed in registers from the staged x tile
//                (or read from the value table for hot blocks).
//   * epilogue : shuffle-reduce over the 4 lanes that share a point, fixed-order sum over the warps, store y.
// Static work assignment, fixed summation order: results are bit-reproducible run to run.
//
// Lane mapping (DMMA fragment layout): gid = lane >> 2, tig = lane & 3.  The lane owns points gid + 8 i (i = 0..3)
// and entries 4 tig + 2 j + {0, 1} (j = 0..1) of the block; accumulator tile (i, j) is the 8 x 8 DMMA tile of points
// 8 i .. 8 i + 7 and the 8 entries { 4 (n >> 1) + 2 j + (n & 1) : n = 0..7 }.
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "smx_fast_device.cuh"

namespace smx {
namespace {

struct alignas(16) ItemStage {
    ItemBuffer item[2];
    unsigned long long bar[2];
};

// One lane of the warp issues the copies of work item c into item buffer `buf` (and, for cold blocks, the x tile).
__device__ __forceinline__ void stage_item(const FastArgs& a, const CUtensorMap* xmap, ItemStage& st, double* xs, int buf, int c,
                                           const int4 dir, long long o, long long p0) {
    const int ksteps = (dir.y + 3) >> 2;
    const unsigned coef_bytes = (unsigned)ksteps * kKStepDoubles * 8;
    const bool cold = !(dir.z & kChunkHot);
    mbar_expect_tx(&st.bar[buf], kMetaInts * 4 + coef_bytes + (cold ? kXTileBytes : 0));
    // metadata + coefficients of (item, set o) are one contiguous record: a single bulk copy
    bulk_copy(&st.item[buf], a.coef + ((size_t)dir.x + (size_t)o * (5 + 4 * ksteps)) * 16, kMetaInts * 4 + coef_bytes, &st.bar[buf]);
    if (cold) tma_load_2d(xs, xmap, dir.w, (int)p0, &st.bar[buf]);
}

template <int NW, int CTAS, int EARLY = 2, bool FLAT = false>
__global__ void __launch_bounds__(NW * 32, CTAS)
fast_eval_kernel(const __grid_constant__ CUtensorMap xmap, const FastArgs a, const double* __restrict__ x, double* __restrict__ y) {
    constexpr int kThreads = NW * 32;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte aligned carve-up (the launch adds 1 KiB of slack)
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    XTile* xtiles = reinterpret_cast<XTile*>(base);                               // [NW]
    ItemStage* stages = reinterpret_cast<ItemStage*>(xtiles + NW);                // [NW]
    double* tab = reinterpret_cast<double*>(stages + NW);                         // [n_tab][kTabPitch] value table
    // FLAT: only the ones row and the hot rows; the products of hot pairs are multiplied on the fly from the factor lists
    // of the item metadata (no product rows => a third of the table, no product pass, one barrier less, two CTAs per SM)
    const int tab_rows = FLAT ? 1 + a.n_hot_rows : a.n_tab;
    int4* s_dir = reinterpret_cast<int4*>(tab + (size_t)tab_rows * kTabPitch);    // [n_chunks] item directory
    int4* s_fac = s_dir + a.n_chunks;                                             // [n_flat] hot rows of each product row
    double* s_eta = reinterpret_cast<double*>(s_fac + (FLAT ? 0 : a.n_flat));     // [n_hot] centres of the hot dimensions
    int2* s_pairs = reinterpret_cast<int2*>(s_eta + a.n_hot);                     // [n_pairs]
    int* s_hot_off = reinterpret_cast<int*>(s_pairs + a.n_pairs);                 // [hot_dims + 1]
    int* s_hot_row = s_hot_off + a.hot_dims + 1;                                  // [n_hot] value-table row of hot pair k

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tig = lane & 3, gid = lane >> 2;
    ItemStage& st = stages[warp];
    double* xs = xtiles[warp].v;
    // this lane's two 16-byte pieces of row gid of the swizzled x tile (rows gid + 8 i are 8 * 128 bytes further on)
    const double* xlo = xs + gid * kBlockWidth + (((2 * tig) ^ gid) << 1);
    const double* xhi = xs + gid * kBlockWidth + (((2 * tig + 1) ^ gid) << 1);

    // ---- once per CTA: small tables to shared memory, mbarriers --------------------------------------------------------
    for (int i = tid; i < a.n_chunks; i += kThreads) s_dir[i] = __ldg(a.chunk_dir + i);
    for (int i = tid; i < a.n_hot; i += kThreads) s_eta[i] = __ldg(a.eta + i);
    for (int i = tid; i < a.n_pairs; i += kThreads) s_pairs[i] = __ldg(a.tab_pairs + i);
    for (int i = tid; i <= a.hot_dims; i += kThreads) s_hot_off[i] = __ldg(a.hot_off + i);
    for (int i = tid; i < a.n_hot; i += kThreads) s_hot_row[i] = 1 + hot_row(__ldg(a.hot_pos + i));
    if (!FLAT)
        for (int i = tid; i < a.n_flat; i += kThreads) s_fac[i] = __ldg(a.tab_factors + i);
    if (lane == 0) {
        mbar_init(&st.bar[0], 1);
        mbar_init(&st.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // hot coordinates of this lane's point: dimension warp + NW * u for u < kHotRegs live in registers, loaded one tile ahead
    double xhot[kHotRegs];
    auto load_hot = [&](long long tile_p0) {
        const double* xrow = x + min(tile_p0 + lane, a.N - 1) * a.ldx;
#pragma unroll
        for (int u = 0; u < kHotRegs; ++u) xhot[u] = (warp + NW * u < a.hot_dims) ? __ldg(xrow + warp + NW * u) : 0.0;
    };
    if ((long long)blockIdx.x < a.num_tiles) load_hot((long long)blockIdx.x * kTile);
    __syncthreads();

    unsigned k_item = 0;  // items this warp has consumed so far: buffer = k & 1, phase parity = (k >> 1) & 1

    for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const long long p0 = tile * kTile;

        // ---- prologue: value table = 1 | 1-D basis values of the hot entries | products of hot pairs, level by level ----
        if (tid < kTile) tab[tid] = 1.0;
        {
            const int slot = t_slot(lane);
            auto hot_dim = [&](int d, double xv) {
                double v = 1.0;
                for (int k = s_hot_off[d]; k < s_hot_off[d + 1]; ++k) {
                    v *= (xv - s_eta[k]);
                    tab[s_hot_row[k] * kTabPitch + slot] = v;
                }
            };
#pragma unroll
            for (int u = 0; u < kHotRegs; ++u)
                if (warp + NW * u < a.hot_dims) hot_dim(warp + NW * u, xhot[u]);
            if (a.hot_dims > NW * kHotRegs) {  // more hot dimensions than the register prefetch covers: load them now
                const double* xrow = x + min(p0 + lane, a.N - 1) * a.ldx;
                for (int d = warp + NW * kHotRegs; d < a.hot_dims; d += NW) hot_dim(d, __ldg(xrow + d));
            }
        }
        __syncthreads();
        if (!FLAT) {
            // products of 2..4 hot pairs in one pass straight from the hot rows (row 0 = 1 pads short products) ..
            // (two points per thread: rows are 16-byte aligned and the products are elementwise)
            const int flat_begin = 1 + a.n_hot_rows, count = a.n_flat * (kTile / 2);
#pragma unroll 4
            for (int idx = tid; idx < count; idx += kThreads) {
                const int k = idx >> 4, s = (idx & 15) * 2;
                const int4 f = s_fac[k];
                const double2 u = *reinterpret_cast<const double2*>(tab + f.x * kTabPitch + s);
                const double2 v = *reinterpret_cast<const double2*>(tab + f.y * kTabPitch + s);
                const double2 w = *reinterpret_cast<const double2*>(tab + f.z * kTabPitch + s);
                const double2 z = *reinterpret_cast<const double2*>(tab + f.w * kTabPitch + s);
                *reinterpret_cast<double2*>(tab + (flat_begin + k) * kTabPitch + s) = make_double2((u.x * v.x) * (w.x * z.x), (u.y * v.y) * (w.y * z.y));
            }
            __syncthreads();
            // .. and, only for hot parts of five and more pairs, level by level from their parents
            for (int l = 5; l < a.n_levels; ++l) {
                const int t_begin = a.level_off[l], cnt = (a.level_off[l + 1] - t_begin) * kTile;
                for (int idx = tid; idx < cnt; idx += kThreads) {
                    const int ti = t_begin + (idx >> 5), s = idx & 31;
                    const int2 pr = s_pairs[ti - 1 - a.n_hot_rows];
                    tab[ti * kTabPitch + s] = tab[pr.x * kTabPitch + s] * tab[pr.y * kTabPitch + s];
                }
                __syncthreads();
            }
        }
        if (tile + gridDim.x < a.num_tiles) load_hot((tile + gridDim.x) * kTile);  // lands during the main loop

        // ---- main: block-sparse contraction, one coefficient set at a time ------------------------------------------------
        const int n_pass = (int)a.d_out;
        for (int pass = 0; pass < n_pass; ++pass) {
            const long long o = pass, set = pass;
            double tot[4] = {0.0, 0.0, 0.0, 0.0};
            const int c_begin = a.warp_off[warp], c_end = a.warp_off[warp + 1];
            if (c_begin < c_end && lane == 0) stage_item(a, &xmap, st, xs, k_item & 1, c_begin, s_dir[c_begin], set, p0);
            for (int c = c_begin; c < c_end; ++c, ++k_item) {
                const int buf = k_item & 1;
                const int4 dir = s_dir[c];
                const ItemBuffer& ib = st.item[buf];
                const int ksteps = (dir.y + 3) >> 2;
                mbar_wait(&st.bar[buf], (k_item >> 1) & 1);

                // fragment loads of the first two k-steps go out first: their latency overlaps the basis-value work below
                // (table rows in the device copy of the metadata are offsets in doubles: row * kTabPitch)
                const int4 r4 = *reinterpret_cast<const int4*>(ib.ridx + 4 * tig);  // this lane's row of every k-step
                const int nf = (dir.z >> 8) & 7;                                    // FLAT: most hot factors of any row
                auto load_a = [&](int s, int row, double2& lo, double2& hi) {       // A fragment of k-step s: 4 points of one row
                    const double* q = tab + 2 * gid;
                    if (!FLAT) {
                        lo = *reinterpret_cast<const double2*>(q + row);
                        hi = *reinterpret_cast<const double2*>(q + row + 16);
                        return;
                    }
                    const int4 f = ib.fac[4 * s + tig];
                    lo = *reinterpret_cast<const double2*>(q + f.x);
                    hi = *reinterpret_cast<const double2*>(q + f.x + 16);
                    if (nf > 1) {
                        const double2 l2 = *reinterpret_cast<const double2*>(q + f.y);
                        const double2 h2 = *reinterpret_cast<const double2*>(q + f.y + 16);
                        lo.x *= l2.x, lo.y *= l2.y, hi.x *= h2.x, hi.y *= h2.y;
                    }
                    if (nf > 2) {
                        const double2 l3 = *reinterpret_cast<const double2*>(q + f.z);
                        const double2 h3 = *reinterpret_cast<const double2*>(q + f.z + 16);
                        const double2 l4 = *reinterpret_cast<const double2*>(q + f.w);
                        const double2 h4 = *reinterpret_cast<const double2*>(q + f.w + 16);
                        lo.x *= l3.x * l4.x, lo.y *= l3.y * l4.y, hi.x *= h3.x * h4.x, hi.y *= h3.y * h4.y;
                    }
                };
                const double2 b0 = *reinterpret_cast<const double2*>(ib.coef + 2 * lane);
                double2 b1 = make_double2(0.0, 0.0), a1lo = b1, a1hi = b1, a0lo, a0hi;
                if (EARLY >= 2) b1 = *reinterpret_cast<const double2*>(ib.coef + kKStepDoubles + 2 * lane);
                load_a(0, r4.x, a0lo, a0hi);
                if (EARLY >= 2) load_a(1, r4.y, a1lo, a1hi);

                // leading basis values pi_e(x_p) of the lane's 4 points x 4 entries (entries 4 tig .. 4 tig + 3)
                double v[4][4];
                if (dir.z & kChunkHot) {
                    const int4 t4 = *reinterpret_cast<const int4*>(ib.tab + 4 * tig);
                    const int tabs[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const double* tr = tab + tabs[e] + 2 * gid;
                        const double2 lo = *reinterpret_cast<const double2*>(tr);
                        const double2 hi = *reinterpret_cast<const double2*>(tr + 16);
                        v[0][e] = lo.x, v[1][e] = lo.y, v[2][e] = hi.x, v[3][e] = hi.y;
                    }
                } else {
                    // 128-byte swizzle: 16-byte piece j of row r sits at piece j ^ (r & 7); r = gid + 8 i => r & 7 = gid
                    const double2 ea = *reinterpret_cast<const double2*>(ib.eta0 + 4 * tig);
                    const double2 eb = *reinterpret_cast<const double2*>(ib.eta0 + 4 * tig + 2);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double2 lo = *reinterpret_cast<const double2*>(xlo + i * (8 * kBlockWidth));
                        const double2 hi = *reinterpret_cast<const double2*>(xhi + i * (8 * kBlockWidth));
                        if (dir.z & kChunkEtaZero) v[i][0] = lo.x, v[i][1] = lo.y, v[i][2] = hi.x, v[i][3] = hi.y;  // pi = x
                        else v[i][0] = lo.x - ea.x, v[i][1] = lo.y - ea.y, v[i][2] = hi.x - eb.x, v[i][3] = hi.y - eb.y;
                    }
                }
                __syncwarp();  // every lane has taken its x values: the x buffer and the other item buffer are free
                if (c + 1 < c_end && lane == 0) stage_item(a, &xmap, st, xs, buf ^ 1, c + 1, s_dir[c + 1], set, p0);

                // acc[i][j] (8 x 8 tiles) = sum over k-steps of A (value-table rows) * B (packed coefficients).
                // Every k-step runs both n-tiles: after the degree-major ordering of the hot entries all but one or two of
                // the (k-step, half block) pairs of a plan carry coefficients, so skipping the empty ones costs more in
                // predicates and accumulator clearing than it saves (the first k-step writes the accumulators).
                double acc[4][2][2];
                {
                    const double af[4] = {a0lo.x, a0lo.y, a0hi.x, a0hi.y};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        dmma_first<0>(acc[i][0], af[i], b0.x);
                        dmma_first<0>(acc[i][1], af[i], b0.y);
                    }
                }
                if (ksteps > 1) {
                    if (EARLY < 2) {
                        b1 = *reinterpret_cast<const double2*>(ib.coef + kKStepDoubles + 2 * lane);
                        load_a(1, r4.y, a1lo, a1hi);
                    }
                    const double af[4] = {a1lo.x, a1lo.y, a1hi.x, a1hi.y};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        dmma_<0>(acc[i][0], af[i], b1.x);
                        dmma_<0>(acc[i][1], af[i], b1.y);
                    }
                }
                if (ksteps > 2) {
                    const int rws[2] = {r4.z, r4.w};
#pragma unroll
                    for (int s = 2; s < 4; ++s) {
                        if (s >= ksteps) break;
                        double2 a01, a23;
                        load_a(s, rws[s - 2], a01, a23);
                        const double2 b = *reinterpret_cast<const double2*>(ib.coef + s * kKStepDoubles + 2 * lane);
                        const double af[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            dmma_<0>(acc[i][0], af[i], b.x);
                            dmma_<0>(acc[i][1], af[i], b.y);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    tot[i] = fma(v[i][0], acc[i][0][0], tot[i]);
                    tot[i] = fma(v[i][1], acc[i][0][1], tot[i]);
                    tot[i] = fma(v[i][2], acc[i][1][0], tot[i]);
                    tot[i] = fma(v[i][3], acc[i][1][1], tot[i]);
                }
            }
            // ---- epilogue: reduce over the 4 lanes that share a point, then over the warps in fixed order ---------------
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                tot[i] += __shfl_xor_sync(0xffffffffu, tot[i], 1);
                tot[i] += __shfl_xor_sync(0xffffffffu, tot[i], 2);
            }
            if (tig == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) xs[gid + 8 * i] = tot[i];  // the warp's x buffer is idle now: reuse it for its partial sums
            }
            __syncthreads();
            if (tid < kTile && p0 + tid < a.N) {
                double s = __ldg(a.c0 + set);
#pragma unroll
                for (int w = 0; w < NW; ++w) s += xtiles[w].v[tid];
                y[(p0 + tid) * a.d_out + o] = s;
            }
            // the x buffers are rewritten (by TMA) only after the barriers of the next prologue, the value table only by that
            // prologue (every warp is past the barrier above): the last output of a tile needs no second barrier
            if (pass + 1 < n_pass) __syncthreads();
        }
    }
}

// ---- lean variant of the FLAT kernel: values only -------------------------------------------------------------------------
// Same data, same shared-memory layout, same arithmetic and summation order as fast_eval_kernel<NW, 2, 0, false, 2, true>;
// what changes is the instruction count of the item loop (the kernel is bound by issue/dependency latency, not by HBM or
// the FP64 pipe):
//   * hot and cold items run separate straight-line bodies, so the 16 leading basis values stay in the registers their
//     LDS.128 delivered them to (the merged body paid 32 register moves per cold item to reconcile the two layouts);
//   * ETA0 (every cold first centre is zero, i.e. the default Leja domain): pi = x, no subtraction at all;
//   * the directory entry of the next item is loaded once (warp-uniform) and carried into the next iteration;
//   * the fragments of the second k-step are only loaded when the item has one (PRE1: before the first DMMAs, else after).
// One lane of the warp issues the copies of an item (lean kernel): the record size comes pre-computed in the directory.
// The shared-memory directory of the lean kernel holds, per item: the global address of its first record (x, y), then
// flags | nf << 8 | record size in 128-byte units << 16 | k-steps << 24 (z; bit 7 = last item of its warp), first column (w).
static_assert(2 * sizeof(LeanStage<false>) <= sizeof(ItemStage) + 32, "the lean layout uses the alignment slack of the carve-up");
static_assert(lean_stage_bytes<false>() <= sizeof(ItemStage) + 32, "smem_bytes() sizes the carve-up with ItemStage");
template <int NW, bool ETA0, bool PRE1, bool ELECT, bool DEEP = false, bool ONE = false>
__global__ void __launch_bounds__(NW * 32, 2)
fast_lean_kernel(const __grid_constant__ CUtensorMap xmap, const FastArgs a, const double* __restrict__ x, double* __restrict__ y) {
    constexpr int kThreads = NW * 32;
    // the dynamic shared memory starts on a 1 KiB boundary of the shared window (no static shared memory in this kernel; checked
    // once below), so every buffer sits at a compile-time offset and no address has to be re-derived inside the item loop
    extern __shared__ __align__(1024) unsigned char smem_lean[];
    unsigned char* base = smem_lean;
    if ((smem_u32(smem_lean) & 1023u) != 0) __trap();
    XTile* xtiles = reinterpret_cast<XTile*>(base);                               // [NW]
    using Stage = LeanStage<DEEP>;
    Stage* stages = reinterpret_cast<Stage*>(xtiles + NW);                        // [NW][2]
    double* tab = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(stages) + NW * lean_stage_bytes<DEEP>());  // [1 + n_hot_rows][kTabPitch]
    int4* s_dir = reinterpret_cast<int4*>(tab + (size_t)(1 + a.n_hot_rows) * kTabPitch);
    double* s_eta = reinterpret_cast<double*>(s_dir + a.n_chunks);
    int2* s_pairs = reinterpret_cast<int2*>(s_eta + a.n_hot);                     // (unused here: same carve-up as smem_bytes)
    int* s_hot_off = reinterpret_cast<int*>(s_pairs + a.n_pairs);
    int* s_hot_row = s_hot_off + a.hot_dims + 1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tig = lane & 3, gid = lane >> 2;
    Stage* st = stages + 2 * warp;
    double* xs = xtiles[warp].v;
    // Lane pointers.  All of them are carried through the item loop by an addition (the three into the item buffers flip
    // between the two buffers, the others add a run-time zero), so that they LIVE in registers: left as loop invariants
    // they were re-derived from the thread index in every iteration (~25 instructions per item at the 128-register cap).
    const double* xlo = xs + gid * kBlockWidth + (((2 * tig) ^ gid) << 1);
    const double* xhi = xs + gid * kBlockWidth + (((2 * tig + 1) ^ gid) << 1);
    const double* tabq = tab + 2 * gid;
    const unsigned char* ibt = reinterpret_cast<const unsigned char*>(&st[0].item) + 16 * tig;    // + offsetof(tab | fac)
    const unsigned char* ibl = reinterpret_cast<const unsigned char*>(&st[0].item) + offsetof(ItemBuffer, coef) + 16 * lane;
    const unsigned char* barp = reinterpret_cast<const unsigned char*>(&st[0].bar);
    int flip = (int)sizeof(Stage);
    const int zero = a.gradient;  // 0 in this kernel; the compiler cannot know
    const double* xsp = xs;       // (warp-uniform: destination of the x tile for the electing variant)

    for (int i = tid; i < a.n_chunks; i += kThreads) {
        const int4 d = __ldg(a.chunk_dir + i);
        const unsigned long long src = reinterpret_cast<unsigned long long>(a.coef) + ((unsigned long long)(unsigned)d.x << 7);
        s_dir[i] = make_int4((int)(unsigned)src, (int)(unsigned)(src >> 32), d.z, d.w);
    }
    for (int i = tid; i < a.n_hot; i += kThreads) s_eta[i] = __ldg(a.eta + i);
    for (int i = tid; i <= a.hot_dims; i += kThreads) s_hot_off[i] = __ldg(a.hot_off + i);
    for (int i = tid; i < a.n_hot; i += kThreads) s_hot_row[i] = 1 + hot_row(__ldg(a.hot_pos + i));
    if (lane == 0) {
        mbar_init(&st[0].bar, 1);
        mbar_init(&st[1].bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    double xhot[kHotRegs];
    auto load_hot = [&](long long tile_p0) {
        const double* xrow = x + min(tile_p0 + lane, a.N - 1) * a.ldx;
#pragma unroll
        for (int u = 0; u < kHotRegs; ++u) xhot[u] = (warp + NW * u < a.hot_dims) ? __ldg(xrow + warp + NW * u) : 0.0;
    };
    if ((long long)blockIdx.x < a.num_tiles) load_hot((long long)blockIdx.x * kTile);
    __syncthreads();
    if (tid < NW && a.warp_off[tid] < a.warp_off[tid + 1]) s_dir[a.warp_off[tid + 1] - 1].z |= kDirLast;
    __syncthreads();

    unsigned k_item = 0;
    const int4* const dir_begin = s_dir + a.warp_off[warp];
    const bool has_items = a.warp_off[warp] < a.warp_off[warp + 1];

    for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int p0 = (int)(tile * kTile);  // (TMA coordinates are 32-bit: N < 2^31, checked at launch)
        if (tid < kTile) tab[tid] = 1.0;
        {
            const int slot = t_slot(lane);
            auto hot_dim = [&](int d, double xv) {
                double v = 1.0;
                for (int k = s_hot_off[d]; k < s_hot_off[d + 1]; ++k) {
                    v *= (xv - s_eta[k]);
                    tab[s_hot_row[k] * kTabPitch + slot] = v;
                }
            };
#pragma unroll
            for (int u = 0; u < kHotRegs; ++u)
                if (warp + NW * u < a.hot_dims) hot_dim(warp + NW * u, xhot[u]);
            if (a.hot_dims > NW * kHotRegs) {
                const double* xrow = x + min((long long)p0 + lane, a.N - 1) * a.ldx;
                for (int d = warp + NW * kHotRegs; d < a.hot_dims; d += NW) hot_dim(d, __ldg(xrow + d));
            }
        }
        __syncthreads();
        if (tile + gridDim.x < a.num_tiles) load_hot((tile + gridDim.x) * kTile);

        const int n_out = ONE ? 1 : a.o_end;  // ONE: single output, the set offset folds away in the staging code
        // (outputs are spread over gridDim.y when there are fewer tiles than CTA slots: small batches, many outputs)
        for (int o = ONE ? 0 : a.o_begin + (int)blockIdx.y; o < n_out; o += ONE ? 1 : (int)gridDim.y) {
            double tot[4] = {0.0, 0.0, 0.0, 0.0};
            int4 dir = make_int4(0, 0, 0, 0);
            const int4* dp = dir_begin;
            bool more = has_items;
            if (has_items) {
                dir = *dp;
                if (ELECT) {
                    if (elect_one()) stage_lean(&xmap, barp - offsetof(Stage, bar), barp, const_cast<double*>(xsp), dir, o, p0);
                } else {
                    if (lane == 0) stage_lean(&xmap, ibt, barp, const_cast<double*>(xlo), dir, o, p0);  // (lane 0: ibt = the buffer, xlo = the x tile)
                }
            }
            for (; more; ++k_item) {
                more = !(dir.z & kDirLast);
                const int4 ndir = *++dp;  // (one slot past the warp's list at its last item: still inside the carve-up, unused)
                const int ksteps = (unsigned)dir.z >> 24;
                const int nf = (dir.z >> 8) & (DEEP ? 15 : 7);
                mbar_wait(const_cast<unsigned long long*>(reinterpret_cast<const unsigned long long*>(barp)), (k_item >> 1) & 1);

                auto load_a = [&](int s, double2& lo, double2& hi) {  // A fragment of k-step s: 4 points of this lane's row
                    const int4 f = *reinterpret_cast<const int4*>(ibt + offsetof(ItemBuffer, fac) + 64 * s);
                    lo = *reinterpret_cast<const double2*>(tabq + f.x);
                    hi = *reinterpret_cast<const double2*>(tabq + f.x + 16);
                    if (nf > 1) {
                        const double2 l2 = *reinterpret_cast<const double2*>(tabq + f.y);
                        const double2 h2 = *reinterpret_cast<const double2*>(tabq + f.y + 16);
                        lo.x *= l2.x, lo.y *= l2.y, hi.x *= h2.x, hi.y *= h2.y;
                    }
                    if (nf > 2) {
                        const double2 l3 = *reinterpret_cast<const double2*>(tabq + f.z);
                        const double2 h3 = *reinterpret_cast<const double2*>(tabq + f.z + 16);
                        const double2 l4 = *reinterpret_cast<const double2*>(tabq + f.w);
                        const double2 h4 = *reinterpret_cast<const double2*>(tabq + f.w + 16);
                        lo.x *= l3.x * l4.x, lo.y *= l3.y * l4.y, hi.x *= h3.x * h4.x, hi.y *= h3.y * h4.y;
                    }
                    if (DEEP && nf > 4) {  // factors 5..8 (the ones row pads): second list, behind the item's coefficients
                        const int4 g = *reinterpret_cast<const int4*>(ibt + offsetof(ItemBuffer, coef) + 8 * kKStepDoubles * ksteps + 64 * s);
                        const int rows4[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const double2 l5 = *reinterpret_cast<const double2*>(tabq + rows4[u]);
                            const double2 h5 = *reinterpret_cast<const double2*>(tabq + rows4[u] + 16);
                            lo.x *= l5.x, lo.y *= l5.y, hi.x *= h5.x, hi.y *= h5.y;
                        }
                    }
                };
                // the whole item, instantiated once for hot and once for cold blocks
                auto item = [&](auto hot_tag) {
                    constexpr bool HOT = decltype(hot_tag)::value;
                    double2 a0lo, a0hi, a1lo = make_double2(0.0, 0.0), a1hi = a1lo, b1 = a1lo;
                    const double2 b0 = *reinterpret_cast<const double2*>(ibl);
                    load_a(0, a0lo, a0hi);
                    if (PRE1 && ksteps > 1) {
                        b1 = *reinterpret_cast<const double2*>(ibl + 8 * kKStepDoubles);
                        load_a(1, a1lo, a1hi);
                    }
                    // leading basis values of the lane's 4 points x 4 entries, as the eight LDS.128 deliver them:
                    //   cold: q0[i] = entries (4 tig, 4 tig + 1), q1[i] = entries (4 tig + 2, 4 tig + 3) of point gid + 8 i
                    //   hot : q0[e] = points (gid, gid + 8),      q1[e] = points (gid + 16, gid + 24)    of entry 4 tig + e
                    // (hot: the values are rows of the value table, which outlives the item - they are read after the DMMAs,
                    // when the A fragments are dead; only their four row offsets are taken from the item buffer now)
                    double2 q0[4], q1[4];
                    int4 t4 = make_int4(0, 0, 0, 0);
                    if (HOT) {
                        t4 = *reinterpret_cast<const int4*>(ibt);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            q0[i] = *reinterpret_cast<const double2*>(xlo + i * (8 * kBlockWidth));
                            q1[i] = *reinterpret_cast<const double2*>(xhi + i * (8 * kBlockWidth));
                        }
                        // (pi = x - eta_0: the subtraction waits until after the DMMAs - its 16 DADDs share the FP64 pipe with
                        // them and would sit in front of the item's first DMMA; the centres stay in the item buffer until then)
                    }
                    __syncwarp();  // every lane has taken its x values: the x buffer and the other item buffer are free
                    if (ELECT) {  // one elected lane, warp-uniform operands: no per-lane-value loops around the copy instructions
                        if (more && elect_one()) stage_lean(&xmap, barp + flip - offsetof(Stage, bar), barp + flip, const_cast<double*>(xsp), ndir, o, p0);
                    } else {
                        if (more && lane == 0) stage_lean(&xmap, ibt + flip, barp + flip, const_cast<double*>(xlo), ndir, o, p0);
                    }
#ifdef SMX_TUNING
                    if (!HOT && (a.ablate & 1) && ksteps == 1) {  // timing experiment: thin cold items stream x but compute nothing
#pragma unroll
                        for (int i = 0; i < 4; ++i) tot[i] += (q0[i].x + q0[i].y) + (q1[i].x + q1[i].y) + a0lo.x + b0.x;
                        return;
                    }
#endif

                    double acc[4][2][2];
                    {
                        const double af[4] = {a0lo.x, a0lo.y, a0hi.x, a0hi.y};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            dmma_first<0>(acc[i][0], af[i], b0.x);
                            dmma_first<0>(acc[i][1], af[i], b0.y);
                        }
                    }
                    if (ksteps > 1) {
                        if (!PRE1) {
                            b1 = *reinterpret_cast<const double2*>(ibl + 8 * kKStepDoubles);
                            load_a(1, a1lo, a1hi);
                        }
                        const double af[4] = {a1lo.x, a1lo.y, a1hi.x, a1hi.y};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            dmma_<0>(acc[i][0], af[i], b1.x);
                            dmma_<0>(acc[i][1], af[i], b1.y);
                        }
#pragma unroll 1
                        for (int s = 2; s < ksteps; ++s) {
                            double2 a01, a23;
                            load_a(s, a01, a23);
                            const double2 b = *reinterpret_cast<const double2*>(ibl + 8 * kKStepDoubles * s);
                            const double ag[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                dmma_<0>(acc[i][0], ag[i], b.x);
                                dmma_<0>(acc[i][1], ag[i], b.y);
                            }
                        }
                    }
                    if (HOT) {
                        const int tabs[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            q0[e] = *reinterpret_cast<const double2*>(tabq + tabs[e]);
                            q1[e] = *reinterpret_cast<const double2*>(tabq + tabs[e] + 16);
                        }
                        const double v[4][4] = {{q0[0].x, q0[1].x, q0[2].x, q0[3].x}, {q0[0].y, q0[1].y, q0[2].y, q0[3].y},
                                                {q1[0].x, q1[1].x, q1[2].x, q1[3].x}, {q1[0].y, q1[1].y, q1[2].y, q1[3].y}};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            tot[i] = fma(v[i][0], acc[i][0][0], tot[i]);
                            tot[i] = fma(v[i][1], acc[i][0][1], tot[i]);
                            tot[i] = fma(v[i][2], acc[i][1][0], tot[i]);
                            tot[i] = fma(v[i][3], acc[i][1][1], tot[i]);
                        }
                    } else {
                        if (!ETA0 && !(dir.z & kChunkEtaZero)) {
                            const double2 ea = *reinterpret_cast<const double2*>(ibt + offsetof(ItemBuffer, eta0) + 16 * tig);
                            const double2 eb = *reinterpret_cast<const double2*>(ibt + offsetof(ItemBuffer, eta0) + 16 * tig + 16);
#pragma unroll
                            for (int i = 0; i < 4; ++i) q0[i].x -= ea.x, q0[i].y -= ea.y, q1[i].x -= eb.x, q1[i].y -= eb.y;
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            tot[i] = fma(q0[i].x, acc[i][0][0], tot[i]);
                            tot[i] = fma(q0[i].y, acc[i][0][1], tot[i]);
                            tot[i] = fma(q1[i].x, acc[i][1][0], tot[i]);
                            tot[i] = fma(q1[i].y, acc[i][1][1], tot[i]);
                        }
                    }
                };
                if (dir.z & kChunkHot) item(std::true_type{});
                else item(std::false_type{});
                dir = ndir;
                ibt += flip, ibl += flip, barp += flip, flip = -flip;
                xlo += zero, xhi += zero, tabq += zero;
                if (ELECT) xsp += zero;
            }
            // ---- epilogue: as in fast_eval_kernel ----------------------------------------------------------------------
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                tot[i] += __shfl_xor_sync(0xffffffffu, tot[i], 1);
                tot[i] += __shfl_xor_sync(0xffffffffu, tot[i], 2);
            }
            if (tig == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) xs[gid + 8 * i] = tot[i];
            }
            __syncthreads();
            if (tid < kTile && (long long)p0 + tid < a.N) {
                double s = __ldg(a.c0 + o);
#pragma unroll
                for (int w = 0; w < NW; ++w) s += xtiles[w].v[tid];
                y[((long long)p0 + tid) * a.d_out + o] = s;
            }
            if (!ONE && o + (int)gridDim.y < n_out) __syncthreads();
        }
    }
}

size_t smem_bytes(const FastDevice& d, int nw, bool flat = false) {
    const size_t rows = flat ? 1 + (size_t)d.n_hot_rows : (size_t)d.n_tab;
    return 1024 + (size_t)nw * (sizeof(XTile) + sizeof(ItemStage)) + (rows * kTabPitch + (size_t)d.n_hot) * sizeof(double) + (flat ? 0 : (size_t)d.n_flat * sizeof(int4)) + 16 +
           (size_t)d.n_chunks * sizeof(int4) + (size_t)d.n_pairs * sizeof(int2) + ((size_t)d.hot_dims + 1 + d.n_hot) * sizeof(int) + 16;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point: the library links no libcuda, so it still loads
// (and reports "no device") on a machine without a driver.
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

template <int NW, int CTAS, int EARLY = 2, bool FLAT = false>
int launch(const CUtensorMap& map, const FastArgs& a, const FastDevice& d, const double* x, double* y, cudaStream_t st) {
    const long long grid = std::min<long long>(a.num_tiles, (long long)d.sm_count * CTAS);
    if (EARLY != 2 || FLAT)
        SMX_CUDA(cudaFuncSetAttribute(fast_eval_kernel<NW, CTAS, EARLY, FLAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(d, NW, FLAT)));
    fast_eval_kernel<NW, CTAS, EARLY, FLAT><<<(unsigned)grid, NW * 32, smem_bytes(d, NW, FLAT), st>>>(map, a, x, y);
    SMX_LAUNCH_CHECK("fast_eval_kernel<%d,%d,%d,%d>", NW, CTAS, (int)EARLY, (int)FLAT);
    return SMX_OK;
}

size_t lean_smem_bytes(const FastDevice& d, int nw, bool deep) {
    return smem_bytes(d, nw, true) + (deep ? (size_t)nw * (lean_stage_bytes<true>() - lean_stage_bytes<false>()) : 0);
}

template <int NW, bool ETA0, bool PRE1, bool ELECT, bool DEEP, bool ONE = false>
int launch_lean2(const CUtensorMap& map, const FastArgs& a, const FastDevice& d, const double* x, double* y, cudaStream_t st) {
    const long long slots = (long long)d.sm_count * 2, grid = std::min<long long>(a.num_tiles, slots);
    // fewer tiles than CTA slots: split the outputs over gridDim.y instead of walking them one after the other
    const long long gy = a.num_tiles < slots ? std::min<long long>(a.o_end - a.o_begin, (slots + a.num_tiles - 1) / a.num_tiles) : 1;
    const size_t smem = lean_smem_bytes(d, NW, DEEP);
    SMX_CUDA(cudaFuncSetAttribute(fast_lean_kernel<NW, ETA0, PRE1, ELECT, DEEP, ONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fast_lean_kernel<NW, ETA0, PRE1, ELECT, DEEP, ONE><<<dim3((unsigned)grid, (unsigned)gy), NW * 32, smem, st>>>(map, a, x, y);
    SMX_LAUNCH_CHECK("fast_lean_kernel<%d,%d,%d,%d,%d,%d>", NW, (int)ETA0, (int)PRE1, (int)ELECT, (int)DEEP, (int)ONE);
    return SMX_OK;
}
template <int NW>
int launch_lean(const CUtensorMap& map, const FastArgs& a, const FastDevice& d, const double* x, double* y, cudaStream_t st) {
    // (staging by an elected lane: measured 1.847 vs 1.888 ms against lane 0; single-output instantiation: 1.805 vs 1.818)
    if (d.deep) return d.eta0_zero ? launch_lean2<NW, true, false, true, true>(map, a, d, x, y, st) : launch_lean2<NW, false, false, true, true>(map, a, d, x, y, st);
    if (a.d_out == 1) return d.eta0_zero ? launch_lean2<NW, true, false, true, false, true>(map, a, d, x, y, st) : launch_lean2<NW, false, false, true, false, true>(map, a, d, x, y, st);
    return d.eta0_zero ? launch_lean2<NW, true, false, true, false>(map, a, d, x, y, st) : launch_lean2<NW, false, false, true, false>(map, a, d, x, y, st);
}

}  // namespace

static int prepare_shape(FastDevice& d);

// Chooses the CTA shape for this plan and opts into the shared memory it needs.
int fast_kernel_prepare(FastDevice& d) {
    const int rc = prepare_shape(d);
    if (rc) return rc;
    d.pipe_warps = 0;
    if (d.flat && !d.deep && !d.multi) {  // single output: the pipelined kernel (smx_fast_pipe.cu) when its buffers fit
        d.pipe_warps = 1;  // eligible; fast_upload() settles the number of workers
    }
    return SMX_OK;
}

static int prepare_shape(FastDevice& d) {
    if (encode_tiled() == nullptr) return fail(SMX_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    int device = 0, smem_optin = 0, smem_sm = 0;
    SMX_CUDA(cudaGetDevice(&device));
    SMX_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    SMX_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
    int want = 0;
    want = tune_int("SMX_FAST_WARPS", 0);  // tuning knob: 4, 8 or 12
    const bool fits4 = 2 * (smem_bytes(d, 4) + 1024) <= (size_t)smem_sm;
    const bool fits8 = smem_bytes(d, 8) <= (size_t)smem_optin;
    const bool fits12 = smem_bytes(d, 12) <= (size_t)smem_optin;
    const bool fits16 = smem_bytes(d, 16) <= (size_t)smem_optin;
    // preferred: no product rows in the table (two CTAs of 8 or 6 warps per SM: the barriers and the prologue of one tile
    // overlap the main loop of another); needs a factor list for every row (hot parts of at most four pairs)
    static const int want_flat = tune_int("SMX_FAST_FLAT", 1);
    d.flat = false;
    d.multi = 0;
    d.rest_warps = 0;
    {   // a few outputs: several coefficient sets per pass (its warp count also fixes the per-warp item lists)
        int sets = 0, warps = 0;
        if (want_flat && want == 0 && multi_kernel_shape(d, smem_optin, &sets, &warps)) {
            d.flat = true;
            d.multi = sets;
            d.warps = warps;
            // more than two passes and outputs left over beyond a multiple of `sets`: those run the single-set kernel in a
            // second launch (its own warp count and item lists) instead of a pass with empty sets
            d.rest_warps = 0;
            if (d.d_out > 2 * sets && d.d_out % sets != 0)
                for (int nw : {8, 6})
                    if (d.rest_warps == 0 && 2 * (smem_bytes(d, nw, true) + 1024) <= (size_t)smem_sm) d.rest_warps = nw;
            return SMX_OK;
        }
    }
    if (d.flat_ok && want_flat && (want == 0 || want == 8 || want == 6)) {
        for (int nw : {8, 6}) {
            if (want != nw && want != 0) continue;
            if (2 * (smem_bytes(d, nw, true) + 1024) <= (size_t)smem_sm) {
                d.flat = true;
                d.warps = nw;
                return SMX_OK;
            }
        }
    }
    // hot parts of five to eight pairs: the lean kernel with eight-factor records (values only; the gradient then runs on the
    // per-summand kernels).  Chosen when no variant with product rows fits in shared memory (SMX_FAST_DEEP=1: always).
    d.deep = false;
    if (!d.flat_ok && d.deep_ok && want_flat) {
        static const int want_deep = tune_int("SMX_FAST_DEEP", -1);
        if (want_deep == 1 || (want_deep != 0 && !(fits4 || fits8 || fits12 || fits16))) {
            for (int nw : {8, 6}) {
                if (2 * (lean_smem_bytes(d, nw, true) + 1024) <= (size_t)smem_sm) {
                    d.flat = d.deep = true;
                    d.warps = nw;
                    return SMX_OK;
                }
            }
        }
    }
    d.warps = 0;
    if (want == 16 && fits16) d.warps = 16;
    else if (want == 12 && fits12) d.warps = 12;
    else if (want == 8 && fits8) d.warps = 8;
    else if (want == 4 && fits4) d.warps = 4;
    else if (fits16) d.warps = 16;
    else if (fits12) d.warps = 12;
    else if (fits8) d.warps = 8;
    else if (fits4) d.warps = 4;
    if (d.warps == 0) return fail(SMX_ERR_UNSUPPORTED, "value table does not fit in shared memory");
    if (d.warps == 4)
        SMX_CUDA(cudaFuncSetAttribute(fast_eval_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(d, 4)));
    if (d.warps == 8)
        SMX_CUDA(cudaFuncSetAttribute(fast_eval_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(d, 8)));
    if (d.warps == 12)
        SMX_CUDA(cudaFuncSetAttribute(fast_eval_kernel<12, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(d, 12)));
    if (d.warps == 16) {
        SMX_CUDA(cudaFuncSetAttribute(fast_eval_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(d, 16)));
        SMX_CUDA(cudaFuncSetAttribute(fast_eval_kernel<16, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(d, 16)));
    }
    return SMX_OK;
}

int make_x_tensor_map(CUtensorMap* map, const double* x, int64_t d_in, int64_t N, int64_t ldx) {
    if (encode_tiled() == nullptr) return fail(SMX_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)d_in, (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)ldx * sizeof(double)};
    const cuuint32_t box[2] = {kBlockWidth, kTile};
    const cuuint32_t estr[2] = {1, 1};
    // L2 promotion: a tile row is 128 bytes; with 256-byte promotion the fetch also brings the same rows of the next column
    // block into L2 (another item of the same tile, a few microseconds later)
    static const int promo = tune_int("SMX_FAST_L2PROMO", 256);  // (measured: 1.790 vs 1.799 ms)
    const CUresult res = encode_tiled()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(x), dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : promo == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                        : promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (res != CUDA_SUCCESS) return fail(SMX_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)res));
    return SMX_OK;
}

int fast_kernel_launch(const FastDevice& d, const FastArgs& a, const double* x, double* y, cudaStream_t st) {
    CUtensorMap map;
    const int rc_map = make_x_tensor_map(&map, x, d.d_in, a.N, a.ldx);
    if (rc_map) return rc_map;
    if (d.pipe_warps > 0 && a.N < (1ll << 31) - kTile) return pipe_kernel_launch(map, d, a, x, y, st);
    if (d.multi) {
        if (d.rest_warps > 0 && a.N < (1ll << 31) - kTile) {
            FastArgs am = a, al = a;
            am.o_end = al.o_begin = (int)(d.d_out / d.multi) * d.multi;
            const int rc = multi_kernel_launch(map, d, am, x, y, st);
            if (rc) return rc;
            al.chunk_dir = reinterpret_cast<const int4*>(d.pipe_dir);
            for (int w = 0; w <= kMaxWarps; ++w) al.warp_off[w] = d.pipe_warp_off[w];
            return d.rest_warps == 8 ? launch_lean<8>(map, al, d, x, y, st) : launch_lean<6>(map, al, d, x, y, st);
        }
        return multi_kernel_launch(map, d, a, x, y, st);
    }
    if (d.flat) {
        // the lean item loop (SMX_FAST_LEAN=0 in a tuning build selects the general kernel, for A/B timing)
        static const int lean = tune_int("SMX_FAST_LEAN", 1);
        if (d.deep && a.N >= (1ll << 31) - kTile) return fail(SMX_ERR_UNSUPPORTED, "more than 2^31 points in one call");
        if ((lean || d.deep) && a.N < (1ll << 31) - kTile) {
            if (d.warps == 8) return launch_lean<8>(map, a, d, x, y, st);
            if (d.warps == 6) return launch_lean<6>(map, a, d, x, y, st);
        }
        return d.warps == 8 ? launch<8, 2, 2, true>(map, a, d, x, y, st) : launch<6, 2, 2, true>(map, a, d, x, y, st);
    }
    if (d.warps == 4) return launch<4, 2>(map, a, d, x, y, st);
    if (d.warps == 8) return launch<8, 1>(map, a, d, x, y, st);
    if (d.warps == 16) {
        // 16 warps leave 128 registers per thread: pre-loading one k-step (not two) keeps the kernel spill free (measured)
        static const int early = tune_int("SMX_FAST_EARLY", 1);
        if (early == 2) return launch<16, 1>(map, a, d, x, y, st);
        return launch<16, 1, 1>(map, a, d, x, y, st);
    }
    return launch<12, 1>(map, a, d, x, y, st);
}

}  // namespace smx
