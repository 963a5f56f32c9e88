// K1, pipelined form: the block-sparse value kernel of smx_fast_kernel.cu as a warp-specialised persistent kernel without
// CTA-wide barriers (single output, FLAT plans).  Replaces reference interpolation.py:281-302 like the other K1 variants:
//
//   I(x_p) = c_0 + sum_{e=(j,a)} pi_e(x_pj) * sum_r C[r][e] * m_r(x_p)
//
// What the barrier-synchronised kernels pay for and this one does not (measured, profiles/r06_k1_experiments.md):
//   * what surrounds the item loop of a tile of 32 points (value-table prologue, two CTA barriers, cross-warp reduction)
//     costs 0.37 ms per 10^6 points at the headline configuration when nothing overlaps it, every tile ends with the
//     slowest warp of its static schedule, and a warp cannot request the first x tile of the next tile before the barriers
//     of the prologue.
// Here ONE CTA per SM runs `nwk` worker warps and one service warp that meet only through mbarriers:
//   * workers: static item lists and the lean kernel's item body (x tile by TMA and record by bulk copy into the warp's own
//     buffers, one item ahead; A fragments multiplied from the value table; DMMA).  A worker's item stream is continuous
//     across tiles: while it computes the last item of tile t it already stages the first item of tile t + 1.
//   * service: builds the value table of tile t + 1 into the OTHER of two table buffers while the workers run tile t (the hot
//     columns of x arrive by its own TMA boxes), adds up the workers' partial sums of tile t - 1 in fixed order and stores y.
//     tab_full[b] (service -> workers) and red_full[b] (workers -> service) are the only cross-warp dependencies of a tile:
//     a fast worker runs up to one tile ahead of the slowest, so the imbalance of the static schedule averages out.
// Deterministic: static lists, fixed summation order.  Results differ from the lean kernel's only in the grouping of the
// per-warp partial sums.
//
// Tried on top of this and NOT kept (all parity-green, all slower; numbers and reasons in profiles/r06_k1_experiments.md, the
// last variant is kept as benchmarks/experiments/smx_fast_pipe_ring.cu.txt): thin cold items (<= 3 rows) as plain FMAs on
// resident coefficients; their x by 256-bit global loads into a register pipeline (in the workers, and in dedicated streamer
// warps); a CTA-wide ring of TMA slots fed by a producer warp; L2 tensor prefetch ahead of the loads.  The common limit:
// HBM latency under this access pattern is ~1.7 us, so the x stream needs 75 - 120 KB in flight per SM all the time, and
// next to two value tables and the item buffers shared memory has room for 60 KB of x tiles.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <type_traits>
#include <vector>

#include "smx_fast_device.cuh"

namespace smx {
namespace {

using Stage = LeanStage<false>;

#ifdef SMX_TUNING
constexpr int kDbgWords = 16 * 8;
#endif

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// byte offsets of the carve-up of the dynamic shared memory (which starts on a 1 KiB boundary)
struct PipeLayout {
    size_t xtiles, xhot, stages, tab, red, dir, pairs, bars, total;
    int tab_doubles;
};
__host__ __device__ inline PipeLayout pipe_layout(int nwk, int n_hot_rows, int n_chunks, int n_hot, int hot_dims) {
    PipeLayout L;
    size_t at = 0;
    L.xtiles = at, at += (size_t)nwk * sizeof(XTile);                  // [nwk] x tiles (TMA destinations, 128-byte swizzle)
    L.xhot = at, at += (size_t)((hot_dims + kBlockWidth - 1) / kBlockWidth) * sizeof(XTile);  // hot columns of the next tile
    L.stages = at, at += (size_t)nwk * lean_stage_bytes<false>();      // [nwk][2] item buffers + their mbarriers
    L.tab_doubles = (1 + n_hot_rows) * kTabPitch;                      // ones row, hot rows
    L.tab = at, at += 2 * (size_t)L.tab_doubles * sizeof(double);      // [2] value tables
    L.red = at, at += 2 * (size_t)nwk * kTile * sizeof(double);        // [2][nwk][32] partial sums of the workers
    L.dir = at, at += ((size_t)n_chunks + 1) * sizeof(int4);           // item directory
    L.pairs = at, at += (size_t)n_hot * sizeof(int4);                   // per hot pair: centre, table row offset, x offset | first << 31
    L.bars = at, at += 5 * sizeof(unsigned long long);                 // tab_full[2], red_full[2], xhot_full
    L.total = at;
    return L;
}

template <bool ETA0>
__global__ void __launch_bounds__(512, 1)
fast_pipe_kernel(const __grid_constant__ CUtensorMap xmap, const FastArgs a, const double* __restrict__ x, double* __restrict__ y) {
    extern __shared__ __align__(1024) unsigned char smem_pipe[];
    if ((smem_u32(smem_pipe) & 1023u) != 0) __trap();
    const int nwk = a.nwk;  // worker warps; warp nwk is the service warp
    const PipeLayout L = pipe_layout(nwk, a.n_hot_rows, a.n_chunks, a.n_hot, a.hot_dims);
    XTile* xtiles = reinterpret_cast<XTile*>(smem_pipe + L.xtiles);
    const double* xhot = reinterpret_cast<const double*>(smem_pipe + L.xhot);
    Stage* stages = reinterpret_cast<Stage*>(smem_pipe + L.stages);
    double* tab = reinterpret_cast<double*>(smem_pipe + L.tab);
    double* red = reinterpret_cast<double*>(smem_pipe + L.red);
    int4* s_dir = reinterpret_cast<int4*>(smem_pipe + L.dir);
    int4* s_pair = reinterpret_cast<int4*>(smem_pipe + L.pairs);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_pipe + L.bars);  // [0..1] tab_full, [2..3] red_full, [4] xhot

    const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31, warp = tid >> 5;

    // ---- once per CTA ----------------------------------------------------------------------------------------------------
    for (int i = tid; i < a.n_chunks; i += nthreads) {
        const int4 d = __ldg(a.chunk_dir + i);
        const unsigned long long src = reinterpret_cast<unsigned long long>(a.coef) + ((unsigned long long)(unsigned)d.x << 7);
        s_dir[i] = make_int4((int)(unsigned)src, (int)(unsigned)(src >> 32), d.z, d.w);
    }
    // hot pairs in dimension-major order (k = hot_off[d] + degree - 1): what the service warp needs to extend the running
    // product of a dimension by one factor, in one 16-byte record
    for (int d = tid; d < a.hot_dims; d += nthreads)
        for (int k = __ldg(a.hot_off + d), k0 = k; k < __ldg(a.hot_off + d + 1); ++k) {
            const double eta = __ldg(a.eta + k);
            // .w: offset of x[row 0][d] inside the staged boxes (doubles, before the swizzle) | swizzle piece << 16 | first << 31
            s_pair[k] = make_int4(__double2loint(eta), __double2hiint(eta), (1 + hot_row(__ldg(a.hot_pos + k))) * kTabPitch,
                                  ((d >> 4) * (kTile * kBlockWidth) + (d & 1)) | (((d >> 1) & 7) << 16) | (k == k0 ? (int)0x80000000 : 0));
        }
    // rows the service warp never writes: the ones row, and the padding rows of the hot block (read by dummy entries)
    for (int i = tid; i < 2 * L.tab_doubles; i += nthreads) tab[i] = 1.0;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], nwk);
        mbar_init(&bars[3], nwk);
        mbar_init(&bars[4], 1);
    }
    if (warp < nwk && lane == 0) {
        mbar_init(&stages[2 * warp].bar, 1);
        mbar_init(&stages[2 * warp + 1].bar, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (tid < nwk) s_dir[a.warp_off[tid + 1] - 1].z |= kDirLast;  // (every worker has at least one item: pipe_kernel_workers)
    __syncthreads();
    if ((long long)blockIdx.x >= a.num_tiles) return;

    if (warp == nwk) {
        // ================================================ service warp ================================================
        // lane = point of the tile.  Per tile: reduce the partial sums of the tile before last (same buffers), then build
        // the value table of this one:  row 0 = 1,  row of hot pair (d, a) = prod_{k < a} (x_d - eta_{d,k}).
        const int slot = t_slot(lane);
        const double c0 = __ldg(a.c0);
        auto reduce = [&](long long t, unsigned b) {  // fixed order: bit-reproducible
            const long long p = t * kTile + lane;
            const double* r = red + (size_t)b * nwk * kTile + lane;
            double part[kMaxWarps];
#pragma unroll
            for (int w = 0; w < kMaxWarps; ++w) part[w] = w < nwk ? r[w * kTile] : 0.0;
            double s = c0;
#pragma unroll
            for (int w = 0; w < kMaxWarps; ++w) s += part[w];
            if (p < a.N) y[p] = s;
        };
        // The hot columns of a tile (32 rows x the first hot_dims columns of x) arrive by TMA, in boxes of 16 columns like the
        // workers' x tiles, a whole tile ahead: requested right after the table of the previous tile has been built.
        const int nbox = (a.hot_dims + kBlockWidth - 1) / kBlockWidth;
        auto request_hot = [&](long long t) {
            if (elect_one()) {
                mbar_expect_tx(&bars[4], (unsigned)nbox * kXTileBytes);
                for (int j = 0; j < nbox; ++j)
                    tma_load_2d(const_cast<double*>(xhot) + (size_t)j * (kTile * kBlockWidth), &xmap, j * kBlockWidth, (int)(t * kTile), &bars[4]);
            }
        };
        request_hot(blockIdx.x);
        // this lane's row of a box: 16-byte piece c of row r sits at piece c ^ (r & 7) (128-byte swizzle)
        const double* xl = xhot + lane * kBlockWidth;
        const int sw = lane & 7;
        unsigned it = 0;
        for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
            const unsigned b = it & 1;
            if (it >= 2) {
                mbar_wait(&bars[2 + b], ((it - 2) >> 1) & 1);
                reduce(tile - 2ll * gridDim.x, b);
            }
            mbar_wait(&bars[4], it & 1);
            double* tb = tab + (size_t)b * L.tab_doubles + slot;
            // eight pairs at a time: their records, then their coordinates, then the running products and the stores (left in
            // one loop the loads of pair k + 1 wait for the store of pair k, which may alias them: 150 cycles per pair)
            double v = 1.0;
            for (int k0 = 0; k0 < a.n_hot; k0 += 8) {
                int4 m[8];
                double xv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) m[u] = s_pair[min(k0 + u, a.n_hot - 1)];
#pragma unroll
                for (int u = 0; u < 8; ++u) xv[u] = xl[(m[u].w & 0xffff) + ((((m[u].w >> 16) & 7) ^ sw) << 1)];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double f = xv[u] - __hiloint2double(m[u].y, m[u].x);
                    v = m[u].w < 0 ? f : v * f;
                    if (k0 + u < a.n_hot) tb[m[u].z] = v;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[b]);
            if (tile + gridDim.x < a.num_tiles) request_hot(tile + gridDim.x);
        }
        // the last two tiles of this CTA
        for (unsigned j = it >= 2 ? it - 2 : 0; j < it; ++j) {
            mbar_wait(&bars[2 + (j & 1)], (j >> 1) & 1);
            reduce((long long)blockIdx.x + (long long)j * gridDim.x, j & 1);
        }
        return;
    }
    if (warp > nwk) return;

    // ==================================================== worker warps ====================================================
    const int tig = lane & 3, gid = lane >> 2;
    Stage* st = stages + 2 * warp;
    double* xs = xtiles[warp].v;
    // Lane pointers, all carried through the item loop by an addition so that they live in registers (smx_fast_kernel.cu).
    const double* xlo = xs + gid * kBlockWidth + (((2 * tig) ^ gid) << 1);
    const double* xhi = xs + gid * kBlockWidth + (((2 * tig + 1) ^ gid) << 1);
    const double* tabq = tab + 2 * gid;
    int tab_flip = L.tab_doubles;
    const unsigned char* ibt = reinterpret_cast<const unsigned char*>(&st[0].item) + 16 * tig;
    const unsigned char* ibl = reinterpret_cast<const unsigned char*>(&st[0].item) + offsetof(ItemBuffer, coef) + 16 * lane;
    const unsigned char* barp = reinterpret_cast<const unsigned char*>(&st[0].bar);
    int flip = (int)sizeof(Stage);
    const int zero = a.gradient;  // 0 in this kernel; the compiler cannot know
    const double* xsp = xs;
    double* redw = red + warp * kTile;
    int red_flip = nwk * kTile;

    const int4* const dir_begin = s_dir + a.warp_off[warp];
#ifdef SMX_TUNING
    long long w_item = 0, w_tab = 0, w_hot = 0, w_cold = 0;
    const long long w_begin = clock64();
#endif
    unsigned k_item = 0;
    long long tile = blockIdx.x;
    int4 dir = *dir_begin;
    if (elect_one()) stage_lean(&xmap, barp - offsetof(Stage, bar), barp, const_cast<double*>(xsp), dir, 0, (int)(tile * kTile));

    for (unsigned it = 0; tile < a.num_tiles; ++it) {
        const long long next_tile = tile + gridDim.x;
        const bool has_next = next_tile < a.num_tiles;
        const int p0 = (int)(tile * kTile), p0n = has_next ? (int)(next_tile * kTile) : 0;  // (N < 2^31: checked at launch)
#ifdef SMX_TUNING
        const long long wt0 = clock64();
#endif
        mbar_wait(&bars[it & 1], (it >> 1) & 1);  // the value table of this tile is complete
#ifdef SMX_TUNING
        w_tab += clock64() - wt0;
#endif

        double tot[4] = {0.0, 0.0, 0.0, 0.0};
        const int4* dp = dir_begin;
        bool more = true;
        while (more) {
            more = !(dir.z & kDirLast);
            dp = more ? dp + 1 : dir_begin;
            const int4 ndir = *dp;  // the next item: of this tile, or the first one of the next tile
            const bool stage_next = more || has_next;
            const int pnext = more ? p0 : p0n;
            const int ksteps = (unsigned)dir.z >> 24;
            const int nf = (dir.z >> 8) & 7;
#ifdef SMX_TUNING
            const long long wi0 = clock64();
#endif
            mbar_wait(const_cast<unsigned long long*>(reinterpret_cast<const unsigned long long*>(barp)), (k_item >> 1) & 1);
#ifdef SMX_TUNING
            const long long wi1 = clock64();
            w_item += wi1 - wi0;
            const int kind = (dir.z & kChunkHot) ? 1 : 2;
#endif

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
            };
            auto item = [&](auto hot_tag) {
                constexpr bool HOT = decltype(hot_tag)::value;
                double2 a0lo, a0hi;
                const double2 b0 = *reinterpret_cast<const double2*>(ibl);
                load_a(0, a0lo, a0hi);
                // leading basis values of the lane's 4 points x 4 entries, as the eight LDS.128 deliver them:
                //   cold: q0[i] = entries (4 tig, 4 tig + 1), q1[i] = entries (4 tig + 2, 4 tig + 3) of point gid + 8 i
                //   hot : q0[e] = points (gid, gid + 8),      q1[e] = points (gid + 16, gid + 24)    of entry 4 tig + e
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
                    // (pi = x - eta_0 is subtracted after the DMMAs, see the lean kernel)
                }
                __syncwarp();  // every lane has taken its x values: the x buffer and the other item buffer are free
                if (stage_next && elect_one())
                    stage_lean(&xmap, barp + flip - offsetof(Stage, bar), barp + flip, const_cast<double*>(xsp), ndir, 0, pnext);

                double acc[4][2][2];
                {
                    const double af[4] = {a0lo.x, a0lo.y, a0hi.x, a0hi.y};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        dmma_first<0>(acc[i][0], af[i], b0.x);
                        dmma_first<0>(acc[i][1], af[i], b0.y);
                    }
                }
#pragma unroll 1
                for (int s = 1; s < ksteps; ++s) {
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
#ifdef SMX_TUNING
            {
                const long long dt = clock64() - wi1;
                if (kind == 1) w_hot += dt;
                else w_cold += dt;
            }
#endif
            dir = ndir;
            ibt += flip, ibl += flip, barp += flip, flip = -flip;
            xlo += zero, xhi += zero;
            xsp += zero;
            ++k_item;
        }
        // ---- this warp's partial sums of the tile: reduce over the 4 lanes that share a point, hand over to the service warp ----
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            tot[i] += __shfl_xor_sync(0xffffffffu, tot[i], 1);
            tot[i] += __shfl_xor_sync(0xffffffffu, tot[i], 2);
        }
        if (tig == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) redw[gid + 8 * i] = tot[i];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[2 + (it & 1)]);
        tabq += tab_flip, tab_flip = -tab_flip;
        redw += red_flip, red_flip = -red_flip;
        tile = next_tile;
    }
#ifdef SMX_TUNING
    if (a.dbg && blockIdx.x == 0 && lane == 0) {
        unsigned long long* o = a.dbg + warp * 8;
        o[0] = (unsigned long long)(clock64() - w_begin), o[1] = (unsigned long long)w_tab, o[2] = (unsigned long long)w_item;
        o[4] = (unsigned long long)w_hot, o[5] = (unsigned long long)w_cold;
    }
#endif
}

size_t pipe_smem_bytes(const FastDevice& d, int nwk) {
    return pipe_layout(nwk, d.n_hot_rows, d.n_chunks, d.n_hot, d.hot_dims).total;
}

template <bool ETA0>
int launch_pipe(const CUtensorMap& map, const FastArgs& a, const FastDevice& d, const double* x, double* y, cudaStream_t st) {
    const size_t smem = pipe_smem_bytes(d, a.nwk);
    SMX_CUDA(cudaFuncSetAttribute(fast_pipe_kernel<ETA0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = std::min<long long>(a.num_tiles, d.sm_count);
    fast_pipe_kernel<ETA0><<<(unsigned)grid, (a.nwk + 1) * 32, smem, st>>>(map, a, x, y);
    SMX_LAUNCH_CHECK("fast_pipe_kernel<%d> workers=%d", (int)ETA0, a.nwk);
    return SMX_OK;
}

}  // namespace

// Worker warps of the pipelined kernel for this plan: as many as fit (at most 15: warp 16 is the service warp; every worker
// needs an item).  0: the kernel cannot run this plan.
int pipe_kernel_workers(const FastDevice& d, int smem_optin) {
    if (!d.flat_ok || d.d_out != 1 || d.n_chunks < 1) return 0;
    const int want = tune_int("SMX_PIPE_WORKERS", 15);  // (0 switches the kernel off)
    for (int nwk = std::min({want, 15, (int)d.n_chunks}); nwk >= 4; --nwk)
        if (pipe_smem_bytes(d, nwk) <= (size_t)smem_optin) return nwk;
    return 0;
}

int pipe_kernel_launch(const CUtensorMap& map, const FastDevice& d, const FastArgs& args, const double* x, double* y, cudaStream_t st) {
    FastArgs a = args;  // the pipelined kernel's own item lists (its worker count differs from the barrier kernels' warp count)
    a.nwk = d.pipe_warps;
    a.chunk_dir = reinterpret_cast<const int4*>(d.pipe_dir);
    for (int w = 0; w <= kMaxWarps; ++w) a.warp_off[w] = d.pipe_warp_off[w];
#ifdef SMX_TUNING
    unsigned long long* dbg = nullptr;
    if (tune_int("SMX_PIPE_DEBUG", 0)) {
        cudaMalloc((void**)&dbg, kDbgWords * 8);
        cudaMemset(dbg, 0, kDbgWords * 8);
        a.dbg = dbg;
    }
#endif
    const int rc = d.eta0_zero ? launch_pipe<true>(map, a, d, x, y, st) : launch_pipe<false>(map, a, d, x, y, st);
#ifdef SMX_TUNING
    if (dbg) {
        std::vector<unsigned long long> h(kDbgWords);
        cudaMemcpy(h.data(), dbg, kDbgWords * 8, cudaMemcpyDeviceToHost);
        cudaFree(dbg);
        for (int w = 0; w < d.pipe_warps; ++w) {
            const unsigned long long* o = &h[(size_t)w * 8];
            std::fprintf(stderr, "worker %2d: total %9llu cycles; waiting: table %5.1f %%, items %5.1f %%; executing: hot items %5.1f %%, cold items %5.1f %%\n",
                         w, o[0], 100.0 * o[1] / o[0], 100.0 * o[2] / o[0], 100.0 * o[4] / o[0], 100.0 * o[5] / o[0]);
        }
    }
#endif
    return rc;
}

}  // namespace smx
