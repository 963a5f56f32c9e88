// GEMM-regime evaluation kernel ("K2") for vector-valued interpolants:   y = c0 + Phi(x) C.
//
// The reference contracts every summand's padded value tensor with d_out as a free index of its einsum
// (barycentric.py:119-123, interpolation.py:283-302).  In the hierarchical basis of smx_plan.h the same quantity is one
// dense product: a row of Phi per point (one column per term, Phi[p][t] = hot part * leading basis value, both read from
// the CTA's value table) times the term-by-output coefficient matrix C.  For large d_out the work is the GEMM, so this
// kernel is organised around the FP64 tensor instruction (mma.sync.m8n8k4.f64, the only FP64 MMA of sm_100a):
//
//   CTA        = 32 points x (NW * NB * 8) outputs; grid = (point tiles, output groups), point tiles fastest so that
//                CTAs running at the same time stream the same slice of C through L2
//   prologue   = value table of the tile in shared memory (hot 1-D basis values, then products level by level)
//   main loop  = stages of NW k-steps (4 terms each).  Every warp builds ONE k-step of the next stage's A fragments
//                (Phi values, already in DMMA fragment order: one conflict-free LDS.128 pair per lane and k-step for the
//                consumers) while it runs the DMMAs of the current stage; one __syncthreads per stage.
//   B operand  = C pre-packed on the host in fragment order ([8 outputs][k-step][lane]); each warp streams the fragments
//                of its own NB output blocks straight from L2 into a small register ring (PF k-steps ahead) - no warp
//                shares B with another warp of the CTA, so shared memory would add nothing.
//   epilogue   = y[p][o] = c0[o] + acc, 16-byte stores.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "smx_dense.cuh"

namespace smx {
namespace {

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// Value table of one tile of 32 points in shared memory: row 0 = 1, hot 1-D basis values, then the products of hot pairs
// level by level (smx_plan.h).  All threads of the CTA; ends with a __syncthreads.
template <int NW>
__device__ __forceinline__ void build_table(const DenseArgs& a, const double* __restrict__ x, long long p0, double* tab) {
    constexpr int kThreads = NW * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < kDenseTile) tab[tid] = 1.0;
    {
        const double* xrow = x + min(p0 + lane, a.N - 1) * a.ldx;
        for (int d = warp; d < a.hot_dims; d += NW) {
            const double xv = __ldg(xrow + d);
            double v = 1.0;
            for (int k = __ldg(a.hot_off + d); k < __ldg(a.hot_off + d + 1); ++k) {
                v *= (xv - __ldg(a.eta + k));
                tab[(1 + hot_row(__ldg(a.hot_pos + k))) * kTabPitch + lane] = v;
            }
        }
    }
    __syncthreads();
    for (int l = 2; l < a.n_levels; ++l) {
        const int t_begin = a.level_off[l], cnt = (a.level_off[l + 1] - t_begin) * kDenseTile;
        for (int idx = tid; idx < cnt; idx += kThreads) {
            const int ti = t_begin + (idx >> 5), s = idx & 31;
            const int2 pr = __ldg(a.tab_pairs + (ti - 1 - a.n_hot_rows));
            tab[ti * kTabPitch + s] = tab[pr.x * kTabPitch + s] * tab[pr.y * kTabPitch + s];
        }
        __syncthreads();
    }
}

template <int NW, int NB, int PF, int CTAS>
__global__ void __launch_bounds__(NW * 32, CTAS)
dense_eval_kernel(const DenseArgs a, const double* __restrict__ x, double* __restrict__ y) {
    static_assert(NW % PF == 0, "the register ring is indexed with the k-step inside a stage");
    constexpr int kThreads = NW * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* abuf = reinterpret_cast<double*>(smem_raw);       // [2][NW][2][32 lanes][2]: A fragments of two stages
    double* tab = abuf + 2 * NW * 128;                         // [n_tab][kTabPitch] value table

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tig = lane & 3, gid = lane >> 2;
    const long long p0 = (long long)blockIdx.x * kDenseTile;

    // ---- which output blocks this warp multiplies ---------------------------------------------------------------------------
    // Dealt in warp order, a column group that does not fill the CTA loads the four FP64 pipes of the SM unevenly (cfg5: 13
    // blocks on 8 warps x 2 -> 4, 4, 3, 2 blocks on the pipes of warps 0/4, 1/5, 2/6, 3/7), and the kernel takes as long as
    // the fullest pipe: measured r08, 12, 13, 14 and 16 blocks all cost the time of 16 (38.6 / 38.6 / 39.1 / 40.2 ms per
    // 5e5 points on cfg5's tables).  A warp's pipe is %warpid & 3 (profiles/smsp_map2.cu; the hardware places the warps of a
    // second resident CTA in another rotation, different from launch to launch, so it has to be read, not assumed).  The
    // partial group is therefore dealt per pipe - floor(blocks / 4) each, the remainder to a run of pipes that starts where
    // the run of the CTA with the previous ticket on this SM ended, so that two resident CTAs put their extra blocks on
    // different pipes (13 blocks: 7, 7, 6, 6 per SM instead of 8, 8, 6, 4; 14 blocks: 7, 7, 7, 7).  Only who computes which columns changes, never the arithmetic of a column.
    // NB = 2 deals HALF blocks (a block for the points 0..15 or 16..31 of the tile: two of the four DMMAs per k-step), so that
    // the pipes differ by at most half a block: 13 blocks = 26 halves -> 7, 7, 6, 6 per CTA and 13, 13, 13, 13 per SM.  A
    // warp then has nbv full blocks and possibly one half block (block hblk, half hsel) in its next accumulator slot.
    int jb0 = 0, nbv = 0, hblk = -1, hsel = 0;
    // (compiled out of the four-block shape, which is at its register limit: its partial group is one of many, and these
    // values alive across the table build cost it 2 % - cfg3 584 -> 596 ms; it takes its blocks in warp order, below)
    if constexpr (NB < 4) {
    const int group_first = (int)blockIdx.y * NW * NB, group_blocks = max(0, min(NW * NB, a.nblk - group_first));
    jb0 = group_first + warp * NB;
    nbv = max(0, min(NB, a.nblk - jb0));
    __shared__ int s_jb0[NW], s_nb[NW], s_pipe[NW], s_hblk[NW], s_hsel[NW];
    if (a.tickets != nullptr && group_blocks < NW * NB) {
        if (lane == 0) {
            unsigned wid;
            asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
            s_pipe[warp] = (int)(wid & 3u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const int ticket = atomicAdd(a.tickets + (smid & 255), 1);
            deal_blocks(NW, NB, s_pipe, group_first, group_blocks, ticket, s_jb0, s_nb, s_hblk, s_hsel);  // (smx_plan.h)
        }
        __syncthreads();
        jb0 = s_jb0[warp], nbv = s_nb[warp], hblk = s_hblk[warp], hsel = s_hsel[warp];
    }
    }
    build_table<NW>(a, x, p0, tab);
    if constexpr (NB == 4) {
        jb0 = (blockIdx.y * NW + warp) * NB;
        nbv = max(0, min(NB, a.nblk - jb0));
    }
    const int nload = NB < 4 ? nbv + (hblk >= 0 ? 1 : 0) : nbv;  // accumulator slots in use: full blocks, then the half block

    // ---- main loop -----------------------------------------------------------------------------------------------------
    // The host pads the term list to whole stages of 16 k-steps with zero coefficients and appends two more stages of
    // zeros, so nothing in this loop needs a bounds check: no branch between the DMMAs of a stage.
    const int n_stage = a.k4 / NW;
    // (block indices are clamped so that every load has a valid address)
    const size_t bstride = (size_t)(a.k4 + kDensePadK4) * 32;
    const double* bbase = a.coef + lane;
    unsigned boff[NB];  // element offsets (the whole matrix has fewer than 2^32 elements: checked at upload)
#pragma unroll
    for (int j = 0; j < NB; ++j) boff[j] = (unsigned)((size_t)(NB < 4 && hblk >= 0 && j == nbv ? hblk : min(jb0 + j, a.nblk - 1)) * bstride);

    // A assembly: this thread owns (k-step `warp` of the stage, fragment lane `lane`): term 4 * k4 + tig, points gid + 8 i
    const double* xt = x + p0 * a.ldx;  // this tile's rows; row offsets of the lane's four points fit 32 bits
    int xoff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xoff[i] = (int)(min((long long)(gid + 8 * i), a.N - 1 - p0) * a.ldx);
    auto meta_of = [&](int stage) { return __ldg(a.meta + 4 * (stage * NW + warp) + tig); };
    // Leading entry on a cold column: pi = x - eta0, straight from x.  load_cold only REQUESTS the coordinates (and the centre);
    // the subtraction happens in assemble, a stage later.  Written as "xc = ldg(..) - e0" the compiler put the DADD right behind
    // the loads, in front of the stage's DMMAs, and every warp waited there for L2 once per stage (ncu r08: 9 % of all samples
    // on those DADDs with long-scoreboard stalls).  Measured r08, ms: cfg3 at full size (16 warps x 4 blocks) 616 -> 584, split-K
    // kernel at cfg4's tables and 16 outputs 2.89 -> 2.76; the skewed 8-warp shapes first lost 2.5 % with it (cfg5 77.1 ->
    // 79.1: the extra DADD lengthened the assembly chain in front of the stage barrier while the fullest FP64 pipe set the pace)
    // and gain 1.3 % since their blocks are dealt per pipe (cfg5 68.85 -> 67.98).
    constexpr bool kLateSub = true;  // (r08, after the per-pipe dealing: also the narrow shapes gain, cfg5 68.85 -> 67.98 ms; before it they lost 2.5 %)
    double xc[4] = {0.0, 0.0, 0.0, 0.0}, xe0 = 0.0;
    auto load_cold = [&](int2 m) {
        if (m.y < 0) {
            const int dim = -1 - m.y;
            const double e0 = __ldg(a.eta0 + dim);
            xe0 = kLateSub ? e0 : 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) xc[i] = kLateSub ? __ldg(xt + xoff[i] + dim) : __ldg(xt + xoff[i] + dim) - e0;
        }
    };
    auto assemble = [&](int2 m, int buf) {
        double v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double lead = m.y >= 0 ? tab[m.y * kTabPitch + gid + 8 * i] : (kLateSub ? xc[i] - xe0 : xc[i]);
            v[i] = tab[m.x * kTabPitch + gid + 8 * i] * lead;
        }
        // two 16-byte halves per lane, each half contiguous over the lanes: conflict-free stores here and loads below
        double2* dst = reinterpret_cast<double2*>(abuf + (buf * NW + warp) * 128) + lane;
        dst[0] = make_double2(v[0], v[1]);
        dst[32] = make_double2(v[2], v[3]);
    };

    // Skewed stages (a.skew): the warps of the upper half of every group of eight assemble their k-step of stage s + 1
    // BEFORE the DMMAs of stage s (from coordinates they requested a stage earlier), the others after them - so between
    // two stage barriers half of the CTA is always in its DMMA phase and the FP64 pipe does not drain while everybody
    // builds fragments at the same time.  Both orders respect the two-buffer protocol: buffer (s + 1) & 1 was last read
    // in stage s - 1, and it is complete at the barrier that ends stage s.
    // Measured (B200, ms per 10^6 points): cfg5 (13 blocks, 8 warps x 2 blocks) 82.6 -> 78.1; no gain for the 16-warp x 4-block
    // shape (cfg3: 611 vs 614), which waits for the pipe and for nothing else and has no registers to spare for the deeper
    // metadata look-ahead - compiled out there.
    constexpr bool kSkew = NB < 4;
    const bool early = kSkew && a.skew != 0 && ((warp >> 2) & 1) != 0;
    int2 m1 = meta_of(0);
    load_cold(m1);
    assemble(m1, 0);
    m1 = meta_of(1);
    int2 m2 = kSkew ? meta_of(2) : m1;
    if (early) load_cold(m1);
    __syncthreads();

    double acc[4][NB][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    double bq[PF][NB];  // register ring of B fragments, PF k-steps ahead
#pragma unroll
    for (int u = 0; u < PF; ++u)
#pragma unroll
        for (int j = 0; j < NB; ++j) bq[u][j] = j < nload ? __ldg(bbase + boff[j] + u * 32) : 0.0;

    // One instantiation per number of blocks the warp really has (the last warp of the last column group has fewer than NB):
    // a DMMA that is predicated off still occupies the FP64 pipe for its 16 cycles (profiles/r06_k1_experiments.md, r08), so
    // "j < nbv" must be resolved at compile time, not by a predicate.
    auto stage_mma = [&](int s, auto nb_tag, auto half_tag) {
        constexpr int NBV = decltype(nb_tag)::value;
        constexpr bool HALF = decltype(half_tag)::value;  // one more block in slot NBV, for two of the four point groups
        const double2* af = reinterpret_cast<const double2*>(abuf + (s & 1) * NW * 128) + lane;
        const double* bp = bbase + (size_t)s * NW * 32;
#pragma unroll
        for (int kk = 0; kk < NW; ++kk) {
            const double2 a01 = af[kk * 64], a23 = af[kk * 64 + 32];
            if (HALF) {
                constexpr int JH = NBV < NB ? NBV : NB - 1;
                const double2 ah = hsel ? a23 : a01;
                dmma(acc[0][JH], ah.x, bq[kk % PF][JH]);
                dmma(acc[1][JH], ah.y, bq[kk % PF][JH]);
                bq[kk % PF][JH] = __ldg(bp + boff[JH] + (kk + PF) * 32);
            }
#pragma unroll
            for (int j = 0; j < NBV; ++j) {
                dmma(acc[0][j], a01.x, bq[kk % PF][j]);
                dmma(acc[1][j], a01.y, bq[kk % PF][j]);
                dmma(acc[2][j], a23.x, bq[kk % PF][j]);
                dmma(acc[3][j], a23.y, bq[kk % PF][j]);
                // refill the slot just consumed: the fragment of k-step kk + PF (no second register set needed).  Only the
                // blocks this warp really has: the kernel runs at the L2 -> SM bandwidth limit (64 bytes of B per DMMA),
                // a duplicate load of a clamped block costs as much as a useful one
                bq[kk % PF][j] = __ldg(bp + boff[j] + (kk + PF) * 32);
            }
        }
    };
    // (the 16-warp x 4-block shape is at its register limit: extra instantiations cost it spills - 615 vs 607 ms at cfg3 - and
    // its only partial warp is the last one of the last of many column groups: that one multiplies its clamped duplicate blocks
    // too and drops them in the epilogue)
    auto stage_mma_any = [&](int s) {
        if (nbv == NB || (NB == 4 && nbv > 0)) stage_mma(s, std::integral_constant<int, NB>(), std::false_type());
        else if (NB == 2 && nbv == 1 && hblk < 0) stage_mma(s, std::integral_constant<int, 1>(), std::false_type());
        else if (NB == 2 && nbv == 1) stage_mma(s, std::integral_constant<int, 1>(), std::true_type());
        else if (NB == 2 && hblk >= 0) stage_mma(s, std::integral_constant<int, 0>(), std::true_type());
    };
    static_assert(NB == 1 || NB == 2 || NB == 4, "partial warps: NB = 2 has its own instantiation, NB = 4 clamps (below)");

    for (int s = 0; s < n_stage; ++s) {
        const int2 mn = meta_of(s + (kSkew ? 3 : 2));
        if (early) {
            assemble(m1, (s + 1) & 1);
            load_cold(m2);  // for stage s + 2: consumed at the start of the next iteration
        } else {
            load_cold(m1);  // for stage s + 1; consumed after the DMMAs below
        }
        stage_mma_any(s);
        if (!early) assemble(m1, (s + 1) & 1);
        if (kSkew) m1 = m2, m2 = mn;
        else m1 = mn;
        __syncthreads();
    }

    // ---- epilogue ------------------------------------------------------------------------------------------------------
    const bool vec = a.colmap == nullptr && (a.ldy & 1) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        if (j >= nload) break;
        const bool is_half = NB < 4 && j >= nbv;  // accumulators 0, 1 of the slot hold point groups 2 hsel, 2 hsel + 1
        const long long col = 8ll * (is_half ? hblk : jb0 + j) + 2 * tig;
        if (col >= a.ncol) continue;
        const bool two = col + 1 < a.ncol;
        const double c0a = __ldg(a.c0 + col), c0b = two ? __ldg(a.c0 + col + 1) : 0.0;
        const long long ya = a.colmap ? __ldg(a.colmap + col) : col, yb = a.colmap && two ? __ldg(a.colmap + col + 1) : col + 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (is_half && i >= 2) break;
            const long long p = p0 + gid + 8 * (is_half ? 2 * hsel + i : i);
            if (p >= a.N) continue;
            double* dst = y + p * a.ldy;
            if (vec) {
                *reinterpret_cast<double2*>(dst + ya) = make_double2(c0a + acc[i][j][0], c0b + acc[i][j][1]);
            } else {
                dst[ya] = c0a + acc[i][j][0];
                if (two) dst[yb] = c0b + acc[i][j][1];
            }
        }
    }
}

// ---- few outputs (up to 8 blocks = 64 columns per CTA): split K over the warps -------------------------------------
// With few output blocks there are not enough column tiles to give every warp its own, and sharing A through shared memory
// costs a barrier per stage.  Here every warp multiplies ALL NBT output blocks of the CTA for its own k-steps
// (warp w: k-steps w, w + NW, ..): it builds its A fragment in registers straight from the value table (no staging, no
// barrier in the main loop), keeps 4 x NBT accumulator tiles, and streams the B fragments of the next k-step into the
// registers the DMMAs have just consumed.  The NW partial results are added in fixed order through shared memory.
// B fragment load that stays where it is written (volatile asm statements keep their order: the refill of a register
// must not be hoisted above the DMMAs that still read it, or the ring needs a second set of registers)
__device__ __forceinline__ double ldg_pinned(const double* p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

template <int NW, int NBT, bool FULL>
__global__ void __launch_bounds__(NW * 32, 1)
dense_splitk_kernel(const DenseArgs a, const double* __restrict__ x, double* __restrict__ y) {
    constexpr int kThreads = NW * 32;
    constexpr int kCols = 8 * NBT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tab = reinterpret_cast<double*>(smem_raw);  // [n_tab][kTabPitch]; afterwards [32][kCols + 2] partial sums

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tig = lane & 3, gid = lane >> 2;
    const long long p0 = (long long)blockIdx.x * kDenseTile;
    build_table<NW>(a, x, p0, tab);

    const int jb0 = blockIdx.y * NBT;
    const int nbv = min(NBT, a.nblk - jb0);
    const size_t bstride = (size_t)(a.k4 + kDensePadK4) * 32;
    const double* bbase = a.coef + lane;
    unsigned boff[NBT];
#pragma unroll
    for (int j = 0; j < NBT; ++j) boff[j] = (unsigned)((size_t)min(jb0 + j, a.nblk - 1) * bstride);

    const double* xt = x + p0 * a.ldx;
    int xoff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xoff[i] = (int)(min((long long)(gid + 8 * i), a.N - 1 - p0) * a.ldx);
    auto meta_of = [&](int g) { return __ldg(a.meta + 4 * g + tig); };  // (arrays are padded: look-ahead needs no check)
    double xc[4] = {0.0, 0.0, 0.0, 0.0}, xe0 = 0.0;
    auto load_cold = [&](int2 m) {  // (requests only; the subtraction waits until assemble - see dense_eval_kernel)
        if (m.y < 0) {
            const int dim = -1 - m.y;
            xe0 = __ldg(a.eta0 + dim);
#pragma unroll
            for (int i = 0; i < 4; ++i) xc[i] = __ldg(xt + xoff[i] + dim);
        }
    };
    auto assemble = [&](int2 m, double (&v)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double lead = m.y >= 0 ? tab[m.y * kTabPitch + gid + 8 * i] : xc[i] - xe0;
            v[i] = tab[m.x * kTabPitch + gid + 8 * i] * lead;
        }
    };

    double acc[4][NBT][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NBT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // software pipeline over this warp's k-steps g, g + NW, ..: A values one step ahead, cold x and metadata two and three
    int g = warp;
    double a_cur[4], a_nxt[4];
    int2 m1 = meta_of(g);
    load_cold(m1);
    assemble(m1, a_cur);
    m1 = meta_of(g + NW);
    load_cold(m1);
    int2 m2 = meta_of(g + 2 * NW);
    double bq[NBT];
#pragma unroll
    for (int j = 0; j < NBT; ++j) bq[j] = __ldg(bbase + boff[j] + (size_t)g * 32);

    for (; g < a.k4; g += NW) {
        assemble(m1, a_nxt);                     // k-step g + NW (its cold x was loaded one iteration ago)
        load_cold(m2);                           // k-step g + 2 NW
        const int2 m3 = meta_of(g + 3 * NW);
        const double* bn = bbase + (size_t)(g + NW) * 32;
#pragma unroll
        for (int j = 0; j < NBT; ++j) {
            if (FULL || j < nbv) {
                dmma(acc[0][j], a_cur[0], bq[j]);
                dmma(acc[1][j], a_cur[1], bq[j]);
                dmma(acc[2][j], a_cur[2], bq[j]);
                dmma(acc[3][j], a_cur[3], bq[j]);
            }
            if (FULL || j < nbv) bq[j] = ldg_pinned(bn + boff[j]);  // refill the registers just consumed (next k-step)
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) a_cur[i] = a_nxt[i];
        m1 = m2;
        m2 = m3;
    }

    // ---- add the warps' partial sums in fixed order (deterministic), then one coalesced store --------------------------
    constexpr int kRedPitch = kCols + 2;
    double* red = tab;
    __syncthreads();  // the value table is dead
#pragma unroll 1
    for (int w = 0; w < NW; ++w) {
        if (warp == w) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NBT; ++j) {
                    double2* dst = reinterpret_cast<double2*>(red + (gid + 8 * i) * kRedPitch + 8 * j + 2 * tig);
                    double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
                    if (w > 0) {
                        const double2 old = *dst;
                        v.x += old.x, v.y += old.y;
                    }
                    *dst = v;
                }
        }
        __syncthreads();
    }
    const int ncols = (int)min((long long)kCols, a.ncol - 8ll * jb0);
    for (int idx = tid; idx < kDenseTile * ncols; idx += kThreads) {
        const int p = idx / ncols, c = idx - p * ncols;
        if (p0 + p >= a.N) break;
        const long long col = 8ll * jb0 + c;
        y[(p0 + p) * a.ldy + (a.colmap ? __ldg(a.colmap + col) : col)] = __ldg(a.c0 + col) + red[p * kRedPitch + c];
    }
}

size_t splitk_smem_bytes(int n_tab, int nbt) {
    return sizeof(double) * std::max((size_t)n_tab * kTabPitch, (size_t)kDenseTile * (8 * nbt + 2));
}

template <int NW, int NBT, bool FULL>
int launch_splitk_(const DenseArgs& a, const double* x, double* y, cudaStream_t st) {
    const size_t smem = splitk_smem_bytes(a.n_tab, NBT);
    static size_t opted = 0;
    if (smem > opted) {
        SMX_CUDA(cudaFuncSetAttribute(dense_splitk_kernel<NW, NBT, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        opted = smem;
    }
    const long long tiles = (a.N + kDenseTile - 1) / kDenseTile;
    const int groups = (a.nblk + NBT - 1) / NBT;
    dense_splitk_kernel<NW, NBT, FULL><<<dim3((unsigned)tiles, (unsigned)groups), NW * 32, smem, st>>>(a, x, y);
    SMX_LAUNCH_CHECK("dense_splitk_kernel<%d,%d,%d>", NW, NBT, (int)FULL);
    return SMX_OK;
}

template <int NW, int NBT>
int launch_splitk(const DenseArgs& a, const double* x, double* y, cudaStream_t st) {
    // every column group full: no per-block predicate in the inner loop
    return a.nblk % NBT == 0 ? launch_splitk_<NW, NBT, true>(a, x, y, st) : launch_splitk_<NW, NBT, false>(a, x, y, st);
}

size_t dense_smem_bytes(int n_tab, int nw) { return sizeof(double) * ((size_t)2 * nw * 128 + (size_t)n_tab * kTabPitch); }

template <int NW, int NB, int PF, int CTAS>
int launch(const DenseArgs& a, const double* x, double* y, cudaStream_t st) {
    const size_t smem = dense_smem_bytes(a.n_tab, NW);
    static size_t opted = 0;
    if (smem > opted) {
        SMX_CUDA(cudaFuncSetAttribute(dense_eval_kernel<NW, NB, PF, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        opted = smem;
    }
    const long long tiles = (a.N + kDenseTile - 1) / kDenseTile;
    const int groups = (a.nblk + NW * NB - 1) / (NW * NB);
    dense_eval_kernel<NW, NB, PF, CTAS><<<dim3((unsigned)tiles, (unsigned)groups), NW * 32, smem, st>>>(a, x, y);
    SMX_LAUNCH_CHECK("dense_eval_kernel<%d,%d,%d,%d>", NW, NB, PF, CTAS);
    return SMX_OK;
}

}  // namespace

bool dense_kernel_fits(int n_tab, int smem_optin) { return dense_smem_bytes(n_tab, 16) <= (size_t)smem_optin; }

// CTA shape.  Many outputs: 16 warps x 4 output blocks (32 points x 512 outputs per CTA; the widest register tile, fewest
// A-fragment loads per DMMA).  Few outputs: 8 warps x 1 or 2 blocks with a deeper B ring, two CTAs per SM when the value
// table leaves room (more independent DMMA chains and loads in flight per SM).
int dense_kernel_launch(const DenseArgs& args, const double* x, double* y, cudaStream_t st) {
    static const int want_nb = tune_int("SMX_DENSE_NB", 0);
    static const int want_nw = tune_int("SMX_DENSE_NW", 0);
    static const int want_ctas = tune_int("SMX_DENSE_CTAS", 0);
    static const int want_skew = tune_int("SMX_DENSE_SKEW", 1);
    DenseArgs a = args;
    a.skew = want_skew;
    int device = 0, smem_sm = 0;
    SMX_CUDA(cudaGetDevice(&device));
    SMX_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
    static const int want_splitk = tune_int("SMX_DENSE_SPLITK", -1);
    // up to 40 output blocks (320 columns): column groups of at most 8 blocks (the accumulators of 32 points x 64 columns
    // are 128 registers per thread), K split over the warps of a CTA
    // (measured: 0.76 vs 0.63 of the DMMA rate at 8 blocks; with two or more groups the staged kernel below wins)
    if (want_splitk != 0 && !want_nb && (a.nblk <= 8 || want_splitk > 0)) {
        // one instantiation per number of blocks per group: a DMMA that is predicated off (block j >= nbv of a wider
        // instantiation) still occupies the FP64 pipe for its 16 cycles - at 2 blocks on the 4-block kernel that was 60 % of
        // the kernel's time (measured r08: cfg4 tables, d_out = 10: 4.98 ms per 10^5 points before)
        const int groups = (a.nblk + 7) / 8, per = (a.nblk + groups - 1) / groups;
        switch (per) {
            case 1: return launch_splitk<16, 1>(a, x, y, st);
            case 2: return launch_splitk<16, 2>(a, x, y, st);
            case 3: return launch_splitk<16, 3>(a, x, y, st);
            case 4: return launch_splitk<16, 4>(a, x, y, st);
            case 5: return launch_splitk<8, 5>(a, x, y, st);
            case 6: return launch_splitk<8, 6>(a, x, y, st);
            case 7: return launch_splitk<8, 7>(a, x, y, st);
            default: return launch_splitk<8, 8>(a, x, y, st);
        }
    }
    const bool two_fit = 2 * (dense_smem_bytes(a.n_tab, 8) + 1024) <= (size_t)smem_sm;
    int nb = a.nblk >= 48 ? 4 : a.nblk > 8 ? 2 : 1, nw = nb == 4 ? 16 : 8;
    // a value table too large for two CTAs of 8 warps per SM (many hot rows, e.g. cfg3's tables): one CTA of 16 warps, one
    // block per warp up to 24 blocks (measured, fraction of the DMMA rate at 13 / 20 / 40 blocks: 8 warps x 2 blocks
    // 0.59 / 0.58 / 0.72, 16 x 1: 0.62 / 0.71 / 0.79, 16 x 2: 0.57 / 0.69 / 0.80)
    if (nb < 4 && !two_fit) nw = 16, nb = a.nblk <= 24 ? 1 : 2;
    if (want_nb) nb = want_nb;
    if (want_nw) nw = want_nw;
    const int ctas = want_ctas ? want_ctas : (two_fit ? 2 : 1);
    if (nb == 4) return nw <= 8 ? launch<8, 4, 4, 1>(a, x, y, st) : launch<16, 4, 2, 1>(a, x, y, st);
    if (nb == 2) {
        if (nw > 8) return launch<16, 2, 4, 1>(a, x, y, st);
        return ctas == 2 && two_fit ? launch<8, 2, 8, 2>(a, x, y, st) : launch<8, 2, 8, 1>(a, x, y, st);
    }
    if (nw > 8) return launch<16, 1, 8, 1>(a, x, y, st);
    return ctas == 2 && two_fit ? launch<8, 1, 8, 2>(a, x, y, st) : launch<8, 1, 8, 1>(a, x, y, st);
}

}  // namespace smx
