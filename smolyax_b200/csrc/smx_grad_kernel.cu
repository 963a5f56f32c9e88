// K3: gradient of the interpolant on the hierarchical layout.  Replaces reference interpolation.py:306-345 (python loop over
// groups and summand batches around jit(vmap(barycentric.evaluate_tensor_product_gradient)), which materialises a dense
// (N, d_out, d_in) tensor PER SUMMAND and adds them up) by one kernel that writes every entry of J (N, d_out, d_in) exactly once.
//
//   dI/dx_j, j cold (only the pair (j, 1) exists, pi = x_j - eta_0):   the row sum  acc[p][e] = sum_r C[r][e] m_r(x_p)  of the
//       value contraction itself - one JOB per cold block of 16 columns, its value items accumulated, 16 columns stored.
//   dI/dx_i, i hot:   a polynomial over the same term set whose coefficients the plan compiler derives from the value
//       coefficients (smx_plan.cpp 7b); only the terms that contain dimension i carry one, so the derivative has work items of
//       its own (cfg2: 41 hot dimensions, 502 items in all, 5 x the value pass) - one JOB per hot dimension, one number per point.
// A job is run by ONE warp of the CTA that owns the tile of 32 points (static longest-processing-time assignment), so nothing
// is ever added to J from two places: no zero-fill pass, no atomics, results bit-reproducible.  Columns without any entry are
// zeroed by the kernel.  The reference returns NaN in dimension j where x_pj sits on an interpolation node of j
// (barycentric.py:152-154): folded into the stores (flags of the hot dimensions from the value-table prologue, the two nodes
// of a cold column's degree-1 rule compared with the staged x tile).
//
// Item machinery, shared-memory layout, lane mapping: the lean value kernel's (smx_fast_kernel.cu) - value table of the tile in
// shared memory, x tile by TMA (128-byte swizzle), record = metadata + coefficients packed as DMMA B fragments by one bulk
// copy, one item ahead, completion on an mbarrier; acc = A (table rows) x B on mma.sync.m8n8k4.f64.
#include <cuda.h>

#include <algorithm>
#include <type_traits>

#include "smx_fast_device.cuh"

namespace smx {
namespace {

// directory flags of a gradient item (.z low byte)
constexpr int kGHot = 1;       // leading entries are hot: their basis values are rows of the value table
constexpr int kGNeedX = 2;     // the x tile of the block is staged with the item
constexpr int kGFirst = 4;     // first item of its job
constexpr int kGLast = 8;      // last item of its job
constexpr int kGEtaZero = 16;  // cold block whose first centres are all zero: pi = x
constexpr int kGColdJob = 32;  // the job stores the row sums of a cold block (else: one number per point for a hot dimension)

using Stage = LeanStage<false>;

struct GradArgs {
    const double* eta;
    const int32_t* hot_off;
    const int32_t* hot_pos;
    const int4* dir;         // per item (warp order): record offset (128-byte units), job, flags | nf << 8 | units << 16 | k-steps << 24, first column
    const int4* jobs;        // per job: kind 0: (store mask, offset of its 32 nodes, first column, 0); kind 1: (column, 0, 0, 1)
    const double* job_c0;    // [job][d_out]
    const double* job_nodes;
    const int32_t* zero_cols;
    const double* records;
    const int32_t* nan_off;  // per dimension: the nodes at which the reference returns NaN
    const double* nan_nodes;
    long long N, ldx, d_in, d_out, num_tiles;
    int n_items, n_jobs, n_zero, n_hot, n_hot_rows, hot_dims;
    int nan_at_nodes;
    int warp_off[kMaxWarps + 1];
};

__device__ __forceinline__ void stage_grad(const CUtensorMap* xmap, const GradArgs& a, const void* item, const void* bar, double* xs, const int4 dir,
                                           int o, int p0) {
    const unsigned units = ((unsigned)dir.z >> 16) & 0xffu, bytes = units << 7;
    unsigned long long* b = const_cast<unsigned long long*>(static_cast<const unsigned long long*>(bar));
    const bool needx = dir.z & kGNeedX;
    mbar_expect_tx(b, bytes + (needx ? kXTileBytes : 0));
    bulk_copy(const_cast<void*>(item), a.records + ((size_t)(unsigned)dir.x + (size_t)o * units) * 16, bytes, b);
    if (needx) tma_load_2d(xs, xmap, dir.w, p0, b);
}

template <int NW, bool ETA0>
__global__ void __launch_bounds__(NW * 32, 2)
grad_kernel(const __grid_constant__ CUtensorMap xmap, const GradArgs a, const double* __restrict__ x, double* __restrict__ J) {
    constexpr int kThreads = NW * 32;
    extern __shared__ __align__(1024) unsigned char smem_grad[];
    if ((smem_u32(smem_grad) & 1023u) != 0) __trap();
    XTile* xtiles = reinterpret_cast<XTile*>(smem_grad);                                      // [NW]
    Stage* stages = reinterpret_cast<Stage*>(xtiles + NW);                                    // [NW][2]
    double* tab = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(stages) + NW * lean_stage_bytes<false>());  // [1 + n_hot_rows][kTabPitch]
    int4* s_dir = reinterpret_cast<int4*>(tab + (size_t)(1 + a.n_hot_rows) * kTabPitch);     // [n_items + 1]
    int4* s_jobs = s_dir + a.n_items + 1;                                                     // [n_jobs]
    double* s_eta = reinterpret_cast<double*>(s_jobs + a.n_jobs);                             // [n_hot]
    int* s_hot_off = reinterpret_cast<int*>(s_eta + a.n_hot);                                 // [hot_dims + 1]
    int* s_hot_row = s_hot_off + a.hot_dims + 1;                                              // [n_hot]
    unsigned* s_nan = reinterpret_cast<unsigned*>(s_hot_row + a.n_hot);                       // [hot_dims] bit p: x[p][dim] sits on a node

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tig = lane & 3, gid = lane >> 2;
    Stage* st = stages + 2 * warp;
    double* xs = xtiles[warp].v;
    const double* xlo = xs + gid * kBlockWidth + (((2 * tig) ^ gid) << 1);
    const double* xhi = xs + gid * kBlockWidth + (((2 * tig + 1) ^ gid) << 1);
    const double* tabq = tab + 2 * gid;

    for (int i = tid; i < a.n_items; i += kThreads) s_dir[i] = __ldg(a.dir + i);
    for (int i = tid; i < a.n_jobs; i += kThreads) s_jobs[i] = __ldg(a.jobs + i);
    for (int i = tid; i < a.n_hot; i += kThreads) s_eta[i] = __ldg(a.eta + i);
    for (int i = tid; i <= a.hot_dims; i += kThreads) s_hot_off[i] = __ldg(a.hot_off + i);
    for (int i = tid; i < a.n_hot; i += kThreads) s_hot_row[i] = 1 + hot_row(__ldg(a.hot_pos + i));
    if (lane == 0) {
        mbar_init(&st[0].bar, 1);
        mbar_init(&st[1].bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    unsigned k_item = 0;  // items this warp has consumed: buffer = k & 1, parity = (k >> 1) & 1
    const int c_begin = a.warp_off[warp], c_end = a.warp_off[warp + 1];
    const int n_out = (int)a.d_out;

    for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int p0 = (int)(tile * kTile);  // (TMA coordinates are 32-bit: N < 2^31, checked at launch)
        // ---- prologue: value table (ones row + 1-D basis values of the hot entries), node flags of the hot dimensions ----
        if (tid < kTile) tab[tid] = 1.0;
        {
            const int slot = t_slot(lane);
            const double* xrow = x + min((long long)p0 + lane, a.N - 1) * a.ldx;
            for (int d = warp; d < a.hot_dims; d += NW) {
                const double xv = __ldg(xrow + d);
                double v = 1.0;
                for (int k = s_hot_off[d]; k < s_hot_off[d + 1]; ++k) {
                    v *= (xv - s_eta[k]);
                    tab[s_hot_row[k] * kTabPitch + slot] = v;
                }
                bool hit = false;
                if (a.nan_at_nodes)
                    for (int k = __ldg(a.nan_off + d); k < __ldg(a.nan_off + d + 1); ++k) hit |= (xv == __ldg(a.nan_nodes + k));
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) s_nan[d] = m;
            }
        }
        // columns of J no job writes (dimensions without any entry): zero
        for (int z = 0; z < a.n_zero; ++z) {
            const int lo = __ldg(a.zero_cols + 2 * z), hi = __ldg(a.zero_cols + 2 * z + 1), w = hi - lo;
            for (long long i = tid; i < (long long)kTile * n_out * w; i += kThreads) {
                const int c = (int)(i % w);
                const long long po = i / w;  // point * n_out + output
                if ((long long)p0 + po / n_out < a.N) J[((long long)p0 * n_out + po) * a.d_in + lo + c] = 0.0;
            }
        }
        __syncthreads();

        // ---- main: this warp's jobs, for every output (outputs spread over gridDim.y when there are few tiles) ----------------
        if (c_begin < c_end) {
            const int o_first = (int)blockIdx.y, o_step = (int)gridDim.y;
            if (o_first < n_out && lane == 0) stage_grad(&xmap, a, &st[k_item & 1].item, &st[k_item & 1].bar, xs, s_dir[c_begin], o_first, p0);
            for (int o = o_first; o < n_out; o += o_step) {
                double acc[4][2][2];   // cold job: row sums of the block, accumulated over its items
                double tot[4] = {0.0, 0.0, 0.0, 0.0};  // hot-dimension job: the derivative at the lane's 4 points (partial over tig)
                for (int c = c_begin; c < c_end; ++c, ++k_item) {
                    const int buf = k_item & 1;
                    const int4 dir = s_dir[c];
                    const ItemBuffer& ib = st[buf].item;
                    const int ksteps = (unsigned)dir.z >> 24;
                    const int nf = (dir.z >> 8) & 7;
                    mbar_wait(&st[buf].bar, (k_item >> 1) & 1);

                    auto load_a = [&](int s, double2& lo, double2& hi) {  // A fragment of k-step s: 4 points of this lane's row
                        const int4 f = ib.fac[4 * s + tig];
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
                    // what the item needs from the x tile, taken before the buffer is handed to the next item
                    double2 q0[4], q1[4];
                    if (dir.z & kGNeedX) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            q0[i] = *reinterpret_cast<const double2*>(xlo + i * (8 * kBlockWidth));
                            q1[i] = *reinterpret_cast<const double2*>(xhi + i * (8 * kBlockWidth));
                        }
                    }
                    int4 t4 = make_int4(0, 0, 0, 0);
                    if (dir.z & kGHot) t4 = *reinterpret_cast<const int4*>(ib.tab + 4 * tig);
                    double2 ea = make_double2(0.0, 0.0), eb = ea;
                    if (!ETA0 && !(dir.z & (kGHot | kGEtaZero | kGColdJob))) {
                        ea = *reinterpret_cast<const double2*>(ib.eta0 + 4 * tig);
                        eb = *reinterpret_cast<const double2*>(ib.eta0 + 4 * tig + 2);
                    }
                    const double2 b0 = *reinterpret_cast<const double2*>(ib.coef + 2 * lane);
                    double2 a0lo, a0hi;
                    load_a(0, a0lo, a0hi);
                    // the first k-step's operands are in registers; the later ones are read from the item buffer below, so the
                    // next item goes into the OTHER buffer; the x tile is free once q0 / q1 are taken
                    __syncwarp();
                    {
                        int cn = c + 1, on = o;
                        if (cn == c_end) cn = c_begin, on = o + o_step;
                        if (on < n_out && lane == 0) stage_grad(&xmap, a, &st[buf ^ 1].item, &st[buf ^ 1].bar, xs, s_dir[cn], on, p0);
                    }

                    double part[4][2][2];
                    {
                        const double af[4] = {a0lo.x, a0lo.y, a0hi.x, a0hi.y};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            dmma_first<0>(part[i][0], af[i], b0.x);
                            dmma_first<0>(part[i][1], af[i], b0.y);
                        }
                    }
#pragma unroll 1
                    for (int s = 1; s < ksteps; ++s) {
                        double2 a01, a23;
                        load_a(s, a01, a23);
                        const double2 b = *reinterpret_cast<const double2*>(ib.coef + s * kKStepDoubles + 2 * lane);
                        const double ag[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            dmma_<0>(part[i][0], ag[i], b.x);
                            dmma_<0>(part[i][1], ag[i], b.y);
                        }
                    }

                    if (dir.z & kGColdJob) {
                        // ---- job of a cold block: acc += part; at its last item the 16 columns are stored ----
                        if (dir.z & kGFirst) {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
#pragma unroll
                                for (int j = 0; j < 2; ++j) acc[i][j][0] = part[i][j][0], acc[i][j][1] = part[i][j][1];
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
#pragma unroll
                                for (int j = 0; j < 2; ++j) acc[i][j][0] += part[i][j][0], acc[i][j][1] += part[i][j][1];
                        }
                        if (dir.z & kGLast) {
                            const int4 job = s_jobs[dir.y];  // (store mask, offset of the nodes, first column, kind)
                            const int m4 = (job.x >> (4 * tig)) & 15;  // this lane's entries 4 tig .. 4 tig + 3
                            double n0[4] = {0, 0, 0, 0}, n1[4] = {0, 0, 0, 0};
                            if (a.nan_at_nodes) {
                                const double* nd = a.job_nodes + job.y + 4 * tig;
#pragma unroll
                                for (int e = 0; e < 4; ++e) n0[e] = __ldg(nd + e), n1[e] = __ldg(nd + 16 + e);
                            }
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const long long p = (long long)p0 + gid + 8 * i;
                                if (p < a.N) {
                                    double* jr = J + (p * n_out + o) * a.d_in + job.z + 4 * tig;
                                    double g4[4] = {acc[i][0][0], acc[i][0][1], acc[i][1][0], acc[i][1][1]};
                                    const double x4[4] = {q0[i].x, q0[i].y, q1[i].x, q1[i].y};
#pragma unroll
                                    for (int e = 0; e < 4; ++e)
                                        if (a.nan_at_nodes && (x4[e] == n0[e] || x4[e] == n1[e])) g4[e] = __longlong_as_double(0x7ff8000000000000ll);
                                    // the lane's four columns are one 32-byte sector of J: written whole when it may be (a
                                    // partial sector makes L2 fetch the rest from HBM before it can write it back)
                                    if (m4 == 15 && (reinterpret_cast<unsigned long long>(jr) & 31) == 0) {
                                        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(jr), "d"(g4[0]), "d"(g4[1]), "d"(g4[2]), "d"(g4[3]) : "memory");
                                    } else {
#pragma unroll
                                        for (int e = 0; e < 4; ++e)
                                            if (m4 >> e & 1) jr[e] = g4[e];
                                    }
                                }
                            }
                        }
                    } else {
                        // ---- job of a hot dimension: tot += sum_e pi_e * part[e]; at its last item one column is stored ----
                        if (dir.z & kGFirst) tot[0] = tot[1] = tot[2] = tot[3] = 0.0;
                        if (dir.z & kGHot) {
                            const int tabs[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const double2 lo = *reinterpret_cast<const double2*>(tabq + tabs[e]);
                                const double2 hi = *reinterpret_cast<const double2*>(tabq + tabs[e] + 16);
                                const double v4[4] = {lo.x, lo.y, hi.x, hi.y};
#pragma unroll
                                for (int i = 0; i < 4; ++i) tot[i] = fma(v4[i], part[i][e >> 1][e & 1], tot[i]);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                tot[i] = fma(q0[i].x - ea.x, part[i][0][0], tot[i]);
                                tot[i] = fma(q0[i].y - ea.y, part[i][0][1], tot[i]);
                                tot[i] = fma(q1[i].x - eb.x, part[i][1][0], tot[i]);
                                tot[i] = fma(q1[i].y - eb.y, part[i][1][1], tot[i]);
                            }
                        }
                        if (dir.z & kGLast) {
                            const int4 job = s_jobs[dir.y];  // (column, ..)
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                tot[i] += __shfl_xor_sync(0xffffffffu, tot[i], 1);
                                tot[i] += __shfl_xor_sync(0xffffffffu, tot[i], 2);
                            }
                            if (tig == 0) {
                                const double c0 = __ldg(a.job_c0 + (size_t)dir.y * n_out + o);
                                const unsigned nanmask = a.nan_at_nodes ? s_nan[job.x] : 0u;
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int pt = gid + 8 * i;
                                    const long long p = (long long)p0 + pt;
                                    if (p < a.N)
                                        J[(p * n_out + o) * a.d_in + job.x] = (nanmask >> pt & 1) ? __longlong_as_double(0x7ff8000000000000ll) : c0 + tot[i];
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();  // the value table is rebuilt by the next tile's prologue
    }
}

size_t grad_smem_bytes(const GradArgs& a, int nw) {
    return (size_t)nw * (sizeof(XTile) + lean_stage_bytes<false>()) + (size_t)(1 + a.n_hot_rows) * kTabPitch * sizeof(double) +
           ((size_t)a.n_items + 1 + a.n_jobs) * sizeof(int4) + (size_t)a.n_hot * sizeof(double) +
           ((size_t)a.hot_dims + 1 + a.n_hot + a.hot_dims) * sizeof(int) + 64;
}

template <int NW, bool ETA0>
int launch_grad(const CUtensorMap& map, const GradArgs& a, int sm_count, const double* x, double* J, cudaStream_t st) {
    const size_t smem = grad_smem_bytes(a, NW);
    SMX_CUDA(cudaFuncSetAttribute(grad_kernel<NW, ETA0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long slots = (long long)sm_count * 2, grid = std::min<long long>(a.num_tiles, slots);
    // fewer tiles than CTA slots: spread the outputs over gridDim.y instead of walking them one after the other
    const long long gy = a.num_tiles < slots ? std::min<long long>(a.d_out, (slots + a.num_tiles - 1) / a.num_tiles) : 1;
    grad_kernel<NW, ETA0><<<dim3((unsigned)grid, (unsigned)gy), NW * 32, smem, st>>>(map, a, x, J);
    SMX_LAUNCH_CHECK("grad_kernel<%d,%d>", NW, (int)ETA0);
    return SMX_OK;
}

}  // namespace

// Warps per CTA for the gradient kernel of this plan: 8 or 6 with two CTAs per SM, 0 if its tables do not fit.
int grad_kernel_warps(const GradDevice& g, const FastDevice& d, int smem_sm) {
    GradArgs a{};
    a.n_items = g.n_items, a.n_jobs = g.n_jobs, a.n_hot = d.n_hot, a.n_hot_rows = d.n_hot_rows, a.hot_dims = d.hot_dims;
    // (an 8-warp instantiation exists no more: at its 128-register cap it spilled 56-88 bytes per thread and lost to 6 warps
    //  wherever both fit - r08: cfg1 0.067 vs 0.058 ms per 10^4 points, cfg2 0.426 vs 0.423 ms per 37 888)
    for (int nw : {6, 4})
        if (2 * (grad_smem_bytes(a, nw) + 1024) <= (size_t)smem_sm) return nw;
    return 0;
}

int grad_kernel_launch(const FastDevice& d, const GradDevice& g, const double* x, int64_t N, int64_t ldx, double* J, bool nan_at_nodes,
                       cudaStream_t st) {
    if (N >= (1ll << 31) - kTile) return fail(SMX_ERR_UNSUPPORTED, "more than 2^31 points in one call");
    CUtensorMap map;
    int rc = make_x_tensor_map(&map, x, d.d_in, N, ldx);
    if (rc) return rc;
    GradArgs a{};
    a.eta = d.eta, a.hot_off = d.hot_off, a.hot_pos = d.hot_pos;
    a.dir = reinterpret_cast<const int4*>(g.dir), a.jobs = reinterpret_cast<const int4*>(g.jobs);
    a.job_c0 = g.job_c0, a.job_nodes = g.job_nodes, a.zero_cols = g.zero_cols, a.records = g.records;
    a.nan_off = d.nan_off, a.nan_nodes = d.nan_nodes;
    a.N = N, a.ldx = ldx, a.d_in = d.d_in, a.d_out = d.d_out, a.num_tiles = (N + kTile - 1) / kTile;
    a.n_items = g.n_items, a.n_jobs = g.n_jobs, a.n_zero = g.n_zero, a.n_hot = d.n_hot, a.n_hot_rows = d.n_hot_rows, a.hot_dims = d.hot_dims;
    a.nan_at_nodes = nan_at_nodes ? 1 : 0;
    for (int w = 0; w <= kMaxWarps; ++w) a.warp_off[w] = g.warp_off[w];
    const bool eta0 = d.eta0_zero;
    switch (g.warps) {
        case 6: return eta0 ? launch_grad<6, true>(map, a, d.sm_count, x, J, st) : launch_grad<6, false>(map, a, d.sm_count, x, J, st);
        case 4: return eta0 ? launch_grad<4, true>(map, a, d.sm_count, x, J, st) : launch_grad<4, false>(map, a, d.sm_count, x, J, st);
    }
    return fail(SMX_ERR_UNSUPPORTED, "the gradient tables of this plan do not fit in shared memory");
}

}  // namespace smx
