// Fast evaluation path (K1): fused 1-D basis + block-sparse Kronecker contraction on the hierarchical layout of
// smx_plan.h.  Replaces the whole of reference interpolation.py:281-302 (python loop over groups and summand
// batches around jit(vmap(barycentric.evaluate_tensor_product_interpolant))) by ONE persistent kernel.
//
//   I(x_p) = c_0 + sum_{e=(j,a)} pi_e(x_pj) * sum_r C[r][e] * m_r(x_p)
//
// Mapping (one CTA = one tile of 32 evaluation points at a time; 4 warps and two CTAs per SM, or 8 warps and one):
//   * prologue   : the CTA fills the value table in shared memory ([row][32 points], 256 B per row): row 0 = 1, then the
//                  1-D basis values pi_e(x_p) of the hot entries (computed from x in registers), then, level by
//                  level, the products of two and more hot pairs (parent row * hot row).
//   * main       : warp w takes the work items (block of 16 leading entries x <= 24 rows) w, w + NW, ..  One item ahead
//                  of its use, the item's coefficient rows (L2 -> smem) and, for cold blocks, its 32 x 16 tile of x
//                  (HBM -> smem, 16-byte cp.async over fully used 128-byte lines; every coordinate of x is read once)
//                  are staged asynchronously into a per-warp double buffer.  A lane owns 4 entries x 4 points
//                  (lane = 4 * point-group + entry-group): it forms its 16 leading basis values pi_e(x) in registers
//                  (from the staged x, or from the value table for hot blocks), then for every row of the item loads
//                  4 coefficients and 4 row values from shared memory (LDS.128, conflict free) for 16 FP64 FMAs.
//   * epilogue   : shuffle-reduce over the 4 entry-groups, fixed-order sum over the warps in shared memory.
// Static work assignment, fixed summation order: results are bit-reproducible run to run.
// Roofline: x is streamed once (8 * d_in bytes per point) against `padded_fma` FP64 FMAs per point.
#include <algorithm>
#include <cstddef>

#include "smx_fast_common.cuh"

namespace smx {
namespace {

// position of tile point t (= 4 * group + pp) inside a 32-double row of the value table: the two halves of a lane's
// four points are 128 bytes apart so that one LDS.128 of the 8 point-groups reads 128 contiguous bytes
__device__ __forceinline__ int m_slot(int t) { return ((t >> 1) & 1) * 16 + (t >> 2) * 2 + (t & 1); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Issue the asynchronous copies of work item c (directory entry `dir`) into item buffer `buf` and the x buffer.
__device__ __forceinline__ void stage_item(const FastArgs& a, const double* __restrict__ x, WarpStage& st, int buf, int c,
                                           const int4 dir, long long o, long long p0, int lane) {
    const int r0 = dir.x, rows = dir.y, flags = dir.z & 15;
    // metadata record (24 pieces of 16 bytes), then the coefficient rows (8 pieces each)
    if (lane < kMetaInts / 4)
        cp_async16(reinterpret_cast<int*>(&st.item[buf]) + 4 * lane, a.chunk_meta + (size_t)c * kMetaInts + 4 * lane);
    for (int id = lane; id < rows * 8; id += 32) {
        const int r = id >> 3, cc = id & 7;
        cp_async16(&st.item[buf].coef[r * kBlockWidth + 2 * cc], a.coef + ((size_t)r0 * a.d_out + (size_t)o * rows + r) * kBlockWidth + 2 * cc);
    }
    if (!(flags & kChunkHot)) {
        // x tile: row `row` of the tile at 128-byte pitch; its 16-byte pieces are XOR-swizzled with bit 2 of the row so
        // that the readers (lane = 4 * group + entry-group, row = 4 * group + pp) do not collide on banks
        if ((flags & kChunkInside) && a.x_vec_ok) {
            const int dim0 = dir.w;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int id = it * 32 + lane, row = id >> 3, cc = id & 7;
                const long long p = min(p0 + row, a.N - 1);
                cp_async16(&st.xs[row * kBlockWidth + ((cc ^ ((row >> 2) & 1)) << 1)], x + p * a.ldx + dim0 + 2 * cc);
            }
        } else {
            const int i = lane & 15;
            const int dim = __ldg(a.ent_dim + (dir.z >> 4) * kBlockWidth + i);
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const int row = it * 2 + (lane >> 4);
                const long long p = min(p0 + row, a.N - 1);
                cp_async8(&st.xs[row * kBlockWidth + ((((i >> 1) ^ ((row >> 2) & 1)) << 1) | (i & 1))], x + p * a.ldx + dim);
            }
        }
    }
    cp_async_commit();
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, NW == 4 ? 2 : 1)
fast_eval_kernel(const FastArgs a, const double* __restrict__ x, double* __restrict__ y) {
    constexpr int kThreads = NW * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tab = reinterpret_cast<double*>(smem_raw);                       // [n_tab][32] value table
    double* ypart = tab + (size_t)a.n_tab * kTile;                            // [NW][32]
    WarpStage* stages = reinterpret_cast<WarpStage*>(ypart + NW * kTile);   // [NW]
    int4* s_dir = reinterpret_cast<int4*>(stages + NW);                       // [n_chunks]
    double* s_eta = reinterpret_cast<double*>(s_dir + a.n_chunks);            // [n_hot] centres of the hot dimensions
    int2* s_pairs = reinterpret_cast<int2*>(s_eta + a.n_hot);                 // [n_pairs]
    int* s_hot_off = reinterpret_cast<int*>(s_pairs + a.n_pairs);             // [hot_dims + 1]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = lane & 3, g = lane >> 2;
    WarpStage& st = stages[warp];

    // ---- once per CTA: the small tables every tile needs move to shared memory ----------------------------------------
    for (int i = tid; i < a.n_chunks; i += kThreads) s_dir[i] = __ldg(a.chunk_dir + i);
    for (int i = tid; i < a.n_hot; i += kThreads) s_eta[i] = __ldg(a.eta + i);
    for (int i = tid; i < a.n_pairs; i += kThreads) s_pairs[i] = __ldg(a.tab_pairs + i);
    for (int i = tid; i <= a.hot_dims; i += kThreads) s_hot_off[i] = __ldg(a.hot_off + i);
    __syncthreads();

    for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const long long p0 = tile * kTile;

        // ---- prologue: value table = 1 | 1-D basis values of the hot entries | products of hot pairs, level by level ----
        // lane = point, warp w takes the hot dimensions w, w + NW, ..; four coordinates are in flight per thread
        if (tid < kTile) tab[tid] = 1.0;
        {
            const double* xrow = x + min(p0 + lane, a.N - 1) * a.ldx;
            const int slot = m_slot(lane);
            for (int d0 = warp; d0 < a.hot_dims; d0 += 4 * NW) {
                double xv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) xv[u] = (d0 + u * NW < a.hot_dims) ? __ldg(xrow + d0 + u * NW) : 0.0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int d = d0 + u * NW;
                    if (d < a.hot_dims) {
                        const int off0 = s_hot_off[d], off1 = s_hot_off[d + 1];
                        double v = 1.0;
                        for (int k = off0; k < off1; ++k) {
                            v *= (xv[u] - s_eta[k]);
                            tab[(1 + k) * kTile + slot] = v;
                        }
                    }
                }
            }
        }
        __syncthreads();
        for (int l = 2; l < a.n_levels; ++l) {
            const int t_begin = a.level_off[l], count = (a.level_off[l + 1] - t_begin) * kTile;
            for (int idx = tid; idx < count; idx += kThreads) {
                const int ti = t_begin + (idx >> 5), s = idx & 31;
                const int2 pr = s_pairs[ti - 1 - a.n_hot];
                tab[ti * kTile + s] = tab[pr.x * kTile + s] * tab[pr.y * kTile + s];
            }
            __syncthreads();
        }

        // ---- main: block-sparse contraction, one output at a time -------------------------------------------------------
        for (long long o = 0; o < a.d_out; ++o) {
            double tot[4] = {0.0, 0.0, 0.0, 0.0};
            if (warp < a.n_chunks) stage_item(a, x, st, 0, warp, s_dir[warp], o, p0, lane);
            int it = 0;
            for (int c = warp; c < a.n_chunks; c += NW, ++it) {
                const int buf = it & 1;
                const int4 dir = s_dir[c];
                const int rows = dir.y, flags = dir.z & 15;
                const ItemBuffer& ib = st.item[buf];
                cp_async_wait<0>();
                __syncwarp();  // the copies of all lanes for this item have landed

                double v[4][4];  // [point][entry] leading basis values pi_e(x_p)
                if (flags & kChunkHot) {
                    const int4 t4 = *reinterpret_cast<const int4*>(ib.tab + 4 * q);
                    const int tabs[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double* tr = tab + tabs[i] * kTile + 2 * g;
                        const double2 lo = *reinterpret_cast<const double2*>(tr);
                        const double2 hi = *reinterpret_cast<const double2*>(tr + 16);
                        v[0][i] = lo.x, v[1][i] = lo.y, v[2][i] = hi.x, v[3][i] = hi.y;
                    }
                } else {
                    const int sw = g & 1;  // (row >> 2) & 1 with row = 4 g + pp
#pragma unroll
                    for (int pp = 0; pp < 4; ++pp) {
                        const double* xr = st.xs + (4 * g + pp) * kBlockWidth;
                        const double2 lo = *reinterpret_cast<const double2*>(xr + (((2 * q) ^ sw) << 1));
                        const double2 hi = *reinterpret_cast<const double2*>(xr + (((2 * q + 1) ^ sw) << 1));
                        v[pp][0] = lo.x, v[pp][1] = lo.y, v[pp][2] = hi.x, v[pp][3] = hi.y;
                    }
                    const double2 ea = *reinterpret_cast<const double2*>(ib.eta0 + 4 * q);
                    const double2 eb = *reinterpret_cast<const double2*>(ib.eta0 + 4 * q + 2);
                    const double e4[4] = {ea.x, ea.y, eb.x, eb.y};
                    if (flags & kChunkContig) {  // 16 degree-1 entries (or zero-coefficient dummies): pi = x - eta_0
#pragma unroll
                        for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[pp][i] -= e4[i];
                    } else {
                        const int4 deg4 = *reinterpret_cast<const int4*>(ib.deg + 4 * q);
                        const int4 eta4 = *reinterpret_cast<const int4*>(ib.etaoff + 4 * q);
                        const int degs[4] = {deg4.x, deg4.y, deg4.z, deg4.w};
                        const int etas[4] = {eta4.x, eta4.y, eta4.z, eta4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (degs[i] == 0) {
#pragma unroll
                                for (int pp = 0; pp < 4; ++pp) v[pp][i] = 1.0;  // dummy entry (coefficients are zero)
                            } else {
                                double xs4[4];
#pragma unroll
                                for (int pp = 0; pp < 4; ++pp) {
                                    xs4[pp] = v[pp][i];
                                    v[pp][i] = xs4[pp] - e4[i];
                                }
                                for (int k = 1; k < degs[i]; ++k) {
                                    const double ek = __ldg(a.eta + etas[i] + k);
#pragma unroll
                                    for (int pp = 0; pp < 4; ++pp) v[pp][i] *= (xs4[pp] - ek);
                                }
                            }
                        }
                    }
                }
                __syncwarp();  // every lane has taken its x values: the x buffer and the other item buffer are free
                if (c + NW < a.n_chunks) stage_item(a, x, st, buf ^ 1, c + NW, s_dir[c + NW], o, p0, lane);

                double acc[4][4];
#pragma unroll
                for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[pp][i] = 0.0;

                const double* cf = ib.coef + 4 * q;
#pragma unroll 2
                for (int r = 0; r < rows; ++r) {
                    const double2 c01 = *reinterpret_cast<const double2*>(cf + r * kBlockWidth);
                    const double2 c23 = *reinterpret_cast<const double2*>(cf + r * kBlockWidth + 2);
                    const double* mr = tab + ib.ridx[r] * kTile + 2 * g;
                    const double2 m01 = *reinterpret_cast<const double2*>(mr);
                    const double2 m23 = *reinterpret_cast<const double2*>(mr + 16);
                    const double cs4[4] = {c01.x, c01.y, c23.x, c23.y};
                    const double ms4[4] = {m01.x, m01.y, m23.x, m23.y};
#pragma unroll
                    for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[pp][i] = fma(cs4[i], ms4[pp], acc[pp][i]);
                }
#pragma unroll
                for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                    for (int i = 0; i < 4; ++i) tot[pp] = fma(v[pp][i], acc[pp][i], tot[pp]);
            }
            // ---- epilogue: reduce over the 4 entry-groups (lanes), then over the warps in fixed order -------------------
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                tot[pp] += __shfl_xor_sync(0xffffffffu, tot[pp], 1);
                tot[pp] += __shfl_xor_sync(0xffffffffu, tot[pp], 2);
            }
            if (q == 0) {
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) ypart[warp * kTile + 4 * g + pp] = tot[pp];
            }
            __syncthreads();
            if (tid < kTile && p0 + tid < a.N) {
                double s = __ldg(a.c0 + o);
#pragma unroll
                for (int w = 0; w < NW; ++w) s += ypart[w * kTile + tid];
                y[(p0 + tid) * a.d_out + o] = s;
            }
            __syncthreads();
        }
    }
}

template <class T>
int upload(const std::vector<T>& v, T** dptr, int64_t& bytes, size_t min_elems = 1) {
    const size_t n = std::max(v.size(), min_elems);
    SMX_CUDA(cudaMalloc((void**)dptr, n * sizeof(T)));
    SMX_CUDA(cudaMemset(*dptr, 0, n * sizeof(T)));
    if (!v.empty()) SMX_CUDA(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    bytes += (int64_t)(n * sizeof(T));
    return SMX_OK;
}

size_t fast_smem_bytes(const FastDevice& d, int nw) {
    return ((size_t)d.n_tab * kTile + (size_t)nw * kTile + (size_t)d.n_hot) * sizeof(double) + (size_t)nw * sizeof(WarpStage) +
           (size_t)d.n_chunks * sizeof(int4) + (size_t)d.n_pairs * sizeof(int2) + ((size_t)d.hot_dims + 1) * sizeof(int) + 16;
}

}  // namespace

int fast_upload(const FastPlan& plan, FastDevice& dev) {
    dev = FastDevice();
    dev.d_in = plan.d_in;
    dev.d_out = plan.d_out;
    dev.n_tab = plan.n_tab;
    dev.n_hot = plan.n_hot;
    dev.n_levels = plan.n_levels;
    dev.n_chunks = plan.n_chunks;
    dev.hot_dims = plan.hot_dims;
    dev.n_pairs = (int32_t)plan.tab_parent.size();
    if (plan.n_levels > kMaxLevels) return fail(SMX_ERR_UNSUPPORTED, "too many active dimensions per term");
    for (size_t l = 0; l < plan.level_off.size() && l < (size_t)kMaxLevels + 2; ++l) dev.level_off[l] = plan.level_off[l];
    int device = 0;
    SMX_CUDA(cudaGetDevice(&device));
    SMX_CUDA(cudaDeviceGetAttribute(&dev.sm_count, cudaDevAttrMultiProcessorCount, device));
    int smem_optin = 0, smem_sm = 0;
    SMX_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    SMX_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
    // two CTAs of 4 warps per SM if the value table is small enough, else one CTA of 8 warps
    if (2 * (fast_smem_bytes(dev, 4) + 1024) <= (size_t)smem_sm) {
        dev.warps = 4;
    } else if (fast_smem_bytes(dev, 8) <= (size_t)smem_optin) {
        dev.warps = 8;
    } else {
        return fail(SMX_ERR_UNSUPPORTED, "value table does not fit in shared memory");
    }
    std::vector<int32_t> dir(plan.chunk_dir);
    for (int32_t c = 0; c < plan.n_chunks; ++c) dir[(size_t)c * 4 + 2] = plan.chunk_flags[c] | (plan.chunk_block[c] << 4);
    std::vector<int32_t> pairs(plan.tab_parent.size() * 2);
    for (size_t i = 0; i < plan.tab_parent.size(); ++i) pairs[2 * i] = plan.tab_parent[i], pairs[2 * i + 1] = plan.tab_hot[i];
    int rc;
    if ((rc = upload(plan.ent_dim, &dev.ent_dim, dev.bytes, 16))) return rc;
    if ((rc = upload(plan.eta, &dev.eta, dev.bytes))) return rc;
    if ((rc = upload(pairs, &dev.tab_pairs, dev.bytes, 2))) return rc;
    if ((rc = upload(plan.hot_off, &dev.hot_off, dev.bytes))) return rc;
    if ((rc = upload(dir, &dev.chunk_dir, dev.bytes, 4))) return rc;
    if ((rc = upload(plan.chunk_meta, &dev.chunk_meta, dev.bytes, 4))) return rc;
    {   // device layout of the coefficients: per work item [output][row][16], so that one (item, output) is contiguous
        std::vector<double> packed(plan.coef.size());
        for (int32_t c = 0; c < plan.n_chunks; ++c) {
            const size_t r0 = plan.chunk_off[c], rows = plan.chunk_off[c + 1] - r0, dout = (size_t)plan.d_out;
            for (size_t r = 0; r < rows; ++r)
                for (size_t o = 0; o < dout; ++o)
                    std::copy_n(&plan.coef[((r0 + r) * dout + o) * kBlockWidth], kBlockWidth,
                                &packed[(r0 * dout + o * rows + r) * kBlockWidth]);
        }
        if ((rc = upload(packed, &dev.coef, dev.bytes))) return rc;
    }
    if ((rc = upload(plan.c0, &dev.c0, dev.bytes))) return rc;
    const size_t smem = fast_smem_bytes(dev, dev.warps);
    if (dev.warps == 4)
        SMX_CUDA(cudaFuncSetAttribute(fast_eval_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
        SMX_CUDA(cudaFuncSetAttribute(fast_eval_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return fast_tma_prepare(dev);
}

void fast_free(FastDevice& d) {
    void* ptrs[] = {d.ent_dim, d.eta, d.tab_pairs, d.hot_off, d.chunk_dir, d.chunk_meta, d.coef, d.c0};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    d = FastDevice();
}

int fast_eval(const FastDevice& d, const double* x, int64_t N, int64_t ldx, double* y, cudaStream_t st) {
    if (N == 0) return SMX_OK;
    const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ldx & 1) == 0;
    if (aligned && d.tma_warps > 0) return fast_eval_tma(d, x, N, ldx, y, st);
    FastArgs a;
    fill_fast_args(d, x, N, ldx, a);
    const size_t smem = fast_smem_bytes(d, d.warps);
    const int per_sm = d.warps == 4 ? 2 : 1;
    const long long grid = std::min<long long>(a.num_tiles, (long long)d.sm_count * per_sm);
    if (d.warps == 4)
        fast_eval_kernel<4><<<(unsigned)grid, 128, smem, st>>>(a, x, y);
    else
        fast_eval_kernel<8><<<(unsigned)grid, 256, smem, st>>>(a, x, y);
    SMX_LAUNCH_CHECK("fast_eval_kernel");
    return SMX_OK;
}

void fill_fast_args(const FastDevice& d, const double* x, int64_t N, int64_t ldx, FastArgs& a) {
    a.ent_dim = d.ent_dim;
    a.eta = d.eta;
    a.tab_pairs = reinterpret_cast<const int2*>(d.tab_pairs);
    a.hot_off = d.hot_off;
    a.chunk_dir = reinterpret_cast<const int4*>(d.chunk_dir);
    a.chunk_meta = d.chunk_meta;
    a.coef = d.coef;
    a.c0 = d.c0;
    a.N = N;
    a.ldx = ldx;
    a.d_out = d.d_out;
    a.num_tiles = (N + kTile - 1) / kTile;
    a.n_hot = d.n_hot;
    a.n_tab = d.n_tab;
    a.n_chunks = d.n_chunks;
    a.n_levels = d.n_levels;
    a.hot_dims = d.hot_dims;
    a.n_pairs = d.n_pairs;
    a.x_vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ldx & 1) == 0) ? 1 : 0;
    for (int l = 0; l < kMaxLevels + 2; ++l) a.level_off[l] = d.level_off[l];
}

}  // namespace smx
