// Fast evaluation path (K1): fused 1-D basis + block-sparse Kronecker contraction on the hierarchical layout of
// smx_plan.h.  Replaces the whole of reference interpolation.py:281-302 (python loop over groups and summand
// batches around jit(vmap(barycentric.evaluate_tensor_product_interpolant))) by ONE persistent kernel.
//
//   I(x_p) = c_0 + sum_{e=(j,a)} pi_e(x_pj) * sum_r C[r][e] * m_r(x_p)
//
// Mapping (one CTA = one tile of 32 evaluation points at a time, 8 warps):
//   * prologue   : the CTA computes the hot 1-D basis values pi_h(x_p) and, level by level, the row products
//                  m_r(x_p) = m_parent(r) * pi_h(r) into shared memory ([row][point], 256 B per row).
//   * main       : warp w takes the work items (entry block of 16 leading entries x <= 24 rows) w, w+8, ..
//                  A lane owns 4 entries x 4 points (lane = 4 * point-group + entry-group): it reads its
//                  16 coordinates of x straight from HBM (each coordinate of x is read once; consecutive entries are
//                  consecutive columns, so the 4 lanes of a point-group cover one 128-byte line), builds the 16
//                  leading basis values pi_e(x) in registers, then for every row of the item loads 4
//                  coefficients (L2, shared by the 8 point-groups) and 4 row products (shared memory, shared by
//                  the 4 entry-groups) for 16 FP64 FMAs.
//   * epilogue   : shuffle-reduce over the 4 entry-groups, fixed-order sum over the 8 warps in shared memory.
// Static work assignment, fixed summation order: results are bit-reproducible run to run.
// Roofline: x is streamed once (8*d_in bytes per point) against `padded_fma` FP64 FMAs per point.
#include <algorithm>

#include "smx_common.cuh"

namespace smx {
namespace {

constexpr int kTile = 32;      // points per tile
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
static_assert(kBlockWidth == 16, "lane mapping below assumes 16 entries per block");

struct FastArgs {
    const int32_t* ent_dim;
    const int32_t* ent_deg;
    const int32_t* ent_eta;
    const double* eta;
    const int32_t* row_parent;
    const int32_t* row_hslot;
    const int32_t* hot_dim;
    const int32_t* hot_deg;
    const int32_t* hot_eta;
    const int32_t* chunk_block;
    const int32_t* chunk_off;
    const int32_t* chunk_rows;
    const double* coef;
    const double* c0;
    long long N, ldx, d_out, num_tiles;
    int n_rows, n_hot, n_chunks, n_levels;
    int level_off[kMaxLevels + 2];
};

// position of tile point t (= 4 * group + pp) inside a 32-double row of the m table: the two halves of a lane's four
// points are 128 bytes apart so that one LDS.128 of the 8 point-groups reads 128 contiguous bytes (no bank conflict)
__device__ __forceinline__ int m_slot(int t) { return ((t >> 1) & 1) * 16 + (t >> 2) * 2 + (t & 1); }

__global__ void __launch_bounds__(kThreads, 2)
fast_eval_kernel(const FastArgs a, const double* __restrict__ x, double* __restrict__ y) {
    extern __shared__ __align__(16) double smem[];
    double* m_tab = smem;                                   // [n_rows][32]
    double* pih = m_tab + (size_t)a.n_rows * kTile;         // [n_hot][32]
    double* ypart = pih + (size_t)a.n_hot * kTile;          // [kWarps][32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = lane & 3, g = lane >> 2;

    for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const long long p0 = tile * kTile;

        // ---- prologue: hot basis values and row products ---------------------------------------------------
        for (int idx = tid; idx < a.n_hot * kTile; idx += kThreads) {
            const int h = idx >> 5, t = idx & 31;
            const long long p = min(p0 + t, a.N - 1);
            const double xv = __ldg(x + p * a.ldx + a.hot_dim[h]);
            const double* eta = a.eta + a.hot_eta[h];
            double v = xv - __ldg(eta);
            for (int k = 1; k < a.hot_deg[h]; ++k) v *= (xv - __ldg(eta + k));
            pih[idx] = v;
        }
        if (tid < kTile) m_tab[tid] = 1.0;
        __syncthreads();
        for (int l = 1; l < a.n_levels; ++l) {
            const int r_begin = a.level_off[l], count = (a.level_off[l + 1] - r_begin) * kTile;
            for (int idx = tid; idx < count; idx += kThreads) {
                const int r = r_begin + (idx >> 5), t = idx & 31, s = m_slot(t);
                m_tab[r * kTile + s] = m_tab[__ldg(a.row_parent + r) * kTile + s] * pih[__ldg(a.row_hslot + r) * kTile + t];
            }
            __syncthreads();
        }

        // ---- main: block-sparse contraction, one output at a time ----------------------------------------------
        for (long long o = 0; o < a.d_out; ++o) {
            double tot[4] = {0.0, 0.0, 0.0, 0.0};
            for (int c = warp; c < a.n_chunks; c += kWarps) {
                const int e0 = __ldg(a.chunk_block + c) * kBlockWidth + 4 * q;
                const int4 dim4 = __ldg(reinterpret_cast<const int4*>(a.ent_dim + e0));
                const int4 deg4 = __ldg(reinterpret_cast<const int4*>(a.ent_deg + e0));
                const int4 eta4 = __ldg(reinterpret_cast<const int4*>(a.ent_eta + e0));
                const int dims[4] = {dim4.x, dim4.y, dim4.z, dim4.w};
                const int degs[4] = {deg4.x, deg4.y, deg4.z, deg4.w};
                const int etas[4] = {eta4.x, eta4.y, eta4.z, eta4.w};

                double v[4][4];  // [point][entry] leading basis values
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    const long long p = min(p0 + 4 * g + pp, a.N - 1);
                    const double* xr = x + p * a.ldx;
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[pp][i] = __ldg(xr + dims[i]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double* eta = a.eta + etas[i];
                    if (degs[i] == 0) {
#pragma unroll
                        for (int pp = 0; pp < 4; ++pp) v[pp][i] = 1.0;  // padding lane (coefficients are zero)
                    } else {
                        const double e_first = __ldg(eta);
                        double xs[4];
#pragma unroll
                        for (int pp = 0; pp < 4; ++pp) {
                            xs[pp] = v[pp][i];
                            v[pp][i] = xs[pp] - e_first;
                        }
                        for (int k = 1; k < degs[i]; ++k) {
                            const double ek = __ldg(eta + k);
#pragma unroll
                            for (int pp = 0; pp < 4; ++pp) v[pp][i] *= (xs[pp] - ek);
                        }
                    }
                }

                double acc[4][4];
#pragma unroll
                for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[pp][i] = 0.0;

                const int r0 = __ldg(a.chunk_off + c), r1 = __ldg(a.chunk_off + c + 1);
                const int my_row = (r0 + lane < r1) ? __ldg(a.chunk_rows + r0 + lane) : 0;  // kChunkRows <= 32
                const double* cf = a.coef + ((size_t)r0 * a.d_out + o) * kBlockWidth + 4 * q;
                for (int r = r0; r < r1; ++r, cf += (size_t)a.d_out * kBlockWidth) {
                    const int row = __shfl_sync(0xffffffffu, my_row, r - r0);
                    const double2 c01 = __ldg(reinterpret_cast<const double2*>(cf));
                    const double2 c23 = __ldg(reinterpret_cast<const double2*>(cf + 2));
                    const double* mr = m_tab + row * kTile + 2 * g;
                    const double2 m01 = *reinterpret_cast<const double2*>(mr);
                    const double2 m23 = *reinterpret_cast<const double2*>(mr + 16);
                    const double cs[4] = {c01.x, c01.y, c23.x, c23.y};
                    const double ms[4] = {m01.x, m01.y, m23.x, m23.y};
#pragma unroll
                    for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[pp][i] = fma(cs[i], ms[pp], acc[pp][i]);
                }
#pragma unroll
                for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                    for (int i = 0; i < 4; ++i) tot[pp] = fma(v[pp][i], acc[pp][i], tot[pp]);
            }
            // ---- epilogue: reduce over the 4 entry-groups (lanes), then over the warps in fixed order ----------
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
                for (int w = 0; w < kWarps; ++w) s += ypart[w * kTile + tid];
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

size_t fast_smem_bytes(const FastDevice& d) {
    return ((size_t)d.n_rows * kTile + (size_t)d.n_hot * kTile + (size_t)kWarps * kTile) * sizeof(double);
}

}  // namespace

int fast_upload(const FastPlan& plan, FastDevice& dev) {
    dev = FastDevice();
    dev.d_in = plan.d_in;
    dev.d_out = plan.d_out;
    dev.n_entries_padded = (int32_t)plan.ent_dim.size();
    dev.n_rows = plan.n_rows;
    dev.n_levels = plan.n_levels;
    dev.n_hot = plan.n_hot;
    dev.n_chunks = plan.n_chunks;
    if (plan.n_levels > kMaxLevels) return fail(SMX_ERR_UNSUPPORTED, "too many active dimensions per term");
    for (int l = 0; l <= plan.n_levels; ++l) dev.level_off[l] = plan.level_off[l];
    int rc;
    // +16 elements of slack on the entry tables: the kernel reads int4 at 4-aligned offsets inside padded blocks
    if ((rc = upload(plan.ent_dim, &dev.ent_dim, dev.bytes, 16))) return rc;
    if ((rc = upload(plan.ent_deg, &dev.ent_deg, dev.bytes, 16))) return rc;
    if ((rc = upload(plan.ent_eta, &dev.ent_eta, dev.bytes, 16))) return rc;
    if ((rc = upload(plan.eta, &dev.eta, dev.bytes))) return rc;
    if ((rc = upload(plan.row_parent, &dev.row_parent, dev.bytes))) return rc;
    if ((rc = upload(plan.row_hslot, &dev.row_hslot, dev.bytes))) return rc;
    if ((rc = upload(plan.hot_dim, &dev.hot_dim, dev.bytes))) return rc;
    if ((rc = upload(plan.hot_deg, &dev.hot_deg, dev.bytes))) return rc;
    if ((rc = upload(plan.hot_eta, &dev.hot_eta, dev.bytes))) return rc;
    if ((rc = upload(plan.chunk_block, &dev.chunk_block, dev.bytes))) return rc;
    if ((rc = upload(plan.chunk_off, &dev.chunk_off, dev.bytes, 2))) return rc;
    if ((rc = upload(plan.chunk_rows, &dev.chunk_rows, dev.bytes))) return rc;
    if ((rc = upload(plan.coef, &dev.coef, dev.bytes))) return rc;
    if ((rc = upload(plan.c0, &dev.c0, dev.bytes))) return rc;
    for (int32_t d : plan.ent_deg) dev.max_ent_deg = std::max(dev.max_ent_deg, (int)d);
    int device = 0;
    SMX_CUDA(cudaGetDevice(&device));
    SMX_CUDA(cudaDeviceGetAttribute(&dev.sm_count, cudaDevAttrMultiProcessorCount, device));
    const size_t smem = fast_smem_bytes(dev);
    if (smem > 227 * 1024) return fail(SMX_ERR_UNSUPPORTED, "row-product table does not fit in shared memory");
    SMX_CUDA(cudaFuncSetAttribute(fast_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return SMX_OK;
}

void fast_free(FastDevice& d) {
    void* ptrs[] = {d.ent_dim, d.ent_deg, d.ent_eta, d.eta, d.row_parent, d.row_hslot, d.hot_dim, d.hot_deg,
                    d.hot_eta, d.chunk_block, d.chunk_off, d.chunk_rows, d.coef, d.c0};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    d = FastDevice();
}

int fast_eval(const FastDevice& d, const double* x, int64_t N, int64_t ldx, double* y, cudaStream_t st) {
    if (N == 0) return SMX_OK;
    FastArgs a;
    a.ent_dim = d.ent_dim;
    a.ent_deg = d.ent_deg;
    a.ent_eta = d.ent_eta;
    a.eta = d.eta;
    a.row_parent = d.row_parent;
    a.row_hslot = d.row_hslot;
    a.hot_dim = d.hot_dim;
    a.hot_deg = d.hot_deg;
    a.hot_eta = d.hot_eta;
    a.chunk_block = d.chunk_block;
    a.chunk_off = d.chunk_off;
    a.chunk_rows = d.chunk_rows;
    a.coef = d.coef;
    a.c0 = d.c0;
    a.N = N;
    a.ldx = ldx;
    a.d_out = d.d_out;
    a.num_tiles = (N + kTile - 1) / kTile;
    a.n_rows = d.n_rows;
    a.n_hot = d.n_hot;
    a.n_chunks = d.n_chunks;
    a.n_levels = d.n_levels;
    for (int l = 0; l < kMaxLevels + 2; ++l) a.level_off[l] = d.level_off[l];
    const size_t smem = fast_smem_bytes(d);
    int per_sm = 0;
    SMX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fast_eval_kernel, kThreads, smem));
    if (per_sm < 1) return fail(SMX_ERR_UNSUPPORTED, "fast kernel does not fit on an SM");
    const long long grid = std::min<long long>(a.num_tiles, (long long)d.sm_count * per_sm);
    fast_eval_kernel<<<(unsigned)grid, kThreads, smem, st>>>(a, x, y);
    SMX_LAUNCH_CHECK("fast_eval_kernel");
    return SMX_OK;
}

}  // namespace smx
