// Fast evaluation path, TMA variant (the one that runs whenever the rows of x are 16-byte aligned).
//
// Same mathematics, plan and work decomposition as smx_fast.cu; what changes is how a work item reaches shared memory:
//   * the 32-point x 16-column tile of x of a cold block is ONE `cp.async.bulk.tensor.2d` (TMA, 128-byte swizzle, rows
//     beyond N and columns beyond d_in are zero-filled by the hardware),
//   * the item's metadata record and its coefficient rows are two `cp.async.bulk` copies,
// all issued by one lane per warp and completed on an mbarrier (expect_tx), one item ahead of its use.  That removes
// the per-copy address arithmetic of the cp.async version (which issued more instructions for staging than for the
// FP64 work) and leaves the warp with: wait, 8 LDS.128 for its 4 x 4 basis values, and 4 LDS.128 + 16 DFMA per row.
//
// Lane mapping: lane = 4 * g + q; entries 4q..4q+3 of the block; points g, g+8, g+16, g+24 of the tile (so that the
// two rows a quarter-warp reads with one LDS.128 differ in bit 0 of the swizzle key and never collide on banks).
#include <cuda.h>

#include <cstdlib>

#include "smx_fast_common.cuh"

namespace smx {
namespace {

constexpr int kXTileBytes = kTile * kBlockWidth * 8;  // 4096

struct alignas(1024) TmaStage {
    double xs[kTile * kBlockWidth];  // TMA destination, 128-byte swizzle => 1024-byte alignment
    ItemBuffer item[2];
    unsigned long long bar[2];       // mbarriers: item buffer b (and the x tile that travels with that item)
};

// tile point t = g + 8 * pp  ->  position inside a 32-double row of the value table (see smx_fast.cu, m_slot)
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

// One lane of the warp issues the copies of work item c into item buffer `buf` (and, for cold blocks, the x tile).
__device__ __forceinline__ void tma_stage_item(const FastArgs& a, const CUtensorMap* xmap, TmaStage& st, int buf, int c, const int4 dir,
                                               long long o, long long p0) {
    const int r0 = dir.x, rows = dir.y, flags = dir.z & 15;
    const unsigned coef_bytes = (unsigned)rows * kBlockWidth * 8;
    const bool cold = !(flags & kChunkHot);
    mbar_expect_tx(&st.bar[buf], kMetaInts * 4 + coef_bytes + (cold ? kXTileBytes : 0));
    bulk_copy(&st.item[buf], a.chunk_meta + (size_t)c * kMetaInts, kMetaInts * 4, &st.bar[buf]);
    if (rows > 0)
        bulk_copy(st.item[buf].coef, a.coef + ((size_t)r0 * a.d_out + (size_t)o * rows) * kBlockWidth, coef_bytes, &st.bar[buf]);
    if (cold) tma_load_2d(st.xs, xmap, dir.w, (int)p0, &st.bar[buf]);
}

template <int NW, int CTAS>
__global__ void __launch_bounds__(NW * 32, CTAS)
fast_eval_tma_kernel(const __grid_constant__ CUtensorMap xmap, const FastArgs a, const double* __restrict__ x, double* __restrict__ y) {
    constexpr int kThreads = NW * 32;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte aligned carve-up (the launch adds 1 KiB of slack)
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    TmaStage* stages = reinterpret_cast<TmaStage*>(base);                        // [NW]
    double* tab = reinterpret_cast<double*>(stages + NW);                         // [n_tab][32] value table
    double* ypart = tab + (size_t)a.n_tab * kTile;                                // [NW][32]
    int4* s_dir = reinterpret_cast<int4*>(ypart + NW * kTile);                    // [n_chunks]
    double* s_eta = reinterpret_cast<double*>(s_dir + a.n_chunks);                // [n_hot]
    int2* s_pairs = reinterpret_cast<int2*>(s_eta + a.n_hot);                     // [n_pairs]
    int* s_hot_off = reinterpret_cast<int*>(s_pairs + a.n_pairs);                 // [hot_dims + 1]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = lane & 3, g = lane >> 2;
    TmaStage& st = stages[warp];

    // ---- once per CTA: small tables to shared memory, mbarriers --------------------------------------------------------
    for (int i = tid; i < a.n_chunks; i += kThreads) s_dir[i] = __ldg(a.chunk_dir + i);
    for (int i = tid; i < a.n_hot; i += kThreads) s_eta[i] = __ldg(a.eta + i);
    for (int i = tid; i < a.n_pairs; i += kThreads) s_pairs[i] = __ldg(a.tab_pairs + i);
    for (int i = tid; i <= a.hot_dims; i += kThreads) s_hot_off[i] = __ldg(a.hot_off + i);
    if (lane == 0) {
        mbar_init(&st.bar[0], 1);
        mbar_init(&st.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    unsigned k_item = 0;  // items this warp has staged/consumed so far: buffer = k & 1, phase parity = (k >> 1) & 1

    for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const long long p0 = tile * kTile;

        // ---- prologue: value table = 1 | 1-D basis values of the hot entries | products of hot pairs, level by level ----
        if (tid < kTile) tab[tid] = 1.0;
        {
            const double* xrow = x + min(p0 + lane, a.N - 1) * a.ldx;
            const int slot = t_slot(lane);
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
            if (warp < a.n_chunks && lane == 0) tma_stage_item(a, &xmap, st, k_item & 1, warp, s_dir[warp], o, p0);
            for (int c = warp; c < a.n_chunks; c += NW, ++k_item) {
                const int buf = k_item & 1;
                const int4 dir = s_dir[c];
                const int rows = dir.y, flags = dir.z & 15;
                const ItemBuffer& ib = st.item[buf];
                mbar_wait(&st.bar[buf], (k_item >> 1) & 1);

                double v[4][4];  // [point][entry] leading basis values pi_e(x_p); point = g + 8 * pp
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
                    // 128-byte swizzle: 16-byte piece j of row r sits at piece j ^ (r & 7); r = g + 8 pp => r & 7 = g
                    const double2 ea = *reinterpret_cast<const double2*>(ib.eta0 + 4 * q);
                    const double2 eb = *reinterpret_cast<const double2*>(ib.eta0 + 4 * q + 2);
#pragma unroll
                    for (int pp = 0; pp < 4; ++pp) {
                        const double* xr = st.xs + (g + 8 * pp) * kBlockWidth;
                        const double2 lo = *reinterpret_cast<const double2*>(xr + (((2 * q) ^ g) << 1));
                        const double2 hi = *reinterpret_cast<const double2*>(xr + (((2 * q + 1) ^ g) << 1));
                        v[pp][0] = lo.x - ea.x, v[pp][1] = lo.y - ea.y, v[pp][2] = hi.x - eb.x, v[pp][3] = hi.y - eb.y;
                    }
                }
                __syncwarp();  // every lane has taken its x values: the x buffer and the other item buffer are free
                if (c + NW < a.n_chunks && lane == 0) tma_stage_item(a, &xmap, st, buf ^ 1, c + NW, s_dir[c + NW], o, p0);

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
                for (int pp = 0; pp < 4; ++pp) ypart[warp * kTile + g + 8 * pp] = tot[pp];
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

size_t tma_smem_bytes(const FastDevice& d, int nw) {
    return 1024 + (size_t)nw * sizeof(TmaStage) + ((size_t)d.n_tab * kTile + (size_t)nw * kTile + (size_t)d.n_hot) * sizeof(double) +
           (size_t)d.n_chunks * sizeof(int4) + (size_t)d.n_pairs * sizeof(int2) + ((size_t)d.hot_dims + 1) * sizeof(int) + 16;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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

template <int NW, int CTAS>
int launch(const CUtensorMap& map, const FastArgs& a, const FastDevice& d, const double* x, double* y, cudaStream_t st) {
    const size_t smem = tma_smem_bytes(d, NW);
    const long long grid = std::min<long long>(a.num_tiles, (long long)d.sm_count * CTAS);
    fast_eval_tma_kernel<NW, CTAS><<<(unsigned)grid, NW * 32, smem, st>>>(map, a, x, y);
    SMX_LAUNCH_CHECK("fast_eval_tma_kernel");
    return SMX_OK;
}

}  // namespace

// Chooses the CTA shape of the TMA kernel for this plan (0 = TMA path not available) and opts into the shared memory.
int fast_tma_prepare(FastDevice& d) {
    d.tma_warps = 0;
    if (encode_tiled() == nullptr) return SMX_OK;
    int device = 0, smem_optin = 0, smem_sm = 0;
    SMX_CUDA(cudaGetDevice(&device));
    SMX_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    SMX_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
    int want = 0;
    if (const char* env = std::getenv("SMX_FAST_WARPS")) want = std::atoi(env);
    if (want == -1) return SMX_OK;  // force the cp.async kernel
    const bool fits4 = 2 * (tma_smem_bytes(d, 4) + 1024) <= (size_t)smem_sm;
    const bool fits8 = tma_smem_bytes(d, 8) <= (size_t)smem_optin;
    const bool fits12 = tma_smem_bytes(d, 12) <= (size_t)smem_optin;
    if (want == 12 && fits12) d.tma_warps = 12;
    else if (want == 8 && fits8) d.tma_warps = 8;
    else if (want == 4 && fits4) d.tma_warps = 4;
    else if (fits4) d.tma_warps = 4;
    else if (fits12) d.tma_warps = 12;
    else if (fits8) d.tma_warps = 8;
    if (d.tma_warps == 4)
        SMX_CUDA(cudaFuncSetAttribute(fast_eval_tma_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem_bytes(d, 4)));
    if (d.tma_warps == 8)
        SMX_CUDA(cudaFuncSetAttribute(fast_eval_tma_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem_bytes(d, 8)));
    if (d.tma_warps == 12)
        SMX_CUDA(cudaFuncSetAttribute(fast_eval_tma_kernel<12, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem_bytes(d, 12)));
    return SMX_OK;
}

int fast_eval_tma(const FastDevice& d, const double* x, int64_t N, int64_t ldx, double* y, cudaStream_t st) {
    FastArgs a;
    fill_fast_args(d, x, N, ldx, a);
    // tensor map of x: (N rows) x (d_in columns) fp64, row pitch ldx * 8 bytes; box = 16 columns x 32 rows
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)d.d_in, (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)ldx * sizeof(double)};
    const cuuint32_t box[2] = {kBlockWidth, kTile};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult res = encode_tiled()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(x), dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (res != CUDA_SUCCESS) return fail(SMX_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)res));
    if (d.tma_warps == 4) return launch<4, 2>(map, a, d, x, y, st);
    if (d.tma_warps == 8) return launch<8, 1>(map, a, d, x, y, st);
    return launch<12, 1>(map, a, d, x, y, st);
}

}  // namespace smx
