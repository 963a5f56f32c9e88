// Block-sparse evaluation kernel for a FEW outputs (2 <= d_out < 32): S coefficient sets per pass.
//
// smx_fast_kernel.cu walks the work items once per output; everything but the coefficients is the same for every output
// (staging of the item, the A fragments multiplied from the value table, the leading basis values from the x tile), and
// that fixed part is ~85 % of the instructions of a pass.  Here one pass serves S outputs: the record copy brings the S
// consecutive records of the item (metadata + coefficients of sets o .. o + S - 1; they are contiguous, see
// fast_upload()), the A fragment of a k-step and the basis values are formed once, and only the B fragment load, the
// 8 DMMAs and the 16 FMAs of `tot` repeat per set.  FLAT layout only (value table = hot rows, product rows multiplied on
// the fly); values only - the gradient's derivative sets are columns of the dense product (smx_dense_kernel.cu).
// Same arithmetic per output as the single-set kernel (only the grouping of the per-warp partial sums differs with the
// number of warps).
#include <algorithm>
#include <cstdlib>

#include "smx_fast_device.cuh"

namespace smx {
namespace {

constexpr int kRecMax = (int)sizeof(ItemBuffer);  // metadata + 4 k-steps of coefficients

template <int S>
struct alignas(16) MultiStage {
    unsigned char rec[2][S * kRecMax];
    unsigned long long bar[2];
};

template <int NW, int S>
__global__ void __launch_bounds__(NW * 32, 1)
fast_multi_kernel(const __grid_constant__ CUtensorMap xmap, const FastArgs a, const double* __restrict__ x, double* __restrict__ y) {
    constexpr int kThreads = NW * 32;
    static_assert(NW >= S, "the epilogue uses one warp per set");
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    XTile* xtiles = reinterpret_cast<XTile*>(base);                                // [NW]
    MultiStage<S>* stages = reinterpret_cast<MultiStage<S>*>(xtiles + NW);         // [NW]
    double* tab = reinterpret_cast<double*>(stages + NW);                          // [1 + n_hot_rows][kTabPitch]
    int4* s_dir = reinterpret_cast<int4*>(tab + (size_t)(1 + a.n_hot_rows) * kTabPitch);
    double* s_eta = reinterpret_cast<double*>(s_dir + a.n_chunks);
    int* s_hot_off = reinterpret_cast<int*>(s_eta + a.n_hot);
    int* s_hot_row = s_hot_off + a.hot_dims + 1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tig = lane & 3, gid = lane >> 2;
    MultiStage<S>& st = stages[warp];
    double* xs = xtiles[warp].v;
    const double* xlo = xs + gid * kBlockWidth + (((2 * tig) ^ gid) << 1);
    const double* xhi = xs + gid * kBlockWidth + (((2 * tig + 1) ^ gid) << 1);

    for (int i = tid; i < a.n_chunks; i += kThreads) s_dir[i] = __ldg(a.chunk_dir + i);
    for (int i = tid; i < a.n_hot; i += kThreads) s_eta[i] = __ldg(a.eta + i);
    for (int i = tid; i <= a.hot_dims; i += kThreads) s_hot_off[i] = __ldg(a.hot_off + i);
    for (int i = tid; i < a.n_hot; i += kThreads) s_hot_row[i] = 1 + hot_row(__ldg(a.hot_pos + i));
    if (lane == 0) {
        mbar_init(&st.bar[0], 1);
        mbar_init(&st.bar[1], 1);
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

    // one lane stages item c for the sets [set0, set0 + ns): ns consecutive records in one bulk copy (+ the x tile)
    auto stage = [&](int buf, const int4 dir, long long set0, int ns, long long p0) {
        const int ksteps = (dir.y + 3) >> 2;
        const unsigned rec_bytes = kMetaInts * 4 + (unsigned)ksteps * kKStepDoubles * 8;
        const bool cold = !(dir.z & kChunkHot);
        mbar_expect_tx(&st.bar[buf], ns * rec_bytes + (cold ? kXTileBytes : 0));
        bulk_copy(st.rec[buf], a.coef + ((size_t)dir.x + (size_t)set0 * (5 + 4 * ksteps)) * 16, ns * rec_bytes, &st.bar[buf]);
        if (cold) tma_load_2d(xs, &xmap, dir.w, (int)p0, &st.bar[buf]);
    };

    unsigned k_item = 0;
    for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const long long p0 = tile * kTile;
        // ---- prologue: ones row + 1-D basis values of the hot entries -------------------------------------------------
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
                const double* xrow = x + min(p0 + lane, a.N - 1) * a.ldx;
                for (int d = warp + NW * kHotRegs; d < a.hot_dims; d += NW) hot_dim(d, __ldg(xrow + d));
            }
        }
        __syncthreads();
        if (tile + gridDim.x < a.num_tiles) load_hot((tile + gridDim.x) * kTile);

        const int n_pass = (a.o_end - a.o_begin + S - 1) / S;
        for (int pass = 0; pass < n_pass; ++pass) {
            const long long set0 = a.o_begin + (long long)pass * S;
            const int ns = (int)min((long long)S, a.o_end - set0);
            double tot[S][4];
#pragma unroll
            for (int j = 0; j < S; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) tot[j][i] = 0.0;
            const int c_begin = a.warp_off[warp], c_end = a.warp_off[warp + 1];
            if (c_begin < c_end && lane == 0) stage(k_item & 1, s_dir[c_begin], set0, ns, p0);
            for (int c = c_begin; c < c_end; ++c, ++k_item) {
                const int buf = k_item & 1;
                const int4 dir = s_dir[c];
                const unsigned char* rec = st.rec[buf];
                const ItemBuffer& ib = *reinterpret_cast<const ItemBuffer*>(rec);
                const int ksteps = (dir.y + 3) >> 2;
                const int rec_doubles = kMetaInts / 2 + ksteps * kKStepDoubles;
                const double* coef = reinterpret_cast<const double*>(rec) + kMetaInts / 2 + 2 * lane;  // + j * rec_doubles + s * 64
                const int nf = (dir.z >> 8) & 7;
                mbar_wait(&st.bar[buf], (k_item >> 1) & 1);

                auto load_a = [&](int s, double2& lo, double2& hi) {  // A fragment of k-step s: 4 points of one row
                    const double* q = tab + 2 * gid;
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
                double2 alo, ahi;
                load_a(0, alo, ahi);

                // leading basis values pi_e(x_p) of the lane's 4 points x 4 entries
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
                    // (raw coordinates: pi = x - eta_0 is subtracted after the DMMAs - its DADDs share their FP64 pipe and would sit
                    // in front of the item's first DMMA; the centres stay in the record buffer until the item ends)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double2 lo = *reinterpret_cast<const double2*>(xlo + i * (8 * kBlockWidth));
                        const double2 hi = *reinterpret_cast<const double2*>(xhi + i * (8 * kBlockWidth));
                        v[i][0] = lo.x, v[i][1] = lo.y, v[i][2] = hi.x, v[i][3] = hi.y;
                    }
                }
                __syncwarp();  // every lane has taken its x values: the x buffer and the other record buffer are free
                if (c + 1 < c_end && lane == 0) stage(buf ^ 1, s_dir[c + 1], set0, ns, p0);
                double acc[S][4][2][2];
                {   // first k-step writes the accumulators
                    const double af[4] = {alo.x, alo.y, ahi.x, ahi.y};
#pragma unroll
                    for (int j = 0; j < S; ++j) {
                        const double2 b = *reinterpret_cast<const double2*>(coef + j * rec_doubles);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            dmma_first<0>(acc[j][i][0], af[i], b.x);
                            dmma_first<0>(acc[j][i][1], af[i], b.y);
                        }
                    }
                }
#pragma unroll
                for (int s = 1; s < 4; ++s) {
                    if (s >= ksteps) break;
                    load_a(s, alo, ahi);
                    const double af[4] = {alo.x, alo.y, ahi.x, ahi.y};
#pragma unroll
                    for (int j = 0; j < S; ++j) {
                        const double2 b = *reinterpret_cast<const double2*>(coef + j * rec_doubles + s * kKStepDoubles);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            dmma_<0>(acc[j][i][0], af[i], b.x);
                            dmma_<0>(acc[j][i][1], af[i], b.y);
                        }
                    }
                }
                if (!(dir.z & (kChunkHot | kChunkEtaZero))) {
                    const double2 ea = *reinterpret_cast<const double2*>(ib.eta0 + 4 * tig);
                    const double2 eb = *reinterpret_cast<const double2*>(ib.eta0 + 4 * tig + 2);
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i][0] -= ea.x, v[i][1] -= ea.y, v[i][2] -= eb.x, v[i][3] -= eb.y;
                }
#pragma unroll
                for (int j = 0; j < S; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        tot[j][i] = fma(v[i][0], acc[j][i][0][0], tot[j][i]);
                        tot[j][i] = fma(v[i][1], acc[j][i][0][1], tot[j][i]);
                        tot[j][i] = fma(v[i][2], acc[j][i][1][0], tot[j][i]);
                        tot[j][i] = fma(v[i][3], acc[j][i][1][1], tot[j][i]);
                    }
            }
            // ---- epilogue: reduce over the 4 lanes that share a point, then over the warps in fixed order ---------------
            // (the warp's x buffer is idle between its last item and the first item of the next pass: it carries the partial sums)
            __syncwarp();
#pragma unroll
            for (int j = 0; j < S; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double t = tot[j][i];
                    t += __shfl_xor_sync(0xffffffffu, t, 1);
                    t += __shfl_xor_sync(0xffffffffu, t, 2);
                    if (tig == 0) xs[j * kTile + gid + 8 * i] = t;
                }
            __syncthreads();
            if (tid < kTile * ns) {
                const int j = tid >> 5, p = tid & 31;
                if (p0 + p < a.N) {
                    double s = __ldg(a.c0 + set0 + j);
#pragma unroll
                    for (int w = 0; w < NW; ++w) s += xtiles[w].v[j * kTile + p];
                    y[(p0 + p) * a.d_out + set0 + j] = s;
                }
            }
            __syncthreads();  // the x buffers are rewritten by the first TMA of the next pass / tile
        }
    }
}

template <int S>
size_t multi_smem_bytes(const FastDevice& d, int nw) {
    return 1024 + (size_t)nw * (sizeof(XTile) + sizeof(MultiStage<S>)) + ((size_t)(1 + d.n_hot_rows) * kTabPitch + (size_t)d.n_hot) * sizeof(double) +
           (size_t)d.n_chunks * sizeof(int4) + ((size_t)d.hot_dims + 1 + d.n_hot) * sizeof(int) + 64;
}

template <int NW, int S>
int launch_multi(const CUtensorMap& map, const FastArgs& a, const FastDevice& d, const double* x, double* y, cudaStream_t st) {
    const size_t smem = multi_smem_bytes<S>(d, NW);
    SMX_CUDA(cudaFuncSetAttribute(fast_multi_kernel<NW, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = std::min<long long>(a.num_tiles, (long long)d.sm_count);
    fast_multi_kernel<NW, S><<<(unsigned)grid, NW * 32, smem, st>>>(map, a, x, y);
    SMX_LAUNCH_CHECK("fast_multi_kernel<%d,%d>", NW, S);
    return SMX_OK;
}

}  // namespace

// Shape for d_out outputs: (sets per pass, warps).  Returns false if the single-set kernel should run (measured at the cfg2
// and cfg4 tables: two sets x 12 warps and three sets x 8 warps gain 13-25 % for 2..6 outputs; beyond that the extra passes
// with a partly filled last group and one CTA per SM eat the gain, and four sets x 6 warps are too few warps).
bool multi_kernel_shape(const FastDevice& d, int smem_optin, int* sets, int* warps) {
    static const int want = tune_int("SMX_FAST_MULTI", -1);  // 0: off; 2, 3: force
    if (want == 0 || !d.flat_ok || d.d_out < 2 || d.d_out >= 32) return false;
    // More than 6 outputs: only where a pass costs much more than its DMMAs - non-zero first centres (16 DADDs per cold item
    // and pass) and hot parts of several pairs (DMULs per k-step and pass), i.e. Gauss-Hermite-like plans.  Measured r08 at cfg4's
    // tables, ms per 1e5 points, one output per pass vs three: 9 outputs 2.31 vs 2.06, 12: 3.08 vs 2.69; at cfg2's tables
    // (zero centres, 2e5 points) 9: 2.67 vs 2.62, 8: 2.38 vs 2.63.  The outputs beyond a multiple of three take the
    // single-set kernel in a second launch (fast_kernel_launch).
    if (want < 0 && d.d_out > 6 && d.eta0_zero) return false;
    // (measured against one output per pass of the lean kernel, cfg2 tables, ms per 1e6 points: 2 outputs 3.46 vs 3.29,
    //  3: 4.57 vs 4.75, 4: 6.52 vs 6.21, 6: 8.74 vs 9.13 - three sets per pass still pay, two no longer do)
    if (want < 0 && (d.d_out == 2 || d.d_out == 4)) return false;
    const int s = want > 0 ? want : 3;
    if (s == 2 && multi_smem_bytes<2>(d, 12) <= (size_t)smem_optin) return *sets = 2, *warps = 12, true;
    if (s == 3 && multi_smem_bytes<3>(d, 8) <= (size_t)smem_optin) return *sets = 3, *warps = 8, true;
    return false;
}

int multi_kernel_launch(const CUtensorMap& map, const FastDevice& d, const FastArgs& a, const double* x, double* y, cudaStream_t st) {
    if (d.multi == 2) return launch_multi<12, 2>(map, a, d, x, y, st);
    return launch_multi<8, 3>(map, a, d, x, y, st);
}

}  // namespace smx
