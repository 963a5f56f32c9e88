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

#include "smx_dense.cuh"

namespace smx {
namespace {

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

template <int NW, int NB, int PF>
__global__ void __launch_bounds__(NW * 32, 1)
dense_eval_kernel(const DenseArgs a, const double* __restrict__ x, double* __restrict__ y) {
    static_assert(NW % PF == 0, "the register ring is indexed with the k-step inside a stage");
    constexpr int kThreads = NW * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* abuf = reinterpret_cast<double*>(smem_raw);       // [2][NW][32 lanes][4]: A fragments of two stages
    double* tab = abuf + 2 * NW * 128;                         // [n_tab][kTabPitch] value table

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tig = lane & 3, gid = lane >> 2;
    const long long p0 = (long long)blockIdx.x * kDenseTile;

    // ---- prologue: value table = 1 | hot basis values | products of hot pairs, level by level ---------------------------
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

    // ---- main loop -----------------------------------------------------------------------------------------------------
    const int n_stage = (a.k4 + NW - 1) / NW;
    // this warp's output blocks (8 outputs each)
    const int jb0 = (blockIdx.y * NW + warp) * NB;
    const int nbv = max(0, min(NB, a.nblk - jb0));
    const double* bsrc = a.coef + ((size_t)jb0 * a.k4) * 32 + lane;
    const size_t bstride = (size_t)a.k4 * 32;

    // A assembly: this thread owns (k-step `warp` of the stage, fragment lane `lane`): term 4 * k4 + tig, points gid + 8 i
    const double* xr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xr[i] = x + min(p0 + gid + 8 * i, a.N - 1) * a.ldx;
    auto meta_of = [&](int stage) {
        const int g = stage * NW + warp;
        return g < a.k4 ? __ldg(a.meta + 4 * g + tig) : make_int2(0, 0);
    };
    double xc[4] = {0.0, 0.0, 0.0, 0.0};
    auto load_cold = [&](int2 m) {  // leading entry on a cold column: pi = x - eta0, straight from x
        if (m.y < 0) {
            const int dim = -1 - m.y;
            const double e0 = __ldg(a.eta0 + dim);
#pragma unroll
            for (int i = 0; i < 4; ++i) xc[i] = __ldg(xr[i] + dim) - e0;
        }
    };
    auto assemble = [&](int2 m, int buf) {
        double v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double lead = m.y >= 0 ? tab[m.y * kTabPitch + gid + 8 * i] : xc[i];
            v[i] = tab[m.x * kTabPitch + gid + 8 * i] * lead;
        }
        double2* dst = reinterpret_cast<double2*>(abuf + ((buf * NW + warp) * 32 + lane) * 4);
        dst[0] = make_double2(v[0], v[1]);
        dst[1] = make_double2(v[2], v[3]);
    };

    int2 m1 = meta_of(0);
    load_cold(m1);
    assemble(m1, 0);
    m1 = meta_of(1);
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
        for (int j = 0; j < NB; ++j) bq[u][j] = (u < a.k4 && j < nbv) ? __ldg(bsrc + j * bstride + (size_t)u * 32) : 0.0;

    for (int s = 0; s < n_stage; ++s) {
        const int2 m2 = meta_of(s + 2);
        load_cold(m1);  // for stage s + 1; consumed after the DMMAs below
        if (nbv > 0) {
            const double* af = abuf + ((s & 1) * NW * 32 + lane) * 4;
#pragma unroll
            for (int kk = 0; kk < NW; ++kk) {
                const int g = s * NW + kk;
                if (g < a.k4) {
                    const double2 a01 = *reinterpret_cast<const double2*>(af + kk * 128);
                    const double2 a23 = *reinterpret_cast<const double2*>(af + kk * 128 + 2);
                    double b[NB];
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        b[j] = bq[kk % PF][j];
                        bq[kk % PF][j] = (g + PF < a.k4 && j < nbv) ? __ldg(bsrc + j * bstride + (size_t)(g + PF) * 32) : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        dmma(acc[0][j], a01.x, b[j]);
                        dmma(acc[1][j], a01.y, b[j]);
                        dmma(acc[2][j], a23.x, b[j]);
                        dmma(acc[3][j], a23.y, b[j]);
                    }
                }
            }
        }
        if (s + 1 < n_stage) assemble(m1, (s + 1) & 1);
        m1 = m2;
        __syncthreads();
    }

    // ---- epilogue ------------------------------------------------------------------------------------------------------
    const bool vec = (a.d_out & 1) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        if (j >= nbv) break;
        const long long col = 8ll * (jb0 + j) + 2 * tig;
        if (col >= a.d_out) continue;
        const double c0a = __ldg(a.c0 + col), c0b = col + 1 < a.d_out ? __ldg(a.c0 + col + 1) : 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long p = p0 + gid + 8 * i;
            if (p >= a.N) continue;
            double* dst = y + p * a.d_out + col;
            if (vec) {
                *reinterpret_cast<double2*>(dst) = make_double2(c0a + acc[i][j][0], c0b + acc[i][j][1]);
            } else {
                dst[0] = c0a + acc[i][j][0];
                if (col + 1 < a.d_out) dst[1] = c0b + acc[i][j][1];
            }
        }
    }
}

size_t dense_smem_bytes(int n_tab, int nw) { return sizeof(double) * ((size_t)2 * nw * 128 + (size_t)n_tab * kTabPitch); }

template <int NW, int NB, int PF>
int launch(const DenseArgs& a, const double* x, double* y, cudaStream_t st) {
    const size_t smem = dense_smem_bytes(a.n_tab, NW);
    static size_t opted = 0;
    if (smem > opted) {
        SMX_CUDA(cudaFuncSetAttribute(dense_eval_kernel<NW, NB, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        opted = smem;
    }
    const long long tiles = (a.N + kDenseTile - 1) / kDenseTile;
    const int groups = (a.nblk + NW * NB - 1) / (NW * NB);
    dense_eval_kernel<NW, NB, PF><<<dim3((unsigned)tiles, (unsigned)groups), NW * 32, smem, st>>>(a, x, y);
    SMX_LAUNCH_CHECK("dense_eval_kernel");
    return SMX_OK;
}

}  // namespace

bool dense_kernel_fits(int n_tab, int smem_optin) { return dense_smem_bytes(n_tab, 16) <= (size_t)smem_optin; }

// CTA shape: NB output blocks per warp (register tile 32 points x 8 NB outputs), NW warps.  Few outputs: narrow warp
// tiles so that every warp has work; many outputs: the widest tile (fewest A-fragment loads per DMMA).
int dense_kernel_launch(const DenseArgs& a, const double* x, double* y, cudaStream_t st) {
    static const int want_nb = std::getenv("SMX_DENSE_NB") ? std::atoi(std::getenv("SMX_DENSE_NB")) : 0;
    static const int want_nw = std::getenv("SMX_DENSE_NW") ? std::atoi(std::getenv("SMX_DENSE_NW")) : 0;
    int nb = want_nb ? want_nb : (a.nblk >= 48 ? 4 : a.nblk > 16 ? 2 : 1);
    int nw = want_nw ? want_nw : std::min(16, std::max(8, ((a.nblk + nb - 1) / nb + 3) / 4 * 4));
    if (nb == 4) return nw <= 8 ? launch<8, 4, 2>(a, x, y, st) : launch<16, 4, 2>(a, x, y, st);
    if (nb == 2) return nw <= 8 ? launch<8, 2, 4>(a, x, y, st) : nw <= 12 ? launch<12, 2, 4>(a, x, y, st) : launch<16, 2, 4>(a, x, y, st);
    return nw <= 8 ? launch<8, 1, 4>(a, x, y, st) : nw <= 12 ? launch<12, 1, 4>(a, x, y, st) : launch<16, 1, 4>(a, x, y, st);
}

}  // namespace smx
