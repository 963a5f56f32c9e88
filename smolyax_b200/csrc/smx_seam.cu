// Per-summand kernels on the REFERENCE layout: the device twins of the two jit(vmap(..)) callables the reference
// creates at interpolation.py:243-248, plus the integral einsum of interpolation.py:389.
//
// These keep the reference's arithmetic per summand — second barycentric form with b~_j = w_j / (x - xi_j)
// (barycentric.py:60-66), one-hot rows at node hits, normalisation by the row sum (:117), NaN derivative at node
// hits (:150-155), B_i = db/sum(b) - b * sum(db)/sum(b) (:212-219), zeta applied after the contraction (:123, :229)
// — but never materialise the (S, N, m) basis arrays or the dense per-summand (N, d_out, d_in) gradient: one thread
// owns one evaluation point, walks the summands of the group (their descriptors and values are warp-uniform
// broadcast loads), keeps the normalised 1-D bases of the current summand in shared memory ([entry][thread], so
// conflict free) and accumulates the group sum in registers.
#include <cmath>

#include "smx_common.cuh"

namespace smx {
namespace {

constexpr int kOutChunk = 4;  // outputs per thread (grid.y walks the chunks of d_out)

// sum_mu F[mu] prod_j basis(j, mu_j) over the exact box mu_j <= deg[j]; F strides in elements.
template <int NA, class Basis>
__device__ __forceinline__ double contract(const double* __restrict__ F, const int (&deg)[NA],
                                           const long long (&fstride)[NA], Basis basis) {
    int mu[NA];
    double pref[NA];  // pref[j] = prod_{i<=j} basis(i, mu_i), j < NA-1
#pragma unroll
    for (int j = 0; j < NA; ++j) mu[j] = 0;
    if constexpr (NA > 1) {
        pref[0] = basis(0, 0);
#pragma unroll
        for (int j = 1; j < NA - 1; ++j) pref[j] = pref[j - 1] * basis(j, 0);
    }
    double total = 0.0;
    long long foff = 0;
    while (true) {
        double partial = 0.0;
        for (int a = 0; a <= deg[NA - 1]; ++a) partial = fma(__ldg(F + foff + a * fstride[NA - 1]), basis(NA - 1, a), partial);
        if constexpr (NA == 1) return partial;
        total = fma(pref[NA > 1 ? NA - 2 : 0], partial, total);
        int j = NA - 2;
        while (j >= 0) {
            if (++mu[j] <= deg[j]) break;
            mu[j] = 0;
            --j;
        }
        if (j < 0) break;
        foff = 0;
#pragma unroll
        for (int i = 0; i < NA - 1; ++i) foff += mu[i] * fstride[i];
#pragma unroll
        for (int i = 0; i < NA - 1; ++i)
            if (i >= j) pref[i] = (i == 0 ? 1.0 : pref[i - 1]) * basis(i, mu[i]);
    }
    return total;
}

template <int NA, int BLOCK>
__global__ void __launch_bounds__(BLOCK) seam_eval_kernel(const double* __restrict__ x, long long N, long long ldx,
                                                           SeamGroup g, long long d_out, double* __restrict__ y) {
    extern __shared__ double sb[];  // [ent_total][BLOCK]
    const int tid = threadIdx.x;
    const long long p = (long long)blockIdx.x * BLOCK + tid;
    const bool active = p < N;
    const long long o0 = (long long)blockIdx.y * kOutChunk;
    const double* xp = x + (active ? p : 0) * ldx;
    double acc[kOutChunk];
#pragma unroll
    for (int i = 0; i < kOutChunk; ++i) acc[i] = 0.0;
    long long fstride[NA];
#pragma unroll
    for (int j = 0; j < NA; ++j) fstride[j] = g.fstride[j];

    for (long long s = 0; s < g.nn; ++s) {
        int deg[NA];
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            deg[j] = (int)__ldg(g.degs + s * NA + j);
            const double xv = xp[__ldg(g.dims + s * NA + j)];
            const double* xi = g.nodes + (s * NA + j) * g.tw;
            const double* w = g.weights + (s * NA + j) * g.tw;
            double* col = sb + (size_t)g.ent_off[j] * BLOCK + tid;
            bool hit = false;
            for (int a = 0; a <= deg[j]; ++a) hit |= (xv - __ldg(xi + a) == 0.0);
            double sum = 0.0;
            for (int a = 0; a <= deg[j]; ++a) {
                const double diff = xv - __ldg(xi + a);
                const double v = hit ? (diff == 0.0 ? 1.0 : 0.0) : __ldg(w + a) / diff;
                col[(size_t)a * BLOCK] = v;
                sum += v;
            }
            for (int a = 0; a <= deg[j]; ++a) col[(size_t)a * BLOCK] /= sum;
        }
        const double zeta = (double)__ldg(g.zetas + s);
        auto basis = [&](int j, int a) { return sb[(size_t)(g.ent_off[j] + a) * BLOCK + tid]; };
#pragma unroll
        for (int i = 0; i < kOutChunk; ++i) {
            if (o0 + i < d_out) {
                const double* Fo = g.F + (s * d_out + o0 + i) * g.fsize;
                acc[i] = fma(zeta, contract<NA>(Fo, deg, fstride, basis), acc[i]);
            }
        }
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < kOutChunk; ++i)
            if (o0 + i < d_out) y[p * d_out + o0 + i] += acc[i];
    }
}

template <int NA, int BLOCK>
__global__ void __launch_bounds__(BLOCK) seam_gradient_kernel(const double* __restrict__ x, long long N, long long ldx,
                                                               long long d_in, SeamGroup g, long long d_out,
                                                               double* __restrict__ J) {
    extern __shared__ double sb[];  // value bases [ent_total][BLOCK], then derivative bases [ent_total][BLOCK]
    const int tid = threadIdx.x;
    const long long p = (long long)blockIdx.x * BLOCK + tid;
    const bool active = p < N;
    const long long o0 = (long long)blockIdx.y * kOutChunk;
    const double* xp = x + (active ? p : 0) * ldx;
    double* sd = sb + (size_t)g.ent_total * BLOCK;
    long long fstride[NA];
#pragma unroll
    for (int j = 0; j < NA; ++j) fstride[j] = g.fstride[j];

    for (long long s = 0; s < g.nn; ++s) {
        int deg[NA];
        long long dim[NA];
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            deg[j] = (int)__ldg(g.degs + s * NA + j);
            dim[j] = __ldg(g.dims + s * NA + j);
            const double xv = xp[dim[j]];
            const double* xi = g.nodes + (s * NA + j) * g.tw;
            const double* w = g.weights + (s * NA + j) * g.tw;
            double* col = sb + (size_t)g.ent_off[j] * BLOCK + tid;
            double* dcol = sd + (size_t)g.ent_off[j] * BLOCK + tid;
            bool hit = false;
            for (int a = 0; a <= deg[j]; ++a) hit |= (xv - __ldg(xi + a) == 0.0);
            double bn = 0.0, dbn = 0.0;
            for (int a = 0; a <= deg[j]; ++a) {
                const double diff = xv - __ldg(xi + a), wa = __ldg(w + a);
                const double v = hit ? (diff == 0.0 ? 1.0 : 0.0) : wa / diff;
                const double sq = diff * diff;
                const double dv = (sq == 0.0) ? nan("") : -wa / sq;
                col[(size_t)a * BLOCK] = v;
                dcol[(size_t)a * BLOCK] = dv;
                bn += v;
                dbn += dv;
            }
            for (int a = 0; a <= deg[j]; ++a) {
                const double b = col[(size_t)a * BLOCK] / bn;
                col[(size_t)a * BLOCK] = b;
                dcol[(size_t)a * BLOCK] = dcol[(size_t)a * BLOCK] / bn - b * dbn / bn;
            }
        }
        const double zeta = (double)__ldg(g.zetas + s);
#pragma unroll
        for (int c = 0; c < NA; ++c) {
            auto basis = [&](int j, int a) {
                const double* base = (j == c) ? sd : sb;
                return base[(size_t)(g.ent_off[j] + a) * BLOCK + tid];
            };
#pragma unroll
            for (int i = 0; i < kOutChunk; ++i) {
                if (o0 + i < d_out) {
                    const double* Fo = g.F + (s * d_out + o0 + i) * g.fsize;
                    const double val = zeta * contract<NA>(Fo, deg, fstride, basis);
                    if (active) J[(p * d_out + o0 + i) * d_in + dim[c]] += val;
                }
            }
        }
    }
}

// partial[o][block] = sum over the block's summands of zeta_s <F[s,o], quad_s>
template <int NA>
__global__ void __launch_bounds__(256) seam_integral_kernel(SeamGroup g, long long d_out, double* __restrict__ partial) {
    const long long o = blockIdx.y;
    long long fstride[NA];
#pragma unroll
    for (int j = 0; j < NA; ++j) fstride[j] = g.fstride[j];
    double acc = 0.0;
    for (long long s = (long long)blockIdx.x * 256 + threadIdx.x; s < g.nn; s += (long long)gridDim.x * 256) {
        int deg[NA];
#pragma unroll
        for (int j = 0; j < NA; ++j) deg[j] = (int)g.degs[s * NA + j];
        auto basis = [&](int j, int a) { return g.quad[(s * NA + j) * g.tw + a]; };
        acc = fma((double)g.zetas[s], contract<NA>(g.F + (s * d_out + o) * g.fsize, deg, fstride, basis), acc);
    }
    __shared__ double warp_sums[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += warp_sums[w];
        partial[o * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void integral_finish_kernel(const double* __restrict__ partial, int nblocks, long long d_out, double* q) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= d_out) return;
    double t = 0.0;
    for (int b = 0; b < nblocks; ++b) t += partial[o * nblocks + b];
    q[o] += t;
}

__global__ void fill_rows_kernel(double* __restrict__ y, long long total, long long d_out, const double* __restrict__ row) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        y[i] = row ? row[i % d_out] : 0.0;
}

__global__ void compute_weights_kernel(const double* __restrict__ nodes, int m, double* __restrict__ w) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        double p = 1.0;
        for (int i = 0; i < m; ++i) {
            double diff = nodes[i] - nodes[j];
            if (diff == 0.0) diff = 1.0;
            p *= 1.0 / diff;
        }
        w[j] = p;
    }
}

// barycentric.py:60-66 / :150-155 for a batch of points of one dimension, all m columns
__global__ void basis_kernel(const double* __restrict__ x, long long N, const double* __restrict__ xi,
                             const double* __restrict__ w, int m, long long nu, int derivative, double* __restrict__ out) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const double xv = x[p];
    bool hit = false;
    for (int a = 0; a < m; ++a) hit |= (xv - xi[a] == 0.0) && (a <= nu);
    for (int a = 0; a < m; ++a) {
        const double diff = xv - xi[a];
        double v;
        if (derivative) {
            const double sq = diff * diff;
            v = (sq == 0.0) ? nan("") : -w[a] / sq;
        } else {
            v = hit ? (diff == 0.0 ? 1.0 : 0.0) : w[a] / diff;
        }
        out[p * m + a] = (a <= nu) ? v : 0.0;
    }
}

int integral_blocks(const SeamGroup& g) {
    long long nb = (g.nn + 255) / 256;
    return (int)(nb < 1 ? 1 : (nb > 592 ? 592 : nb));
}

template <int NA>
int launch_eval(const double* x, int64_t N, int64_t ldx, const SeamGroup& g, int64_t d_out, double* y, cudaStream_t st) {
    const dim3 grid_y((unsigned)((d_out + kOutChunk - 1) / kOutChunk));
    if ((size_t)g.ent_total * 128 * 8 <= 200 * 1024) {
        const size_t smem = (size_t)g.ent_total * 128 * sizeof(double);
        SMX_CUDA(cudaFuncSetAttribute(seam_eval_kernel<NA, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        seam_eval_kernel<NA, 128><<<dim3((unsigned)((N + 127) / 128), grid_y.x), 128, smem, st>>>(x, N, ldx, g, d_out, y);
    } else if ((size_t)g.ent_total * 32 * 8 <= 200 * 1024) {
        const size_t smem = (size_t)g.ent_total * 32 * sizeof(double);
        SMX_CUDA(cudaFuncSetAttribute(seam_eval_kernel<NA, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        seam_eval_kernel<NA, 32><<<dim3((unsigned)((N + 31) / 32), grid_y.x), 32, smem, st>>>(x, N, ldx, g, d_out, y);
    } else {
        return fail(SMX_ERR_UNSUPPORTED, "group has too many nodes per summand for the per-summand kernel");
    }
    SMX_LAUNCH_CHECK("seam_eval_kernel");
    return SMX_OK;
}

template <int NA>
int launch_gradient(const double* x, int64_t N, int64_t ldx, int64_t d_in, const SeamGroup& g, int64_t d_out, double* J,
                    cudaStream_t st) {
    const unsigned gy = (unsigned)((d_out + kOutChunk - 1) / kOutChunk);
    if ((size_t)g.ent_total * 2 * 128 * 8 <= 200 * 1024) {
        const size_t smem = (size_t)g.ent_total * 2 * 128 * sizeof(double);
        SMX_CUDA(cudaFuncSetAttribute(seam_gradient_kernel<NA, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        seam_gradient_kernel<NA, 128><<<dim3((unsigned)((N + 127) / 128), gy), 128, smem, st>>>(x, N, ldx, d_in, g, d_out, J);
    } else if ((size_t)g.ent_total * 2 * 32 * 8 <= 200 * 1024) {
        const size_t smem = (size_t)g.ent_total * 2 * 32 * sizeof(double);
        SMX_CUDA(cudaFuncSetAttribute(seam_gradient_kernel<NA, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        seam_gradient_kernel<NA, 32><<<dim3((unsigned)((N + 31) / 32), gy), 32, smem, st>>>(x, N, ldx, d_in, g, d_out, J);
    } else {
        return fail(SMX_ERR_UNSUPPORTED, "group has too many nodes per summand for the per-summand gradient kernel");
    }
    SMX_LAUNCH_CHECK("seam_gradient_kernel");
    return SMX_OK;
}

}  // namespace

int make_seam_group(const smx_group_desc* d, SeamGroup& g) {
    if (!d) return fail(SMX_ERR_INVALID_ARG, "null group descriptor");
    if (d->n < 1 || d->n > kSeamMaxN) return fail(SMX_ERR_UNSUPPORTED, "number of active dimensions per summand must be in 1..8");
    if (d->nn < 0 || !d->tau) return fail(SMX_ERR_INVALID_ARG, "group descriptor: bad nn or tau");
    if (d->nn > 0 && (!d->F || !d->nodes || !d->weights || !d->dims || !d->degs || !d->zetas))
        return fail(SMX_ERR_INVALID_ARG, "group descriptor has null arrays");
    g.n = d->n;
    g.nn = d->nn;
    g.fsize = 1;
    g.tw = 0;
    g.ent_total = 0;
    for (int j = 0; j < kSeamMaxN; ++j) {
        g.shape[j] = 1;
        g.fstride[j] = 0;
        g.ent_off[j] = 0;
    }
    for (int j = 0; j < d->n; ++j) {
        if (d->tau[j] < 1 || d->tau[j] > 4096) return fail(SMX_ERR_INVALID_ARG, "tau out of range");
        g.shape[j] = (int)d->tau[j] + 1;
        g.ent_off[j] = g.ent_total;
        g.ent_total += g.shape[j];
        if (g.shape[j] > g.tw) g.tw = g.shape[j];
    }
    for (int j = d->n - 1; j >= 0; --j) {
        g.fstride[j] = g.fsize;
        g.fsize *= g.shape[j];
    }
    g.F = d->F;
    g.nodes = d->nodes;
    g.weights = d->weights;
    g.dims = (const long long*)d->dims;
    g.degs = (const long long*)d->degs;
    g.zetas = (const long long*)d->zetas;
    g.quad = d->quad;
    return SMX_OK;
}

int seam_eval(const double* x, int64_t N, int64_t ldx, const SeamGroup& g, int64_t d_out, double* y, cudaStream_t st) {
    if (N == 0 || g.nn == 0) return SMX_OK;
#define CALL_EVAL(NA) launch_eval<NA>(x, N, ldx, g, d_out, y, st)
    switch (g.n) {
        case 1: return CALL_EVAL(1);
        case 2: return CALL_EVAL(2);
        case 3: return CALL_EVAL(3);
        case 4: return CALL_EVAL(4);
        case 5: return CALL_EVAL(5);
        case 6: return CALL_EVAL(6);
        case 7: return CALL_EVAL(7);
        default: return CALL_EVAL(8);
    }
#undef CALL_EVAL
}

int seam_gradient(const double* x, int64_t N, int64_t ldx, int64_t d_in, const SeamGroup& g, int64_t d_out, double* J,
                  cudaStream_t st) {
    if (N == 0 || g.nn == 0) return SMX_OK;
#define CALL_GRAD(NA) launch_gradient<NA>(x, N, ldx, d_in, g, d_out, J, st)
    switch (g.n) {
        case 1: return CALL_GRAD(1);
        case 2: return CALL_GRAD(2);
        case 3: return CALL_GRAD(3);
        case 4: return CALL_GRAD(4);
        case 5: return CALL_GRAD(5);
        case 6: return CALL_GRAD(6);
        case 7: return CALL_GRAD(7);
        default: return CALL_GRAD(8);
    }
#undef CALL_GRAD
}

int64_t seam_integral_workspace(const SeamGroup& g, int64_t d_out) { return (int64_t)integral_blocks(g) * d_out; }

int seam_integral(const SeamGroup& g, int64_t d_out, double* q, double* workspace, int64_t workspace_doubles,
                  cudaStream_t st) {
    if (g.nn == 0) return SMX_OK;
    if (!g.quad) return fail(SMX_ERR_INVALID_ARG, "group has no quadrature-weight table");
    const int nb = integral_blocks(g);
    if (workspace_doubles < (int64_t)nb * d_out) return fail(SMX_ERR_INVALID_ARG, "integral workspace too small");
    const dim3 grid((unsigned)nb, (unsigned)d_out);
    switch (g.n) {
        case 1: seam_integral_kernel<1><<<grid, 256, 0, st>>>(g, d_out, workspace); break;
        case 2: seam_integral_kernel<2><<<grid, 256, 0, st>>>(g, d_out, workspace); break;
        case 3: seam_integral_kernel<3><<<grid, 256, 0, st>>>(g, d_out, workspace); break;
        case 4: seam_integral_kernel<4><<<grid, 256, 0, st>>>(g, d_out, workspace); break;
        case 5: seam_integral_kernel<5><<<grid, 256, 0, st>>>(g, d_out, workspace); break;
        case 6: seam_integral_kernel<6><<<grid, 256, 0, st>>>(g, d_out, workspace); break;
        case 7: seam_integral_kernel<7><<<grid, 256, 0, st>>>(g, d_out, workspace); break;
        default: seam_integral_kernel<8><<<grid, 256, 0, st>>>(g, d_out, workspace); break;
    }
    SMX_LAUNCH_CHECK("seam_integral_kernel");
    integral_finish_kernel<<<(unsigned)((d_out + 127) / 128), 128, 0, st>>>(workspace, nb, d_out, q);
    SMX_LAUNCH_CHECK("integral_finish_kernel");
    return SMX_OK;
}

int fill_rows(double* y, int64_t N, int64_t d_out, const double* row, cudaStream_t st) {
    const long long total = N * d_out;
    if (total == 0) return SMX_OK;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    fill_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(y, total, d_out, row);
    SMX_LAUNCH_CHECK("fill_rows_kernel");
    return SMX_OK;
}

int device_basis(const double* x, int64_t N, const double* xi, const double* w, int64_t m, int64_t nu, int derivative,
                 double* out, cudaStream_t st) {
    if (N <= 0 || m <= 0) return SMX_OK;
    basis_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(x, N, xi, w, (int)m, nu, derivative, out);
    SMX_LAUNCH_CHECK("basis_kernel");
    return SMX_OK;
}

int device_compute_weights(const double* nodes, int64_t m, double* w, cudaStream_t st) {
    if (m <= 0) return SMX_OK;
    compute_weights_kernel<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(nodes, (int)m, w);
    SMX_LAUNCH_CHECK("compute_weights_kernel");
    return SMX_OK;
}

}  // namespace smx
