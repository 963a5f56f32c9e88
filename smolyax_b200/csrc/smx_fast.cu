// Fast evaluation path, host side: device upload of the plan (smx_plan.h), coefficient packing, launch.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "smx_dense.cuh"
#include "smx_fast_common.cuh"

namespace smx {
namespace {

template <class T>
int upload(const std::vector<T>& v, T** dptr, int64_t& bytes, size_t min_elems = 1) {
    const size_t n = std::max(v.size(), min_elems);
    SMX_CUDA(cudaMalloc((void**)dptr, n * sizeof(T)));
    SMX_CUDA(cudaMemset(*dptr, 0, n * sizeof(T)));
    if (!v.empty()) SMX_CUDA(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    bytes += (int64_t)(n * sizeof(T));
    return SMX_OK;
}

// Coefficients of one (work item, output) in the order the DMMA B fragments are read: k-step s (rows 4s..4s+3),
// lane = 4 * gid + tig, n-tile j:   packed[s][lane][j] = C[row 4s + tig][entry 4 * (gid >> 1) + 2 * j + (gid & 1)]
// (rows beyond the item's are zero).  One LDS.128 per lane and k-step, consecutive lanes at consecutive addresses.
void pack_coefficients(const FastPlan& plan, std::vector<double>& packed, std::vector<int32_t>& kstep_off) {
    const size_t dout = (size_t)plan.n_sets;  // every coefficient set (outputs, then derivative sets) is packed alike
    kstep_off.assign((size_t)plan.n_chunks + 1, 0);
    for (int32_t c = 0; c < plan.n_chunks; ++c) {
        const int32_t rows = plan.chunk_off[c + 1] - plan.chunk_off[c];
        kstep_off[c + 1] = kstep_off[c] + (int32_t)(((rows + 3) / 4) * dout);
    }
    packed.assign((size_t)kstep_off.back() * kKStepDoubles, 0.0);
    for (int32_t c = 0; c < plan.n_chunks; ++c) {
        const size_t r0 = plan.chunk_off[c], rows = plan.chunk_off[c + 1] - r0, ksteps = (rows + 3) / 4;
        for (size_t o = 0; o < dout; ++o)
            for (size_t s = 0; s < ksteps; ++s)
                for (int lane = 0; lane < 32; ++lane)
                    for (int j = 0; j < 2; ++j) {
                        const size_t row = 4 * s + (lane & 3);
                        const int gid = lane >> 2, entry = 4 * (gid >> 1) + 2 * j + (gid & 1);
                        if (row < rows)
                            packed[((size_t)kstep_off[c] + o * ksteps + s) * kKStepDoubles + 2 * lane + j] =
                                plan.coef[((r0 + row) * dout + o) * kBlockWidth + entry];
                    }
    }
}

}  // namespace

int fast_upload(const FastPlan& plan, FastDevice& dev) {
    dev = FastDevice();
    dev.d_in = plan.d_in;
    dev.d_out = plan.d_out;
    dev.n_tab = plan.n_tab;
    dev.n_hot = plan.n_hot;
    dev.n_hot_rows = plan.n_hot_rows;
    dev.n_levels = plan.n_levels;
    dev.n_chunks = plan.n_chunks;
    dev.hot_dims = plan.hot_dims;
    dev.n_pairs = (int32_t)plan.tab_parent.size();
    dev.n_flat = (int32_t)(plan.tab_factors.size() / 4);
    if (plan.n_levels > kMaxLevels) return fail(SMX_ERR_UNSUPPORTED, "too many active dimensions per term");
    for (size_t l = 0; l < plan.level_off.size() && l < (size_t)kMaxLevels + 2; ++l) dev.level_off[l] = plan.level_off[l];
    bool sparse_ok = plan.has_sparse;
    for (int32_t c = 0; c < plan.n_chunks; ++c) {
        if (!(plan.chunk_flags[c] & (kChunkHot | kChunkContig))) sparse_ok = false;
        if (!(plan.chunk_flags[c] & kChunkHot)) dev.has_cold = true;
        if (!(plan.chunk_flags[c] & (kChunkHot | kChunkEtaZero))) dev.eta0_zero = false;
    }
    int device = 0, smem_optin = 0;
    SMX_CUDA(cudaGetDevice(&device));
    SMX_CUDA(cudaDeviceGetAttribute(&dev.sm_count, cudaDevAttrMultiProcessorCount, device));
    SMX_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    int rc;
    std::vector<int32_t> pairs(plan.tab_parent.size() * 2);
    for (size_t i = 0; i < plan.tab_parent.size(); ++i) pairs[2 * i] = plan.tab_parent[i], pairs[2 * i + 1] = plan.tab_hot[i];
    if ((rc = upload(plan.eta, &dev.eta, dev.bytes))) return rc;
    if ((rc = upload(pairs, &dev.tab_pairs, dev.bytes, 2))) return rc;
    if ((rc = upload(plan.hot_off, &dev.hot_off, dev.bytes))) return rc;
    if ((rc = upload(plan.hot_pos, &dev.hot_pos, dev.bytes))) return rc;
    if ((rc = upload(plan.c0, &dev.c0, dev.bytes))) return rc;
    if ((rc = upload(plan.nan_off, &dev.nan_off, dev.bytes, 2))) return rc;
    if ((rc = upload(plan.nan_nodes, &dev.nan_nodes, dev.bytes))) return rc;
    // (the kernel addresses the dense matrix with 32-bit element offsets)
    if ((plan.has_dense || plan.has_dense_grad) && dense_kernel_fits(plan.n_tab, smem_optin) &&
        plan.dense_coef.size() < (size_t)UINT32_MAX && plan.dense_grad_coef.size() < (size_t)UINT32_MAX) {
        dev.dense_k4 = plan.dense_k4;
        if ((rc = upload(plan.dense_meta, &dev.dense_meta, dev.bytes, 2))) return rc;
        if ((rc = upload(plan.dense_eta0, &dev.dense_eta0, dev.bytes))) return rc;
        if (plan.has_dense) {
            if ((rc = upload(plan.dense_coef, &dev.dense_coef, dev.bytes))) return rc;
            dev.has_dense = true;
        }
        if (plan.has_dense_grad && sparse_ok) {  // (the cold columns of J come from the block-sparse kernel)
            if ((rc = upload(plan.dense_grad_coef, &dev.dense_grad_coef, dev.bytes))) return rc;
            if ((rc = upload(plan.dense_grad_c0, &dev.dense_grad_c0, dev.bytes))) return rc;
            if ((rc = upload(plan.dense_grad_col, &dev.dense_grad_col, dev.bytes))) return rc;
            dev.has_dense_grad = true;
        }
    }
    // (else: the product table is too large for the GEMM-regime kernel.  Values run the block-sparse kernel, one output per
    // pass; if the derivative sets were only built as dense columns, grad_ok stays false and the gradient runs per summand.)
    if (!sparse_ok) {
        if (!dev.has_dense)
            return fail(SMX_ERR_UNSUPPORTED, plan.has_sparse ? "plan has a cold block that is not a contiguous tile of x"
                                                             : "value table does not fit in shared memory");
        return SMX_OK;
    }
    dev.flat_ok = plan.flat_ok;
    dev.deep_ok = !plan.flat_ok && plan.deep_ok;  // hot parts of five to eight pairs: records with a second factor list
    if ((rc = fast_kernel_prepare(dev))) return dev.has_dense ? SMX_OK : rc;

    std::vector<double> packed;
    std::vector<int32_t> kstep_off;
    pack_coefficients(plan, packed, kstep_off);
    // One record per (work item, coefficient set): metadata (640 B; every value-table row index pre-multiplied by kTabPitch,
    // i.e. an offset in doubles) followed by the set's packed coefficients, so that the kernel stages an item with ONE bulk
    // copy.  rec_off[c] = offset of item c's first record in units of 128 bytes; the records of the other sets follow at a
    // stride of 5 + 4 * ksteps units (+ 2 for the second factor list of deep records).
    const size_t nsets = (size_t)plan.n_sets, extra = dev.deep ? 2 : 0;
    std::vector<int32_t> rec_off((size_t)plan.n_chunks);
    if (dev.pipe_warps > 0) dev.pipe_warps = pipe_kernel_workers(dev, smem_optin);  // (marker set by fast_kernel_prepare: eligible)
    {
        size_t units = 0;
        for (int32_t c = 0; c < plan.n_chunks; ++c)
            units += nsets * (5 + extra + 4 * (size_t)((plan.chunk_dir[(size_t)c * 4 + 1] + 3) / 4));
        if (units >= (size_t)INT32_MAX) return fail(SMX_ERR_UNSUPPORTED, "too many coefficient sets for the block-sparse form");
        std::vector<double> records(units * 16, 0.0);
        std::vector<int32_t> meta(kMetaInts), fac2(64);
        size_t at = 0;
        for (int32_t c = 0; c < plan.n_chunks; ++c) {
            const size_t ksteps = (size_t)((plan.chunk_dir[(size_t)c * 4 + 1] + 3) / 4), k0 = (size_t)kstep_off[c];
            std::copy_n(&plan.chunk_meta[(size_t)c * kMetaInts], kMetaInts, meta.begin());
            for (int i = 0; i < 16; ++i) meta[i] *= kTabPitch, meta[48 + i] *= kTabPitch;
            for (int i = 96; i < 160; ++i) meta[i] *= kTabPitch;
            if (dev.deep)
                for (int i = 0; i < 64; ++i) fac2[i] = plan.chunk_fac2[(size_t)c * 64 + i] * kTabPitch;
            rec_off[c] = (int32_t)at;
            for (size_t o = 0; o < nsets; ++o) {
                std::memcpy(&records[at * 16], meta.data(), kMetaInts * 4);
                std::memcpy(&records[(at + 5) * 16], &packed[(k0 + o * ksteps) * kKStepDoubles], ksteps * kKStepDoubles * 8);
                if (dev.deep) std::memcpy(&records[(at + 5 + 4 * ksteps) * 16], fac2.data(), 256);
                at += 5 + extra + 4 * ksteps;
            }
        }
        std::vector<double>().swap(packed);
        if ((rc = upload(records, &dev.coef, dev.bytes, 2))) return rc;
    }
    // Balanced static schedule for a CTA shape of nw warps: longest-processing-time assignment of the work items to the
    // warps (cost ~ fixed part + k-steps), then every warp alternates big and small items so that tensor-heavy and
    // streaming items are always in flight together.  The directory is stored in that order: per item the offset of its
    // first record, its rows, flags | nf << 8 | record size in 128-byte units << 16 | k-steps << 24, first column of x.
    auto build_directory = [&](int nw, std::vector<int32_t>& dir, int32_t* warp_off, bool pipe) {
        dir.assign((size_t)plan.n_chunks * 4, 0);
        // cost model of an item: fixed + per k-step + extra for streaming x (measured at the headline configuration, ms per
        // 10^6 points: barrier kernels 2,1,0; pipelined kernel 2,1,0: 1.684, 2,1,1: 1.651, 1,1,1: 1.659, 3,1,1: 1.694, 2,1,2: 1.781)
        double ca = 2.0, cb = 1.0, cc = pipe ? 1.0 : 0.0;
        if (const char* env = tune_str("SMX_FAST_COST")) std::sscanf(env, "%lf,%lf,%lf", &ca, &cb, &cc);
        auto cost = [&](int32_t c) {
            return ca + cb * (double)((plan.chunk_off[c + 1] - plan.chunk_off[c] + 3) / 4) + ((plan.chunk_flags[c] & kChunkHot) ? 0.0 : cc);
        };
        std::vector<int32_t> by_cost((size_t)plan.n_chunks);
        for (int32_t c = 0; c < plan.n_chunks; ++c) by_cost[c] = c;
        std::stable_sort(by_cost.begin(), by_cost.end(), [&](int32_t x1, int32_t x2) { return cost(x1) > cost(x2); });
        std::vector<std::vector<int32_t>> lists((size_t)nw);
        std::vector<double> load((size_t)nw, 0.0);
        for (int32_t c : by_cost) {
            const size_t w = std::min_element(load.begin(), load.end()) - load.begin();
            lists[w].push_back(c);
            load[w] += cost(c);
        }
        size_t pos = 0;
        for (int w = 0; w < nw; ++w) {
            warp_off[w] = (int32_t)pos;
            const std::vector<int32_t>& l = lists[w];  // sorted by cost, descending
            for (size_t lo = 0, hi = l.size(), k = 0; lo < hi; ++k) {
                const int32_t c = (k & 1) ? l[--hi] : l[lo++];
                const int32_t ks = (plan.chunk_dir[(size_t)c * 4 + 1] + 3) / 4;
                dir[pos * 4 + 0] = rec_off[c];
                dir[pos * 4 + 1] = plan.chunk_dir[(size_t)c * 4 + 1];
                dir[pos * 4 + 2] = (plan.chunk_dir[(size_t)c * 4 + 2] & 0xffff) | ((5 + 4 * ks + (dev.deep ? 2 : 0)) << 16) | (ks << 24);
                dir[pos * 4 + 3] = plan.chunk_dir[(size_t)c * 4 + 3];
                ++pos;
            }
        }
        for (int w = nw; w <= kMaxWarps; ++w) warp_off[w] = (int32_t)pos;
    };
    {
        std::vector<int32_t> dir;
        build_directory(dev.warps, dir, dev.warp_off, false);
        if ((rc = upload(dir, &dev.chunk_dir, dev.bytes, 4))) return rc;
        if (dev.pipe_warps > 0) {  // the pipelined kernel has its own worker count, hence its own item lists
            build_directory(dev.pipe_warps, dir, dev.pipe_warp_off, true);
            if ((rc = upload(dir, &dev.pipe_dir, dev.bytes, 4))) return rc;
        }
    }
    if ((rc = upload(plan.tab_factors, &dev.tab_factors, dev.bytes, 4))) return rc;
    dev.n_sets = plan.n_sets;
    dev.n_gd = (int32_t)plan.grad_dims.size();
    dev.grad_ok = !dev.deep && (dev.has_dense_grad || (plan.n_sets == plan.d_out * (1 + (int64_t)plan.grad_dims.size()) &&
                                         (dev.n_gd > 0 || plan.hot_dims == 0 || plan.n_hot == 0)));
    if ((rc = upload(plan.grad_dims, &dev.grad_dims, dev.bytes))) return rc;
    dev.has_sparse = true;
    return SMX_OK;
}

void fast_free(FastDevice& d) {
    void* ptrs[] = {d.eta, d.tab_pairs, d.tab_factors, d.hot_off, d.hot_pos, d.chunk_dir, d.pipe_dir, d.chunk_meta, d.coef, d.c0,
                    d.grad_dims, d.nan_off, d.nan_nodes, d.dense_meta, d.dense_eta0, d.dense_coef,
                    d.dense_grad_coef, d.dense_grad_c0, d.dense_grad_col};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    d = FastDevice();
}

namespace {

// J[p, :, dim] = NaN where x[p, dim] sits on an interpolation node of that dimension (reference barycentric.py:152-154)
__global__ void nan_at_nodes_kernel(const double* __restrict__ x, long long N, long long ldx, long long d_in, long long d_out,
                                    const int32_t* __restrict__ nan_off, const double* __restrict__ nan_nodes, double* __restrict__ J) {
    const long long total = N * d_in;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / d_in, dim = i - p * d_in;
        const double xv = x[p * ldx + dim];
        bool hit = false;
        for (int k = nan_off[dim]; k < nan_off[dim + 1]; ++k) hit |= (xv == nan_nodes[k]);
        if (hit)
            for (long long o = 0; o < d_out; ++o) J[(p * d_out + o) * d_in + dim] = nan("");
    }
}

int run_fast(const FastDevice& d, const double* x, int64_t N, int64_t ldx, double* out, int gradient, cudaStream_t st) {
    // TMA needs 16-byte aligned rows.  Anything else (odd pitch, odd base address) is first packed into an aligned
    // scratch copy on the same stream; callers that care about the last few percent pass aligned rows.
    double* scratch = nullptr;
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (ldx & 1) != 0) {
        const int64_t pitch = (d.d_in + 1) & ~(int64_t)1;
        SMX_CUDA(cudaMallocAsync((void**)&scratch, sizeof(double) * (size_t)N * pitch, st));
        SMX_CUDA(cudaMemcpy2DAsync(scratch, sizeof(double) * pitch, x, sizeof(double) * ldx, sizeof(double) * d.d_in, (size_t)N,
                                   cudaMemcpyDeviceToDevice, st));
        x = scratch;
        ldx = pitch;
    }
    FastArgs a;
    a.eta = d.eta;
    a.tab_pairs = reinterpret_cast<const int2*>(d.tab_pairs);
    a.tab_factors = reinterpret_cast<const int4*>(d.tab_factors);
    a.n_flat = d.n_flat;
    a.hot_off = d.hot_off;
    a.hot_pos = d.hot_pos;
    a.chunk_dir = reinterpret_cast<const int4*>(d.chunk_dir);
    a.coef = d.coef;
    a.c0 = d.c0;
    a.N = N;
    a.ldx = ldx;
    a.d_out = d.d_out;
    a.num_tiles = (N + kTile - 1) / kTile;
    a.gradient = gradient;
    a.n_gd = d.has_dense_grad ? 0 : d.n_gd;  // dense derivative columns: this kernel only runs the function's own sets
    a.grad_dims = d.grad_dims;
    a.d_in = d.d_in;
    a.n_hot = d.n_hot;
    a.n_hot_rows = d.n_hot_rows;
    for (int w = 0; w <= kMaxWarps; ++w) a.warp_off[w] = d.warp_off[w];
    a.n_tab = d.n_tab;
    a.n_chunks = d.n_chunks;
    a.n_levels = d.n_levels;
    a.hot_dims = d.hot_dims;
    a.n_pairs = d.n_pairs;
    a.flat = d.flat ? 1 : 0;
    a.ablate = tune_int("SMX_ABL_KERNEL", 0);
    a.dbg = nullptr;
    a.nwk = 0;
    for (int l = 0; l < kMaxLevels + 2; ++l) a.level_off[l] = d.level_off[l];
    const int rc = fast_kernel_launch(d, a, x, out, st);
    if (scratch) cudaFreeAsync(scratch, st);
    return rc;
}

}  // namespace

int fast_eval(const FastDevice& d, const double* x, int64_t N, int64_t ldx, double* y, cudaStream_t st) {
    if (N == 0) return SMX_OK;
    if (!d.has_sparse) return fail(SMX_ERR_UNSUPPORTED, "plan has no block-sparse form");
    return run_fast(d, x, N, ldx, y, 0, st);
}

namespace {

int run_dense(const FastDevice& d, const double* x, int64_t N, int64_t ldx, double* out, bool gradient, cudaStream_t st) {
    DenseArgs a;
    a.eta = d.eta;
    a.tab_pairs = reinterpret_cast<const int2*>(d.tab_pairs);
    a.hot_off = d.hot_off;
    a.hot_pos = d.hot_pos;
    a.meta = reinterpret_cast<const int2*>(d.dense_meta);
    a.eta0 = d.dense_eta0;
    a.coef = gradient ? d.dense_grad_coef : d.dense_coef;
    a.c0 = gradient ? d.dense_grad_c0 : d.c0;
    a.colmap = gradient ? d.dense_grad_col : nullptr;
    a.N = N;
    a.ldx = ldx;
    a.ncol = gradient ? d.d_out * d.n_gd : d.d_out;
    a.ldy = gradient ? d.d_out * d.d_in : d.d_out;
    a.k4 = d.dense_k4;
    a.nblk = (int)((a.ncol + 7) / 8);
    a.n_tab = d.n_tab;
    a.n_hot_rows = d.n_hot_rows;
    a.n_levels = d.n_levels;
    a.hot_dims = d.hot_dims;
    a.skew = 0;  // (dense_kernel_launch decides)
    for (int l = 0; l < kMaxLevels + 2; ++l) a.level_off[l] = d.level_off[l];
    return dense_kernel_launch(a, x, out, st);
}

}  // namespace

int dense_eval(const FastDevice& d, const double* x, int64_t N, int64_t ldx, double* y, cudaStream_t st) {
    if (N == 0) return SMX_OK;
    if (!d.has_dense) return fail(SMX_ERR_UNSUPPORTED, "plan has no dense form");
    return run_dense(d, x, N, ldx, y, false, st);
}

int fast_gradient(const FastDevice& d, const double* x, int64_t N, int64_t ldx, double* J, bool nan_at_nodes, cudaStream_t st) {
    if (N == 0) return SMX_OK;
    if (!d.has_sparse || !d.grad_ok) return fail(SMX_ERR_UNSUPPORTED, "plan has no derivative sets");
    // dimensions without any entry keep derivative zero; everything else is written by the kernel
    SMX_CUDA(cudaMemsetAsync(J, 0, sizeof(double) * (size_t)N * d.d_out * d.d_in, st));
    // block-sparse kernel: derivatives w.r.t. the cold columns (row sums) and, unless they are dense columns, the
    // derivative sets of the hot dimensions; dense kernel: the derivative sets as columns of Phi G
    int rc = SMX_OK;
    if ((!d.has_dense_grad || d.has_cold) && (rc = run_fast(d, x, N, ldx, J, 1, st))) return rc;
    if (d.has_dense_grad && (rc = run_dense(d, x, N, ldx, J, true, st))) return rc;
    if (nan_at_nodes) {
        const long long total = (long long)N * d.d_in;
        const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)d.sm_count * 16);
        nan_at_nodes_kernel<<<blocks, 256, 0, st>>>(x, N, ldx, d.d_in, d.d_out, d.nan_off, d.nan_nodes, J);
        SMX_LAUNCH_CHECK("nan_at_nodes_kernel");
    }
    return SMX_OK;
}

}  // namespace smx
