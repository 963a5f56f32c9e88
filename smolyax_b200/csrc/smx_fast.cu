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
    if (plan.has_dense && dense_kernel_fits(plan.n_tab, smem_optin) && plan.dense_coef.size() < (size_t)UINT32_MAX) {
        dev.dense_k4 = plan.dense_k4;
        if ((rc = upload(plan.dense_meta, &dev.dense_meta, dev.bytes, 2))) return rc;
        if ((rc = upload(plan.dense_eta0, &dev.dense_eta0, dev.bytes))) return rc;
        if (plan.has_dense) {
            if ((rc = upload(plan.dense_coef, &dev.dense_coef, dev.bytes))) return rc;
            if ((rc = upload(std::vector<int32_t>(256, 0), &dev.dense_tickets, dev.bytes))) return rc;
            dev.has_dense = true;
        }
    }
    // (else: the product table is too large for the GEMM-regime kernel: values run the block-sparse kernel, one output per pass)
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
        // 10^6 points: barrier kernels 2,1,0; pipelined kernel r06: 2,1,0: 1.684, 2,1,1: 1.651, 1,1,1: 1.659, 3,1,1: 1.694;
        // r08, after the bank-group layout made the k-steps cheaper (tuning build, benchmarks/k1_sweep.sh): 2,1,1: 1.614,
        // 3,1,1: 1.614, 3,1,1.5: 1.586, 2.5,1,2 / 3,1,2 / 3.5,1,2: 1.569, 3,1,2.5: 1.601, 4,1,2: 1.581, 6,1,3: 1.631)
        double ca = pipe ? 3.0 : 2.0, cb = 1.0, cc = pipe ? 2.0 : 0.0, cd = 0.0;  // (cd: per k-step and factor beyond the first)
        if (const char* env = tune_str("SMX_FAST_COST")) std::sscanf(env, "%lf,%lf,%lf,%lf", &ca, &cb, &cc, &cd);
        auto cost = [&](int32_t c) {
            const double ks = (double)((plan.chunk_off[c + 1] - plan.chunk_off[c] + 3) / 4);
            const double nf = (double)((plan.chunk_dir[(size_t)c * 4 + 2] >> 8) & 15);
            return ca + (cb + cd * std::max(0.0, nf - 1.0)) * ks + ((plan.chunk_flags[c] & kChunkHot) ? 0.0 : cc);
        };
        std::vector<int32_t> by_cost((size_t)plan.n_chunks);
        for (int32_t c = 0; c < plan.n_chunks; ++c) by_cost[c] = c;
        std::stable_sort(by_cost.begin(), by_cost.end(), [&](int32_t x1, int32_t x2) { return cost(x1) > cost(x2); });
        std::vector<std::vector<int32_t>> lists((size_t)nw);
        // Pipelined kernel: warp w runs on SM sub-partition w % 4 (profiles/smsp_map.cu) and the service warp is warp nw, so the
        // workers that share its sub-partition have one DMMA-issuing competitor fewer.  Giving them proportionally more work
        // (speed > 1) was measured and does not pay (r08, ms per 10^6 points at speed 0.8 / 0.9 / 1.0 / 1.15 / 1.25 / 1.5:
        // 1.650 / 1.601 / 1.614 / 1.678 / 1.716 / 1.937): the knob stays for tuning builds, the default is no bonus.
        std::vector<double> load((size_t)nw, 0.0), speed((size_t)nw, 1.0);
        if (pipe) {
            double bonus = 1.0;
            if (const char* env = tune_str("SMX_PIPE_BONUS")) bonus = std::atof(env);
            for (int w = 0; w < nw; ++w)
                if (w % 4 == nw % 4) speed[(size_t)w] = bonus;
        }
        for (int32_t c : by_cost) {
            size_t w = 0;
            for (size_t w2 = 1; w2 < (size_t)nw; ++w2)
                if ((load[w2] + cost(c)) / speed[w2] < (load[w] + cost(c)) / speed[w]) w = w2;
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
        } else if (dev.rest_warps > 0) {  // .. and so has the lean kernel that takes the outputs the three-set kernel leaves over
            build_directory(dev.rest_warps, dir, dev.pipe_warp_off, false);
            if ((rc = upload(dir, &dev.pipe_dir, dev.bytes, 4))) return rc;
        }
    }
    if ((rc = upload(plan.tab_factors, &dev.tab_factors, dev.bytes, 4))) return rc;
    dev.has_sparse = true;
    return SMX_OK;
}

void fast_free(FastDevice& d) {
    void* ptrs[] = {d.eta, d.tab_pairs, d.tab_factors, d.hot_off, d.hot_pos, d.chunk_dir, d.pipe_dir, d.chunk_meta, d.coef, d.c0,
                    d.nan_off, d.nan_nodes, d.dense_meta, d.dense_eta0, d.dense_coef, d.dense_tickets};
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

// TMA needs 16-byte aligned rows.  Anything else (odd pitch, odd base address) is first packed into an aligned scratch copy
// on the same stream; callers that care about the last few percent pass aligned rows.  *scratch is freed by the caller.
int aligned_rows(int64_t d_in, const double*& x, int64_t N, int64_t& ldx, double** scratch, cudaStream_t st) {
    *scratch = nullptr;
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (ldx & 1) != 0) {
        const int64_t pitch = (d_in + 1) & ~(int64_t)1;
        SMX_CUDA(cudaMallocAsync((void**)scratch, sizeof(double) * (size_t)N * pitch, st));
        SMX_CUDA(cudaMemcpy2DAsync(*scratch, sizeof(double) * pitch, x, sizeof(double) * ldx, sizeof(double) * d_in, (size_t)N,
                                   cudaMemcpyDeviceToDevice, st));
        x = *scratch;
        ldx = pitch;
    }
    return SMX_OK;
}

int run_fast(const FastDevice& d, const double* x, int64_t N, int64_t ldx, double* out, cudaStream_t st) {
    double* scratch = nullptr;
    int rca = aligned_rows(d.d_in, x, N, ldx, &scratch, st);
    if (rca) return rca;
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
    a.o_begin = 0, a.o_end = (int)d.d_out;
    a.num_tiles = (N + kTile - 1) / kTile;
    a.gradient = 0;
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
    return run_fast(d, x, N, ldx, y, st);
}

namespace {

int run_dense(const FastDevice& d, const double* x, int64_t N, int64_t ldx, double* out, cudaStream_t st) {
    DenseArgs a;
    a.eta = d.eta;
    a.tab_pairs = reinterpret_cast<const int2*>(d.tab_pairs);
    a.hot_off = d.hot_off;
    a.hot_pos = d.hot_pos;
    a.meta = reinterpret_cast<const int2*>(d.dense_meta);
    a.eta0 = d.dense_eta0;
    a.coef = d.dense_coef;
    a.c0 = d.c0;
    a.colmap = nullptr;
    a.N = N;
    a.ldx = ldx;
    a.ncol = d.d_out;
    a.ldy = d.d_out;
    a.k4 = d.dense_k4;
    a.nblk = (int)((a.ncol + 7) / 8);
    a.n_tab = d.n_tab;
    a.n_hot_rows = d.n_hot_rows;
    a.n_levels = d.n_levels;
    a.hot_dims = d.hot_dims;
    a.skew = 0;  // (dense_kernel_launch decides)
    a.tickets = d.dense_tickets;
    for (int l = 0; l < kMaxLevels + 2; ++l) a.level_off[l] = d.level_off[l];
    return dense_kernel_launch(a, x, out, st);
}

}  // namespace

int dense_eval(const FastDevice& d, const double* x, int64_t N, int64_t ldx, double* y, cudaStream_t st) {
    if (N == 0) return SMX_OK;
    if (!d.has_dense) return fail(SMX_ERR_UNSUPPORTED, "plan has no dense form");
    return run_dense(d, x, N, ldx, y, st);
}

int fast_gradient(const FastDevice& d, const GradDevice& g, const double* x, int64_t N, int64_t ldx, double* J, bool nan_at_nodes, cudaStream_t st) {
    if (N == 0) return SMX_OK;
    if (!g.present) return fail(SMX_ERR_UNSUPPORTED, "plan has no gradient jobs");
    double* scratch = nullptr;
    int rc = aligned_rows(d.d_in, x, N, ldx, &scratch, st);
    if (rc) return rc;
    rc = grad_kernel_launch(d, g, x, N, ldx, J, nan_at_nodes, st);  // writes every entry of J exactly once
    if (scratch) cudaFreeAsync(scratch, st);
    return rc;
}

// Gradient jobs of the plan -> device: records (metadata + coefficients packed as DMMA B fragments, one per item and output),
// the jobs assigned to the warps of a CTA by longest processing time, the directory in the order the warps walk it.
int grad_upload(const FastPlan& plan, const FastDevice& dev, GradDevice& g) {
    g = GradDevice();
    const GradPlan& G = plan.grad;
    if (!G.present || !plan.flat_ok || !dev.has_sparse) return SMX_OK;  // (no gradient kernel for eight-factor records yet)
    const int32_t n_items = (int32_t)G.item_off.size() - 1, n_jobs = (int32_t)G.job_kind.size();
    const size_t dout = (size_t)plan.d_out;
    g.n_items = n_items, g.n_jobs = n_jobs, g.n_zero = (int32_t)G.zero_cols.size() / 2;
    int device = 0, smem_sm = 0;
    SMX_CUDA(cudaGetDevice(&device));
    SMX_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
    g.warps = grad_kernel_warps(g, dev, smem_sm);
    if (g.warps == 0) return SMX_OK;
    // records
    std::vector<int32_t> rec_off((size_t)n_items), units((size_t)n_items);
    size_t total = 0;
    for (int32_t it = 0; it < n_items; ++it) {
        const size_t ks = (size_t)((G.item_off[it + 1] - G.item_off[it] + 3) / 4);
        rec_off[it] = (int32_t)total, units[it] = (int32_t)(5 + 4 * ks);
        total += dout * (5 + 4 * ks);
        if (total >= (size_t)INT32_MAX) return SMX_OK;  // too many outputs for one set of gradient tables: the host layer splits them
    }
    std::vector<double> records(total * 16, 0.0);
    std::vector<int32_t> meta(kMetaInts);
    for (int32_t it = 0; it < n_items; ++it) {
        const size_t r0 = (size_t)G.item_off[it], rows = (size_t)G.item_off[it + 1] - r0, ks = (rows + 3) / 4;
        std::copy_n(&G.item_meta[(size_t)it * kMetaInts], kMetaInts, meta.begin());
        for (int i = 0; i < 16; ++i) meta[i] *= kTabPitch, meta[48 + i] *= kTabPitch;
        for (int i = 96; i < 160; ++i) meta[i] *= kTabPitch;
        for (size_t o = 0; o < dout; ++o) {
            double* rec = &records[((size_t)rec_off[it] + o * units[it]) * 16];
            std::memcpy(rec, meta.data(), kMetaInts * 4);
            double* packed = rec + 5 * 16;
            for (size_t s2 = 0; s2 < ks; ++s2)
                for (int lane = 0; lane < 32; ++lane)
                    for (int j = 0; j < 2; ++j) {
                        const size_t row = 4 * s2 + (lane & 3);
                        const int gid = lane >> 2, entry = 4 * (gid >> 1) + 2 * j + (gid & 1);
                        if (row < rows) packed[s2 * kKStepDoubles + 2 * lane + j] = G.coef[((r0 + row) * dout + o) * kBlockWidth + entry];
                    }
        }
    }
    int rc;
    if ((rc = upload(records, &g.records, g.bytes, 2))) return rc;
    std::vector<double>().swap(records);
    // jobs -> warps (longest processing time first); cost of an item as in the value path
    auto item_cost = [&](int32_t it) { return 2.0 + (double)((G.item_off[it + 1] - G.item_off[it] + 3) / 4); };
    std::vector<double> job_cost((size_t)n_jobs, 0.0);
    for (int32_t jb = 0; jb < n_jobs; ++jb)
        for (int32_t it = G.job_off[jb]; it < G.job_off[jb + 1]; ++it) job_cost[jb] += item_cost(it);
    std::vector<int32_t> by_cost((size_t)n_jobs);
    for (int32_t jb = 0; jb < n_jobs; ++jb) by_cost[jb] = jb;
    std::stable_sort(by_cost.begin(), by_cost.end(), [&](int32_t a1, int32_t a2) { return job_cost[a1] > job_cost[a2]; });
    std::vector<std::vector<int32_t>> lists((size_t)g.warps);
    std::vector<double> load((size_t)g.warps, 0.0);
    for (int32_t jb : by_cost) {
        const size_t w = std::min_element(load.begin(), load.end()) - load.begin();
        lists[w].push_back(jb);
        load[w] += job_cost[jb];
    }
    std::vector<int32_t> dir((size_t)n_items * 4), jobs((size_t)n_jobs * 4, 0);
    size_t pos = 0;
    int32_t cold_ordinal = 0;
    std::vector<int32_t> node_off((size_t)n_jobs, 0);
    for (int32_t jb = 0; jb < n_jobs; ++jb)
        if (G.job_kind[jb] == 0) node_off[jb] = 32 * cold_ordinal++;
    for (int w = 0; w < g.warps; ++w) {
        g.warp_off[w] = (int32_t)pos;
        for (int32_t jb : lists[w]) {
            const bool cold_job = G.job_kind[jb] == 0;
            for (int32_t it = G.job_off[jb]; it < G.job_off[jb + 1]; ++it) {
                const int32_t* d4 = &G.item_dir[(size_t)it * 4];
                const int32_t ks = (d4[1] + 3) / 4, nf = (d4[2] >> 8) & 15;
                const bool hot = d4[2] & kChunkHot, first = it == G.job_off[jb], last = it + 1 == G.job_off[jb + 1];
                int32_t flags = (hot ? 1 : 0) | (first ? 4 : 0) | (last ? 8 : 0) | ((d4[2] & kChunkEtaZero) ? 16 : 0) | (cold_job ? 32 : 0);
                if (cold_job ? last : !hot) flags |= 2;  // x tile: the cold job's node test; the leading basis values of a cold block
                dir[pos * 4 + 0] = rec_off[it];
                dir[pos * 4 + 1] = jb;
                dir[pos * 4 + 2] = flags | (nf << 8) | (units[it] << 16) | (ks << 24);
                dir[pos * 4 + 3] = d4[3];
                if (last && cold_job) jobs[(size_t)jb * 4 + 2] = d4[3];
                ++pos;
            }
            jobs[(size_t)jb * 4 + 0] = G.job_target[jb];
            jobs[(size_t)jb * 4 + 1] = node_off[jb];
            jobs[(size_t)jb * 4 + 3] = G.job_kind[jb];
        }
    }
    for (int w = g.warps; w <= kMaxWarps; ++w) g.warp_off[w] = (int32_t)pos;
    if ((rc = upload(dir, &g.dir, g.bytes, 4))) return rc;
    if ((rc = upload(jobs, &g.jobs, g.bytes, 4))) return rc;
    if ((rc = upload(G.job_c0, &g.job_c0, g.bytes))) return rc;
    if ((rc = upload(G.job_nodes, &g.job_nodes, g.bytes))) return rc;
    if ((rc = upload(G.zero_cols, &g.zero_cols, g.bytes, 2))) return rc;
    g.present = true;
    return SMX_OK;
}

void grad_free(GradDevice& g) {
    void* ptrs[] = {g.records, g.dir, g.jobs, g.job_c0, g.job_nodes, g.zero_cols};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    g = GradDevice();
}

}  // namespace smx
