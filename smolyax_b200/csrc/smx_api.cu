// C-ABI of include/smolyax_b200.h: handle life cycle, dispatch to the kernels, host-buffer pipeline.
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>

#include "smx_common.cuh"

namespace smx {

static thread_local std::string t_error;
std::atomic<int64_t> g_launches{0};
thread_local char t_last_kernel[96] = "";

void set_error(const std::string& msg) { t_error = msg; }
int fail(int status, const std::string& msg) {
    t_error = msg;
    return status;
}
int cuda_fail(cudaError_t e, const char* what) {
    t_error = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();  // clear the sticky-less error state
    return e == cudaErrorMemoryAllocation ? SMX_ERR_OUT_OF_MEMORY : SMX_ERR_CUDA;
}

}  // namespace smx

using namespace smx;

struct smx_interp {
    int device = 0;
    int64_t d_in = 0, d_out = 0;
    smx_info info{};
    bool has_fast = false, has_groups = false, grad_finite = false, compact = false;
    bool no_groups_by_depth = false;  // reference layout not uploaded: a summand has more active dimensions than the per-summand kernels take
    double* d_integral = nullptr;  // compact handles: the integral, computed at create time
    FastDevice fast;
    GradDevice grad;  // gradient jobs of the fast path (present when the plan could build them)
    std::vector<SeamGroup> groups;
    std::vector<void*> owned;  // device allocations behind `groups`
    double* d_offset = nullptr;
    double* integral_ws = nullptr;
    int64_t integral_ws_doubles = 0;
    // host pipeline (lazy)
    static constexpr int kStages = 3;
    cudaStream_t streams[kStages] = {nullptr, nullptr, nullptr};
    double* stage_x[kStages] = {nullptr, nullptr, nullptr};
    double* stage_y[kStages] = {nullptr, nullptr, nullptr};
    int64_t stage_points = 0;
    std::mutex host_mutex;
};

namespace {

// The entry points that switch the CUDA device (create, destroy, the host pipeline) put the caller's device back on return.
struct DeviceGuard {
    int saved = -1;
    DeviceGuard() {
        if (cudaGetDevice(&saved) != cudaSuccess) {
            saved = -1;
            cudaGetLastError();
        }
    }
    ~DeviceGuard() {
        if (saved >= 0) cudaSetDevice(saved);
    }
};

template <class T>
int upload_array(smx_interp* h, const T* host, size_t count, const T** out) {
    *out = nullptr;
    if (count == 0 || host == nullptr) return SMX_OK;
    void* d = nullptr;
    SMX_CUDA(cudaMalloc(&d, count * sizeof(T)));
    h->owned.push_back(d);
    h->info.device_bytes += (int64_t)(count * sizeof(T));
    SMX_CUDA(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *out = static_cast<const T*>(d);
    return SMX_OK;
}

int check_device(int device, int* resolved) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(SMX_ERR_NO_DEVICE, "no CUDA device visible: smolyax_b200 has no CPU fallback");
    }
    if (device < 0) SMX_CUDA(cudaGetDevice(&device));
    if (device >= count) return fail(SMX_ERR_INVALID_ARG, "device ordinal out of range");
    SMX_CUDA(cudaSetDevice(device));
    int major = 0;
    SMX_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) return fail(SMX_ERR_NO_DEVICE, "device is not sm_100 (B200): kernels are built for sm_100a only");
    *resolved = device;
    return SMX_OK;
}

void release(smx_interp* h) {
    if (!h) return;
    DeviceGuard guard;
    cudaSetDevice(h->device);
    fast_free(h->fast);
    grad_free(h->grad);
    for (void* p : h->owned) cudaFree(p);
    if (h->d_offset) cudaFree(h->d_offset);
    if (h->d_integral) cudaFree(h->d_integral);
    if (h->integral_ws) cudaFree(h->integral_ws);
    for (int i = 0; i < smx_interp::kStages; ++i) {
        if (h->stage_x[i]) cudaFree(h->stage_x[i]);
        if (h->stage_y[i]) cudaFree(h->stage_y[i]);
        if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
    }
    delete h;
}

int ensure_stages(smx_interp* h, int64_t points, int64_t ldx) {
    if (h->stage_points >= points && h->stage_x[0]) return SMX_OK;
    for (int i = 0; i < smx_interp::kStages; ++i) {
        if (h->stage_x[i]) cudaFree(h->stage_x[i]);
        if (h->stage_y[i]) cudaFree(h->stage_y[i]);
        h->stage_x[i] = h->stage_y[i] = nullptr;
        if (!h->streams[i]) SMX_CUDA(cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking));
        SMX_CUDA(cudaMalloc((void**)&h->stage_x[i], (size_t)points * h->d_in * sizeof(double)));
        SMX_CUDA(cudaMalloc((void**)&h->stage_y[i], (size_t)points * h->d_out * sizeof(double)));
    }
    (void)ldx;
    h->stage_points = points;
    return SMX_OK;
}

PlanOptions plan_options(uint32_t flags, int64_t d_out, bool sparse_wanted) {
    PlanOptions opt;
    static const int dense_min = tune_int("SMX_DENSE_MIN", 32);
    opt.dense = (flags & SMX_DENSE_PATH) || (!(flags & SMX_NO_DENSE_PATH) && d_out >= dense_min);
    if (!opt.dense && !(flags & SMX_NO_DENSE_PATH) && sparse_wanted && d_out >= 8 && d_out < dense_min)
        opt.dense = opt.dense_if_cheaper = true;  // decided by the plan compiler from the index set (smx_plan.h)
    // the block-sparse form stores every coefficient set padded to 16-entry blocks: beyond a few thousand outputs it
    // only costs memory once the dense form exists (it would still serve the gradient)
    opt.sparse = sparse_wanted && !(opt.dense && d_out > 2048);
    opt.gradient = opt.sparse;
    return opt;
}

// plan -> device tables; fills the statistics of the handle
int adopt_plan(smx_interp* h, const FastPlan& plan) {
    h->info.n_summands = plan.n_summands;
    h->info.w_raw = plan.w_raw;
    h->info.w_pad = plan.w_pad;
    h->info.n_terms = plan.n_terms;
    h->info.n_entries = plan.n_entries;
    h->info.n_rows = plan.n_rows;
    h->info.n_chunks = plan.n_chunks;
    h->info.padded_fma = plan.padded_fma;
    h->info.nested = plan.nested ? 1 : 0;
    const int rc = fast_upload(plan, h->fast);
    if (rc) return rc;
    h->has_fast = true;
    h->info.device_bytes += h->fast.bytes;
    h->info.has_dense_path = h->fast.has_dense;
    h->info.dense_terms = 4ll * h->fast.dense_k4;
    const int rcg = grad_upload(plan, h->fast, h->grad);
    if (rcg) return rcg;
    h->info.device_bytes += h->grad.bytes;
    h->info.grad_jobs = h->grad.present ? h->grad.n_jobs : 0;
    h->info.grad_items = h->grad.present ? h->grad.n_items : 0;
    return SMX_OK;
}

int upload_offset(smx_interp* h, const double* offset) {
    std::vector<double> off((size_t)h->d_out, 0.0);
    if (offset) std::memcpy(off.data(), offset, sizeof(double) * h->d_out);
    SMX_CUDA(cudaMalloc((void**)&h->d_offset, sizeof(double) * h->d_out));
    SMX_CUDA(cudaMemcpy(h->d_offset, off.data(), sizeof(double) * h->d_out, cudaMemcpyHostToDevice));
    h->info.device_bytes += (int64_t)sizeof(double) * h->d_out;
    return SMX_OK;
}

}  // namespace

extern "C" {

int smx_create_compact(int64_t d_in, int64_t d_out, const double* offset, const smx_compact_desc* desc, uint32_t flags,
                       int device, smx_interp** out) {
    if (!desc || !out) return fail(SMX_ERR_INVALID_ARG, "smx_create_compact: null argument");
    *out = nullptr;
    if (d_in <= 0 || d_out <= 0) return fail(SMX_ERR_INVALID_ARG, "smx_create_compact: d_in and d_out must be positive");
    int dev = 0, rc;
    DeviceGuard guard;
    if ((rc = check_device(device, &dev))) return rc;
    std::unique_ptr<smx_interp, void (*)(smx_interp*)> h(new smx_interp(), release);
    h->device = dev;
    h->d_in = h->info.d_in = d_in;
    h->d_out = h->info.d_out = d_out;
    h->compact = true;
    h->grad_finite = (flags & SMX_GRAD_FINITE_AT_NODES) != 0;
    CompactView cv;
    cv.n_summands = desc->n_summands;
    cv.n_active = desc->n_active;
    cv.slot_off = desc->slot_off;
    cv.dims = desc->dims;
    cv.degs = desc->degs;
    cv.node_off = desc->node_off;
    cv.node_pool = desc->node_pool;
    cv.quad_pool = desc->quad_pool;
    cv.zetas = desc->zetas;
    cv.val_off = desc->val_off;
    cv.val_index = desc->val_index;
    cv.values = desc->values;
    cv.n_values = desc->n_values;
    {
        FastPlan plan;
        const std::string err = build_fast_plan_compact(d_in, d_out, offset, cv, plan, plan_options(flags, d_out, true));
        if (err.rfind("ill-conditioned", 0) == 0)  // (no per-summand kernels behind a compact handle: the caller uses smx_create)
            return fail(SMX_ERR_UNSUPPORTED, "smx_create_compact: " + err + "; use the reference layout (smx_create), which falls back to the per-summand kernels");
        if (!err.empty()) return fail(SMX_ERR_INVALID_ARG, "smx_create_compact: " + err);
        if ((rc = adopt_plan(h.get(), plan))) return rc;
    }
    if (desc->quad_pool || desc->n_summands == 0) {
        std::vector<double> Q;
        const std::string err = integrate_compact(d_out, offset, cv, Q);
        if (!err.empty()) return fail(SMX_ERR_INVALID_ARG, "smx_create_compact: " + err);
        SMX_CUDA(cudaMalloc((void**)&h->d_integral, sizeof(double) * d_out));
        SMX_CUDA(cudaMemcpy(h->d_integral, Q.data(), sizeof(double) * d_out, cudaMemcpyHostToDevice));
        h->info.device_bytes += (int64_t)sizeof(double) * d_out;
    }
    if ((rc = upload_offset(h.get(), offset))) return rc;
    h->info.has_fast_path = 1;
    *out = h.release();
    return SMX_OK;
}

int smx_create(const smx_interp_desc* desc, int device, smx_interp** out) {
    if (!desc || !out) return fail(SMX_ERR_INVALID_ARG, "smx_create: null argument");
    *out = nullptr;
    if (desc->d_in <= 0 || desc->d_out <= 0) return fail(SMX_ERR_INVALID_ARG, "smx_create: d_in and d_out must be positive");
    if (desc->n_groups < 0 || (desc->n_groups > 0 && !desc->groups)) return fail(SMX_ERR_INVALID_ARG, "smx_create: bad group list");
    int dev = 0, rc;
    DeviceGuard guard;
    if ((rc = check_device(device, &dev))) return rc;

    std::unique_ptr<smx_interp, void (*)(smx_interp*)> h(new smx_interp(), release);
    h->device = dev;
    h->d_in = desc->d_in;
    h->d_out = desc->d_out;
    h->info.d_in = desc->d_in;
    h->info.d_out = desc->d_out;

    std::vector<GroupView> views;
    for (int32_t g = 0; g < desc->n_groups; ++g) {
        const smx_group_desc& d = desc->groups[g];
        if (d.n < 1 || !d.tau || d.nn < 0) return fail(SMX_ERR_INVALID_ARG, "smx_create: malformed group descriptor");
        GroupView v;
        v.n = d.n;
        v.nn = d.nn;
        v.tau.assign(d.tau, d.tau + d.n);
        v.F = d.F;
        v.nodes = d.nodes;
        v.weights = d.weights;
        v.dims = d.dims;
        v.degs = d.degs;
        v.zetas = d.zetas;
        v.quad = d.quad;
        views.push_back(v);
    }

    h->grad_finite = (desc->flags & SMX_GRAD_FINITE_AT_NODES) != 0;
    bool want_fast = !(desc->flags & SMX_NO_FAST_PATH);
    bool want_groups = (desc->flags & SMX_KEEP_GROUPS) != 0 || !want_fast;
    if (want_fast) {
        FastPlan plan;
        const std::string err = build_fast_plan(desc->d_in, desc->d_out, desc->offset, views, plan, plan_options(desc->flags, desc->d_out, true));
        const bool ill = err.rfind("ill-conditioned", 0) == 0;  // high-degree non-nested rule: per-summand kernels (reference arithmetic)
        if (!err.empty() && !ill) return fail(SMX_ERR_INVALID_ARG, "smx_create: " + err);
        rc = ill ? (int)SMX_ERR_UNSUPPORTED : adopt_plan(h.get(), plan);
        if (ill) set_error("smx_create: " + err);
        if (rc == SMX_ERR_UNSUPPORTED) {
            fast_free(h->fast);  // fall back to the per-summand kernels; still a CUDA path
            grad_free(h->grad);
            want_groups = true;
        } else if (rc) {
            return rc;
        }
    }
    // Summands with more active dimensions than the per-summand kernels take (8): with a fast path the handle does without
    // the reference layout on the device - values and gradients come from the fast path, the integral (which does not
    // depend on x) is computed here in extended precision, as for compact handles.
    bool deep_groups = false;
    for (const GroupView& v : views) deep_groups = deep_groups || v.n > kSeamMaxN;
    if (want_groups && deep_groups) {
        if (!h->has_fast) return fail(SMX_ERR_UNSUPPORTED, "smx_create: more than 8 active dimensions per summand and no fast path for this layout");
        want_groups = false;
        bool quad = !views.empty();
        for (const GroupView& v : views) quad = quad && v.quad != nullptr;
        if (quad) {
            std::vector<double> Q;
            const std::string err = integrate_groups(desc->d_out, desc->offset, views, Q);
            if (!err.empty()) return fail(SMX_ERR_INVALID_ARG, "smx_create: " + err);
            SMX_CUDA(cudaMalloc((void**)&h->d_integral, sizeof(double) * desc->d_out));
            SMX_CUDA(cudaMemcpy(h->d_integral, Q.data(), sizeof(double) * desc->d_out, cudaMemcpyHostToDevice));
            h->info.device_bytes += (int64_t)sizeof(double) * desc->d_out;
        }
        h->no_groups_by_depth = true;
    }
    if (want_groups) {
        for (size_t gi = 0; gi < views.size(); ++gi) {
            const smx_group_desc& d = desc->groups[gi];
            const GroupView& v = views[gi];
            smx_group_desc dd = d;
            const size_t slots = (size_t)d.nn * d.n, tw = (size_t)v.tw();
            if ((rc = upload_array(h.get(), d.F, (size_t)d.nn * desc->d_out * v.fsize(), &dd.F))) return rc;
            if ((rc = upload_array(h.get(), d.nodes, slots * tw, &dd.nodes))) return rc;
            if ((rc = upload_array(h.get(), d.weights, slots * tw, &dd.weights))) return rc;
            if ((rc = upload_array(h.get(), d.dims, slots, &dd.dims))) return rc;
            if ((rc = upload_array(h.get(), d.degs, slots, &dd.degs))) return rc;
            if ((rc = upload_array(h.get(), d.zetas, (size_t)d.nn, &dd.zetas))) return rc;
            if ((rc = upload_array(h.get(), d.quad, slots * tw, &dd.quad))) return rc;
            SeamGroup sg;
            if ((rc = make_seam_group(&dd, sg))) return rc;
            h->groups.push_back(sg);
            h->integral_ws_doubles = std::max(h->integral_ws_doubles, seam_integral_workspace(sg, desc->d_out));
            if (!want_fast) {
                h->info.n_summands += d.nn;
                h->info.w_pad += d.nn * v.fsize();
            }
        }
        h->has_groups = true;
    }
    if ((rc = upload_offset(h.get(), desc->offset))) return rc;
    h->info.has_fast_path = h->has_fast;
    h->info.has_groups = h->has_groups;
    *out = h.release();
    return SMX_OK;
}

int smx_destroy(smx_interp* h) {
    release(h);
    return SMX_OK;
}

int smx_get_info(const smx_interp* h, smx_info* info) {
    if (!h || !info) return fail(SMX_ERR_INVALID_ARG, "smx_get_info: null argument");
    *info = h->info;
    return SMX_OK;
}

int smx_eval(smx_interp* h, const double* x, int64_t N, int64_t ldx, double* y, void* stream) {
    if (!h || N < 0) return fail(SMX_ERR_INVALID_ARG, "smx_eval: bad arguments");
    if (N == 0) return SMX_OK;
    if (!x || !y || ldx < h->d_in) return fail(SMX_ERR_INVALID_ARG, "smx_eval: null buffer or ldx < d_in");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (h->has_fast && h->fast.has_dense) return dense_eval(h->fast, x, N, ldx, y, st);
    if (h->has_fast && h->fast.has_sparse) return fast_eval(h->fast, x, N, ldx, y, st);
    int rc;
    if ((rc = fill_rows(y, N, h->d_out, h->d_offset, st))) return rc;
    for (const SeamGroup& g : h->groups)
        if ((rc = seam_eval(x, N, ldx, g, h->d_out, y, st))) return rc;
    return SMX_OK;
}

int smx_gradient(smx_interp* h, const double* x, int64_t N, int64_t ldx, double* J, void* stream) {
    if (!h || N < 0) return fail(SMX_ERR_INVALID_ARG, "smx_gradient: bad arguments");
    if (N == 0) return SMX_OK;
    if (!x || !J || ldx < h->d_in) return fail(SMX_ERR_INVALID_ARG, "smx_gradient: null buffer or ldx < d_in");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (h->has_fast && h->grad.present) return fast_gradient(h->fast, h->grad, x, N, ldx, J, !h->grad_finite, st);
    if (h->compact)
        return fail(SMX_ERR_UNSUPPORTED, "smx_gradient: the derivative coefficient sets of this handle were not built (d_out too large)");
    if (h->no_groups_by_depth)
        return fail(SMX_ERR_UNSUPPORTED, "smx_gradient: no derivative sets on this handle and more than 8 active dimensions per summand");
    if (!h->has_groups && h->info.n_summands > 0)
        return fail(SMX_ERR_INVALID_ARG, "smx_gradient: no derivative sets and no reference layout (SMX_KEEP_GROUPS) on this handle");
    SMX_CUDA(cudaMemsetAsync(J, 0, sizeof(double) * (size_t)N * h->d_out * h->d_in, st));
    int rc;
    for (const SeamGroup& g : h->groups)
        if ((rc = seam_gradient(x, N, ldx, h->d_in, g, h->d_out, J, st))) return rc;
    return SMX_OK;
}

int smx_integral(smx_interp* h, double* q, void* stream) {
    if (!h || !q) return fail(SMX_ERR_INVALID_ARG, "smx_integral: null argument");
    if (h->compact || h->no_groups_by_depth) {
        if (!h->d_integral) return fail(SMX_ERR_INVALID_ARG, "smx_integral: handle was created without quadrature weights");
        SMX_CUDA(cudaMemcpyAsync(q, h->d_integral, sizeof(double) * h->d_out, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
        return SMX_OK;
    }
    if (!h->has_groups && h->info.n_summands > 0)
        return fail(SMX_ERR_INVALID_ARG, "smx_integral: handle was created without SMX_KEEP_GROUPS");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    if ((rc = fill_rows(q, 1, h->d_out, h->d_offset, st))) return rc;
    if (h->integral_ws_doubles > 0 && !h->integral_ws) {
        SMX_CUDA(cudaMalloc((void**)&h->integral_ws, sizeof(double) * h->integral_ws_doubles));
        h->info.device_bytes += (int64_t)sizeof(double) * h->integral_ws_doubles;
    }
    for (const SeamGroup& g : h->groups)
        if ((rc = seam_integral(g, h->d_out, q, h->integral_ws, h->integral_ws_doubles, st))) return rc;
    return SMX_OK;
}

static int64_t default_chunk_points(const smx_interp* h);

int smx_prepare(smx_interp* h, int64_t n_points) {
    if (!h || n_points < 0) return fail(SMX_ERR_INVALID_ARG, "smx_prepare: bad arguments");
    if (n_points == 0) return SMX_OK;
    std::lock_guard<std::mutex> lock(h->host_mutex);
    DeviceGuard guard;
    SMX_CUDA(cudaSetDevice(h->device));
    return ensure_stages(h, std::min(default_chunk_points(h), n_points), h->d_in);
}

int smx_eval_host(smx_interp* h, const double* x_host, int64_t N, int64_t ldx, double* y_host, int64_t chunk_points) {
    if (!h || N < 0) return fail(SMX_ERR_INVALID_ARG, "smx_eval_host: bad arguments");
    if (N == 0) return SMX_OK;
    if (!x_host || !y_host || ldx < h->d_in) return fail(SMX_ERR_INVALID_ARG, "smx_eval_host: null buffer or ldx < d_in");
    std::lock_guard<std::mutex> lock(h->host_mutex);
    DeviceGuard guard;
    SMX_CUDA(cudaSetDevice(h->device));
    if (chunk_points <= 0) chunk_points = default_chunk_points(h);
    chunk_points = std::min(chunk_points, N);
    int rc;
    if ((rc = ensure_stages(h, chunk_points, ldx))) return rc;
    // on any failure: wait for the copies already queued (they read and write the caller's buffers) before returning
    auto drain = [&](int status) {
        for (int s = 0; s < smx_interp::kStages; ++s)
            if (h->streams[s]) cudaStreamSynchronize(h->streams[s]);
        return status;
    };
    int64_t done = 0;
    for (int64_t c = 0; done < N; ++c, done += chunk_points) {
        const int s = (int)(c % smx_interp::kStages);
        const int64_t n = std::min(chunk_points, N - done);
        cudaStream_t st = h->streams[s];
        // stream order protects the stage buffers: the next use of stage s is queued behind this one
        cudaError_t e;
        if (ldx == h->d_in)  // contiguous rows: one linear copy
            e = cudaMemcpyAsync(h->stage_x[s], x_host + done * ldx, sizeof(double) * (size_t)n * h->d_in, cudaMemcpyHostToDevice, st);
        else
            e = cudaMemcpy2DAsync(h->stage_x[s], sizeof(double) * h->d_in, x_host + done * ldx, sizeof(double) * ldx,
                                  sizeof(double) * h->d_in, (size_t)n, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return drain(cuda_fail(e, "smx_eval_host: copy of x to the device"));
        if ((rc = smx_eval(h, h->stage_x[s], n, h->d_in, h->stage_y[s], st))) return drain(rc);
        e = cudaMemcpyAsync(y_host + done * h->d_out, h->stage_y[s], sizeof(double) * (size_t)n * h->d_out, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return drain(cuda_fail(e, "smx_eval_host: copy of y to the host"));
    }
    for (int s = 0; s < smx_interp::kStages; ++s) {
        const cudaError_t e = cudaStreamSynchronize(h->streams[s]);
        if (e != cudaSuccess) return drain(cuda_fail(e, "smx_eval_host: cudaStreamSynchronize"));
    }
    return SMX_OK;
}

int smx_group_eval(const double* x, int64_t N, int64_t ldx, int64_t d_in, const smx_group_desc* g, int64_t d_out,
                   double* y, int accumulate, void* stream) {
    if (N < 0 || d_out <= 0 || ldx < d_in) return fail(SMX_ERR_INVALID_ARG, "smx_group_eval: bad sizes");
    if (N == 0) return SMX_OK;
    if (!x || !y) return fail(SMX_ERR_INVALID_ARG, "smx_group_eval: null buffer");
    SeamGroup sg;
    int rc;
    if ((rc = make_seam_group(g, sg))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!accumulate && (rc = fill_rows(y, N, d_out, nullptr, st))) return rc;
    return seam_eval(x, N, ldx, sg, d_out, y, st);
}

int smx_group_gradient(const double* x, int64_t N, int64_t ldx, int64_t d_in, const smx_group_desc* g, int64_t d_out,
                       double* J, int accumulate, void* stream) {
    if (N < 0 || d_out <= 0 || ldx < d_in) return fail(SMX_ERR_INVALID_ARG, "smx_group_gradient: bad sizes");
    if (N == 0) return SMX_OK;
    if (!x || !J) return fail(SMX_ERR_INVALID_ARG, "smx_group_gradient: null buffer");
    SeamGroup sg;
    int rc;
    if ((rc = make_seam_group(g, sg))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!accumulate) SMX_CUDA(cudaMemsetAsync(J, 0, sizeof(double) * (size_t)N * d_out * d_in, st));
    return seam_gradient(x, N, ldx, d_in, sg, d_out, J, st);
}

int smx_group_integral(const smx_group_desc* g, int64_t d_out, double* q, int accumulate, void* stream) {
    if (d_out <= 0 || !q) return fail(SMX_ERR_INVALID_ARG, "smx_group_integral: bad arguments");
    SeamGroup sg;
    int rc;
    if ((rc = make_seam_group(g, sg))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!accumulate && (rc = fill_rows(q, 1, d_out, nullptr, st))) return rc;
    const int64_t ws = seam_integral_workspace(sg, d_out);
    double* work = nullptr;
    SMX_CUDA(cudaMallocAsync((void**)&work, sizeof(double) * ws, st));
    rc = seam_integral(sg, d_out, q, work, ws, st);
    cudaFreeAsync(work, st);
    return rc;
}

int smx_compute_weights(const double* nodes, int64_t m, double* w, void* stream) {
    if (m < 0 || (m > 0 && (!nodes || !w))) return fail(SMX_ERR_INVALID_ARG, "smx_compute_weights: bad arguments");
    return device_compute_weights(nodes, m, w, static_cast<cudaStream_t>(stream));
}

int smx_basis(const double* x, int64_t N, const double* xi, const double* w, int64_t m, int64_t nu, int derivative,
              double* out, void* stream) {
    if (N < 0 || m < 0 || (N > 0 && m > 0 && (!x || !xi || !w || !out)))
        return fail(SMX_ERR_INVALID_ARG, "smx_basis: bad arguments");
    return device_basis(x, N, xi, w, m, nu, derivative, out, static_cast<cudaStream_t>(stream));
}

int64_t smx_launch_count(void) { return g_launches.load(); }
const char* smx_last_kernel(void) { return t_last_kernel; }
const char* smx_last_error(void) { return t_error.c_str(); }
int smx_version(void) { return 200; }
const char* smx_arch(void) { return "sm_100a"; }
#ifndef SMX_BUILD_STAMP
#define SMX_BUILD_STAMP "unstamped build"
#endif
const char* smx_build_info(void) {
#ifdef SMX_TUNING
    return SMX_BUILD_STAMP "; SMX_TUNING";
#else
    return SMX_BUILD_STAMP;
#endif
}

}  // extern "C"

// ~32 MiB per stage keeps the copy engines and the SMs busy at the same time (measured 8 .. 1024 MiB: 149 .. 175 ms per
// 8 GB at cfg2, i.e. the PCIe link - 54 GB/s - whatever the chunk), rounded up to whole waves of 32-point tiles (two CTAs
// per SM; the GEMM-regime kernel runs ceil(d_out / 128 .. 512) CTAs per tile): a chunk that fills 44 % of the CTA slots
// takes as long as a full one.  It does not matter while the link is the bound (cfg2), it does when the kernel is (cfg5,
// 32 MiB = 4 194 points = 131 of 296 slots: 467 ms per 10^6 points end to end against 78 ms of kernel time).
static int64_t default_chunk_points(const smx_interp* h) {
    const long long chunk_mb = tune_int("SMX_HOST_CHUNK_MB", 32);
    int64_t chunk_points = std::max<int64_t>(1024, (chunk_mb << 20) / (int64_t)(std::max(h->d_in, h->d_out) * sizeof(double)));
    if (h->has_fast && h->fast.sm_count > 0) {
        const int64_t per_tile = h->d_out >= 384 ? (h->d_out + 511) / 512 : (h->d_out + 127) / 128;
        const int64_t wave = std::max<int64_t>(1, ((int64_t)h->fast.sm_count * 2 + per_tile - 1) / per_tile) * 32;
        chunk_points = (chunk_points + wave - 1) / wave * wave;
    }
    return chunk_points;
}
