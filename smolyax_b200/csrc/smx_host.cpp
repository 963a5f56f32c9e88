// Host-side combinatorics for the Smolyak operator: the multi-index set Lambda(k,t), the
// Smolyak coefficients zeta, node-set cardinalities.  No CUDA in this translation unit; it
// is compiled into libsmolyax_host.so (g++ -ffp-contract=off) so the CPU-only tests can
// load it without a GPU runtime.
//
// The results must equal the reference's bit for bit, and that includes the floating-point
// path the reference's depth-first searches take (reference: src/smolyax/indices.py):
//   * indexset / non_zero_indices_and_zetas carry a *remaining* budget and subtract
//     j*k[i] from it                                              (indices.py:58-67, :252-258)
//   * indexset_cardinality carries a *used* budget and adds j*k[i] (indices.py:95-118)
//   * smolyak_coefficient subtracts k[i] once per included dim     (indices.py:152-168)
// and the LIFO visiting order (skip-branch pushed first, then j = 1,2,..; popped largest
// j first).  Everything here is written from that specification, with an arena of
// parent-linked nodes instead of Python tuples.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct Frame {
    int64_t dim;     // next dimension to branch on
    double rem;      // remaining (or used) budget
    int32_t node;    // arena id of the multi-index head, -1 = empty head
};

struct HeadNode {
    int32_t parent;
    int32_t dim;
    int32_t deg;
    int32_t len;  // number of active dims up to and including this node
};

// zeta for the head whose remaining budget is rem_t: sum over e in {0,1}^d (restricted to
// admissible e) of (-1)^{|e|}, parity tracked by bit flips.  indices.py:152-168.
int64_t zeta_dfs(const double* k, int64_t d, double rem_t, int64_t parity) {
    struct F {
        int64_t i;
        double rt;
        int64_t p;
    };
    std::vector<F> st;
    st.reserve(256);
    st.push_back({0, rem_t, parity});
    int64_t total = 0;
    while (!st.empty()) {
        F f = st.back();
        st.pop_back();
        if (f.i >= d) {
            total += 1 - (f.p << 1);
            continue;
        }
        if (f.i + 1 < d && k[f.i + 1] < f.rt)
            st.push_back({f.i + 1, f.rt, f.p});
        else
            total += 1 - (f.p << 1);
        const double cost = k[f.i];
        if (cost < f.rt) st.push_back({f.i + 1, f.rt - cost, f.p ^ 1});
    }
    return total;
}

struct IndexResult {
    // CSR over emitted multi-indices, in emission order
    std::vector<int64_t> offsets{0};
    std::vector<int32_t> dims;
    std::vector<int32_t> degs;
    std::vector<int64_t> zetas;  // only filled by the non-zero-zeta walk
};

void emit(IndexResult& r, const std::vector<HeadNode>& arena, int32_t node) {
    const int len = node < 0 ? 0 : arena[node].len;
    const size_t base = r.dims.size();
    r.dims.resize(base + len);
    r.degs.resize(base + len);
    int pos = len - 1;
    for (int32_t n = node; n >= 0; n = arena[n].parent, --pos) {
        r.dims[base + pos] = arena[n].dim;
        r.degs[base + pos] = arena[n].deg;
    }
    r.offsets.push_back(static_cast<int64_t>(r.dims.size()));
}

}  // namespace

extern "C" {

// ---- opaque result objects -------------------------------------------------------------
void smxh_result_free(void* h) { delete static_cast<IndexResult*>(h); }
int64_t smxh_result_count(void* h) { return (int64_t)static_cast<IndexResult*>(h)->offsets.size() - 1; }
int64_t smxh_result_nnz(void* h) { return (int64_t)static_cast<IndexResult*>(h)->dims.size(); }
void smxh_result_copy(void* h, int64_t* offsets, int32_t* dims, int32_t* degs, int64_t* zetas) {
    auto* r = static_cast<IndexResult*>(h);
    std::memcpy(offsets, r->offsets.data(), r->offsets.size() * sizeof(int64_t));
    if (!r->dims.empty()) {
        std::memcpy(dims, r->dims.data(), r->dims.size() * sizeof(int32_t));
        std::memcpy(degs, r->degs.data(), r->degs.size() * sizeof(int32_t));
    }
    if (zetas && !r->zetas.empty()) std::memcpy(zetas, r->zetas.data(), r->zetas.size() * sizeof(int64_t));
}

// Lambda(k,t) in sparse form, reference order.  Mirrors indices.py:46-69.
void* smxh_indexset(const double* k, int64_t d, double t) {
    auto* res = new IndexResult();
    std::vector<HeadNode> arena;
    std::vector<Frame> st;
    st.push_back({0, t, -1});
    while (!st.empty()) {
        Frame f = st.back();
        st.pop_back();
        if (f.dim >= d || k[f.dim] >= f.rem) {
            emit(*res, arena, f.node);
            continue;
        }
        st.push_back({f.dim + 1, f.rem, f.node});
        const double ki = k[f.dim];
        const int32_t plen = f.node < 0 ? 0 : arena[f.node].len;
        for (int64_t j = 1; (double)j * ki < f.rem; ++j) {
            arena.push_back({f.node, (int32_t)f.dim, (int32_t)j, plen + 1});
            st.push_back({f.dim + 1, f.rem - (double)j * ki, (int32_t)arena.size() - 1});
        }
    }
    return res;
}

// |Lambda(k,t)| without building it.  Mirrors indices.py:95-118 (note: *used* budget).
int64_t smxh_indexset_cardinality(const double* k, int64_t d, double t) {
    struct F {
        int64_t dim;
        double used;
    };
    std::vector<F> st;
    st.push_back({0, 0.0});
    int64_t count = 0;
    while (!st.empty()) {
        F f = st.back();
        st.pop_back();
        if (f.dim >= d) {
            ++count;
            continue;
        }
        const double remaining = t - f.used;
        if (f.dim + 1 < d && k[f.dim + 1] < remaining)
            st.push_back({f.dim + 1, f.used});
        else
            ++count;
        for (int64_t j = 1; f.used + (double)j * k[f.dim] < t; ++j) st.push_back({f.dim + 1, f.used + (double)j * k[f.dim]});
    }
    return count;
}

int64_t smxh_smolyak_coefficient(const double* k, int64_t d, double rem_t, int64_t parity) {
    return zeta_dfs(k, d, rem_t, parity);
}

// Number of distinct nodes for non-nested rules: sum of prod(nu_j+1) over nu with zeta != 0.
// Mirrors indices.py:183-211.
int64_t smxh_nodeset_cardinality_non_nested(const double* k, int64_t d, double t) {
    struct F {
        int64_t dim;
        double rem;
        int64_t parity;
        int64_t prod;
    };
    std::vector<F> st;
    st.push_back({0, t, 0, 1});
    int64_t total = 0;
    while (!st.empty()) {
        F f = st.back();
        st.pop_back();
        const bool can_skip = f.dim < d && f.dim + 1 < d && k[f.dim + 1] < f.rem;
        if (f.dim >= d || !can_skip) {
            if (zeta_dfs(k, d, f.rem, f.parity) != 0) total += f.prod;
        }
        if (f.dim < d) {
            if (can_skip) st.push_back({f.dim + 1, f.rem, f.parity, f.prod});
            const double cost = k[f.dim];
            for (int64_t j = 1; cost * (double)j < f.rem; ++j)
                st.push_back({f.dim + 1, f.rem - cost * (double)j, f.parity ^ (j & 1), f.prod * (j + 1)});
        }
    }
    return total;
}

// Multi-indices with zeta != 0 and their zetas, in the order the reference's walk emits them
// (callers bin by number of active dims).  Mirrors indices.py:238-259.
void* smxh_nonzero_indices_and_zetas(const double* k, int64_t d, double t) {
    auto* res = new IndexResult();
    std::vector<HeadNode> arena;
    std::vector<Frame> st;
    st.push_back({0, t, -1});
    while (!st.empty()) {
        Frame f = st.back();
        st.pop_back();
        const bool can_skip = f.dim < d && f.dim + 1 < d && k[f.dim + 1] < f.rem;
        if (f.dim >= d || !can_skip) {
            const int64_t z = zeta_dfs(k, d, f.rem, 0);
            if (z != 0) {
                emit(*res, arena, f.node);
                res->zetas.push_back(z);
            }
        }
        if (f.dim < d) {
            if (can_skip) st.push_back({f.dim + 1, f.rem, f.node});
            const double ki = k[f.dim];
            const int32_t plen = f.node < 0 ? 0 : arena[f.node].len;
            for (int64_t j = 1; (double)j * ki < f.rem; ++j) {
                arena.push_back({f.node, (int32_t)f.dim, (int32_t)j, plen + 1});
                st.push_back({f.dim + 1, f.rem - (double)j * ki, (int32_t)arena.size() - 1});
            }
        }
    }
    return res;
}

}  // extern "C"
