// Plan compiler: reference per-group layout -> hierarchical block-sparse device layout (see smx_plan.h).
// Host-only C++ (long double arithmetic); compiled into both libsmolyax_host.so and libsmolyax_b200.so.
#include "smx_plan.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <thread>
#include <unordered_map>

namespace smx {
namespace {

typedef long double ld;
typedef std::vector<int64_t> Key;  // sorted codes dim * kCode + deg
constexpr int64_t kCode = 65536;

struct KeyHash {
    size_t operator()(const Key& k) const {
        uint64_t h = 1469598103934665603ull;
        for (int64_t v : k) {
            h ^= (uint64_t)v;
            h *= 1099511628211ull;
        }
        return (size_t)h;
    }
};

struct PairInfo {
    int dim, deg;
    const double* nodes;   // deg+1 nodes of this (dim, deg), as stored in the layout
    std::vector<ld> minv;  // (deg+1)^2: Newton coefficient k = sum_m minv[k*(deg+1)+m] * value_m
};

// Gauss-Jordan inverse with partial pivoting in long double.
bool invert(std::vector<ld> a, int n, std::vector<ld>& inv) {
    inv.assign((size_t)n * n, 0.0L);
    for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0L;
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r)
            if (fabsl(a[(size_t)r * n + c]) > fabsl(a[(size_t)p * n + c])) p = r;
        if (a[(size_t)p * n + c] == 0.0L || !std::isfinite((double)a[(size_t)p * n + c])) return false;
        if (p != c)
            for (int j = 0; j < n; ++j) {
                std::swap(a[(size_t)p * n + j], a[(size_t)c * n + j]);
                std::swap(inv[(size_t)p * n + j], inv[(size_t)c * n + j]);
            }
        const ld d = 1.0L / a[(size_t)c * n + c];
        for (int j = 0; j < n; ++j) {
            a[(size_t)c * n + j] *= d;
            inv[(size_t)c * n + j] *= d;
        }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const ld f = a[(size_t)r * n + c];
            if (f == 0.0L) continue;
            for (int j = 0; j < n; ++j) {
                a[(size_t)r * n + j] -= f * a[(size_t)c * n + j];
                inv[(size_t)r * n + j] -= f * inv[(size_t)c * n + j];
            }
        }
    }
    return true;
}

// Greedy Leja ordering: start next to the centroid, then maximise the product of distances to the chosen ones.
std::vector<double> leja_order(const double* pts, int n) {
    std::vector<double> rest(pts, pts + n), out;
    double mean = 0;
    for (double v : rest) mean += v / n;
    size_t best = 0;
    for (size_t i = 1; i < rest.size(); ++i)
        if (std::fabs(rest[i] - mean) < std::fabs(rest[best] - mean)) best = i;
    out.push_back(rest[best]);
    rest.erase(rest.begin() + best);
    while (!rest.empty()) {
        best = 0;
        ld bestv = -1;
        for (size_t i = 0; i < rest.size(); ++i) {
            ld prod = 1;
            for (double c : out) prod *= fabsl((ld)rest[i] - (ld)c);
            if (prod > bestv) {
                bestv = prod;
                best = i;
            }
        }
        out.push_back(rest[best]);
        rest.erase(rest.begin() + best);
    }
    return out;
}

// Largest absolute error of a cardinal function l_c(x) = prod_{j != c} (x - xi_j) / (xi_c - xi_j) when it is evaluated the
// way the kernels do it - sum_k minv[k][c] pi_k(x) with the Newton products pi_k and the sum in fp64 - measured against the
// product formula in long double on 48 uniform samples of the node span and the midpoints between neighbouring nodes.
double newton_form_error(const double* nodes, const double* eta, int m, const std::vector<ld>& minv) {
    std::vector<double> xs, sorted(nodes, nodes + m);
    std::sort(sorted.begin(), sorted.end());
    for (int i = 0; i <= 48; ++i) xs.push_back(sorted.front() + (sorted.back() - sorted.front()) * i / 48.0);
    for (int i = 0; i + 1 < m; ++i) xs.push_back(0.5 * (sorted[i] + sorted[i + 1]));
    std::vector<double> c((size_t)m * m), pi((size_t)m);
    for (size_t i = 0; i < c.size(); ++i) c[i] = (double)minv[i];
    double worst = 0.0;
    for (double x : xs) {
        pi[0] = 1.0;
        for (int k = 1; k < m; ++k) pi[k] = pi[k - 1] * (x - eta[k - 1]);
        for (int cidx = 0; cidx < m; ++cidx) {
            double v = 0.0;
            for (int k = 0; k < m; ++k) v = std::fma(c[(size_t)k * m + cidx], pi[k], v);
            ld exact = 1.0L;
            for (int j = 0; j < m; ++j)
                if (j != cidx) exact *= ((ld)x - (ld)nodes[j]) / ((ld)nodes[cidx] - (ld)nodes[j]);
            worst = std::max(worst, (double)fabsl((ld)v - exact));
        }
    }
    return worst;
}
// Accepted loss (absolute, cardinal functions are O(1) between the nodes).  Nested Leja rules stay below 2e-14 up to degree
// 150; Gauss-Hermite rules re-expressed on the Leja-ordered nodes of their highest degree pass up to degree ~14 (6e-15 at
// degree 10, 1.5e-11 at 20, 5e-8 at 30).  Beyond the limit smx_create falls back to the per-summand barycentric kernels.
constexpr double kNewtonErrorMax = 2.0e-13;

// One summand (multi-index nu with zeta != 0), wherever its tables live: in the reference's padded per-group layout or
// in the compact node-indexed form of smx_create_compact.
struct Summand {
    int n = 0;
    int64_t zeta = 0;
    int64_t mu_off = 0;  // offset into mu_terms
    int64_t count = 1;   // prod (deg+1)
    const int64_t* dims = nullptr;
    const int64_t* degs = nullptr;
    const double* nodes[kMaxLevels] = {nullptr};  // per slot: its deg+1 nodes
    const double* quad[kMaxLevels] = {nullptr};   // per slot: its quadrature weights (optional)
    // values: either the summand's block of the padded F (element strides per axis, stride between outputs) ..
    const double* F = nullptr;
    int64_t fstride[kMaxLevels] = {0};
    int64_t ostride = 0;
    // .. or, per entry of the exact-shape tensor (C order over the slots), a row of the (n_values, d_out) value table
    const int64_t* vidx = nullptr;
    const double* values = nullptr;
    int64_t vstride = 0;
};

std::string compact_summands(int64_t d_out, const CompactView& cv, std::vector<Summand>& summands) {
    if (cv.n_summands < 0 || (cv.n_summands > 0 && (!cv.n_active || !cv.slot_off || !cv.dims || !cv.degs || !cv.node_off ||
                                                    !cv.node_pool || !cv.zetas || !cv.val_off || !cv.val_index || !cv.values)))
        return "compact descriptor has null arrays";
    // CSR offsets start at 0 (the descriptor carries no array lengths: an offset array shifted as a whole would otherwise pass
    // the per-summand difference checks and read past the end of val_index / dims)
    if (cv.n_summands > 0 && (cv.slot_off[0] != 0 || cv.val_off[0] != 0)) return "slot_off / val_off must start at 0";
    for (int64_t s = 0; s < cv.n_summands; ++s) {
        Summand sm;
        sm.n = cv.n_active[s];
        if (sm.n <= 0 || sm.n > kMaxLevels) return "summand with unsupported number of active dimensions";
        if (cv.slot_off[s + 1] - cv.slot_off[s] != sm.n) return "slot_off does not match n_active";
        sm.zeta = cv.zetas[s];
        sm.dims = cv.dims + cv.slot_off[s];
        sm.degs = cv.degs + cv.slot_off[s];
        int64_t count = 1;
        for (int j = 0; j < sm.n; ++j) {
            sm.nodes[j] = cv.node_pool + cv.node_off[cv.slot_off[s] + j];
            sm.quad[j] = cv.quad_pool ? cv.quad_pool + cv.node_off[cv.slot_off[s] + j] : nullptr;
            if (sm.degs[j] < 1) return "sorted_degs entry out of range";
            count *= (sm.degs[j] + 1);
        }
        if (cv.val_off[s + 1] - cv.val_off[s] != count) return "val_off does not match the summand's shape";
        sm.vidx = cv.val_index + cv.val_off[s];
        for (int64_t i = 0; i < count; ++i)
            if (sm.vidx[i] < 0 || sm.vidx[i] >= cv.n_values) return "val_index entry out of range";
        sm.values = cv.values;
        sm.vstride = d_out;
        summands.push_back(sm);
    }
    return "";
}


}  // namespace

static std::string build_from_summands(int64_t d_in, int64_t d_out, const double* offset, std::vector<Summand>& summands,
                                       int64_t w_pad, FastPlan& plan, const PlanOptions& opt);

static std::string group_summands(int64_t d_out, const std::vector<GroupView>& groups, std::vector<Summand>& summands, int64_t& w_pad) {
    for (const GroupView& G : groups) {
        if (G.n <= 0 || G.n > kMaxLevels) return "group with unsupported number of active dimensions";
        if ((int)G.tau.size() != G.n) return "tau has wrong length";
        if (!G.F || !G.nodes || !G.dims || !G.degs || !G.zetas) return "group descriptor has null arrays";
        const int64_t tw = G.tw(), fsize = G.fsize();
        w_pad += G.nn * fsize;
        for (int64_t s = 0; s < G.nn; ++s) {
            Summand sm;
            sm.n = G.n;
            sm.zeta = G.zetas[s];
            sm.dims = G.dims + s * G.n;
            sm.degs = G.degs + s * G.n;
            for (int j = G.n - 1; j >= 0; --j) {
                if (sm.degs[j] > G.tau[j]) return "sorted_degs entry out of range";
                sm.nodes[j] = G.nodes + (s * G.n + j) * tw;
                sm.quad[j] = G.quad ? G.quad + (s * G.n + j) * tw : nullptr;
                sm.fstride[j] = (j == G.n - 1) ? 1 : sm.fstride[j + 1] * (G.tau[j + 1] + 1);
            }
            sm.F = G.F + s * d_out * fsize;
            sm.ostride = fsize;
            summands.push_back(sm);
        }
    }
    return "";
}

std::string build_fast_plan(int64_t d_in, int64_t d_out, const double* offset, const std::vector<GroupView>& groups,
                            FastPlan& plan, const PlanOptions& opt) {
    std::vector<Summand> summands;
    int64_t w_pad = 0;
    const std::string err = group_summands(d_out, groups, summands, w_pad);
    if (!err.empty()) return err;
    return build_from_summands(d_in, d_out, offset, summands, w_pad, plan, opt);
}

std::string build_fast_plan_compact(int64_t d_in, int64_t d_out, const double* offset, const CompactView& cv, FastPlan& plan,
                                    const PlanOptions& opt) {
    std::vector<Summand> summands;
    const std::string err = compact_summands(d_out, cv, summands);
    if (!err.empty()) return err;
    return build_from_summands(d_in, d_out, offset, summands, 0, plan, opt);
}

// Smolyak quadrature of the summands in long double:  Q[o] = offset[o] + sum_s zeta_s sum_mu value(mu)[o] prod_j quad_j[mu_j]
static std::string integrate_summands(int64_t d_out, const double* offset, const std::vector<Summand>& summands, std::vector<double>& Q) {
    std::vector<ld> acc((size_t)d_out, 0.0L);
    for (int64_t o = 0; o < d_out; ++o) acc[o] = offset ? (ld)offset[o] : 0.0L;
    for (const Summand& sm : summands) {
        int mu[kMaxLevels] = {0};
        int64_t count = 1;
        for (int j = 0; j < sm.n; ++j) {
            if (!sm.quad[j]) return "no quadrature weights in the descriptor";
            count *= sm.degs[j] + 1;
        }
        for (int64_t idx = 0; idx < count; ++idx) {
            ld w = (ld)sm.zeta;
            int64_t foff = 0;
            for (int j = 0; j < sm.n; ++j) w *= (ld)sm.quad[j][mu[j]], foff += mu[j] * sm.fstride[j];
            if (sm.vidx) {
                const double* v = sm.values + sm.vidx[idx] * sm.vstride;
                for (int64_t o = 0; o < d_out; ++o) acc[o] += w * (ld)v[o];
            } else {
                for (int64_t o = 0; o < d_out; ++o) acc[o] += w * (ld)sm.F[o * sm.ostride + foff];
            }
            for (int j = sm.n - 1; j >= 0; --j) {
                if (++mu[j] <= sm.degs[j]) break;
                mu[j] = 0;
            }
        }
    }
    Q.resize((size_t)d_out);
    for (int64_t o = 0; o < d_out; ++o) Q[o] = (double)acc[o];
    return "";
}

std::string integrate_compact(int64_t d_out, const double* offset, const CompactView& cv, std::vector<double>& Q) {
    std::vector<Summand> summands;
    const std::string err = compact_summands(d_out, cv, summands);
    if (!err.empty()) return err;
    return integrate_summands(d_out, offset, summands, Q);
}

static std::string group_summands(int64_t d_out, const std::vector<GroupView>& groups, std::vector<Summand>& summands, int64_t& w_pad);

std::string integrate_groups(int64_t d_out, const double* offset, const std::vector<GroupView>& groups, std::vector<double>& Q) {
    std::vector<Summand> summands;
    int64_t w_pad = 0;
    const std::string err = group_summands(d_out, groups, summands, w_pad);
    if (!err.empty()) return err;
    return integrate_summands(d_out, offset, summands, Q);
}

static std::string build_from_summands(int64_t d_in, int64_t d_out, const double* offset, std::vector<Summand>& summands,
                                       int64_t w_pad, FastPlan& plan, const PlanOptions& opt) {
    const bool with_gradient = opt.gradient && opt.sparse;  // (the cold derivatives always come from the sparse form)
    plan = FastPlan();
    plan.d_in = d_in;
    plan.d_out = d_out;
    plan.w_pad = w_pad;
    if (d_in <= 0 || d_out <= 0) return "d_in and d_out must be positive";

    // ---- 1. unique (dim, deg) pairs, per-dimension maximal degree ------------------------------------------
    std::map<std::pair<int, int>, int> pair_id;
    std::vector<PairInfo> pairs;
    std::vector<int> maxdeg((size_t)d_in, 0);
    int64_t total_mu = 0;
    for (Summand& sm : summands) {
        sm.mu_off = total_mu;
        sm.count = 1;
        for (int j = 0; j < sm.n; ++j) {
            const int64_t dim = sm.dims[j], deg = sm.degs[j];
            if (dim < 0 || dim >= d_in) return "sorted_dims entry out of range";
            if (deg < 1 || deg >= kCode) return "sorted_degs entry out of range";
            for (int i = 0; i < j; ++i)
                if (sm.dims[i] == dim) return "duplicate dimension inside one summand";
            auto key = std::make_pair((int)dim, (int)deg);
            if (!pair_id.count(key)) {
                pair_id[key] = (int)pairs.size();
                pairs.push_back({(int)dim, (int)deg, sm.nodes[j], {}});
            }
            maxdeg[dim] = std::max(maxdeg[dim], (int)deg);
            sm.count *= (deg + 1);
        }
        total_mu += sm.count;
    }
    plan.n_summands = (int64_t)summands.size();
    plan.w_raw = total_mu;

    // ---- 2. nestedness and Newton centres per dimension ---------------------------------------------------
    std::vector<int32_t> eta_off((size_t)d_in + 1, 0);
    for (int64_t d = 0; d < d_in; ++d) eta_off[d + 1] = eta_off[d] + maxdeg[d];
    plan.eta.assign((size_t)eta_off[d_in], 0.0);
    bool nested = true;
    for (const PairInfo& p : pairs) {
        const PairInfo& top = pairs[pair_id[{p.dim, maxdeg[p.dim]}]];
        if (std::memcmp(p.nodes, top.nodes, sizeof(double) * (p.deg + 1)) != 0) nested = false;
    }
    plan.nested = nested;
    for (int64_t d = 0; d < d_in; ++d) {
        if (maxdeg[d] == 0) continue;
        const PairInfo& top = pairs[pair_id[{(int)d, maxdeg[d]}]];
        if (nested) {
            std::copy(top.nodes, top.nodes + maxdeg[d], plan.eta.begin() + eta_off[d]);
        } else {
            std::vector<double> ord = leja_order(top.nodes, maxdeg[d] + 1);
            std::copy(ord.begin(), ord.begin() + maxdeg[d], plan.eta.begin() + eta_off[d]);
        }
    }

    // ---- 3. nodal values -> Newton coefficients, per pair -------------------------------------------------
    std::map<std::vector<double>, double> cond_cache;  // (nodes, centres) -> error of the fp64 Newton form (pairs share node sets)
    for (PairInfo& p : pairs) {
        const int m = p.deg + 1;
        std::vector<ld> A((size_t)m * m);
        const double* eta = plan.eta.data() + eta_off[p.dim];
        for (int r = 0; r < m; ++r) {
            ld v = 1.0L;
            for (int k = 0; k < m; ++k) {
                A[(size_t)r * m + k] = v;  // pi_k(xi_r)
                if (k + 1 < m) v *= ((ld)p.nodes[r] - (ld)eta[k]);
            }
        }
        if (!invert(A, m, p.minv)) return "interpolation nodes of one slot are not distinct";
        // The change of basis is exact in long double, but the kernels evaluate the Newton form in fp64: for non-nested
        // rules of high degree (Gauss-Hermite beyond degree ~15) the Newton coefficients of the cardinal functions grow
        // until their cancellation costs more digits than the path may lose.  Measure it: every cardinal function through
        // its Newton form in fp64 against its product formula in long double, on samples over the span of the nodes.
        if (m > 6) {
            std::vector<double> key(p.nodes, p.nodes + m);
            key.insert(key.end(), eta, eta + (m - 1));
            auto it = cond_cache.find(key);
            double err = 0.0;
            if (it != cond_cache.end()) {
                err = it->second;
            } else {
                err = newton_form_error(p.nodes, eta, m, p.minv);
                cond_cache.emplace(std::move(key), err);
            }
            plan.newton_error = std::max(plan.newton_error, err);
        }
    }
    if (plan.newton_error > kNewtonErrorMax) {
        char msg[200];
        std::snprintf(msg, sizeof msg, "ill-conditioned: the hierarchical (Newton) form of a 1-D rule of this layout loses %.1e of a cardinal function in "
                      "fp64 (limit %.1e)", plan.newton_error, kNewtonErrorMax);
        return msg;
    }

    // ---- 4. enumerate terms: every mu <= nu of every summand ------------------------------------------------
    std::unordered_map<Key, int32_t, KeyHash> term_id;
    std::vector<Key> term_key;
    term_id.reserve((size_t)total_mu);
    term_id[Key()] = 0;
    term_key.push_back(Key());
    std::vector<int32_t> mu_terms((size_t)total_mu);
    {
        Key key;
        std::vector<std::pair<int64_t, int>> act;
        for (const Summand& sm : summands) {
            int mu[kMaxLevels] = {0};
            for (int64_t idx = 0; idx < sm.count; ++idx) {
                act.clear();
                for (int j = 0; j < sm.n; ++j)
                    if (mu[j] > 0) act.push_back({sm.dims[j], mu[j]});
                std::sort(act.begin(), act.end());
                key.clear();
                for (auto& a : act) key.push_back(a.first * kCode + a.second);
                auto it = term_id.find(key);
                int32_t id;
                if (it == term_id.end()) {
                    id = (int32_t)term_key.size();
                    term_id.emplace(key, id);
                    term_key.push_back(key);
                } else {
                    id = it->second;
                }
                mu_terms[(size_t)(sm.mu_off + idx)] = id;
                for (int j = sm.n - 1; j >= 0; --j) {  // odometer, last axis fastest (C order of F)
                    if (++mu[j] <= sm.degs[j]) break;
                    mu[j] = 0;
                }
            }
        }
    }
    const int64_t T = (int64_t)term_key.size();
    plan.n_terms = T;

    // ---- 5. coefficients C[T][d_out] in long double ---------------------------------------------------------
    std::vector<ld> C((size_t)T * d_out, 0.0L);
    if (offset)
        for (int64_t o = 0; o < d_out; ++o) C[o] = (ld)offset[o];
    {
        const int64_t och_max = 256;
        const int64_t n_och = (d_out + och_max - 1) / och_max;
        auto work = [&](int64_t c_begin, int64_t c_end) {
            std::vector<ld> Gt, tmp;
            for (int64_t ch = c_begin; ch < c_end; ++ch) {
                const int64_t o0 = ch * och_max, och = std::min(och_max, d_out - o0);
                for (const Summand& sm : summands) {
                    const int n = sm.n;
                    int m[kMaxLevels], pid[kMaxLevels];
                    int64_t mstride[kMaxLevels];
                    for (int j = n - 1; j >= 0; --j) {
                        m[j] = (int)sm.degs[j] + 1;
                        pid[j] = pair_id.at({(int)sm.dims[j], m[j] - 1});
                        mstride[j] = (j == n - 1) ? 1 : mstride[j + 1] * m[j + 1];
                    }
                    Gt.resize((size_t)sm.count * och);
                    int mu[kMaxLevels] = {0};
                    for (int64_t idx = 0; idx < sm.count; ++idx) {
                        if (sm.vidx) {
                            const double* v = sm.values + sm.vidx[idx] * sm.vstride + o0;
                            for (int64_t oo = 0; oo < och; ++oo) Gt[(size_t)(idx * och + oo)] = (ld)v[oo];
                        } else {
                            int64_t foff = 0;
                            for (int j = 0; j < n; ++j) foff += mu[j] * sm.fstride[j];
                            for (int64_t oo = 0; oo < och; ++oo) Gt[(size_t)(idx * och + oo)] = (ld)sm.F[(o0 + oo) * sm.ostride + foff];
                        }
                        for (int j = n - 1; j >= 0; --j) {
                            if (++mu[j] < m[j]) break;
                            mu[j] = 0;
                        }
                    }
                    for (int j = 0; j < n; ++j) {  // mode product with minv along axis j
                        const int mj = m[j];
                        const std::vector<ld>& Mi = pairs[pid[j]].minv;
                        const int64_t inner = mstride[j], outer = sm.count / (inner * mj);
                        tmp.resize((size_t)mj * och);
                        for (int64_t u = 0; u < outer; ++u)
                            for (int64_t v = 0; v < inner; ++v) {
                                const int64_t base = u * mj * inner + v;
                                for (int a = 0; a < mj; ++a)
                                    std::memcpy(&tmp[(size_t)a * och], &Gt[(size_t)((base + a * inner) * och)], sizeof(ld) * och);
                                for (int k = 0; k < mj; ++k) {
                                    ld* dst = &Gt[(size_t)((base + k * inner) * och)];
                                    for (int64_t oo = 0; oo < och; ++oo) {
                                        ld acc = 0.0L;
                                        for (int a = 0; a < mj; ++a) acc += Mi[(size_t)k * mj + a] * tmp[(size_t)a * och + oo];
                                        dst[oo] = acc;
                                    }
                                }
                            }
                    }
                    const ld z = (ld)sm.zeta;
                    for (int64_t idx = 0; idx < sm.count; ++idx) {
                        ld* dst = &C[(size_t)mu_terms[(size_t)(sm.mu_off + idx)] * d_out + o0];
                        const ld* src = &Gt[(size_t)(idx * och)];
                        for (int64_t oo = 0; oo < och; ++oo) dst[oo] += z * src[oo];
                    }
                }
            }
        };
        unsigned hw = std::thread::hardware_concurrency();
        int64_t nthr = std::max<int64_t>(1, std::min<int64_t>(hw ? hw : 1, n_och));
        if (nthr == 1) {
            work(0, n_och);
        } else {
            std::vector<std::thread> pool;
            for (int64_t t = 0; t < nthr; ++t)
                pool.emplace_back(work, t * n_och / nthr, (t + 1) * n_och / nthr);
            for (auto& th : pool) th.join();
        }
    }

    // ---- 6. hot parts: every term minus its leading (largest-dimension) pair ---------------------------------
    std::unordered_map<Key, int32_t, KeyHash> row_id;  // hot part -> provisional row id (0 = empty product)
    std::vector<Key> row_key{Key()};
    row_id[Key()] = 0;
    int hot_dim_max = -1;
    std::vector<int32_t> term_row((size_t)T, 0);
    for (int32_t t = 1; t < T; ++t) {
        const Key& key = term_key[t];
        Key pre;
        int32_t id = 0;
        for (size_t i = 0; i + 1 < key.size(); ++i) {  // all prefixes are rows too (parents)
            pre.push_back(key[i]);
            hot_dim_max = std::max(hot_dim_max, (int)(key[i] / kCode));
            auto it = row_id.find(pre);
            if (it == row_id.end()) {
                id = (int32_t)row_key.size();
                row_id.emplace(pre, id);
                row_key.push_back(pre);
            } else {
                id = it->second;
            }
        }
        term_row[t] = id;
    }
    plan.n_rows = (int32_t)row_key.size();

    // ---- 7. leading entries: hot prefix, then cold blocks = 16 consecutive columns of x from an even column -----------
    // The hot prefix holds every entry of the dimensions up to the last one that occurs inside a hot part or has a
    // degree above 1; beyond it every dimension has at most the entry (dim, 1), so a cold block is a contiguous tile
    // of x (dummy entries with zero coefficients fill unused columns).
    for (int64_t d = 0; d < d_in; ++d)
        if (maxdeg[d] >= 2) hot_dim_max = std::max(hot_dim_max, (int)d);
    plan.hot_dims = hot_dim_max + 1;
    std::unordered_map<int64_t, int32_t> ent_index;  // (dim, deg) code -> entry
    auto push_entry = [&](int32_t dim, int32_t deg, int32_t tab) {
        plan.ent_dim.push_back(dim);
        plan.ent_deg.push_back(deg);
        plan.ent_eta.push_back(eta_off[dim]);
        plan.ent_tab.push_back(tab);
        plan.ent_eta0.push_back(maxdeg[dim] > 0 ? plan.eta[eta_off[dim]] : 0.0);
    };
    {   // hot entries degree-major: entries of equal degree have nearly the same admissible hot parts, so the row
        // union of a block stays close to the rows each of its entries needs (measured: 36 % fewer padded slots)
        std::vector<std::pair<int32_t, int32_t>> hot_pairs;  // (deg, dim)
        for (int64_t d = 0; d <= hot_dim_max; ++d)
            for (int a = 1; a <= maxdeg[d]; ++a) hot_pairs.push_back({a, (int32_t)d});
        std::sort(hot_pairs.begin(), hot_pairs.end());
        plan.hot_pos.assign(hot_pairs.size(), 0);
        for (auto& pr : hot_pairs) {
            const int32_t e = (int32_t)plan.ent_dim.size();
            ent_index[(int64_t)pr.second * kCode + pr.first] = e;
            plan.hot_pos[(size_t)eta_off[pr.second] + pr.first - 1] = e;
            push_entry(pr.second, pr.first, 1 + hot_row(e));
            ++plan.n_hot;
            ++plan.n_entries;
        }
    }
    while (plan.ent_dim.size() % kBlockWidth) push_entry(0, 0, 0);  // cold entries start on a block boundary
    for (int64_t next = hot_dim_max + 1; next < d_in;) {
        while (next < d_in && maxdeg[next] == 0) ++next;  // skip columns nobody uses
        if (next >= d_in) break;
        const int64_t col0 = next & ~(int64_t)1;
        for (int64_t col = col0; col < col0 + kBlockWidth; ++col) {
            const bool real = col >= next && col < d_in && maxdeg[col] == 1;
            if (real) {
                ent_index[col * kCode + 1] = (int32_t)plan.ent_dim.size();
                ++plan.n_entries;
            }
            push_entry((int32_t)std::min<int64_t>(col, d_in - 1), real ? 1 : 0, 0);
        }
        next = col0 + kBlockWidth;
    }
    // hot entry h (in entry order) sits in table row 1 + h; map (dim, deg) code -> table row
    std::unordered_map<int64_t, int32_t> hot_tab;
    for (size_t e = 0; e < plan.ent_dim.size(); ++e)
        if (plan.ent_tab[e] > 0) hot_tab[(int64_t)plan.ent_dim[e] * kCode + plan.ent_deg[e]] = plan.ent_tab[e];

    // ---- 7b. derivative coefficients, sparse --------------------------------------------------------------------------
    // d/dx_i of the interpolant is a polynomial over the same term set: a term with the pair (i, b) contributes to the
    // terms with (i, k), k < b, through  pi_b' = sum_k D[b][k] pi_k  (D from pi_b = pi_{b-1} (x - eta_{b-1})).  Only hot
    // dimensions need it (cold ones: pi = x - eta_0, pi' = 1, the derivative is a row sum of the value contraction), and
    // only the terms that contain the dimension carry a coefficient: Cd[h] maps term -> d_out coefficients.
    std::vector<std::unordered_map<int32_t, std::vector<ld>>> Cd;
    if (with_gradient) {
        for (int64_t d = 0; d <= hot_dim_max; ++d)
            if (maxdeg[d] > 0) plan.grad_dims.push_back((int32_t)d);
        // (memory of the sparse sets: terms that contain the dimension x degree x outputs; refuse beyond ~3 GB - the host
        //  layer then differentiates by blocks of output columns)
        double cells = 0.0;
        for (int32_t t = 1; t < T; ++t)
            for (int64_t code : term_key[t])
                if (code / kCode <= hot_dim_max) cells += (double)(code % kCode);
        if (cells * (double)d_out * (sizeof(ld) + 24.0) > 3.0e9) plan.grad_dims.clear();
        Cd.assign(plan.grad_dims.size(), {});
        std::vector<int32_t> h_of_dim((size_t)d_in, -1);
        for (size_t h = 0; h < plan.grad_dims.size(); ++h) h_of_dim[plan.grad_dims[h]] = (int32_t)h;
        std::vector<std::vector<std::vector<ld>>> Dm((size_t)hot_dim_max + 2);  // per hot dimension: D[b][k]
        for (size_t h = 0; h < plan.grad_dims.size(); ++h) {
            const int d = plan.grad_dims[h], D = maxdeg[d];
            const double* eta = plan.eta.data() + eta_off[d];
            auto& M = Dm[(size_t)d];
            M.assign((size_t)D + 1, std::vector<ld>((size_t)D + 1, 0.0L));
            for (int b = 1; b <= D; ++b) {
                for (int k = 0; k < b - 1; ++k) {
                    M[b][k + 1] += M[b - 1][k];
                    M[b][k] += M[b - 1][k] * ((ld)eta[k] - (ld)eta[b - 1]);
                }
                M[b][b - 1] += 1.0L;
            }
        }
        Key key2;
        for (int32_t t = 1; t < T && !plan.grad_dims.empty(); ++t) {
            const Key& key = term_key[t];
            const ld* src = &C[(size_t)t * d_out];
            for (size_t pos = 0; pos < key.size(); ++pos) {
                const int d = (int)(key[pos] / kCode), b = (int)(key[pos] % kCode);
                if (d > hot_dim_max) continue;
                const int32_t h = h_of_dim[d];
                for (int k = 0; k < b; ++k) {
                    const ld w = Dm[(size_t)d][b][k];
                    if (w == 0.0L) continue;
                    key2 = key;
                    if (k == 0) key2.erase(key2.begin() + pos);
                    else key2[pos] = (int64_t)d * kCode + k;
                    const auto it = term_id.find(key2);
                    if (it == term_id.end()) return "internal error: index set is not downward closed";
                    std::vector<ld>& dst = Cd[(size_t)h][it->second];
                    if (dst.empty()) dst.assign((size_t)d_out, 0.0L);
                    for (int64_t o = 0; o < d_out; ++o) dst[o] += w * src[o];
                }
            }
        }
    }
    plan.n_sets = (int32_t)d_out;
    plan.c0.resize((size_t)d_out);
    for (int64_t o = 0; o < d_out; ++o) plan.c0[o] = (double)C[(size_t)o];

    // ---- 8. value table: level-1 rows alias the hot entries, level >= 2 rows are appended level by level ---------
    const int32_t R = plan.n_rows;
    std::vector<int32_t> row_tab((size_t)R, 0);
    int maxlevel = 0;
    for (int32_t r = 0; r < R; ++r) maxlevel = std::max(maxlevel, (int)row_key[r].size());
    plan.n_levels = maxlevel + 1;
    plan.level_off.assign((size_t)plan.n_levels + 2, 0);
    plan.n_hot_rows = (plan.n_hot + kBlockWidth - 1) / kBlockWidth * kBlockWidth;
    int32_t next = 1 + plan.n_hot_rows;
    for (int32_t r = 0; r < R; ++r)
        if (row_key[r].size() == 1) row_tab[r] = hot_tab.at(row_key[r][0]);
    for (int l = 2; l <= maxlevel; ++l) {
        plan.level_off[l] = next;
        for (int32_t r = 0; r < R; ++r) {  // provisional ids are in creation order: parents precede children
            if ((int)row_key[r].size() != l) continue;
            Key parent(row_key[r].begin(), row_key[r].end() - 1);
            row_tab[r] = next++;
            plan.tab_parent.push_back(row_tab[row_id.at(parent)]);
            plan.tab_hot.push_back(hot_tab.at(row_key[r].back()));
            if (l <= 4) {  // rows of up to four pairs are also stored as a flat product of hot rows (one pass, no level order)
                for (int f = 0; f < 4; ++f) plan.tab_factors.push_back(f < l ? hot_tab.at(row_key[r][f]) : 0);
            }
            if (l <= 8) {  // .. and every row of up to eight pairs with its full factor list (kernels without product rows)
                for (int f = 0; f < 8; ++f) plan.tab_factors8.push_back(f < l ? hot_tab.at(row_key[r][f]) : 0);
            }
        }
        plan.level_off[l + 1] = next;
    }
    for (int l = maxlevel + 1; l < (int)plan.level_off.size(); ++l) plan.level_off[l] = next;
    if (maxlevel < 2) plan.level_off.assign(plan.level_off.size(), next);
    plan.n_tab = next;

    // ---- (last step, shared by both forms) nodes per dimension: the reference's gradient is NaN where a coordinate sits
    // on one of them
    auto finish_nodes = [&]() -> std::string {
        std::vector<std::vector<double>> per_dim((size_t)d_in);
        for (const PairInfo& pr : pairs) per_dim[pr.dim].insert(per_dim[pr.dim].end(), pr.nodes, pr.nodes + pr.deg + 1);
        plan.nan_off.assign((size_t)d_in + 1, 0);
        for (int64_t d = 0; d < d_in; ++d) {
            std::sort(per_dim[d].begin(), per_dim[d].end());
            per_dim[d].erase(std::unique(per_dim[d].begin(), per_dim[d].end()), per_dim[d].end());
            plan.nan_nodes.insert(plan.nan_nodes.end(), per_dim[d].begin(), per_dim[d].end());
            plan.nan_off[d + 1] = (int32_t)plan.nan_nodes.size();
        }
        return "";
    };

    // ---- 9. block-sparse coefficient matrix and work items ------------------------------------------------------
    struct Nz {
        int32_t block, row, lane, term;
    };
    auto nz_less = [](const Nz& a, const Nz& b) {
        if (a.block != b.block) return a.block < b.block;
        if (a.row != b.row) return a.row < b.row;
        return a.lane < b.lane;
    };
    std::vector<Nz> nz;
    nz.reserve((size_t)T);
    std::vector<Nz> term_nz((size_t)T, Nz{0, 0, 0, 0});  // where term t sits in the (block, row, lane) structure
    for (int32_t t = 1; t < T; ++t) {
        const int64_t lead = term_key[t].back();
        const int32_t e = ent_index.at(lead);
        term_nz[t] = {e / kBlockWidth, row_tab[term_row[t]], e % kBlockWidth, t};
        nz.push_back(term_nz[t]);
    }
    std::sort(nz.begin(), nz.end(), nz_less);
    // ---- 9a. dense form ------------------------------------------------------------------------------------------------
    if (opt.dense) {
        const int64_t K = (int64_t)nz.size();
        plan.dense_k4 = (int32_t)(((K + 3) / 4 + kDenseStageK4 - 1) / kDenseStageK4 * kDenseStageK4);
        const int64_t k4s = plan.dense_k4 + kDensePadK4;       // allocated k-steps
        plan.dense_meta.assign((size_t)k4s * 8, 0);            // padding terms: ones row * ones row, zero coefficients
        plan.dense_eta0.assign((size_t)d_in, 0.0);
        for (int64_t d = 0; d < d_in; ++d)
            if (maxdeg[d] > 0) plan.dense_eta0[d] = plan.eta[eta_off[d]];
        const int64_t nblk = (d_out + 7) / 8;
        plan.dense_coef.assign((size_t)nblk * k4s * 32, 0.0);
        for (int64_t i = 0; i < K; ++i) {
            const int32_t e = nz[i].block * kBlockWidth + nz[i].lane;
            plan.dense_meta[2 * i] = nz[i].row;
            plan.dense_meta[2 * i + 1] = plan.ent_tab[e] > 0 ? plan.ent_tab[e] : -1 - plan.ent_dim[e];
            const int64_t k4 = i >> 2, tig = i & 3;
            const ld* src = &C[(size_t)nz[i].term * d_out];
            for (int64_t o = 0; o < d_out; ++o)
                plan.dense_coef[(size_t)(((o >> 3) * k4s + k4) * 32 + 4 * (o & 7) + tig)] = (double)src[o];
        }
        plan.has_dense = true;
    }
    plan.hot_off.assign((size_t)plan.hot_dims + 1, 0);
    for (int32_t d = 0; d < plan.hot_dims; ++d) plan.hot_off[d + 1] = plan.hot_off[d] + maxdeg[d];
    if (plan.hot_off.back() != plan.n_hot) return "internal error: hot prefix mismatch";
    if (!opt.sparse) return finish_nodes();
    plan.has_sparse = true;
    auto block_flags = [&](int32_t b) {
        int flags = kChunkHot | kChunkContig;
        const int32_t e0 = b * kBlockWidth;
        for (int i = 0; i < kBlockWidth; ++i) {
            const int32_t e = e0 + i;
            if (plan.ent_deg[e] > 0 && plan.ent_tab[e] == 0) flags &= ~kChunkHot;
            if (plan.ent_deg[e] > 1) flags &= ~kChunkContig;
            if (plan.ent_dim[e] != std::min<int64_t>(plan.ent_dim[e0] + i, d_in - 1)) flags &= ~kChunkContig;
        }
        if (plan.ent_dim[e0] & 1) flags &= ~kChunkContig;
        if (flags & kChunkHot) flags &= ~kChunkContig;
        if ((flags & kChunkContig) && plan.ent_dim[e0] + kBlockWidth <= d_in) flags |= kChunkInside;
        return flags;
    };
    struct Chunk {
        int32_t block, flags;
        std::vector<int32_t> rows;
        std::vector<double> coef;
    };
    // A table row as the kernels without product rows see it: the hot rows whose product it is (itself, if it is one).
    // Shared-memory bank of a row's 16-byte pieces: (row * kTabPitch * 2) mod 32 words = 8 * (row mod 4) - the four lanes
    // (tig) that fetch the four rows of a DMMA k-step in ONE LDS.128 are conflict free iff their rows differ modulo 4.
    const int32_t flat_begin = 1 + plan.n_hot_rows, n_flat = (int32_t)(plan.tab_factors.size() / 4);
    const int32_t n_flat8 = (int32_t)(plan.tab_factors8.size() / 8);
    auto row_factors = [&](int32_t row, int32_t* out) -> int {
        if (row < flat_begin) {
            out[0] = row;
            return 1;
        }
        const int32_t k = row - flat_begin;
        int n = 0;
        if (k < n_flat8)
            for (int i = 0; i < 8; ++i)
                if (plan.tab_factors8[(size_t)k * 8 + i]) out[n++] = plan.tab_factors8[(size_t)k * 8 + i];
        return n;  // 0: deeper than eight pairs (product rows in the table only)
    };
    auto residues = [&](int32_t row) -> unsigned {
        int32_t f[8];
        const int n = row_factors(row, f);
        unsigned m = 0;
        for (int i = 0; i < n; ++i) m |= 1u << (f[i] & 3);
        return n ? m : 1u << (row & 3);
    };
    // The four rows of one DMMA k-step (the four lanes `tig` of one LDS.128) with `nf` factor slots each: which factor of a row
    // goes to which slot is free (a product; unused slots read the ones row), so choose it such that, slot by slot, the four
    // table rows differ modulo 4 as far as possible.  Cost = shared-memory wavefronts per quarter warp, summed over the slots
    // (nf = the best case).  Coordinate descent over the rows (each tries all its placements); more than four slots: the
    // factors stay in their order.  Deterministic.
    auto arrange4 = [&](const int32_t* rows4, int n_rows, int nf, int32_t (*out)[8]) -> int {
        int32_t fac[4][8];
        int cnt[4];
        for (int q = 0; q < 4; ++q) {
            cnt[q] = q < n_rows ? row_factors(rows4[q], fac[q]) : 0;
            if (cnt[q] == 1 && fac[q][0] == 0) cnt[q] = 0;  // the ones row itself
            for (int i = 0; i < 8; ++i) out[q][i] = i < cnt[q] ? fac[q][i] : 0;
        }
        auto cost = [&]() {
            int c = 0;
            for (int pos = 0; pos < nf; ++pos) {
                int worst = 1;
                for (int q = 1; q < 4; ++q) {
                    int same = 1;  // distinct rows before q in the same bank group, q included
                    bool dup = false;
                    for (int q2 = 0; q2 < q; ++q2) {
                        if (out[q2][pos] == out[q][pos]) dup = true;
                        else if (((out[q2][pos] ^ out[q][pos]) & 3) == 0) {
                            bool first = true;
                            for (int q3 = 0; q3 < q2; ++q3) first = first && out[q3][pos] != out[q2][pos];
                            same += first;
                        }
                    }
                    if (!dup) worst = std::max(worst, same);
                }
                c += worst;
            }
            return c;
        };
        int best = cost();
        if (nf > 4) return best;
        for (int pass = 0; pass < 4 && best > nf; ++pass) {
            bool better = false;
            for (int q = 0; q < 4; ++q) {
                if (cnt[q] == 0 || nf == 1) continue;
                int slot[4], keep[4];
                for (int i = 0; i < nf; ++i) slot[i] = i < nf - cnt[q] ? -1 : i - (nf - cnt[q]), keep[i] = out[q][i];
                do {  // all placements of the cnt factors into the nf slots
                    for (int i = 0; i < nf; ++i) out[q][i] = slot[i] < 0 ? 0 : fac[q][slot[i]];
                    const int c = cost();
                    if (c < best) {
                        best = c, better = true;
                        for (int i = 0; i < nf; ++i) keep[i] = out[q][i];
                    }
                } while (std::next_permutation(slot, slot + nf));
                for (int i = 0; i < nf; ++i) out[q][i] = keep[i];
            }
            if (!better) break;
        }
        return best;
    };
    auto item_slots = [&](const std::vector<int32_t>& item_rows) {
        int32_t f[8];
        int nf = 1;
        for (int32_t r : item_rows) nf = std::max(nf, row_factors(r, f));
        return nf;
    };
    // (block, row, lane)-sorted non-zeros -> work items of at most kChunkRows rows; coefficient of term t and output o from `coef_fn`
    auto build_chunks = [&](const std::vector<Nz>& nzl, auto coef_fn, std::vector<Chunk>& chunks) {
        for (size_t i = 0; i < nzl.size();) {
            const int32_t b = nzl[i].block;
            size_t j = i;
            std::vector<std::pair<int32_t, std::pair<size_t, size_t>>> rows;  // row -> [first, last)
            while (j < nzl.size() && nzl[j].block == b) {
                size_t k2 = j;
                while (k2 < nzl.size() && nzl[k2].block == b && nzl[k2].row == nzl[j].row) ++k2;
                rows.push_back({nzl[j].row, {j, k2}});
                j = k2;
            }
            // split into chunks of at most kChunkRows rows, evenly in whole k-steps of four rows (18 rows: 12 + 6 = 3 + 2 k-steps;
            // cut in the middle, 9 + 9, they would pad to 3 + 3)
            const size_t nr = rows.size(), nks = (nr + 3) / 4, nch = (nks + kChunkRows / 4 - 1) / (kChunkRows / 4);
            for (size_t c = 0; c < nch; ++c) {
                const size_t r0 = std::min(nr, 4 * (c * nks / nch)), r1 = std::min(nr, 4 * ((c + 1) * nks / nch));
                Chunk ck;
                ck.block = b;
                ck.flags = block_flags(b) | (nch > 1 ? kChunkSplit : 0);
                ck.coef.assign((r1 - r0) * (size_t)d_out * kBlockWidth, 0.0);
                // order the rows so that every group of four (one DMMA k-step) can put four hot rows with different residues
                // modulo 4 first in their factor lists (pack_item then orders the factors): conflict-free A-fragment loads
                std::vector<size_t> order;
                {
                    std::vector<size_t> rest;
                    for (size_t r = r0; r < r1; ++r) rest.push_back(r);
                    while (!rest.empty()) {
                        unsigned used = 0;
                        for (int slot = 0; slot < 4 && !rest.empty(); ++slot) {
                            int best = -1, best_n = 99;  // the row with the fewest residues still free (but at least one)
                            for (size_t q = 0; q < rest.size(); ++q) {
                                const int n = __builtin_popcount(residues(rows[rest[q]].first) & ~used);
                                if (n > 0 && n < best_n) best = (int)q, best_n = n;
                            }
                            if (best < 0) best = 0;
                            const unsigned free_res = residues(rows[rest[(size_t)best]].first) & ~used;
                            used |= free_res & (~free_res + 1);
                            order.push_back(rest[(size_t)best]);
                            rest.erase(rest.begin() + best);
                        }
                    }
                }
                {   // .. then swap rows between k-steps while that removes wavefronts (arrange4 = the cost pack_item will realise)
                    std::vector<int32_t> ids;
                    for (size_t r : order) ids.push_back(rows[r].first);
                    const int nf = item_slots(ids), ng = (int)((ids.size() + 3) / 4);
                    int32_t scratch[4][8];
                    auto group_cost = [&](int g) { return arrange4(ids.data() + 4 * g, std::min<int>(4, (int)ids.size() - 4 * g), nf, scratch); };
                    std::vector<int> gc((size_t)ng);
                    int total = 0;
                    for (int g = 0; g < ng; ++g) total += gc[(size_t)g] = group_cost(g);
                    for (int pass = 0; pass < 3 && total > ng * nf && nf <= 4; ++pass) {
                        bool better = false;
                        for (size_t i1 = 0; i1 < ids.size(); ++i1)
                            for (size_t i2 = i1 + 1; i2 < ids.size(); ++i2) {
                                const int g1 = (int)(i1 / 4), g2 = (int)(i2 / 4);
                                if (g1 == g2 || (gc[(size_t)g1] == nf && gc[(size_t)g2] == nf)) continue;
                                std::swap(ids[i1], ids[i2]);
                                const int c1 = group_cost(g1), c2 = group_cost(g2);
                                if (c1 + c2 < gc[(size_t)g1] + gc[(size_t)g2]) {
                                    total += c1 + c2 - gc[(size_t)g1] - gc[(size_t)g2];
                                    gc[(size_t)g1] = c1, gc[(size_t)g2] = c2;
                                    std::swap(order[i1], order[i2]);
                                    better = true;
                                } else {
                                    std::swap(ids[i1], ids[i2]);
                                }
                            }
                        if (!better) break;
                    }
                }
                for (size_t pos = 0; pos < order.size(); ++pos) {
                    const size_t r = order[pos];
                    ck.rows.push_back(rows[r].first);
                    for (size_t q = rows[r].second.first; q < rows[r].second.second; ++q)
                        for (int64_t o = 0; o < d_out; ++o)
                            ck.coef[(pos * (size_t)d_out + o) * kBlockWidth + nzl[q].lane] = coef_fn(nzl[q].term, o);
                }
                chunks.push_back(std::move(ck));
            }
            i = j;
        }
    };
    std::vector<Chunk> chunks;
    build_chunks(nz, [&](int32_t t, int64_t o) { return (double)C[(size_t)t * d_out + o]; }, chunks);
    const std::vector<Chunk> value_chunks_by_block = chunks;  // (block order: the gradient's cold jobs walk them block by block)
#ifdef SMX_TUNING
    {   // timing experiments (results are WRONG on purpose; tuning builds only): drop classes of work items to measure what
        // each class costs.  SMX_ABL_DROP bits: 1 = cold items of at most SMX_ABL_THIN_ROWS rows, 2 = other cold items, 4 = hot
        const int drop = tune_int("SMX_ABL_DROP", 0), thin_rows = tune_int("SMX_ABL_THIN_ROWS", 4);
        if (drop) {
            std::vector<Chunk> kept;
            for (Chunk& ck : chunks) {
                const bool hot = ck.flags & kChunkHot, thin = !hot && (int)ck.rows.size() <= thin_rows && !(ck.flags & kChunkSplit);
                const int cls = hot ? 4 : thin ? 1 : 2;
                if (!(drop & cls)) kept.push_back(std::move(ck));
            }
            chunks.swap(kept);
        }
    }
#endif
    // Order: big (FP64-heavy) and small (streaming) items alternate, so that every warp of a CTA always has both kinds
    // in flight (warp w takes items w, w + n_warps, ..).
    std::stable_sort(chunks.begin(), chunks.end(), [](const Chunk& a, const Chunk& b) { return a.rows.size() > b.rows.size(); });
    {
        std::vector<Chunk> mixed;
        size_t lo = 0, hi = chunks.size();
        const size_t group = 12;  // = warps of the default CTA shape: every warp alternates big and small items
        bool front = true;
        while (lo < hi) {
            for (size_t g2 = 0; g2 < group && lo < hi; ++g2) mixed.push_back(std::move(front ? chunks[lo++] : chunks[--hi]));
            front = !front;
        }
        chunks.swap(mixed);
    }
    plan.n_chunks = (int32_t)chunks.size();
    plan.chunk_off.push_back(0);
    for (const Chunk& ck : chunks) {
        plan.chunk_block.push_back(ck.block);
        plan.chunk_flags.push_back(ck.flags);
        plan.chunk_rows.insert(plan.chunk_rows.end(), ck.rows.begin(), ck.rows.end());
        plan.coef.insert(plan.coef.end(), ck.coef.begin(), ck.coef.end());
        plan.chunk_off.push_back((int32_t)plan.chunk_rows.size());
    }

    // ---- 10. kernel-side packing: directory + one contiguous metadata record per work item ------------------------
    // item = (entry block `block`, `rows` table rows at rows_ptr, flags) -> dir (4 ints), meta (kMetaInts ints), fac2 (64 ints);
    // returns the k-mask.  Also records in the plan whether four resp. eight factors per row suffice.
    auto pack_item = [&](int32_t block, int32_t flags_in, const int32_t* rows_ptr, int32_t rows, const double* coef, int32_t first_slot,
                         int32_t* dir, int32_t* meta, int32_t* fac2, int32_t* flags_out) -> int32_t {
        const int32_t e0 = block * kBlockWidth;
        // which (k-step, half block) pairs carry any coefficient at all (for any output): the others are skipped
        int32_t kmask = 0;
        for (int32_t r = 0; r < rows; ++r)
            for (int64_t o = 0; o < d_out; ++o)
                for (int i = 0; i < kBlockWidth; ++i)
                    if (coef[((size_t)r * d_out + o) * kBlockWidth + i] != 0.0) {
                        // entry i belongs to n-tile j = (i >> 1) & 1 (lane mapping: entry = 4 * (n >> 1) + 2 * j + (n & 1))
                        kmask |= 1 << (2 * (r >> 2) + ((i >> 1) & 1));
                    }
        // every row slot as the product of up to four hot rows (row 0 = the ones row pads): the kernel variant without
        // product rows in the value table multiplies them on the fly.  nf = most factors of any row of the item.
        // The order of a row's factors is free (a product): per k-step (row slots 4 s .. 4 s + 3 = the four lanes of one
        // LDS.128) and factor position, the four hot rows are chosen with different residues modulo 4 where the rows allow it.
        int32_t nf = 1;
        (void)n_flat;
        for (int32_t r = 0; r < rows; ++r) {
            int32_t f[8];
            const int cnt = row_factors(rows_ptr[r], f);
            if (rows_ptr[r] >= flat_begin && cnt > 4) plan.flat_ok = false;  // five or more pairs: the four-factor kernels cannot run ..
            if (cnt == 0) plan.flat_ok = false, plan.deep_ok = false;          // .. more than eight: nor the eight-factor one
            nf = std::max(nf, cnt);
        }
        for (int32_t s0 = 0; s0 < kBlockWidth; s0 += 4) {
            int32_t placed[4][8];
            arrange4(rows_ptr + s0, std::max(0, std::min(4, rows - s0)), nf, placed);
            for (int q = 0; q < 4; ++q)
                for (int pos = 0; pos < 8; ++pos) (pos < 4 ? meta[96 + 4 * (s0 + q) + pos] : fac2[4 * (s0 + q) + pos - 4]) = placed[q][pos];
        }
        {
            int n_half[2] = {0, 0};
            for (int32_t r = 0; r < rows; ++r)
                for (int hb = 0; hb < 2; ++hb) {
                    bool any = false;
                    for (int64_t o = 0; o < d_out; ++o)
                        for (int i = 0; i < kBlockWidth; ++i)
                            if (((i >> 1) & 1) == hb && coef[((size_t)r * d_out + o) * kBlockWidth + i] != 0.0) any = true;
                    n_half[hb] += any;
                }
            plan.bank_stats[4] += __builtin_popcount((unsigned)kmask);
            plan.bank_stats[5] += (n_half[0] + 3) / 4 + (n_half[1] + 3) / 4;
        }
        for (int32_t s0 = 0; s0 < rows; s0 += 4) {  // statistics: shared-memory wavefronts (per quarter warp) of the factor loads
            ++plan.bank_stats[3];
            for (int pos = 0; pos < nf; ++pos) {
                int32_t v[4];
                for (int q = 0; q < 4; ++q) v[q] = pos < 4 ? meta[96 + 4 * (s0 + q) + pos] : fac2[4 * (s0 + q) + pos - 4];
                int worst = 0, worst_real = 0;
                for (int r4 = 0; r4 < 4; ++r4) {
                    int n = 0, n_real = 0;
                    for (int q = 0; q < 4; ++q) {
                        bool seen = false;
                        for (int q2 = 0; q2 < q; ++q2) seen = seen || v[q2] == v[q];
                        if (!seen && (v[q] & 3) == r4) ++n, n_real += v[q] != 0;
                    }
                    worst = std::max(worst, n), worst_real = std::max(worst_real, n_real);
                }
                ++plan.bank_stats[0], plan.bank_stats[1] += worst, plan.bank_stats[2] += std::max(1, worst_real);
            }
        }
        int32_t flags = flags_in;
        bool eta_zero = !(flags & kChunkHot);
        for (int i = 0; i < kBlockWidth; ++i) eta_zero = eta_zero && plan.ent_eta0[e0 + i] == 0.0;
        if (eta_zero) flags |= kChunkEtaZero;
        *flags_out = flags;
        dir[0] = first_slot, dir[1] = rows, dir[2] = flags | (nf << 8), dir[3] = plan.ent_dim[e0];
        double eta0[kBlockWidth];
        for (int i = 0; i < kBlockWidth; ++i) {
            meta[i] = plan.ent_tab[e0 + i];
            meta[16 + i] = plan.ent_deg[e0 + i];
            meta[32 + i] = plan.ent_eta[e0 + i];
            // row indices transposed for the kernel: position 4 * k + s holds row 4 * s + k (k-step s, A-fragment column k)
            meta[48 + 4 * (i & 3) + (i >> 2)] = i < rows ? rows_ptr[i] : 0;
            eta0[i] = plan.ent_eta0[e0 + i];
        }
        std::memcpy(meta + 64, eta0, sizeof(eta0));
        return kmask;
    };
    plan.chunk_dir.resize((size_t)plan.n_chunks * 4);
    plan.chunk_meta.assign((size_t)plan.n_chunks * kMetaInts, 0);
    plan.chunk_fac2.assign((size_t)plan.n_chunks * 64, 0);
    plan.chunk_kmask.assign((size_t)plan.n_chunks, 0);
    plan.padded_fma = 0;
    for (int32_t c = 0; c < plan.n_chunks; ++c) {
        const int32_t r0 = plan.chunk_off[c], rows = plan.chunk_off[c + 1] - r0;
        plan.chunk_kmask[c] = pack_item(plan.chunk_block[c], plan.chunk_flags[c], &plan.chunk_rows[r0], rows,
                                        &plan.coef[(size_t)r0 * d_out * kBlockWidth], r0, &plan.chunk_dir[(size_t)c * 4],
                                        &plan.chunk_meta[(size_t)c * kMetaInts], &plan.chunk_fac2[(size_t)c * 64], &plan.chunk_flags[c]);
        plan.padded_fma += 32 * __builtin_popcount((unsigned)plan.chunk_kmask[c]);
    }

    const double dense_blocks = (d_out + 7) / 8 == 4 ? 5.0 : (double)((d_out + 7) / 8);
    bool eta_zero = true;  // (non-zero first centres: the block-sparse side runs three outputs per pass, ~9 % cheaper per output)
    for (int32_t c = 0; c < plan.n_chunks; ++c)
        if (!(plan.chunk_flags[c] & (kChunkHot | kChunkEtaZero))) eta_zero = false;
    if (opt.dense && opt.dense_if_cheaper && plan.has_dense &&
        (double)plan.n_terms * (0.955 + 0.353 * dense_blocks) >= (eta_zero ? 0.1 : 0.091) * (double)plan.padded_fma * (double)d_out) {
        plan.has_dense = false;  // (PlanOptions::dense_if_cheaper: one pass per output of the block-sparse kernel is cheaper)
        plan.dense_k4 = 0;
        std::vector<int32_t>().swap(plan.dense_meta);
        std::vector<double>().swap(plan.dense_coef);
    }

    // ---- 11. gradient: jobs -----------------------------------------------------------------------------------------------
    // A job produces columns of J for every point of a tile and is run by ONE warp, so nothing is ever added to J:
    //   kind 0, one per cold block: the row sums acc[p][e] = sum_r C[r][e] m_r(p) of the block's value items ARE dI/dx of
    //           its 16 columns (pi_e = x - eta_0) - stored straight away (`target` = which of the 16 may be stored: real
    //           entries, and columns nobody uses; NOT the padding that belongs to a neighbouring block or lies beyond d_in);
    //   kind 1, one per hot dimension i: the derivative polynomial over the terms that carry a coefficient (section 7b), as
    //           work items of its own -> one number per point, column `target` = i.
    // Columns no job writes (dimensions without any entry) are listed in zero_cols and zeroed by the kernel.
    if (with_gradient && (!plan.grad_dims.empty() || plan.n_hot == 0)) {
        GradPlan& G = plan.grad;
        std::vector<char> written((size_t)d_in, 0);
        auto add_items = [&](const std::vector<Chunk>& cks, size_t begin, size_t end) {
            for (size_t c = begin; c < end; ++c) {
                const Chunk& ck = cks[c];
                const int32_t r0 = (int32_t)G.rows.size();
                G.rows.insert(G.rows.end(), ck.rows.begin(), ck.rows.end());
                G.coef.insert(G.coef.end(), ck.coef.begin(), ck.coef.end());
                G.item_off.push_back((int32_t)G.rows.size());
                G.item_dir.resize(G.item_dir.size() + 4);
                G.item_meta.resize(G.item_meta.size() + kMetaInts, 0);
                std::vector<int32_t> fac2(64, 0);
                int32_t flags = 0;
                pack_item(ck.block, ck.flags, ck.rows.data(), (int32_t)ck.rows.size(), ck.coef.data(), r0, &G.item_dir[G.item_dir.size() - 4],
                          &G.item_meta[G.item_meta.size() - kMetaInts], fac2.data(), &flags);
            }
        };
        G.item_off.push_back(0);
        G.job_off.push_back(0);
        for (size_t c = 0; c < value_chunks_by_block.size();) {  // kind 0
            const int32_t b = value_chunks_by_block[c].block;
            size_t c1 = c;
            while (c1 < value_chunks_by_block.size() && value_chunks_by_block[c1].block == b) ++c1;
            if (!(value_chunks_by_block[c].flags & kChunkHot)) {
                add_items(value_chunks_by_block, c, c1);
                const int32_t e0 = b * kBlockWidth, col0 = plan.ent_dim[e0];
                int32_t mask = 0;
                for (int i = 0; i < kBlockWidth; ++i) {
                    const int64_t col = (int64_t)col0 + i;
                    const bool real = plan.ent_deg[e0 + i] > 0;
                    // (a padding entry's column is this block's to store only if no entry at all lives there)
                    if (col < d_in && (real || maxdeg[col] == 0) && !written[col]) mask |= 1 << i, written[col] = 1;
                }
                // the two nodes of each column's degree-1 rule: the reference's gradient is NaN there
                for (int half = 0; half < 2; ++half)
                    for (int i = 0; i < kBlockWidth; ++i) {
                        const int64_t col = std::min<int64_t>((int64_t)col0 + i, d_in - 1);
                        const auto it = pair_id.find({(int)col, 1});
                        G.job_nodes.push_back(it == pair_id.end() ? std::nan("") : pairs[it->second].nodes[half]);
                    }
                G.job_kind.push_back(0);
                G.job_target.push_back(mask);
                G.job_off.push_back((int32_t)G.item_off.size() - 1);
            }
            c = c1;
        }
        for (size_t h = 0; h < plan.grad_dims.size(); ++h) {  // kind 1
            std::vector<Nz> nzh;
            for (const auto& kv : Cd[h])
                if (kv.first != 0) nzh.push_back(term_nz[kv.first]);
            std::sort(nzh.begin(), nzh.end(), nz_less);
            std::vector<Chunk> cks;
            const auto& Ch = Cd[h];
            build_chunks(nzh, [&](int32_t t, int64_t o) { return (double)Ch.at(t)[(size_t)o]; }, cks);
            for (Chunk& ck : cks) ck.flags &= ~kChunkSplit;  // (a hot-dimension job adds its items' contributions anyway)
            add_items(cks, 0, cks.size());
            G.job_kind.push_back(1);
            G.job_target.push_back(plan.grad_dims[h]);
            G.job_off.push_back((int32_t)G.item_off.size() - 1);
            written[plan.grad_dims[h]] = 1;
            const auto it0 = Ch.find(0);
            for (int64_t o = 0; o < d_out; ++o) G.job_c0.push_back(it0 == Ch.end() ? 0.0 : (double)it0->second[(size_t)o]);
        }
        // (kind 1 constants are indexed by job: pad the kind 0 jobs' slots)
        {
            std::vector<double> c0((size_t)G.job_kind.size() * d_out, 0.0);
            size_t at = 0;
            for (size_t jb = 0; jb < G.job_kind.size(); ++jb)
                if (G.job_kind[jb] == 1) {
                    std::copy_n(&G.job_c0[at], (size_t)d_out, &c0[jb * d_out]);
                    at += (size_t)d_out;
                }
            G.job_c0.swap(c0);
        }
        for (int64_t d = 0; d < d_in;) {  // column ranges nobody writes
            if (written[d]) {
                ++d;
                continue;
            }
            int64_t e = d;
            while (e < d_in && !written[e]) ++e;
            G.zero_cols.push_back((int32_t)d);
            G.zero_cols.push_back((int32_t)e);
            d = e;
        }
        G.present = true;
    }

    return finish_nodes();
}

void eval_plan_host(const FastPlan& plan, const double* x, int64_t N, int64_t ldx, double* y) {
    const int64_t d_out = plan.d_out;
    std::vector<double> tab((size_t)plan.n_tab, 1.0);
    std::vector<double> acc((size_t)d_out);
    auto pi = [&](const double* xp, int32_t e) {
        double v = 1.0;
        for (int k = 0; k < plan.ent_deg[e]; ++k) v *= (xp[plan.ent_dim[e]] - plan.eta[plan.ent_eta[e] + k]);
        return v;
    };
    for (int64_t p = 0; p < N; ++p) {
        const double* xp = x + p * ldx;
        tab[0] = 1.0;
        for (size_t e = 0; e < plan.ent_dim.size(); ++e)
            if (plan.ent_tab[e] > 0) tab[plan.ent_tab[e]] = pi(xp, (int32_t)e);
        for (int32_t t = 1 + plan.n_hot_rows; t < plan.n_tab; ++t)
            tab[t] = tab[plan.tab_parent[t - 1 - plan.n_hot_rows]] * tab[plan.tab_hot[t - 1 - plan.n_hot_rows]];
        double* yp = y + p * d_out;
        for (int64_t o = 0; o < d_out; ++o) yp[o] = plan.c0[o];
        for (int32_t c = 0; c < plan.n_chunks; ++c) {
            const int32_t b = plan.chunk_block[c];
            for (int lane = 0; lane < kBlockWidth; ++lane) {
                const int32_t e = b * kBlockWidth + lane;
                const double v = (plan.chunk_flags[c] & kChunkHot) ? tab[plan.ent_tab[e]] : pi(xp, e);
                std::fill(acc.begin(), acc.end(), 0.0);
                for (int32_t i = plan.chunk_off[c]; i < plan.chunk_off[c + 1]; ++i)
                    for (int64_t o = 0; o < d_out; ++o)
                        acc[o] = std::fma(plan.coef[((size_t)i * plan.n_sets + o) * kBlockWidth + lane], tab[plan.chunk_rows[i]], acc[o]);
                for (int64_t o = 0; o < d_out; ++o) yp[o] = std::fma(v, acc[o], yp[o]);
            }
        }
    }
}

void eval_plan_dense_host(const FastPlan& plan, const double* x, int64_t N, int64_t ldx, double* y) {
    const int64_t d_out = plan.d_out;
    std::vector<double> tab((size_t)plan.n_tab, 1.0);
    auto pi = [&](const double* xp, int32_t e) {
        double v = 1.0;
        for (int k = 0; k < plan.ent_deg[e]; ++k) v *= (xp[plan.ent_dim[e]] - plan.eta[plan.ent_eta[e] + k]);
        return v;
    };
    for (int64_t p = 0; p < N; ++p) {
        const double* xp = x + p * ldx;
        tab[0] = 1.0;
        for (size_t e = 0; e < plan.ent_dim.size(); ++e)
            if (plan.ent_tab[e] > 0) tab[plan.ent_tab[e]] = pi(xp, (int32_t)e);
        for (int32_t t = 1 + plan.n_hot_rows; t < plan.n_tab; ++t)
            tab[t] = tab[plan.tab_parent[t - 1 - plan.n_hot_rows]] * tab[plan.tab_hot[t - 1 - plan.n_hot_rows]];
        double* yp = y + p * d_out;
        for (int64_t o = 0; o < d_out; ++o) {
            double acc = 0.0;
            for (int64_t i = 0; i < 4 * (int64_t)plan.dense_k4; ++i) {
                const int32_t ia = plan.dense_meta[2 * i], ib = plan.dense_meta[2 * i + 1];
                const double phi = tab[ia] * (ib >= 0 ? tab[ib] : xp[-1 - ib] - plan.dense_eta0[-1 - ib]);
                acc = std::fma(phi, plan.dense_coef[(size_t)(((o >> 3) * (plan.dense_k4 + kDensePadK4) + (i >> 2)) * 32 + 4 * (o & 7) + (i & 3))], acc);
            }
            yp[o] = plan.c0[o] + acc;
        }
    }
}

void eval_plan_gradient_host(const FastPlan& plan, const double* x, int64_t N, int64_t ldx, double* J) {
    // the gradient jobs of the plan, evaluated in fp64 in the order of the kernel (smx_grad_kernel.cu); finite at the nodes
    const int64_t d_out = plan.d_out, d_in = plan.d_in;
    const GradPlan& G = plan.grad;
    std::vector<double> tab((size_t)plan.n_tab, 1.0);
    auto pi = [&](const double* xp, int32_t e) {
        double v = 1.0;
        for (int k = 0; k < plan.ent_deg[e]; ++k) v *= (xp[plan.ent_dim[e]] - plan.eta[plan.ent_eta[e] + k]);
        return v;
    };
    for (int64_t p = 0; p < N; ++p) {
        const double* xp = x + p * ldx;
        double* Jp = J + p * d_out * d_in;
        std::fill(Jp, Jp + d_out * d_in, std::nan(""));  // every entry must be written by exactly one job (or zero_cols)
        for (size_t e = 0; e < plan.ent_dim.size(); ++e)
            if (plan.ent_tab[e] > 0) tab[plan.ent_tab[e]] = pi(xp, (int32_t)e);
        for (int32_t t = 1 + plan.n_hot_rows; t < plan.n_tab; ++t)
            tab[t] = tab[plan.tab_parent[t - 1 - plan.n_hot_rows]] * tab[plan.tab_hot[t - 1 - plan.n_hot_rows]];
        for (size_t z = 0; z + 1 < G.zero_cols.size(); z += 2)
            for (int64_t o = 0; o < d_out; ++o)
                for (int32_t c = G.zero_cols[z]; c < G.zero_cols[z + 1]; ++c) Jp[o * d_in + c] = 0.0;
        for (size_t jb = 0; jb + 1 < G.job_off.size(); ++jb) {
            for (int64_t o = 0; o < d_out; ++o) {
                double acc16[kBlockWidth] = {0.0};
                double tot = G.job_c0[jb * d_out + o];
                int32_t col0 = 0;
                for (int32_t it = G.job_off[jb]; it < G.job_off[jb + 1]; ++it) {
                    const int32_t* dir = &G.item_dir[(size_t)it * 4];
                    const int32_t* meta = &G.item_meta[(size_t)it * kMetaInts];
                    const bool hot = dir[2] & kChunkHot;
                    col0 = dir[3];
                    for (int lane = 0; lane < kBlockWidth; ++lane) {
                        double acc = 0.0;
                        for (int32_t i = G.item_off[it]; i < G.item_off[it + 1]; ++i)
                            acc = std::fma(G.coef[((size_t)i * d_out + o) * kBlockWidth + lane], tab[G.rows[i]], acc);
                        if (G.job_kind[jb] == 0) {
                            acc16[lane] += acc;
                        } else {
                            double eta0;
                            std::memcpy(&eta0, meta + 64 + 2 * lane, sizeof(double));
                            const double v = hot ? tab[meta[lane]] : (meta[16 + lane] > 0 ? xp[std::min<int64_t>(col0 + lane, d_in - 1)] - eta0 : 1.0);
                            tot = std::fma(v, acc, tot);
                        }
                    }
                }
                if (G.job_kind[jb] == 0) {
                    for (int lane = 0; lane < kBlockWidth; ++lane)
                        if (G.job_target[jb] >> lane & 1) Jp[o * d_in + col0 + lane] = acc16[lane];
                } else {
                    Jp[o * d_in + G.job_target[jb]] = tot;
                }
            }
        }
    }
}

}  // namespace smx

// ---- C entry points for the CPU-only test-suite (exported from libsmolyax_host.so) --------------------------------
extern "C" {

struct smxh_group {
    int32_t n;
    int64_t nn;
    const int64_t* tau;
    const double* F;
    const double* nodes;
    const double* weights;
    const int64_t* dims;
    const int64_t* degs;
    const int64_t* zetas;
    const double* quad;
};

static thread_local std::string g_plan_error;
const char* smxh_plan_error() { return g_plan_error.c_str(); }

// option bits: 1 = derivative jobs, 2 = block-sparse form, 4 = dense form
static smx::PlanOptions plan_options(int32_t bits) {
    smx::PlanOptions o;
    o.gradient = bits & 1, o.sparse = bits & 2, o.dense = bits & 4;
    return o;
}

void* smxh_plan_build_opt(int64_t d_in, int64_t d_out, const double* offset, int32_t n_groups, const smxh_group* groups,
                          int32_t option_bits) {
    std::vector<smx::GroupView> gv;
    for (int32_t g = 0; g < n_groups; ++g) {
        smx::GroupView v;
        v.n = groups[g].n;
        v.nn = groups[g].nn;
        v.tau.assign(groups[g].tau, groups[g].tau + groups[g].n);
        v.F = groups[g].F;
        v.nodes = groups[g].nodes;
        v.weights = groups[g].weights;
        v.dims = groups[g].dims;
        v.degs = groups[g].degs;
        v.zetas = groups[g].zetas;
        v.quad = groups[g].quad;
        gv.push_back(v);
    }
    auto* plan = new smx::FastPlan();
    g_plan_error = smx::build_fast_plan(d_in, d_out, offset, gv, *plan, plan_options(option_bits));
    if (!g_plan_error.empty()) {
        delete plan;
        return nullptr;
    }
    return plan;
}
void* smxh_plan_build(int64_t d_in, int64_t d_out, const double* offset, int32_t n_groups, const smxh_group* groups) {
    return smxh_plan_build_opt(d_in, d_out, offset, n_groups, groups, 3);
}
// Same field order as smx_compact_desc (include/smolyax_b200.h) and smx::CompactView.
struct smxh_compact {
    int64_t n_summands;
    const int32_t* n_active;
    const int64_t* slot_off;
    const int64_t* dims;
    const int64_t* degs;
    const int64_t* node_off;
    const double* node_pool;
    const double* quad_pool;
    const int64_t* zetas;
    const int64_t* val_off;
    const int64_t* val_index;
    const double* values;
    int64_t n_values;
};
static smx::CompactView compact_view(const smxh_compact* c) {
    smx::CompactView v;
    v.n_summands = c->n_summands;
    v.n_active = c->n_active;
    v.slot_off = c->slot_off;
    v.dims = c->dims;
    v.degs = c->degs;
    v.node_off = c->node_off;
    v.node_pool = c->node_pool;
    v.quad_pool = c->quad_pool;
    v.zetas = c->zetas;
    v.val_off = c->val_off;
    v.val_index = c->val_index;
    v.values = c->values;
    v.n_values = c->n_values;
    return v;
}
void* smxh_plan_build_compact(int64_t d_in, int64_t d_out, const double* offset, const smxh_compact* desc, int32_t option_bits) {
    auto* plan = new smx::FastPlan();
    g_plan_error = smx::build_fast_plan_compact(d_in, d_out, offset, compact_view(desc), *plan, plan_options(option_bits));
    if (!g_plan_error.empty()) {
        delete plan;
        return nullptr;
    }
    return plan;
}
int smxh_integrate_compact(int64_t d_out, const double* offset, const smxh_compact* desc, double* Q) {
    std::vector<double> q;
    g_plan_error = smx::integrate_compact(d_out, offset, compact_view(desc), q);
    if (!g_plan_error.empty()) return 1;
    std::memcpy(Q, q.data(), sizeof(double) * (size_t)d_out);
    return 0;
}
void smxh_plan_free(void* p) { delete static_cast<smx::FastPlan*>(p); }
// stats: [n_terms, n_entries, n_rows, n_hot, n_chunks, padded_fma, n_levels, nested, n_summands, w_raw, w_pad]
void smxh_plan_stats(void* p, int64_t* out) {
    auto* pl = static_cast<smx::FastPlan*>(p);
    int64_t v[11] = {pl->n_terms, pl->n_entries, pl->n_rows, pl->n_hot, pl->n_chunks, pl->padded_fma,
                     pl->n_levels, pl->nested ? 1 : 0, pl->n_summands, pl->w_raw, pl->w_pad};
    (void)pl->n_tab;
    std::memcpy(out, v, sizeof(v));
}
void smxh_plan_bank_stats(void* p, int64_t* out) { std::memcpy(out, static_cast<smx::FastPlan*>(p)->bank_stats, 6 * sizeof(int64_t)); }
// (test hook) deal_blocks of smx_plan.h: out = jb0[nw], nbv[nw], hblk[nw], hsel[nw]
void smxh_deal_blocks(int32_t nw, int32_t nb, const int32_t* pipe, int32_t group_first, int32_t group_blocks, int32_t ticket, int32_t* out) {
    smx::deal_blocks(nw, nb, pipe, group_first, group_blocks, ticket, out, out + nw, out + 2 * nw, out + 3 * nw);
}
// gradient jobs: out = [n_jobs, n_items, k-steps of all items, cold jobs, most items of a job, most k-steps of a job, zero ranges]
void smxh_plan_grad_stats(void* p, int64_t* out) {
    auto* pl = static_cast<smx::FastPlan*>(p);
    const smx::GradPlan& G = pl->grad;
    int64_t ks = 0, cold = 0, maxi = 0, maxk = 0;
    for (size_t j = 0; j + 1 < G.job_off.size(); ++j) {
        int64_t k = 0;
        for (int32_t it = G.job_off[j]; it < G.job_off[j + 1]; ++it) k += (G.item_off[it + 1] - G.item_off[it] + 3) / 4;
        ks += k;
        cold += G.job_kind[j] == 0;
        maxi = std::max<int64_t>(maxi, G.job_off[j + 1] - G.job_off[j]);
        maxk = std::max(maxk, k);
    }
    int64_t v[7] = {(int64_t)G.job_kind.size(), (int64_t)G.item_off.size() - 1, ks, cold, maxi, maxk, (int64_t)G.zero_cols.size() / 2};
    std::memcpy(out, v, sizeof(v));
}
// per work item: rows, flags, kmask (statistics for tests and tuning); returns the number of items
int32_t smxh_plan_items(void* p, int32_t* out, int32_t capacity) {
    auto* pl = static_cast<smx::FastPlan*>(p);
    for (int32_t c = 0; c < pl->n_chunks && c < capacity; ++c) {
        out[3 * c] = pl->chunk_off[c + 1] - pl->chunk_off[c];
        out[3 * c + 1] = pl->chunk_flags[c];
        out[3 * c + 2] = pl->chunk_kmask[c];
    }
    return pl->n_chunks;
}
// Verification aid, CPU tests only (see smx_plan.h).
void smxh_plan_eval_host(void* p, const double* x, int64_t N, int64_t ldx, double* y) {
    smx::eval_plan_host(*static_cast<smx::FastPlan*>(p), x, N, ldx, y);
}
void smxh_plan_eval_dense_host(void* p, const double* x, int64_t N, int64_t ldx, double* y) {
    smx::eval_plan_dense_host(*static_cast<smx::FastPlan*>(p), x, N, ldx, y);
}
void smxh_plan_gradient_host(void* p, const double* x, int64_t N, int64_t ldx, double* J) {
    smx::eval_plan_gradient_host(*static_cast<smx::FastPlan*>(p), x, N, ldx, J);
}
}
