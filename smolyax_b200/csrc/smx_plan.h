// Host-side "plan": the device layout of the fast evaluation path, built from the reference's per-group layout.
//
// Mathematics.  The Smolyak interpolant  I(x) = sum_nu zeta_nu (x)_j I^{nu_j}[f](x)
// (reference interpolation.py:283-302) is a polynomial in span{ prod_j x_j^{a_j} : a in Lambda }.  The plan
// re-expresses it, exactly, in the product basis of monic Newton polynomials
//        pi_{j,a}(x_j) = prod_{i<a} (x_j - eta_{j,i}),      Psi_alpha(x) = prod_{(j,a) in alpha} pi_{j,a}(x_j),
//        I(x) = c_0 + sum_{alpha in Lambda, alpha != 0} c_alpha Psi_alpha(x),
// where eta_j are the interpolation nodes of dimension j themselves (the nested sequence, or the Leja-ordered
// nodes of the highest degree for non-nested rules).  The coefficients come from the value tensors F by the
// 1-D change of basis "nodal values at the (deg+1) barycentric nodes -> Newton coefficients" along every axis,
// times zeta, accumulated over summands in long double.  Each term is then split into its *leading entry*
// (the pair with the largest dimension) and its *hot part* (the rest):
//        I(x) = c_0 + sum_{e=(j,a)} pi_e(x_j) * sum_r  C[r][e] * m_r(x),       m_r = prod of the hot part r,
// a block-sparse (rows r x entries e) contraction that the kernel in smx_fast.cu evaluates with the entries of a
// block across lanes, the points in registers, and the row products m_r in shared memory.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

namespace smx {

// Tuning knobs are read from the environment ONLY in builds with -DSMX_TUNING (SMX_TUNING=1 python -m smolyax_b200._build,
// used by benchmarks/ for A/B timing).  The product library ignores the environment: every knob has its measured default.
inline int tune_int(const char* name, int dflt) {
#ifdef SMX_TUNING
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
#else
    (void)name;
    return dflt;
#endif
}
inline const char* tune_str(const char* name) {
#ifdef SMX_TUNING
    return std::getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}

struct GroupView {
    int n = 0;
    int64_t nn = 0;
    std::vector<int64_t> tau;  // length n
    const double* F = nullptr;
    const double* nodes = nullptr;
    const double* weights = nullptr;
    const int64_t* dims = nullptr;
    const int64_t* degs = nullptr;
    const int64_t* zetas = nullptr;
    const double* quad = nullptr;
    int64_t fsize() const {
        int64_t f = 1;
        for (auto t : tau) f *= (t + 1);
        return f;
    }
    int64_t tw() const {
        int64_t m = 0;
        for (auto t : tau) m = t > m ? t : m;
        return m + 1;
    }
};

// The same information in compact form (smx_create_compact): summands as CSR over their slots, values as rows of one
// (n_values, d_out) table shared by all summands (nested rules: one row per sparse-grid node).
struct CompactView {
    int64_t n_summands = 0;
    const int32_t* n_active = nullptr;   // (n_summands)
    const int64_t* slot_off = nullptr;   // (n_summands + 1)
    const int64_t* dims = nullptr;       // (slots)
    const int64_t* degs = nullptr;       // (slots)
    const int64_t* node_off = nullptr;   // (slots) offset of the slot's deg+1 nodes in node_pool (and weights in quad_pool)
    const double* node_pool = nullptr;
    const double* quad_pool = nullptr;   // optional
    const int64_t* zetas = nullptr;      // (n_summands)
    const int64_t* val_off = nullptr;    // (n_summands + 1)
    const int64_t* val_index = nullptr;  // per entry of the exact-shape value tensor (C order over the slots): row of `values`
    const double* values = nullptr;      // (n_values, d_out)
    int64_t n_values = 0;
};

constexpr int kBlockWidth = 16;    // entries per block: 4 lane-groups x 4 entries per lane
constexpr int kChunkRows = 16;     // rows per work item
constexpr int kMaxLevels = 16;

constexpr int kMetaInts = 4 * 16 + 2 * 16 + 4 * 16;  // 640 bytes
constexpr int kDenseStageK4 = 16;  // dense form: the term list is padded to whole stages of 16 k-steps (64 terms) ..
constexpr int kDensePadK4 = 64;    // .. and every array carries four more stages of zeros (look-ahead without bounds checks)
constexpr int kTabPitch = 36;      // doubles per value-table row in shared memory (32 points + 32 bytes of skew:
                                   // rows r, r' with r != r' (mod 4) never share a bank in the DMMA A-fragment loads)

// chunk flags
constexpr int kChunkHot = 1;      // all 16 entries are hot: their basis values are rows of the value table
constexpr int kChunkContig = 2;   // 16 degree-1 entries on consecutive columns of x starting at an even column
constexpr int kChunkInside = 4;   // .. and the 16 columns all exist (first column + 16 <= d_in)
constexpr int kChunkSplit = 8;    // the block's rows are spread over more than one item
constexpr int kChunkEtaZero = 16; // cold block whose first centres are all zero (default Leja domain): pi = x

// Gradient of the fast path (smx_plan.cpp section 11, kernel smx_grad_kernel.cu): jobs, each run by one warp, each writing its
// own columns of J.  Items are laid out like the value path's (entry block x <= 16 table rows, coefficients [row][output][16]).
struct GradPlan {
    bool present = false;
    std::vector<int32_t> item_off;    // size n_items + 1: row slots of item i are [item_off[i], item_off[i + 1])
    std::vector<int32_t> rows;        // value-table row of every row slot
    std::vector<double> coef;         // [row slot][d_out][kBlockWidth]
    std::vector<int32_t> item_dir;    // 4 ints per item: first row slot, rows, flags | nf << 8, first column of x
    std::vector<int32_t> item_meta;   // kMetaInts per item (as FastPlan::chunk_meta)
    std::vector<int32_t> job_off;     // size n_jobs + 1: items of job j
    std::vector<int32_t> job_kind;    // 0: cold block (row sums = derivatives w.r.t. its 16 columns), 1: hot dimension
    std::vector<int32_t> job_target;  // kind 0: bit e set = column (first column + e) is stored by this job; kind 1: the dimension
    std::vector<double> job_c0;       // [job][d_out] constant term of a hot dimension's derivative polynomial (kind 0: zeros)
    std::vector<double> job_nodes;    // kind 0 jobs, in order: 32 doubles each - node 0 and node 1 of the 16 columns' degree-1 rule
    std::vector<int32_t> zero_cols;   // pairs [lo, hi): columns of J that no job writes (dimensions without any entry)
};

struct FastPlan {
    int64_t d_in = 0, d_out = 0;
    bool nested = false;
    int64_t n_summands = 0, w_raw = 0, w_pad = 0, n_terms = 0;

    // Leading entries (dimension, degree), sorted by dimension then degree; the *hot* entries (every entry of a
    // dimension that occurs inside some hot part) form a prefix, padded with degree-0 dummies to a block boundary;
    // cold blocks are padded so that they start at an even column of x.  Dummies have zero coefficients.
    int32_t n_entries = 0;               // real entries
    int32_t n_hot = 0;                   // hot prefix (real entries)
    int32_t n_hot_rows = 0;              // value-table rows reserved for the hot entries (n_hot rounded up to a block)
    int32_t hot_dims = 0;                // the hot entries live on columns [0, hot_dims) of x
    std::vector<int32_t> ent_dim;        // column of x (dummies: a valid column)
    std::vector<int32_t> ent_deg;        // a >= 1, 0 for dummies (pi = 1)
    std::vector<int32_t> ent_eta;        // offset of the dimension's centres in `eta`
    std::vector<int32_t> ent_tab;        // row of the value table holding pi_e (hot entries), 0 otherwise
    std::vector<double> ent_eta0;        // first centre (pi_{j,1}(x) = x - eta0)
    std::vector<double> eta;             // centres, concatenated per dimension

    // Value table, one row of 32 points per index:  [0] = 1,  [1 + hot_row(e)] = pi of hot entry e.  The shared-memory bank
    // group of a row is its index modulo 4 (kTabPitch); hot_row() rotates every group of four entries by its number, so
    // that BOTH four consecutive entries (the cheapest hot pairs: the rows the cold items multiply with, one per lane of
    // an A-fragment load) AND the entries e, e + 4, e + 8, e + 12 of a block (what the four lanes of a hot item read as
    // "their e-th entry") sit in four different bank groups (measured before: 18 % of all shared-memory wavefronts of the
    // headline kernel were conflicts of the first kind),
    // [1 + n_hot_rows ..) = products of >= 2 hot pairs ("rows" of level >= 2), each parent * hot:
    int32_t n_tab = 1;
    int32_t n_levels = 1;                // highest number of pairs in a hot part, plus one
    std::vector<int32_t> tab_parent;     // for index >= n_hot+1 (level >= 2): table indices of the two factors
    std::vector<int32_t> tab_hot;
    std::vector<int32_t> level_off;      // table indices of level l (l >= 2) are [level_off[l], level_off[l+1])
    std::vector<int32_t> tab_factors;    // 4 ints per row of level 2..4: the hot rows whose product it is (0 = the ones row)

    // work items: (entry block, slice of its row list); coefficients [row slot][d_out][kBlockWidth]
    int32_t n_chunks = 0;
    std::vector<int32_t> chunk_block;
    std::vector<int32_t> chunk_flags;
    std::vector<int32_t> chunk_off;      // size n_chunks+1, offsets into chunk_rows / coefficient row slots
    std::vector<int32_t> chunk_rows;     // value-table index of every row slot
    std::vector<double> coef;            // [row slot][output][kBlockWidth]
    int32_t n_sets = 0;                  // = d_out (coefficient sets per row slot)
    std::vector<int32_t> grad_dims;      // the hot dimensions with an entry: their derivatives are jobs of kind 1 in `grad`
    // per dimension, the nodes at which the reference's gradient is NaN (barycentric.py:152-154): CSR over dimensions
    std::vector<int32_t> nan_off;
    std::vector<double> nan_nodes;
    // the same information packed for the kernel: one directory entry and one metadata record per work item
    std::vector<int32_t> chunk_dir;      // 4 ints per item: first row slot, number of rows, flags | nf << 8 (nf = most hot
                                         // factors of any of its rows), first column of x
    bool flat_ok = true;                 // every row is a product of at most four hot rows (factor lists in the metadata)
    bool deep_ok = true;                 // .. of at most eight: the "deep" records carry a second factor list per row slot
    std::vector<int32_t> tab_factors8;   // 8 ints per row of level 2..8: its hot rows (0 = the ones row)
    std::vector<int32_t> chunk_fac2;     // 64 ints per item: factors 5..8 of row slot i at [4 * i .. 4 * i + 3] (0 = ones row)
    std::vector<int32_t> chunk_kmask;    // bit 2 s + j set: k-step s has a non-zero coefficient in entries 8 j .. 8 j + 7
    std::vector<int32_t> chunk_meta;     // kMetaInts ints per item: tab[16], deg[16], eta offset[16], row index[16], eta0[16]
                                         // (doubles), then per row slot the four hot rows whose product it is (int4)
    std::vector<int32_t> hot_off;        // prefix sums of the degrees of the hot dimensions, size hot_dims + 1
    std::vector<int32_t> hot_pos;        // entry index of hot pair (d, a), indexed hot_off[d] + a - 1 (hot entries are
                                         // stored degree-major: blocks of equal degree share most of their rows)
    std::vector<double> c0;              // (d_out) constant term, includes the offset
    int64_t bank_stats[6] = {0, 0, 0, 0, 0, 0};  // A-fragment loads of all items: [loads, wavefront groups, the same without the ones row, k-steps,
                                                //  (k-step, half block) pairs executed, fewest possible for the items' rows]
    int64_t padded_fma = 0;              // FMAs per point and output the kernel executes: 32 per non-empty (k-step, half block)
    int32_t n_rows = 0;                  // distinct hot parts (statistics)
    bool has_sparse = false;             // work items + coefficient sets above are filled
    GradPlan grad;                       // derivative jobs (opt.gradient)
    double newton_error = 0.0;           // worst fp64 error of a cardinal function through its Newton form (conditioning check)

    // Dense (GEMM-regime) form for large d_out:  y = c0 + Phi(x) C  with one column of Phi per term,
    // Phi[p][t] = tab[hot part of t][p] * pi_{leading entry of t}(x_p)  and  C (terms x d_out).  Terms are ordered hot
    // entries first (by block, row), then by column of x; padded with zero rows to whole k-steps of 4 terms.
    bool has_dense = false;
    int32_t dense_k4 = 0;                // k-steps, a multiple of kDenseStageK4; the arrays hold kDensePadK4 more (zeros)
    std::vector<int32_t> dense_meta;     // 2 ints per term: table row of the hot part; table row of the leading entry
                                         // (hot) or  -1 - column of x  (cold: pi = x - eta0[column])
    std::vector<double> dense_eta0;      // (d_in) first centre of every dimension
    std::vector<double> dense_coef;      // [ceil(d_out / 8)][dense_k4 + kDensePadK4][32]: DMMA B fragments, lane = 4 * gid + tig
                                         // holds C[4 * k4 + tig][8 * jb + gid]
};

struct PlanOptions {
    bool gradient = true;   // derivative jobs (GradPlan; needs the block-sparse form)
    bool sparse = true;     // block-sparse work items (K1; values and gradients)
    bool dense = false;     // dense term matrix (K2; values, large d_out)
    // dense asked for by the d_out rule only (8 <= d_out < 32): kept if it is the cheaper form for THIS index set, dropped
    // otherwise.  Measured on B200 (benchmarks/crossover.py, profiles/r08_crossover.txt; cfg2 and cfg4 tables), seconds per
    // point: split-K dense kernel  n_terms * (0.955e-12 + 3.53e-13 * blocks)  (assembling A, plus the DMMAs of ceil(d_out / 8)
    // blocks of 8 outputs; the 16-warp x 4-block instantiation behaves like 5 blocks), block-sparse kernel, one pass per
    // output:  1.0e-13 * padded_fma * d_out  (0.91e-13 for plans with non-zero first centres, which run three outputs per pass).
    bool dense_if_cheaper = false;
};

// ---- staged dense kernel: which output blocks a warp of the CTA multiplies (smx_dense_kernel.cu) --------------------------------
// A column group of `group_blocks` blocks (8 outputs each) that does not fill the CTA (nw warps x nb blocks) is dealt per FP64
// pipe: pipe[w] = %warpid & 3 of warp w.  Every pipe gets floor(units / 4) units, the remainder goes to a run of pipes that
// starts at ticket * remainder (consecutive tickets of one SM continue where the previous CTA stopped); unit = a block, or - for
// nb = 2 - half a block (the points 0..15 or 16..31 of the tile).  Warp w then has nbv[w] full blocks jb0[w] .. and, if
// hblk[w] >= 0, half hsel[w] of block hblk[w].  Every block of the group is covered exactly once whatever the placement of the
// warps (tests/test_plan.py runs this function on the host over all group sizes, tickets and placements).
#ifdef __CUDACC__
__host__ __device__
#endif
inline void deal_blocks(int nw, int nb, const int* pipe, int group_first, int group_blocks, int ticket, int* jb0, int* nbv, int* hblk,
                        int* hsel) {
    int want[4], cap[4] = {0, 0, 0, 0};
    for (int w = 0; w < nw; ++w) cap[pipe[w] & 3] += nb;
    for (int q = 0; q < 4; ++q) want[q] = group_blocks / 4;
    const int rem = group_blocks % 4, first = ticket * rem;
    for (int r = 0; r < rem; ++r) ++want[(first + r) & 3];
    int left = 0;  // what a pipe cannot take (fewer warps of this CTA on it than on the others) goes where there is room
    for (int q = 0; q < 4; ++q)
        if (want[q] > cap[q]) left += want[q] - cap[q], want[q] = cap[q];
    for (int q = 0; q < 4 && left > 0; ++q) {
        const int t = (first + rem + q) & 3, room = cap[t] - want[t], add = left < room ? left : room;
        want[t] += add, left -= add;
    }
    int at = group_first;
    for (int w = 0; w < nw; ++w) {  // a pipe's share goes to its warps in warp order, nb blocks at most each
        const int q = pipe[w] & 3, n = nb < want[q] ? nb : want[q];
        want[q] -= n;
        jb0[w] = at, nbv[w] = n, hblk[w] = -1, hsel[w] = 0;
        at += n;
    }
    if (nb != 2 || nw > 16) return;
    // the same in halves; kept only if it works out (every half placed, capacities respected)
    int h[4], fullrem[4], halfrem[4], jb[16], nn[16], hb[16], hs[16];
    const int total = 2 * group_blocks, rem_h = total % 4, first_h = ticket * rem_h;
    bool ok = true;
    for (int q = 0; q < 4; ++q) h[q] = total / 4;
    for (int r = 0; r < rem_h; ++r) ++h[(first_h + r) & 3];
    int n_half = 0, n_full = 0;
    for (int q = 0; q < 4; ++q) {
        ok = ok && h[q] <= 2 * cap[q];
        fullrem[q] = h[q] / 2, halfrem[q] = h[q] & 1;
        n_full += fullrem[q], n_half += halfrem[q];
    }
    ok = ok && (n_half % 2 == 0) && n_full + n_half / 2 == group_blocks;
    int at_full = group_first, placed = 0;
    for (int w = 0; w < nw && ok; ++w) {
        const int q = pipe[w] & 3, n = nb < fullrem[q] ? nb : fullrem[q];
        fullrem[q] -= n;
        jb[w] = at_full, nn[w] = n, hb[w] = -1, hs[w] = 0;
        at_full += n;
        if (halfrem[q] && n < nb) {  // split blocks are the last ones of the group; their halves alternate
            hb[w] = group_first + n_full + placed / 2, hs[w] = placed & 1;
            ++placed, halfrem[q] = 0;
        }
    }
    for (int q = 0; q < 4; ++q) ok = ok && fullrem[q] == 0 && halfrem[q] == 0;
    if (ok && placed == n_half)
        for (int w = 0; w < nw; ++w) jb0[w] = jb[w], nbv[w] = nn[w], hblk[w] = hb[w], hsel[w] = hs[w];
}

// value-table row (minus one) of hot entry h
#ifdef __CUDACC__
__host__ __device__
#endif
inline int32_t hot_row(int32_t h) { return (h & ~3) | (((h >> 2) + h) & 3); }

// Builds the plan.  Returns "" on success, otherwise an error message (invalid layout, singular node set ..).
std::string build_fast_plan(int64_t d_in, int64_t d_out, const double* offset, const std::vector<GroupView>& groups,
                            FastPlan& plan, const PlanOptions& opt = PlanOptions());

std::string build_fast_plan_compact(int64_t d_in, int64_t d_out, const double* offset, const CompactView& cv, FastPlan& plan,
                                    const PlanOptions& opt = PlanOptions());
// Smolyak quadrature of a compact descriptor on the host, in long double (the integral does not depend on x).
std::string integrate_compact(int64_t d_out, const double* offset, const CompactView& cv, std::vector<double>& Q);
// same for the reference's per-group layout (needs the quadrature tables of the groups)
std::string integrate_groups(int64_t d_out, const double* offset, const std::vector<GroupView>& groups, std::vector<double>& Q);

// Verification aid for the CPU-only test-suite: evaluates the plan on the host in fp64 in the same order as
// the kernel.  NOT a product path — nothing in smolyax_b200/ calls it; see tests/test_plan.py.
void eval_plan_host(const FastPlan& plan, const double* x, int64_t N, int64_t ldx, double* y);
// same for the dense form (mirrors the arithmetic of the GEMM-regime kernel)
void eval_plan_dense_host(const FastPlan& plan, const double* x, int64_t N, int64_t ldx, double* y);
// same for the gradient sets: J (N, d_out, d_in), finite at nodes
void eval_plan_gradient_host(const FastPlan& plan, const double* x, int64_t N, int64_t ldx, double* J);

}  // namespace smx
