/*
 * smolyax_b200 — C ABI of the B200 (sm_100a) evaluation path of the barycentric Smolyak interpolant.
 *
 * The reference (JoWestermann/smolyax) has no FFI layer: its boundary is the Python class
 * SmolyakBarycentricInterpolator plus one internal seam, the two jit(vmap(..)) callables created at
 * interpolation.py:243-248 and invoked per group at interpolation.py:293-301 / :334-342.  This header is what a
 * binding for that path would bind (ctypes in smolyax_b200/_lib.py; a jax.ffi / cgo / JNI stub is shown in
 * INTEGRATION.md).  Plain pointers and sizes only; every compute entry takes a cudaStream_t (as void*) and is
 * asynchronous on it; every entry returns an smx_status and never throws.  File:line citations are into
 * /root/reference/src/smolyax/.
 */
#ifndef SMOLYAX_B200_H
#define SMOLYAX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    SMX_OK = 0,
    SMX_ERR_INVALID_ARG = 1,   /* the reference raises AssertionError / ValueError here */
    SMX_ERR_CUDA = 2,          /* CUDA runtime failure (message has cudaGetErrorString) */
    SMX_ERR_OUT_OF_MEMORY = 3,
    SMX_ERR_UNSUPPORTED = 4,   /* shape outside the compiled limits */
    SMX_ERR_NO_DEVICE = 5      /* no sm_100 device: there is no CPU fallback */
} smx_status;

/*
 * One group of summands with the same number n of active dimensions, in the layout the reference keeps on the
 * device after set_f (interpolation.py:170-183, 203, 230-235):
 *   F        (nn, d_out, tau[0]+1, .., tau[n-1]+1)  C order, zero padded
 *   nodes    (nn, n, taumax+1)    interpolation nodes per slot, zero padded
 *   weights  (nn, n, taumax+1)    barycentric weights per slot (barycentric.py:13-31), zero padded
 *   dims     (nn, n) int64        active dimensions, sorted by degree descending (interpolation.py:175)
 *   degs     (nn, n) int64        their degrees
 *   zetas    (nn)    int64        Smolyak coefficients
 *   quad     (nn, n, taumax+1)    quadrature weights per slot (interpolation.py:361-379); may be NULL
 * Pointers are HOST pointers for smx_create and DEVICE pointers for the smx_group_* seam twins.
 */
typedef struct {
    int32_t n;
    int64_t nn;
    const int64_t* tau; /* host pointer in both uses, length n */
    const double* F;
    const double* nodes;
    const double* weights;
    const int64_t* dims;
    const int64_t* degs;
    const int64_t* zetas;
    const double* quad;
} smx_group_desc;

/* Flags for smx_interp_desc.flags */
#define SMX_KEEP_GROUPS 1u   /* also upload the reference layout (needed by gradient / integral / barycentric mode) */
#define SMX_NO_FAST_PATH 2u  /* do not build the hierarchical fast path: smx_eval runs the per-summand kernels */
#define SMX_GRAD_FINITE_AT_NODES 4u /* fast gradient: return the true (finite) derivative where a coordinate sits on a
                                       node, instead of the NaN the reference produces there (barycentric.py:152-154) */

#define SMX_DENSE_PATH 8u     /* build the GEMM-regime form (dense term matrix, FP64 tensor instruction) even for small d_out */
#define SMX_NO_DENSE_PATH 16u /* never build it; default: values use it when d_out >= 32  (DESIGN.md "K2") */

typedef struct {
    int64_t d_in;
    int64_t d_out;
    const double* offset; /* (d_out) zeta_0 * f(zero), interpolation.py:151-166; NULL = zeros */
    int32_t n_groups;
    const smx_group_desc* groups;
    uint32_t flags;
} smx_interp_desc;

/*
 * The same interpolant WITHOUT the reference's zero padding and without repeating shared function values: what
 * set_f (interpolation.py:115-239) knows before it pads.  The padded F_n of the named configuration
 * "d_in = 100, d_out = 10^4, n = 10^4" would be 29 GB of host memory; this form of it is 0.8 GB.
 *   summand s (multi-index with zeta != 0, n_active[s] >= 1 active dimensions) owns slots
 *   [slot_off[s], slot_off[s+1]):  dims / degs per slot (any order inside the summand), node_off = offset of the
 *   slot's deg+1 nodes in node_pool (and of its deg+1 quadrature weights in quad_pool, optional);
 *   its value tensor has shape (deg_1+1, .., deg_n+1) in slot order, C order; entry i of it is row
 *   val_index[val_off[s] + i] of `values` (n_values, d_out) - for nested rules one row per sparse-grid node
 *   (= one evaluation of f, interpolation.py:160-163, 217-224), shared by every summand that contains the node.
 * All pointers are HOST pointers.
 */
typedef struct {
    int64_t n_summands;
    const int32_t* n_active;  /* (n_summands) */
    const int64_t* slot_off;  /* (n_summands + 1), slot_off[0] = 0 */
    const int64_t* dims;      /* (slots) */
    const int64_t* degs;      /* (slots) */
    const int64_t* node_off;  /* (slots) */
    const double* node_pool;
    const double* quad_pool;  /* NULL: smx_integral is not available on the handle */
    const int64_t* zetas;     /* (n_summands) */
    const int64_t* val_off;   /* (n_summands + 1), val_off[0] = 0 */
    const int64_t* val_index;
    const double* values;     /* (n_values, d_out) row-major */
    int64_t n_values;
} smx_compact_desc;

typedef struct smx_interp smx_interp; /* opaque; owns every device table it needs */

/* ---- life cycle ------------------------------------------------------------------------------------------
 * smx_create replaces the upload at interpolation.py:230-235: it takes the reference-layout HOST arrays and
 * re-packs them into the device layout (DESIGN.md "Data layout in HBM").  device = CUDA ordinal (-1: current). */
int smx_create(const smx_interp_desc* desc, int device, smx_interp** out);
/* Same from the compact description.  The handle serves smx_eval, smx_gradient (when the derivative sets fit, else
 * SMX_ERR_UNSUPPORTED) and smx_integral (evaluated once at create time, in extended precision on the host: it does not
 * depend on x).  flags: SMX_GRAD_FINITE_AT_NODES, SMX_DENSE_PATH, SMX_NO_DENSE_PATH. */
int smx_create_compact(int64_t d_in, int64_t d_out, const double* offset, const smx_compact_desc* desc, uint32_t flags,
                       int device, smx_interp** out);
int smx_destroy(smx_interp* h);

/* ---- the path --------------------------------------------------------------------------------------------
 * x: (N, ldx) row-major device doubles, ldx >= d_in.  Results are written (not accumulated).
 * smx_eval      replaces SmolyakBarycentricInterpolator.__call__   interpolation.py:264-304 -> y (N, d_out)
 * smx_gradient  replaces SmolyakBarycentricInterpolator.gradient   interpolation.py:306-345 -> J (N, d_out, d_in)
 *               (NaN in J[p,:,dim] when x[p,dim] sits on a node of that dimension, as barycentric.py:152-154,
 *               unless the handle was created with SMX_GRAD_FINITE_AT_NODES)
 * smx_integral  replaces SmolyakBarycentricInterpolator.integral   interpolation.py:347-390 -> q (d_out)        */
int smx_eval(smx_interp* h, const double* x, int64_t N, int64_t ldx, double* y, void* stream);
int smx_gradient(smx_interp* h, const double* x, int64_t N, int64_t ldx, double* J, void* stream);
int smx_integral(smx_interp* h, double* q, void* stream);

/* Same as smx_eval but x and y are HOST buffers (pinned for full speed): the copy in, the kernels and the copy
 * out are pipelined over chunks of `chunk_points` rows (0 = default) on internal streams; returns when y is
 * complete.  This is the reference's `np.asarray(interp(X))` (benchmarking/benchmark.py:131) in one call. */
int smx_eval_host(smx_interp* h, const double* x_host, int64_t N, int64_t ldx, double* y_host, int64_t chunk_points);
/* Sizes the staging buffers and streams of smx_eval_host for batches of n_points rows ahead of the first call - the
 * counterpart of the reference's `n_inputs` constructor argument, which triggers a warm-up call (interpolation.py:250-251). */
int smx_prepare(smx_interp* h, int64_t n_points);

/* ---- stateless seam twins (DEVICE pointers inside the descriptor) ------------------------------------------
 * smx_group_eval      = sum over s of jit(vmap(evaluate_tensor_product_interpolant))   barycentric.py:69-123,
 *                       interpolation.py:293-302:   y (N, d_out) += sum_s zeta_s I_s(x)
 * smx_group_gradient  = sum over s of jit(vmap(evaluate_tensor_product_gradient))      barycentric.py:158-229,
 *                       interpolation.py:334-343:   J (N, d_out, d_in) += sum_s zeta_s grad I_s(x)
 * smx_group_integral  = the einsum at interpolation.py:389:   q (d_out) += sum_s zeta_s <F_s, quad_s>
 * accumulate = 0 overwrites the output first.  These are the functions to register as XLA FFI handlers.       */
int smx_group_eval(const double* x, int64_t N, int64_t ldx, int64_t d_in, const smx_group_desc* g, int64_t d_out,
                   double* y, int accumulate, void* stream);
int smx_group_gradient(const double* x, int64_t N, int64_t ldx, int64_t d_in, const smx_group_desc* g, int64_t d_out,
                       double* J, int accumulate, void* stream);
int smx_group_integral(const smx_group_desc* g, int64_t d_out, double* q, int accumulate, void* stream);

/* barycentric.py:13-31 compute_weights on the device: w_j = prod_{i != j} 1 / (nodes_i - nodes_j). */
int smx_compute_weights(const double* nodes, int64_t m, double* w, void* stream);

/* barycentric.py:34-66 evaluate_basis_unnormalized (derivative = 0) and :126-155
 * evaluate_basis_gradient_unnormalized (derivative = 1) for N points of one dimension:
 * x (N), xi and w (m), degree nu  ->  out (N, m); columns beyond nu are zero. */
int smx_basis(const double* x, int64_t N, const double* xi, const double* w, int64_t m, int64_t nu, int derivative,
              double* out, void* stream);

/* ---- introspection ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t d_in, d_out;
    int64_t n_summands;     /* multi-indices with zeta != 0 and n >= 1 */
    int64_t w_raw;          /* sum over summands of prod(deg+1): entries of the unpadded value tensors */
    int64_t w_pad;          /* same for the reference's padded layout */
    int64_t n_terms;        /* fast path: number of product-basis terms (= |Lambda|) */
    int64_t n_entries;      /* fast path: leading (dim, degree) entries */
    int64_t n_rows;         /* fast path: distinct hot parts */
    int64_t n_chunks;       /* fast path: (entry block, row range) work items */
    int64_t padded_fma;     /* fast path: lane-FMAs per point and output in the block-sparse contraction */
    int64_t device_bytes;   /* HBM held by the handle */
    int32_t has_fast_path, has_groups, nested;
    int32_t has_dense_path; /* GEMM-regime form present: smx_eval uses it */
    int64_t dense_terms;    /* its K (terms, padded to whole stages of 64) */
    int64_t grad_jobs;      /* > 0: smx_gradient runs the fast path - that many jobs per tile and output (one per cold block of 16
                               columns of J, one per hot dimension) .. */
    int64_t grad_items;     /* .. of that many work items in all */
} smx_info;
int smx_get_info(const smx_interp* h, smx_info* info);

/* Number of kernels this library has launched in the calling process (bench.py reports it as gpu_launches). */
int64_t smx_launch_count(void);
/* Name and template arguments of the last kernel the calling thread launched through this library, e.g.
 * "dense_eval_kernel<16,4,2,1>" (tests assert that a configuration ran on the kernel shape it was written for; bench.py
 * records it next to every timing).  "" before the first launch. */
const char* smx_last_kernel(void);

const char* smx_last_error(void); /* thread-local message of the last failing call */
int smx_version(void);            /* 100 * major + minor; the binding checks it against the struct layouts it was written for */
const char* smx_arch(void);       /* "sm_100a" */
const char* smx_build_info(void); /* build stamp: compiler version, hash of the sources the library was built from, flags */

#ifdef __cplusplus
}
#endif
#endif /* SMOLYAX_B200_H */
