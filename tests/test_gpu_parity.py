"""GPU parity (run with -m gpu on a B200): the CUDA paths, called through the C-ABI, against (a) the outputs of the
unmodified reference stored in tests/golden and (b) the CPU oracle on the same inputs.

Tolerances (fp64):
  * per-summand (second barycentric form) kernels vs oracle: same formulas, different summation order ->
    1e-13 of the summand magnitude  sum_nu |zeta_nu I_nu f|.
  * fast path vs reference: 1e-12 of the summand magnitude — the bound north_star states, measured in the only
    norm in which two correct fp64 implementations can meet it (the reference's own rounding noise
    is 1e-13..4e-11 pointwise), plus "at least as close to the 80-bit referee as the reference is".
"""
import ctypes

import numpy as np
import pytest
import torch

from helpers import (ALL_CASES, BIG_CASES, COMPACT_CASES, ILL_CONDITIONED, LAYOUT_CASES, REFERENCE_LAYOUT_CASES, golden_layout, interpolator_inputs, load, long_double,
                     scaled_error)

pytestmark = pytest.mark.gpu


def _build(case, **extra):
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    g = load(case)
    kwargs, f = interpolator_inputs(g)
    return g, SmolyakBarycentricInterpolator(**kwargs, f=f, **extra)


@pytest.mark.parametrize("case", ALL_CASES)
def test_values_fast_path(case):
    from oracle import oracle

    g, ip = _build(case)
    # (a non-nested rule of high degree is refused by the plan compiler's conditioning check: per-summand kernels then)
    assert ip.device_info()["has_fast_path"] == (0 if case in ILL_CONDITIONED else 1)
    x = g["x"]
    y = ip(x)
    assert isinstance(y, np.ndarray) and y.shape == g["y_ref"].shape
    assert scaled_error(y, g["y_ref"], g["cond_abs"]) < 1e-12
    if case not in COMPACT_CASES:
        y_orc = oracle.evaluate(ip.reference_layout(), x)
        assert scaled_error(y, y_orc, g["cond_abs"]) < 1e-12
    y_ld = long_double(g, "y")
    scale = np.max(np.abs(g["y_ref"]))
    err_new = np.max(np.abs((y - y_ld).astype(float))) / scale
    err_ref = np.max(np.abs((g["y_ref"] - y_ld).astype(float))) / scale
    assert err_new <= max(err_ref, 5e-14)
    # device-resident input gives the same bits as the host pipeline; single point keeps the (1, d_out) shape
    y_dev = ip(torch.from_numpy(x).cuda())
    assert y_dev.is_cuda and np.array_equal(y_dev.cpu().numpy(), y)
    assert ip(x[0]).shape == (1, ip.d_out) and np.array_equal(ip(x[0])[0], y[0])


@pytest.mark.parametrize("case", REFERENCE_LAYOUT_CASES)
def test_values_barycentric_kernels(case):
    from oracle import oracle

    g, ip = _build(case, method="barycentric")
    assert ip.device_info()["has_fast_path"] == 0
    y = ip(g["x"])
    y_orc = oracle.evaluate(ip.reference_layout(), g["x"])
    assert scaled_error(y, y_orc, g["cond_abs"]) < 1e-13
    assert scaled_error(y, g["y_ref"], g["cond_abs"]) < 1e-13


@pytest.mark.parametrize("case", ALL_CASES)
def test_gradient(case):
    from oracle import oracle

    g, ip = _build(case)
    J_ref = g["J_ref"]
    x = g["x"][: len(J_ref)]
    J = ip.gradient(x)
    assert J.shape == J_ref.shape == (len(x), ip.d_out, ip.d_in)
    assert np.array_equal(np.isnan(J), np.isnan(J_ref))  # NaN where a coordinate sits on a node, as the reference
    ok = ~np.isnan(J_ref)
    scale = max(1.0, float(np.max(np.abs(J_ref[ok])))) if ok.any() else 1.0
    assert np.max(np.abs(J[ok] - J_ref[ok]), initial=0.0) <= 1e-9 * scale
    if case not in COMPACT_CASES:
        J_orc = oracle.gradient(ip.reference_layout(), x)
        assert np.max(np.abs(J[ok] - J_orc[ok]), initial=0.0) <= 1e-9 * scale


@pytest.mark.parametrize("case", REFERENCE_LAYOUT_CASES)
def test_gradient_barycentric_kernels_and_finite_option(case):
    """The per-summand gradient kernels (method="barycentric") against the oracle, and the fast gradient with
    nan_at_nodes=False: finite everywhere, equal to the default result wherever that one is not NaN."""
    from oracle import oracle

    g, ip = _build(case, method="barycentric")
    J_ref = g["J_ref"]
    x = g["x"][: len(J_ref)]
    J = ip.gradient(x)
    J_orc = oracle.gradient(ip.reference_layout(), x)
    assert np.array_equal(np.isnan(J), np.isnan(J_orc))
    ok = ~np.isnan(J_orc)
    scale = max(1.0, float(np.max(np.abs(J_orc[ok])))) if ok.any() else 1.0
    assert np.max(np.abs(J[ok] - J_orc[ok]), initial=0.0) <= 1e-10 * scale
    if case in ILL_CONDITIONED:
        return  # (no hierarchical form, hence no finite-at-nodes gradient: the handle runs the reference's arithmetic)
    _, fin = _build(case, nan_at_nodes=False)
    Jf = fin.gradient(x)
    assert np.isfinite(Jf).all()
    J_ld = long_double(g, "J")
    err_new = np.max(np.abs((Jf - J_ld)[ok].astype(float)), initial=0.0) / scale
    err_ref = np.max(np.abs((J_ref - J_ld)[ok].astype(float)), initial=0.0) / scale
    assert err_new <= max(err_ref, 1e-12)


@pytest.mark.parametrize("case", ALL_CASES)
def test_integral(case):
    g, ip = _build(case)
    Q = ip.integral()
    assert Q.shape == (ip.d_out,)
    Q_ld = long_double(g, "Q")
    err_new = np.max(np.abs((Q - Q_ld).astype(float)))
    err_ref = np.max(np.abs((g["Q_ref"] - Q_ld).astype(float)))
    assert err_new <= max(10 * err_ref, 1e-11 * max(1.0, float(np.max(np.abs(g["Q_ref"])))))


@pytest.mark.parametrize("case", LAYOUT_CASES)
def test_seam_twins_on_reference_layout(case):
    """smx_group_eval / smx_group_gradient / smx_group_integral with DEVICE pointers to the reference's own arrays."""
    from oracle import oracle
    from smolyax_b200 import _lib

    g = load(case)
    layout = golden_layout(g)
    d_out, x = int(g["d_out"]), g["x"]
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in layout.items()}
    xd = torch.from_numpy(x).cuda()
    y = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(layout["offset"], (len(x), d_out)))).cuda()
    J = torch.zeros((len(x), d_out, x.shape[1]), dtype=torch.float64, device="cuda")
    q = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(layout["offset"], (d_out,)))).cuda()
    arr, n_groups, keep = _lib.pack_groups(dev, lambda a, dt: (a, a.data_ptr()))
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for i in range(n_groups):
        one = ctypes.byref(arr[i]) if False else ctypes.pointer(arr[i])
        _lib.check(_lib.lib.smx_group_eval(xd.data_ptr(), len(x), x.shape[1], x.shape[1], one, d_out, y.data_ptr(), 1, stream))
        _lib.check(_lib.lib.smx_group_gradient(xd.data_ptr(), len(x), x.shape[1], x.shape[1], one, d_out, J.data_ptr(), 1, stream))
        _lib.check(_lib.lib.smx_group_integral(one, d_out, q.data_ptr(), 1, stream))
    torch.cuda.synchronize()
    assert scaled_error(y.cpu().numpy(), oracle.evaluate(layout, x), g["cond_abs"]) < 1e-13
    J_orc = oracle.gradient(layout, x)
    Jh = J.cpu().numpy()
    assert np.array_equal(np.isnan(Jh), np.isnan(J_orc))
    ok = ~np.isnan(J_orc)
    assert np.max(np.abs(Jh[ok] - J_orc[ok]), initial=0.0) <= 1e-10 * max(1.0, float(np.max(np.abs(J_orc[ok]), initial=0.0)))
    assert np.allclose(q.cpu().numpy(), oracle.integral(layout), rtol=1e-11, atol=1e-12)


def test_barycentric_module_functions():
    """The reference's barycentric.py API, one summand at a time (barycentric.py:13,34,69,126,158)."""
    from oracle import oracle
    from smolyax_b200 import barycentric

    g = load("medium_00")
    layout = golden_layout(g)
    x = g["x"]
    n = 2
    F, nd, wt = layout[f"F_{n}"], layout[f"nodes_{n}"], layout[f"weights_{n}"]
    dims, degs, zetas = layout[f"dims_{n}"], layout[f"degs_{n}"], layout[f"zetas_{n}"]
    for s in range(min(5, len(zetas))):
        one = {"offset": np.zeros(F.shape[1]), f"F_{n}": F[s:s + 1], f"nodes_{n}": nd[s:s + 1], f"weights_{n}": wt[s:s + 1],
               f"dims_{n}": dims[s:s + 1], f"degs_{n}": degs[s:s + 1], f"zetas_{n}": zetas[s:s + 1]}
        val = barycentric.evaluate_tensor_product_interpolant(x, F[s], nd[s], wt[s], dims[s], degs[s], int(zetas[s]))
        assert np.allclose(val, oracle.evaluate(one, x), rtol=1e-13, atol=1e-14)
        grad = barycentric.evaluate_tensor_product_gradient(x, F[s], nd[s], wt[s], dims[s], degs[s], int(zetas[s]))
        ref = oracle.gradient(one, x)
        assert np.array_equal(np.isnan(grad), np.isnan(ref))
        assert np.allclose(grad[~np.isnan(ref)], ref[~np.isnan(ref)], rtol=1e-11, atol=1e-12)
    pts = nd[0, 0, : degs[0, 0] + 1]
    w = barycentric.compute_weights(pts)
    assert np.allclose(w, oracle.compute_weights(pts), rtol=1e-15) and np.allclose(w, wt[0, 0, : len(pts)], rtol=1e-14)
    xs = x[:, dims[0, 0]]
    xs = np.concatenate([xs[~np.isin(xs, pts)], pts[:1]])  # regular points, then one that sits on node 0
    b = barycentric.evaluate_basis_unnormalized(xs[:, None], nd[0, 0], wt[0, 0], int(degs[0, 0]))
    db = barycentric.evaluate_basis_gradient_unnormalized(xs[:, None], nd[0, 0], wt[0, 0], int(degs[0, 0]))
    m = nd.shape[2]
    assert b.shape == db.shape == (len(xs), m)
    k = int(degs[0, 0]) + 1
    assert np.allclose(b[:-1, :k], wt[0, 0, :k] / (xs[:-1, None] - nd[0, 0, :k])) and np.all(b[:, k:] == 0)
    assert np.array_equal(b[-1, :k], np.eye(k)[0]) and np.isnan(db[-1, 0]) and np.all(db[:, k:] == 0)
    assert np.allclose(db[:-1, :k], -wt[0, 0, :k] / (xs[:-1, None] - nd[0, 0, :k]) ** 2)


# ---------------------------------------------------------------------------------------------------------------------
# GEMM-regime form (K2) and the compact create entry
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [c for c in ALL_CASES if c not in ILL_CONDITIONED])
def test_values_dense_path(case):
    """smx_eval through the dense term matrix + FP64 tensor instruction (forced here; chosen by itself for d_out >= 32):
    same bound against the reference as the block-sparse path, ragged batch sizes, bitwise repeatable."""
    g, ip = _build(case, dense=True)
    info = ip.device_info()
    assert info["has_dense_path"] == 1 and info["dense_terms"] >= info["n_terms"] - 1
    x = g["x"]
    y = ip(x)
    assert scaled_error(y, g["y_ref"], g["cond_abs"]) < 1e-12
    y_ld = long_double(g, "y")
    scale = np.max(np.abs(g["y_ref"]))
    err_new = np.max(np.abs((y - y_ld).astype(float))) / scale
    err_ref = np.max(np.abs((g["y_ref"] - y_ld).astype(float))) / scale
    assert err_new <= max(err_ref, 5e-14)
    _, sparse = _build(case, dense=False)
    assert sparse.device_info()["has_dense_path"] == 0
    assert scaled_error(y, sparse(x), g["cond_abs"]) < 1e-13
    for n in (1, 31, 33, min(len(x), 77)):
        assert np.array_equal(ip(x[:n]), y[:n])
    xd = torch.from_numpy(x).cuda()
    assert np.array_equal(ip(xd).cpu().numpy(), y)
    # gradient of a dense handle: the same gradient jobs as any other handle (the GEMM-regime form only serves the values);
    # against the reference and, bit for bit, against the handle without the dense form
    J_ref = g["J_ref"]
    xs = x[: len(J_ref)]
    J = ip.gradient(xs)
    assert np.array_equal(np.isnan(J), np.isnan(J_ref))
    ok = ~np.isnan(J_ref)
    scale = max(1.0, float(np.max(np.abs(J_ref[ok])))) if ok.any() else 1.0
    assert np.max(np.abs(J[ok] - J_ref[ok]), initial=0.0) <= 1e-9 * scale
    assert info["grad_jobs"] > 0 and sparse.device_info()["grad_jobs"] == info["grad_jobs"]
    assert np.array_equal(J, sparse.gradient(xs), equal_nan=True)


@pytest.mark.parametrize("case,kernel,grad_kernel", [
    ("cfg1", "fast_pipe_kernel<1>", "grad_kernel<"),                # d_out = 1: the pipelined single-output kernel
    ("cfg2", "fast_pipe_kernel<1>", "grad_kernel<"),                # the headline configuration
    ("cfg4", "fast_", "grad_kernel<"),                              # Gauss-Hermite, 10 outputs
    ("cfg3_dout16", "dense_splitk_kernel<16,2,1>", "grad_kernel<"),  # 16 outputs = two blocks: K split over the warps of a CTA
    ("cfg5", "dense_eval_kernel<8,2,8,2>", "grad_kernel<"),         # d_in = 1000, d_out = 100: 8 warps x 2 blocks, skewed stages
    ("cfg3_dout520", "dense_eval_kernel<16,4,2,1>", "grad_kernel<"),  # cfg3's tables at 520 outputs: 16 warps x 4 blocks
])
def test_benchmark_tables_run_on_the_kernel_written_for_them(case, kernel, grad_kernel):
    """BASELINE's configurations on their REAL tables (index set, nodes, target family; outputs of the unmodified reference in
    the fixture): the operator picks the kernel instantiation that was written for the shape (smx_last_kernel names it),
    and that kernel - not a stand-in of another shape - meets the bounds: against the reference, against the CPU oracle on
    up to 64 sampled output columns, against the 80-bit referee; ragged batch sizes (the fixtures hold 64 .. 257 points,
    never a multiple of the 32-point tile for the wide ones) give the same bits as the full batch."""
    from oracle import oracle
    from smolyax_b200 import _lib

    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    g, ip = _build(case)  # (layout="auto": cfg3 at 520 outputs goes through smx_create_compact)
    x = g["x"]
    y = ip(x)
    assert _lib.last_kernel().startswith(kernel), _lib.last_kernel()
    if ip.d_out >= 100 and case not in COMPACT_CASES:  # the compact entry builds the same plan: same kernel, same bits
        _, compact = _build(case, layout="compact")
        assert np.array_equal(compact(x), y) and _lib.last_kernel().startswith(kernel)
        del compact
    assert scaled_error(y, g["y_ref"], g["cond_abs"]) < 1e-12
    y_ld = long_double(g, "y")
    scale = np.max(np.abs(g["y_ref"]))
    assert np.max(np.abs((y - y_ld).astype(float))) / scale <= max(np.max(np.abs((g["y_ref"] - y_ld).astype(float))) / scale, 5e-14)
    cols = np.unique(np.linspace(0, ip.d_out - 1, min(ip.d_out, 64)).astype(int))
    kwargs, f = interpolator_inputs(g)  # the reference layout of the sampled outputs alone, for the oracle
    kwargs["d_out"] = len(cols)
    narrow = SmolyakBarycentricInterpolator(**kwargs, f=lambda p: np.asarray(f(p))[..., cols], layout="reference", method="barycentric")
    y_orc = oracle.evaluate(narrow.reference_layout(), x)
    del narrow
    assert scaled_error(y[:, cols], y_orc, g["cond_abs"][:, cols]) < 1e-12
    for n in sorted({1, 31, 33, len(x) - 1}):
        assert np.array_equal(ip(x[:n]), y[:n])
    assert np.array_equal(ip(torch.from_numpy(x).cuda()).cpu().numpy(), y)
    # gradient on the same tables: the job kernel, NaN pattern and values as the reference
    J_ref = g["J_ref"]
    J = ip.gradient(x[: len(J_ref)])
    assert _lib.last_kernel().startswith(grad_kernel), _lib.last_kernel()
    assert np.array_equal(np.isnan(J), np.isnan(J_ref))
    ok = ~np.isnan(J_ref)
    assert np.max(np.abs(J[ok] - J_ref[ok]), initial=0.0) <= 1e-9 * max(1.0, float(np.max(np.abs(J_ref[ok]))))


@pytest.mark.parametrize("d_in,d_out,n_target,rule", [(10, 203, 300, "leja"), (6, 520, 120, "leja"), (8, 100, 200, "gh"),
                                                       (12, 1031, 150, "leja")])
def test_wide_outputs_against_oracle(d_in, d_out, n_target, rule):
    """Vector-valued targets wide enough for every CTA shape of the dense kernel (1, 2 and 4 output blocks per warp,
    partial last block, odd d_out): against the CPU oracle on the reference layout."""
    from oracle import oracle
    from smolyax_b200 import indices, nodes, workloads
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    w = workloads.Workload("wide", rule, d_in, d_out, n_target, 0)
    # (f is called point by point: a batched matrix product may round differently depending on the row order)
    ip = SmolyakBarycentricInterpolator(node_gen=w.generator(), k=w.k(), t=w.threshold(), d_out=d_out, f=w.target(),
                                        layout="reference")
    assert ip.device_info()["has_dense_path"] == 1  # chosen without being asked for
    x = w.points(301, seed=5)
    y = ip(x)
    y_orc = oracle.evaluate(ip.reference_layout(), x)
    np.testing.assert_allclose(y, y_orc, rtol=0, atol=1e-12 * max(1.0, float(np.max(np.abs(y_orc)))) * 50)
    sparse = SmolyakBarycentricInterpolator(node_gen=w.generator(), k=w.k(), t=w.threshold(), d_out=d_out, dense=False,
                                            layout="reference")
    sparse.set_layout(ip.reference_layout())
    np.testing.assert_allclose(y, sparse(x), rtol=0, atol=1e-13 * max(1.0, float(np.max(np.abs(y_orc)))) * 50)
    # compact handle: same plan, hence the same bits; integral from the host quadrature
    compact = SmolyakBarycentricInterpolator(node_gen=w.generator(), k=w.k(), t=w.threshold(), d_out=d_out, f=w.target(),
                                             layout="compact")
    assert np.array_equal(compact(x), y)
    np.testing.assert_allclose(compact.integral(), ip.integral(), rtol=1e-10, atol=1e-10)
    with pytest.raises(AssertionError):
        compact.reference_layout()


@pytest.mark.parametrize("case", ["small_02", "small_07", "medium_03", "cfg1", "cfg3_dout16", "gh_d100"])
def test_compact_layout_handle(case):
    """smx_create_compact: values, gradient (same coefficient sets, same bits as the reference-layout handle) and the
    integral (host quadrature in extended precision)."""
    g, ref = _build(case, layout="reference")
    _, ip = _build(case, layout="compact")
    info, info_ref = ip.device_info(), ref.device_info()
    assert info["has_groups"] == 0 and info["n_terms"] == info_ref["n_terms"] and info["w_raw"] == info_ref["w_raw"]
    x = g["x"]
    assert np.array_equal(ip(x), ref(x))
    J_ref = g["J_ref"]
    xs = x[: len(J_ref)]
    J, J2 = ip.gradient(xs), ref.gradient(xs)
    # (every entry of J is written by exactly one job of one warp, in a fixed order: bit for bit, run to run and handle to handle)
    assert np.array_equal(J, J2, equal_nan=True) and np.array_equal(J, ip.gradient(xs), equal_nan=True)
    Q_ld = long_double(g, "Q")
    err_new = np.max(np.abs((ip.integral() - Q_ld).astype(float)))
    err_ref = np.max(np.abs((g["Q_ref"] - Q_ld).astype(float)))
    assert err_new <= max(err_ref, 1e-13 * max(1.0, float(np.max(np.abs(g["Q_ref"])))))


@pytest.mark.parametrize("d_in,d_out,n_target,rule", [(9, 2, 250, "leja"), (12, 4, 300, "leja"), (7, 5, 200, "gh"), (30, 6, 500, "leja")])
def test_few_outputs_multi_set_kernel(d_in, d_out, n_target, rule):
    """2..6 outputs run the block-sparse kernel with several coefficient sets per pass (smx_fast_multi.cu): against the
    CPU oracle, ragged batch sizes, and column by column against single-output handles of the same targets."""
    from oracle import oracle
    from smolyax_b200 import workloads
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    w = workloads.Workload("few", rule, d_in, d_out, n_target, 0)
    fam = w.target()
    ip = SmolyakBarycentricInterpolator(node_gen=w.generator(), k=w.k(), t=w.threshold(), d_out=d_out, f=fam)
    assert ip.device_info()["has_dense_path"] == 0
    x = w.points(205, seed=9)
    y = ip(x)
    y_orc = oracle.evaluate(ip.reference_layout(), x)
    scale = max(1.0, float(np.max(np.abs(y_orc))))
    assert np.max(np.abs(y - y_orc)) <= 5e-11 * scale  # (the oracle carries the reference's Sigma|zeta| rounding noise)
    for n in (1, 31, 33, 64):
        assert np.array_equal(ip(x[:n]), y[:n])
    for o in (0, d_out - 1):
        one = SmolyakBarycentricInterpolator(node_gen=w.generator(), k=w.k(), t=w.threshold(), d_out=1, f=lambda z, o=o: fam(z)[..., o])
        assert np.max(np.abs(one(x)[:, 0] - y[:, o])) <= 1e-13 * scale
    J = ip.gradient(x[:16])
    J_orc = oracle.gradient(ip.reference_layout(), x[:16])
    ok = ~np.isnan(J_orc)
    assert np.array_equal(np.isnan(J), np.isnan(J_orc)) and np.max(np.abs(J[ok] - J_orc[ok])) <= 1e-9 * max(1.0, float(np.max(np.abs(J_orc[ok]))))


def _benchmark_grid_interpolator(d_in, d_out, n_target, **extra):
    """A cell of the reference's heat-map grid (benchmarking/benchmark.py:33-36,180): k_j = log((j + 2)^r / theta)."""
    from smolyax_b200 import indices, nodes, workloads
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    k = np.array([np.log(((j + 2) ** workloads.BASE_R) / workloads.BASE_THETA) for j in range(d_in)])
    gen = nodes.Leja(dim=d_in)
    t = indices.find_approximate_threshold(k, n_target, gen.is_nested)
    f = workloads.TargetFamily(d_in, d_out)
    return f, SmolyakBarycentricInterpolator(node_gen=gen, k=k, t=t, d_out=d_out, f=f, batched_f=True, **extra)


@pytest.mark.parametrize("layout", ["reference", "compact"])
@pytest.mark.parametrize("d_out", [1, 3, 40])
def test_values_low_dimensional_high_cardinality(d_out, layout):
    """d_in = 10, |Lambda| = 6000 (reference benchmark grid): terms with up to seven active dimensions, i.e. hot parts of
    five and six pairs, and a product table far beyond shared memory.  Must stay on the fast path (eight-factor records),
    for any d_out (the GEMM-regime kernel needs the product table, so d_out = 40 runs the block-sparse kernel too)."""
    from oracle import oracle

    f, ip = _benchmark_grid_interpolator(10, d_out, 6000, layout=layout)
    info = ip.device_info()
    assert info["has_fast_path"] == 1
    x = np.random.default_rng(5).uniform(-1.0, 1.0, size=(777, 10))
    y = ip(x)
    # the oracle needs the reference layout: assemble it separately for the compact handle
    _, ip_ref = (f, ip) if layout == "reference" else _benchmark_grid_interpolator(10, d_out, 6000, layout="reference",
                                                                                   method="barycentric")
    ref_layout = ip_ref.reference_layout()
    y_orc = oracle.evaluate(ref_layout, x)
    # summand magnitude sum_nu |zeta_nu I_nu f| is O(sum |zeta|) * |f| here; the fast path is far more accurate than the
    # reference arithmetic, so compare against f as well (the interpolation error at 6000 nodes is ~1e-9)
    assert np.max(np.abs(y - y_orc)) < 1e-10 * max(1.0, np.max(np.abs(y_orc)))
    assert np.sqrt(np.mean((y - f(x)) ** 2) / np.mean(f(x) ** 2)) < 1e-7
    assert np.array_equal(ip(torch.from_numpy(x).cuda()).cpu().numpy(), y)


def test_gradient_wider_than_one_set_of_derivative_tables():
    """d_out > 2 048 on the compact layout: the handle keeps only the GEMM-regime value tables, ``smx_gradient`` answers
    SMX_ERR_UNSUPPORTED and ``gradient`` works through blocks of output columns (linearity in f).  Checked against the
    per-summand kernels on the reference layout (which are checked against the oracle above)."""
    from smolyax_b200 import _lib, workloads
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    d_in, d_out = 4, 2056  # (four dimensions: no hot part deeper than four pairs)
    w = workloads.Workload("wide_grad", "leja", d_in, d_out, 60, 0)
    ip = SmolyakBarycentricInterpolator(node_gen=w.generator(), k=w.k(), t=w.threshold(), d_out=d_out, f=w.target(), layout="compact")
    x = w.points(37, seed=3)
    xt = torch.from_numpy(x).cuda()
    J_raw = torch.empty((len(x), d_out, d_in), dtype=torch.float64, device="cuda")
    status = _lib.lib.smx_gradient(ip._handle, xt.data_ptr(), len(x), d_in, J_raw.data_ptr(), None)
    assert status == _lib.SMX_ERR_UNSUPPORTED  # (the C ABI itself still says so: the blocks are the host layer's)
    J = ip.gradient(x)
    assert J.shape == (len(x), d_out, d_in)
    ref = SmolyakBarycentricInterpolator(node_gen=w.generator(), k=w.k(), t=w.threshold(), d_out=d_out, f=w.target(),
                                         layout="reference", method="barycentric")
    J_ref = ref.gradient(x)
    scale = max(1.0, float(np.nanmax(np.abs(J_ref))))
    assert np.array_equal(np.isnan(J), np.isnan(J_ref))
    assert np.max(np.abs(np.nan_to_num(J) - np.nan_to_num(J_ref))) <= 1e-9 * scale  # (the bound of test_gradient)
    assert np.allclose(ip.gradient(xt).cpu().numpy(), J, rtol=0, atol=1e-12 * scale, equal_nan=True)  # device input
