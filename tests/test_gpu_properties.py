"""Size-independent properties on the GPU, including BASELINE-size runs: polynomial exactness (the reference's own
known-answer test, tests/setup.py:74-130 + tests/test_smolyak.py), interpolation at the nodes, linearity in f,
host-pipeline == device path, strided inputs, the empty index set."""
import ctypes

import numpy as np
import pytest
import torch
from numpy.polynomial import hermite, legendre

from smolyax_b200 import indices, nodes, workloads

pytestmark = pytest.mark.gpu


def _interp(**kw):
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    return SmolyakBarycentricInterpolator(**kw)


class ProductPolynomial:
    """f_o(x) = prod_j P_{nu_o,j}(x_j) with nu_o in Lambda: Legendre (Leja) / Hermite (Gauss-Hermite) factors on the
    reference domain, so the Smolyak interpolant reproduces f exactly."""

    def __init__(self, gen, k, t, d_out, rng):
        self.gen = gen
        lam = indices.indexset(k, t)
        self.nus = []
        for _ in range(d_out):
            full = [0] * len(k)
            for dim, deg in lam[int(rng.integers(len(lam)))]:
                full[dim] = deg
            self.nus.append(full)

    def _p(self, g, n, x, der=0):
        z = g.scale_back(x)
        c = [0] * n + [1]
        if isinstance(g, nodes.Leja1D):
            if der:
                scale = 1 if g.domain is None else (g.domain[1] - g.domain[0]) / 2
                return legendre.legval(z, legendre.legder(c)) / scale
            return legendre.legval(z, c)
        if der:
            return hermite.hermval(z, hermite.hermder(c) / g.scaling)
        return hermite.hermval(z, c)

    def __call__(self, x):
        x = np.atleast_2d(x)
        out = np.array([np.prod([self._p(g, n, x[:, j]) for j, (g, n) in enumerate(zip(self.gen, nu))], axis=0) for nu in self.nus]).T
        return out[0] if out.shape[0] == 1 else out

    def gradient(self, x):
        x = np.atleast_2d(x)
        J = np.zeros((x.shape[0], len(self.nus), x.shape[1]))
        for o, nu in enumerate(self.nus):
            vals = [self._p(g, n, x[:, j]) for j, (g, n) in enumerate(zip(self.gen, nu))]
            for j, (g, n) in enumerate(zip(self.gen, nu)):
                others = np.prod([v for i, v in enumerate(vals) if i != j], axis=0) if len(vals) > 1 else 1.0
                J[:, o, j] = self._p(g, n, x[:, j], der=1) * others
        return J

    def integral(self):
        return np.array([float(all(n == 0 for n in nu)) for nu in self.nus])


def test_polynomial_exactness_values_gradient_integral():
    rng = np.random.default_rng(11)
    for trial in range(10):
        d = int(rng.integers(1, 5))
        if trial % 2 == 0:
            gen = nodes.Leja(domains=np.sort(rng.random((d, 2)), axis=1) + [[0, 0.05]] * d)
        else:
            gen = nodes.GaussHermite(rng.standard_normal(d), 0.2 + rng.random(d))
        k = np.sort(rng.uniform(1, 10, d))
        k = k / k[0]
        d_out, t = int(rng.integers(1, 4)), float(rng.uniform(1, 8))
        f = ProductPolynomial(gen, k, t, d_out, rng)
        for method in ("auto", "barycentric"):
            ip = _interp(node_gen=gen, k=k, t=t, d_out=d_out, f=f, method=method)
            np.random.seed(trial)
            x = gen.get_random(int(rng.integers(1, 5)))
            assert np.allclose(f(x), ip(x), rtol=1e-9, atol=1e-10)
            assert np.allclose(f.gradient(x), ip.gradient(x), rtol=1e-8, atol=1e-8)
            Q = ip.integral()
            assert Q.shape == (d_out,) and np.allclose(Q, f.integral(), atol=1e-10)


def test_interpolation_property_at_grid_nodes():
    """ip(xi_mu) == f(xi_mu) at sparse-grid nodes: every coordinate sits on a node (one-hot rows in the
    barycentric kernels, plain polynomial evaluation in the fast path)."""
    wl = workloads.Workload("t", "leja", 6, 2, 300, 0)
    gen, k = wl.generator(), wl.k()
    t = wl.threshold()
    f = wl.target()
    pts = []
    for nu in indices.indexset(k, t)[:200]:
        x = np.zeros(6)
        for dim, deg in nu:
            x[dim] = gen[dim](deg)[deg]
        pts.append(x)
    x = np.array(pts)
    for method in ("auto", "barycentric"):
        ip = _interp(node_gen=gen, k=k, t=t, d_out=2, f=f, method=method)
        assert np.allclose(ip(x), f(x), rtol=1e-12, atol=1e-13)


def test_linearity_in_f_and_empty_index_set():
    wl = workloads.Workload("t", "gh", 5, 1, 200, 0)
    gen, k, t = wl.generator(), wl.k(), wl.threshold()
    f1 = lambda x: np.sin(x.sum())
    f2 = lambda x: np.exp(-0.1 * x[0]) * x[-1]
    x = wl.points(257, seed=3)
    a = _interp(node_gen=gen, k=k, t=t, d_out=1, f=f1)(x)
    b = _interp(node_gen=gen, k=k, t=t, d_out=1, f=f2)(x)
    c = _interp(node_gen=gen, k=k, t=t, d_out=1, f=lambda x: 2.0 * f1(x) - 3.0 * f2(x))(x)
    assert np.allclose(c, 2.0 * a - 3.0 * b, rtol=1e-11, atol=1e-12)
    # Lambda = {0}: the reference trips an assert in __validate_input; the drop-in returns f(zero)
    tiny = _interp(node_gen=gen, k=k, t=0.5, d_out=1, f=f1)
    assert tiny.n_f_evals == 1 and np.allclose(tiny(x), f1(np.zeros(5)))
    assert np.allclose(tiny.gradient(x[:3]), 0.0) and np.allclose(tiny.integral(), f1(np.zeros(5)))


def test_wrong_shapes_raise_like_the_reference():
    wl = workloads.Workload("t", "leja", 4, 1, 50, 0)
    gen, k, t = wl.generator(), wl.k(), wl.threshold()
    ip = _interp(node_gen=gen, k=k, t=t, d_out=1)
    with pytest.raises(AssertionError, match="set_f"):
        ip(np.zeros((2, 4)))
    ip.set_f(f=wl.target())
    with pytest.raises(AssertionError):
        ip(np.zeros((2, 5)))
    with pytest.raises(AssertionError):
        ip.gradient(np.zeros((3, 3)))


def test_strided_input_and_c_abi_directly():
    from smolyax_b200 import _lib

    wl = workloads.Workload("t", "leja", 37, 3, 800, 0)
    gen, k, t = wl.generator(), wl.k(), wl.threshold()
    ip = _interp(node_gen=gen, k=k, t=t, d_out=3, f=wl.target())
    x = wl.points(1001, seed=5)
    y = ip(x)
    wide = torch.zeros((1001, 48), dtype=torch.float64, device="cuda")
    wide[:, :37] = torch.from_numpy(x).cuda()
    out = torch.empty((1001, 3), dtype=torch.float64, device="cuda")
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib.smx_eval(ip._handle, wide.data_ptr(), 1001, 48, out.data_ptr(), stream))
    assert np.array_equal(out.cpu().numpy(), y)
    assert np.array_equal(ip(wide[:, :37]).cpu().numpy(), y)  # non-contiguous view, stride(0) = 48
    assert _lib.lib.smx_eval(ip._handle, wide.data_ptr(), 10, 5, out.data_ptr(), stream) == 1  # ldx < d_in
    # rows that are not 16-byte aligned (odd pitch) take the cp.async kernel instead of the TMA kernel: same values
    odd = torch.zeros((1001, 49), dtype=torch.float64, device="cuda")
    odd[:, :37] = torch.from_numpy(x).cuda()
    assert np.allclose(ip(odd[:, :37]).cpu().numpy(), y, rtol=1e-13, atol=1e-14)
    shifted = torch.zeros(1001 * 48 + 1, dtype=torch.float64, device="cuda")[1:].view(1001, 48)  # base not 16-byte aligned
    shifted[:, :37] = torch.from_numpy(x).cuda()
    assert np.allclose(ip(shifted[:, :37]).cpu().numpy(), y, rtol=1e-13, atol=1e-14)
    before = _lib.lib.smx_launch_count()
    ip(wide[:, :37])
    assert _lib.lib.smx_launch_count() == before + 1  # one fused kernel per call


def test_headline_config_full_size_consistency():
    """cfg2 tables (d_in=1000, n=10^4) at 2*10^5 points: the host pipeline and the device path agree bit for bit,
    results match the oracle on a sample, and the output is finite everywhere."""
    from oracle import oracle

    wl = workloads.CONFIGS["cfg2"]
    ip = _interp(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=1, f=wl.target())
    info = ip.device_info()
    assert info["n_terms"] == 9999 and info["n_summands"] == 8751
    gen = torch.Generator(device="cuda").manual_seed(0)
    xd = torch.rand((200_000, 1000), dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    yd = ip(xd)
    assert torch.isfinite(yd).all()
    xh = xd.cpu().numpy()
    assert np.array_equal(ip(xh), yd.cpu().numpy())
    sample = np.r_[0:64, 199_936:200_000]
    y_orc = oracle.evaluate(ip.reference_layout(), xh[sample])
    f_true = wl.target()(xh[sample]).reshape(-1, 1)
    assert np.max(np.abs(yd.cpu().numpy()[sample] - y_orc)) < 1e-9  # the oracle's own noise is ~1e-11 here
    assert np.max(np.abs(y_orc - f_true)) < 1e-2  # and both approximate f


def test_dense_path_strided_unaligned_streams_and_empty_batch():
    """GEMM-regime kernels through the C ABI directly: padded / odd / unaligned row pitches of x, an unaligned y, an
    empty batch, two streams at once on one handle, one launch per call."""
    from smolyax_b200 import _lib

    wl = workloads.Workload("t", "leja", 37, 75, 600, 0)  # 75 outputs: odd count, partial last block, staged kernel
    ip = _interp(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=75, f=wl.target(), batched_f=True)
    assert ip.device_info()["has_dense_path"] == 1
    x = wl.points(333, seed=11)
    y = ip(x)
    lib, h = _lib.lib, ip._handle
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for pitch, shift in ((48, 0), (49, 0), (48, 1), (37, 1)):
        buf = torch.zeros(333 * pitch + shift, dtype=torch.float64, device="cuda")[shift:].view(333, pitch)
        buf[:, :37] = torch.from_numpy(x).cuda()
        out = torch.zeros(333 * 75 + 1, dtype=torch.float64, device="cuda")
        for off in (0, 1):  # y 16-byte aligned or not
            o = out[off:off + 333 * 75].view(333, 75)
            _lib.check(lib.smx_eval(h, buf.data_ptr(), 333, pitch, o.data_ptr(), stream))
            assert np.array_equal(o.cpu().numpy(), y), (pitch, shift, off)
    assert lib.smx_eval(h, 0, 0, 37, 0, stream) == 0  # N = 0: nothing to do, no buffers needed
    assert ip(np.zeros((0, 37))).shape == (0, 75)
    # two streams, same handle (tables are read-only after create)
    xd = torch.from_numpy(x).cuda()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        y1 = ip(xd)
    with torch.cuda.stream(s2):
        y2 = ip(xd)
    torch.cuda.synchronize()
    assert np.array_equal(y1.cpu().numpy(), y) and np.array_equal(y2.cpu().numpy(), y)
    before = lib.smx_launch_count()
    ip(xd)
    assert lib.smx_launch_count() == before + 1


def test_compact_create_rejects_malformed_descriptors_and_flags():
    from smolyax_b200 import _lib

    wl = workloads.Workload("t", "gh", 6, 3, 80, 0)  # non-nested rule through the compact entry
    ip = _interp(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=3, f=wl.target(), layout="compact")
    ref = _interp(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=3, f=wl.target(), layout="reference")
    x = wl.points(65, seed=2)
    assert np.array_equal(ip(x), ref(x))
    np.testing.assert_allclose(ip.integral(), ref.integral(), rtol=1e-12, atol=1e-13)
    good = ip._layout
    for key, bad in (("val_index", good["val_index"] + good["values"].shape[0]),       # row out of range
                     ("dims", np.where(good["dims"] == good["dims"].max(), 6, good["dims"])),  # dimension out of range
                     ("val_off", good["val_off"] + 1),                                 # shape mismatch
                     ("degs", np.zeros_like(good["degs"]))):                           # degree 0 is not an active slot
        layout = dict(good)
        layout[key] = bad
        with pytest.raises(AssertionError):
            _lib.create_compact(layout, 6, 3, 0)
    no_quad = dict(good)
    no_quad["quad_pool"] = None
    h = _lib.create_compact(no_quad, 6, 3, 0)
    q = torch.empty(3, dtype=torch.float64, device="cuda")
    assert _lib.lib.smx_integral(h, q.data_ptr(), None) == 1  # created without quadrature weights
    _lib.lib.smx_destroy(h)
    # SMX_NO_DENSE_PATH / SMX_DENSE_PATH are honoured
    assert _lib.info(_lib.create_compact(good, 6, 3, _lib.SMX_DENSE_PATH))["has_dense_path"] == 1
    assert _lib.info(_lib.create_compact(good, 6, 3, _lib.SMX_NO_DENSE_PATH))["has_dense_path"] == 0


def test_block_sparse_kernel_small_batches_many_outputs_and_knobs():
    """The lean block-sparse kernel: (a) a small batch spreads its outputs over gridDim.y, a large one walks them inside
    the CTA - same bits either way; (c) non-zero first centres (custom Leja domain) take the variant with the subtraction
    and still reproduce a polynomial exactly.  (The kernels' tuning knobs are compiled out of the product library.)"""

    d_in, d_out = 24, 10  # ten passes of the block-sparse kernel (dense=False: for so few terms the plan compiler would pick K2)
    k = workloads.anisotropy(d_in)
    gen = nodes.Leja(dim=d_in)
    t = indices.find_approximate_threshold(k, 700, True)
    f = workloads.TargetFamily(d_in, d_out)
    ip = _interp(node_gen=gen, k=k, t=t, d_out=d_out, f=f, batched_f=True, dense=False)
    assert ip.device_info()["has_fast_path"] == 1 and ip.device_info()["has_dense_path"] == 0
    x = torch.rand((40_000, d_in), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) * 2 - 1
    y_big = ip(x)                      # 1250 tiles: gridDim.y = 1
    y_small = ip(x[:1000])             # 32 tiles: outputs spread over gridDim.y
    assert torch.equal(y_small, y_big[:1000])
    assert torch.equal(ip(x[:33]), y_big[:33])
    # (c) custom domain: first centres != 0
    rng = np.random.default_rng(11)
    dom = np.stack([rng.uniform(-3, -1, 6), rng.uniform(0.5, 2, 6)], axis=1)
    gen_c = nodes.Leja(domains=dom)
    k_c = workloads.anisotropy(6)
    t_c = indices.find_approximate_threshold(k_c, 200, True)
    fp = ProductPolynomial(gen_c, k_c, t_c, 3, rng)
    ip_c = _interp(node_gen=gen_c, k=k_c, t=t_c, d_out=3, f=fp)
    xc = rng.uniform(dom[:, 0], dom[:, 1], size=(500, 6))
    assert np.allclose(ip_c(xc), fp(xc), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("d_out", [72, 100, 203])
def test_dense_kernel_block_dealing_never_changes_a_bit(d_out):
    """The staged dense kernel deals the output blocks of a column group that does not fill the CTA to its warps per FP64 pipe
    (%warpid), in half blocks, rotated by a per-SM ticket that advances with every CTA - so consecutive launches compute the
    same columns on different warps.  Only who computes a column may change, never its arithmetic: many launches, batch sizes
    with ragged last tiles, all bit-identical; and equal to the block-sparse kernels within the usual bound."""
    w = workloads.Workload("deal", "leja", 40, d_out, 600, 0)
    ip = _interp(node_gen=w.generator(), k=w.k(), t=w.threshold(), d_out=d_out, f=w.target(), batched_f=True, dense=True)
    assert ip.device_info()["has_dense_path"] == 1
    x = torch.from_numpy(w.points(3000, seed=7)).cuda()
    y0 = ip(x)
    for _ in range(6):
        assert torch.equal(ip(x), y0)
    for n in (1, 33, 1000, 2999):
        assert torch.equal(ip(x[:n]), y0[:n])
    sparse = _interp(node_gen=w.generator(), k=w.k(), t=w.threshold(), d_out=d_out, f=w.target(), batched_f=True, dense=False)
    scale = float(y0.abs().max())
    assert float((sparse(x[:256]) - y0[:256]).abs().max()) <= 1e-12 * max(1.0, scale) * 50


def test_out_buffer_of_each_input_kind():
    """``__call__(x, out=...)``: the caller's result buffer (device, page-locked host, NumPy) receives the same bits as a
    fresh one; a buffer of the wrong shape, type or kind is refused like any other bad argument (AssertionError)."""
    from smolyax_b200 import workloads
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    w = workloads.Workload("out", "leja", 12, 3, 200, 0)
    ip = SmolyakBarycentricInterpolator(node_gen=w.generator(), k=w.k(), t=w.threshold(), d_out=3, f=w.target())
    x = w.points(1000, seed=11)
    y = ip(x)
    out_np = np.full((1000, 3), np.nan)
    assert ip(x, out=out_np) is out_np and np.array_equal(out_np, y)
    xt = torch.from_numpy(x).cuda()
    out_dev = torch.full((1000, 3), float("nan"), dtype=torch.float64, device="cuda")
    assert ip(xt, out=out_dev) is out_dev and np.array_equal(out_dev.cpu().numpy(), y)
    xp = torch.from_numpy(x).pin_memory()
    out_pin = torch.full((1000, 3), float("nan"), dtype=torch.float64).pin_memory()
    assert ip(xp, out=out_pin) is out_pin and np.array_equal(out_pin.numpy(), y)
    for bad in (np.empty((999, 3)), np.empty((1000, 3), dtype=np.float32), np.empty((3, 1000)).T, out_dev):
        with pytest.raises(AssertionError):
            ip(x, out=bad)
    with pytest.raises(AssertionError):
        ip(xt, out=out_pin)


@pytest.mark.parametrize("mode", ["reference", "compact"])
def test_tables_from_a_file_give_the_same_bits(tmp_path, mode):
    """``save_layout`` / ``load_layout``: a handle built from the stored tables (no call of ``f``) evaluates and integrates
    to the same bits as the handle ``set_f`` built - values, gradient and integral."""
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    w = workloads.Workload("file", "leja", 15, 2, 400, 0)
    ip = SmolyakBarycentricInterpolator(node_gen=w.generator(), k=w.k(), t=w.threshold(), d_out=2, f=w.target(), layout=mode)
    ip.save_layout(tmp_path / "tables.npz")
    again = SmolyakBarycentricInterpolator(node_gen=w.generator(), k=w.k(), t=w.threshold(), d_out=2)
    again.load_layout(tmp_path / "tables.npz")
    x = w.points(777, seed=5)
    assert np.array_equal(ip(x), again(x))
    J, J2 = ip.gradient(x[:50]), again.gradient(x[:50])
    assert np.array_equal(J, J2, equal_nan=True)
    assert np.array_equal(ip.integral(), again.integral())


@pytest.mark.parametrize("rule", ["leja", "gh"])
def test_single_output_pipelined_kernel_ragged_batches_and_repeatability(rule):
    """d_out = 1 runs the warp-specialised kernel (workers + service warp, csrc/smx_fast_pipe.cu): against the CPU oracle in the
    summand-magnitude norm, for batch sizes below one tile, ragged, one tile per CTA and several tiles per CTA (the item
    stream then crosses tile boundaries), and with the same bits run to run and whatever the batch the point sits in.
    Gauss-Hermite takes the variant with non-zero first centres."""
    from oracle import oracle

    w = workloads.Workload("pipe", rule, 60, 1, 900, 0)
    ip = _interp(node_gen=w.generator(), k=w.k(), t=w.threshold(), d_out=1, f=w.target(), batched_f=True, layout="reference")
    info = ip.device_info()
    assert info["has_fast_path"] == 1 and info["has_dense_path"] == 0
    n_big = 148 * 32 * 3 + 17
    x = w.points(n_big, seed=5)
    xd = torch.from_numpy(x).cuda()
    y = ip(xd)
    assert torch.equal(ip(xd), y)
    sample = np.r_[0:200, n_big - 200:n_big]
    layout = ip.reference_layout()
    ref = oracle.evaluate(layout, x[sample])
    mag = oracle.evaluate({k: (np.abs(v) if k.startswith("zetas_") or k == "offset" else v) for k, v in layout.items()}, x[sample])
    assert np.max(np.abs(y.cpu().numpy()[sample] - ref) / np.abs(mag)) < 1e-12
    for n in (1, 5, 31, 32, 33, 148 * 32, 148 * 32 + 1):
        assert torch.equal(ip(xd[:n]), y[:n]), n
    assert np.array_equal(ip(x[:1000]), y[:1000].cpu().numpy())  # host pipeline: same kernel, same bits


def test_high_degree_gauss_hermite_falls_back_to_the_reference_arithmetic():
    """Non-nested rules of high degree: the Newton form of the fast path would lose digits without any warning (at degree 40
    more than the interpolation error).  The plan compiler measures that and refuses; the handle then evaluates with the
    per-summand barycentric kernels and agrees with the oracle - values, gradient and integral - like any other."""
    from oracle import oracle

    gen = nodes.GaussHermite(dim=2)
    k = [1.0, 10.0]
    f = workloads.TargetFamily(2, 1)
    for t, fast in ((9.5, 1), (41.5, 0)):
        ip = _interp(node_gen=gen, k=k, t=t, d_out=1, f=f, layout="reference")
        assert ip.device_info()["has_fast_path"] == fast, t
        x = np.random.default_rng(3).standard_normal((300, 2)) / np.sqrt(2.0)
        layout = ip.reference_layout()
        y, ref = ip(x), oracle.evaluate(layout, x)
        assert np.max(np.abs(y - ref)) < 1e-11 * max(1.0, float(np.max(np.abs(ref)))), t
        J, J_ref = ip.gradient(x[:40]), oracle.gradient(layout, x[:40])
        assert np.max(np.abs(J - J_ref)) < 1e-9 * max(1.0, float(np.max(np.abs(J_ref)))), t
        assert np.max(np.abs(ip.integral() - oracle.integral(layout))) < 1e-11
    # the compact form has no per-summand kernels behind it: refused with a message that says what to do
    with pytest.raises(Exception, match="ill-conditioned"):
        _interp(node_gen=gen, k=k, t=41.5, d_out=1, f=f, layout="compact")


def test_more_than_eight_active_dimensions_per_summand():
    """Summands with nine active dimensions (low d_in, high cardinality: |Lambda| ~ 5e4 and up, too slow to build in a test,
    so the reference-layout arrays are written down directly): the per-summand kernels stop at eight, the fast path does
    not - smx_create takes the layout all the same, values come from the fast path, the integral from the host at create
    time; the gradient is either computed by the fast path or refused (SMX_ERR_UNSUPPORTED), never silently wrong."""
    from oracle import oracle
    from smolyax_b200.interpolation import _host_weights

    rng = np.random.default_rng(7)
    d_in, d_out = 12, 2
    layout = {"offset": rng.standard_normal(d_out)}
    pts = np.array([0.0, 1.0])
    for n, nn in ((2, 5), (9, 3)):
        layout[f"F_{n}"] = rng.standard_normal((nn, d_out) + (2,) * n)
        layout[f"nodes_{n}"] = np.tile(pts, (nn, n, 1))
        layout[f"weights_{n}"] = np.tile(_host_weights(pts), (nn, n, 1))
        layout[f"quad_{n}"] = np.tile(np.array([0.75, 0.25]), (nn, n, 1))
        layout[f"dims_{n}"] = np.stack([np.sort(rng.choice(d_in, n, replace=False)) for _ in range(nn)]).astype(np.int64)
        layout[f"degs_{n}"] = np.ones((nn, n), dtype=np.int64)
        layout[f"zetas_{n}"] = rng.integers(-3, 4, nn).astype(np.int64)
    ip = _interp(node_gen=nodes.Leja(dim=d_in), k=[1.0] * d_in, t=1.5, d_out=d_out)
    ip.set_layout(layout)
    info = ip.device_info()
    assert info["has_fast_path"] == 1 and info["has_groups"] == 0
    x = rng.uniform(-1, 1, (300, d_in))
    ref = oracle.evaluate(layout, x)
    assert np.max(np.abs(ip(x) - ref)) < 1e-12 * max(1.0, float(np.max(np.abs(ref))))
    assert np.max(np.abs(ip.integral() - oracle.integral(layout))) < 1e-12 * max(1.0, float(np.max(np.abs(oracle.integral(layout)))))
    from smolyax_b200._lib import SmolyaxCudaError

    try:  # refused (no derivative sets for eight-factor records) or right - never silently wrong
        J = ip.gradient(x[:64])
    except SmolyaxCudaError as exc:
        assert "unsupported" in str(exc)
    else:
        J_ref = oracle.gradient(layout, x[:64])
        assert np.max(np.abs(J - J_ref)) < 1e-11 * max(1.0, float(np.max(np.abs(J_ref))))
