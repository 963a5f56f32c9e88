"""The plan compiler (csrc/smx_plan.cpp): reference layout -> hierarchical block-sparse layout.  Runs on the CPU
through libsmolyax_host.so; `smxh_plan_eval_host` mirrors the kernel's arithmetic and is a verification aid only."""
import ctypes

import numpy as np
import pytest

from smolyax_b200 import _build
from smolyax_b200.interpolation import SmolyakBarycentricInterpolator
from helpers import ALL_CASES, ILL_CONDITIONED, LAYOUT_CASES, golden_layout, interpolator_inputs, load, long_double, scaled_error


class _Group(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("nn", ctypes.c_int64)] + [
        (name, ctypes.c_void_p) for name in ("tau", "F", "nodes", "weights", "dims", "degs", "zetas", "quad")]


_host = ctypes.CDLL(str(_build.build_host()))
_host.smxh_plan_build.restype = ctypes.c_void_p
_host.smxh_plan_build.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
_host.smxh_plan_build_opt.restype = ctypes.c_void_p
_host.smxh_plan_build_opt.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32]
_host.smxh_plan_build_compact.restype = ctypes.c_void_p
_host.smxh_plan_build_compact.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
_host.smxh_integrate_compact.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
_host.smxh_plan_eval_dense_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
_host.smxh_plan_error.restype = ctypes.c_char_p
_host.smxh_plan_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
_host.smxh_plan_eval_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
_host.smxh_plan_gradient_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
_host.smxh_plan_free.argtypes = [ctypes.c_void_p]
STATS = ("n_terms", "n_entries", "n_rows", "n_hot", "n_chunks", "padded_fma", "n_levels", "nested", "n_summands", "w_raw", "w_pad")


GRADIENT, SPARSE, DENSE = 1, 2, 4  # option bits of smxh_plan_build_opt


class _Compact(ctypes.Structure):  # smx_compact_desc
    _fields_ = [("n_summands", ctypes.c_int64)] + [(name, ctypes.c_void_p) for name in (
        "n_active", "slot_off", "dims", "degs", "node_off", "node_pool", "quad_pool", "zetas", "val_off", "val_index",
        "values")] + [("n_values", ctypes.c_int64)]


def _compact_desc(layout):
    desc, keep = _Compact(), []
    desc.n_summands = len(layout["zetas"])
    for name, _ in _Compact._fields_[1:-1]:
        a = np.ascontiguousarray(layout[name], dtype={"n_active": np.int32, "node_pool": np.float64, "quad_pool": np.float64,
                                                       "values": np.float64}.get(name, np.int64))
        keep.append(a)
        setattr(desc, name, a.ctypes.data)
    desc.n_values = layout["values"].shape[0]
    return desc, keep


class Plan:
    def __init__(self, layout, d_in, d_out, options=GRADIENT | SPARSE):
        self.d_out = d_out
        off = np.ascontiguousarray(np.broadcast_to(np.asarray(layout["offset"], dtype=np.float64), (d_out,)))
        if layout.get("compact"):
            desc, self._keep = _compact_desc(layout)
            self.h = _host.smxh_plan_build_compact(d_in, d_out, off.ctypes.data, ctypes.addressof(desc), options)
            self._finish()
            return
        ns = sorted(int(k.split("_")[1]) for k in layout if k.startswith("zetas_"))
        arr = (_Group * max(len(ns), 1))()
        self._keep = []
        for i, n in enumerate(ns):
            F = np.ascontiguousarray(layout[f"F_{n}"], dtype=np.float64)
            vals = [np.ascontiguousarray(np.array(F.shape[2:], dtype=np.int64) - 1), F]
            vals += [np.ascontiguousarray(layout[f"{k}_{n}"], dtype=np.float64) for k in ("nodes", "weights")]
            vals += [np.ascontiguousarray(layout[f"{k}_{n}"], dtype=np.int64) for k in ("dims", "degs", "zetas")]
            self._keep.append(vals)
            arr[i].n, arr[i].nn = n, F.shape[0]
            for name, v in zip(("tau", "F", "nodes", "weights", "dims", "degs", "zetas"), vals):
                setattr(arr[i], name, v.ctypes.data)
        self.h = _host.smxh_plan_build_opt(d_in, d_out, off.ctypes.data, len(ns), ctypes.addressof(arr), options)
        self._finish()

    def _finish(self):
        self.error = None if self.h else _host.smxh_plan_error().decode()
        if self.h:
            st = np.zeros(len(STATS), dtype=np.int64)
            _host.smxh_plan_stats(self.h, st.ctypes.data)
            self.stats = dict(zip(STATS, st.tolist()))

    def __call__(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros((len(x), self.d_out))
        _host.smxh_plan_eval_host(self.h, x.ctypes.data, len(x), x.shape[1], y.ctypes.data)
        return y

    def dense(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros((len(x), self.d_out))
        _host.smxh_plan_eval_dense_host(self.h, x.ctypes.data, len(x), x.shape[1], y.ctypes.data)
        return y

    def gradient(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        J = np.zeros((len(x), self.d_out, x.shape[1]))
        _host.smxh_plan_gradient_host(self.h, x.ctypes.data, len(x), x.shape[1], J.ctypes.data)
        return J

    def __del__(self):
        if getattr(self, "h", None):
            _host.smxh_plan_free(self.h)


def _layout_of(g, case=None):
    if any(k.startswith("layout_F_") for k in g.files):
        return golden_layout(g)
    kwargs, f = interpolator_inputs(g)
    return SmolyakBarycentricInterpolator(**kwargs)._assemble(f, {})[0]


@pytest.mark.parametrize("case", ALL_CASES)
def test_plan_reproduces_reference_values(case):
    g = load(case)
    if case in LAYOUT_CASES:
        layout = golden_layout(g)
    else:
        kwargs, f = interpolator_inputs(g)
        layout, _ = SmolyakBarycentricInterpolator(**kwargs)._assemble(f, {})
    plan = Plan(layout, g["x"].shape[1], int(g["d_out"]))
    if case in ILL_CONDITIONED:  # refused, loudly: smx_create falls back to the per-summand kernels
        assert plan.error is not None and plan.error.startswith("ill-conditioned"), plan.error
        return
    assert plan.error is None, plan.error
    y = plan(g["x"])
    # parity with the reference: within 1e-12 of the summand magnitude (the reference's own rounding noise is larger)
    assert scaled_error(y, g["y_ref"], g["cond_abs"]) < 1e-12
    # accuracy: at least as close to the 80-bit referee as the reference itself (up to a few ulps of the result)
    y_ld = long_double(g, "y")
    scale = np.max(np.abs(g["y_ref"]))
    err_new = np.max(np.abs((y - y_ld).astype(float))) / scale
    err_ref = np.max(np.abs((g["y_ref"] - y_ld).astype(float))) / scale
    assert err_new <= max(err_ref, 5e-14)
    st = plan.stats
    assert st["n_terms"] >= st["n_entries"] and st["padded_fma"] >= st["n_terms"] - 1
    if str(g["rule"]) == "leja":
        assert st["nested"] == 1  # (a Gauss-Hermite case with a single degree per dimension is trivially "nested")


@pytest.mark.parametrize("case", ALL_CASES)
def test_plan_gradient_sets_reproduce_reference_gradients(case):
    """The gradient jobs of the plan (smx_plan.cpp section 11: per hot dimension the derivative polynomial over the terms
    that carry a coefficient, per cold block the row sums of its value items; every entry of J written by exactly one job):
    equal to the reference's gradient wherever that one is finite, finite at the nodes, and at least as close to the 80-bit
    referee."""
    g = load(case)
    if case in LAYOUT_CASES:
        layout = golden_layout(g)
    else:
        kwargs, f = interpolator_inputs(g)
        layout, _ = SmolyakBarycentricInterpolator(**kwargs)._assemble(f, {})
    if case in ILL_CONDITIONED:
        pytest.skip("the plan compiler refuses this layout (conditioning check): no derivative sets")
    J_ref = g["J_ref"]
    x = g["x"][: len(J_ref)]
    J = Plan(layout, x.shape[1], int(g["d_out"])).gradient(x)
    assert np.isfinite(J).all()
    ok = ~np.isnan(J_ref)
    scale = max(1.0, float(np.max(np.abs(J_ref[ok]))))
    assert np.max(np.abs(J[ok] - J_ref[ok])) <= 1e-10 * scale
    J_ld = long_double(g, "J")
    err_new = np.max(np.abs((J - J_ld)[ok].astype(float))) / scale
    err_ref = np.max(np.abs((J_ref - J_ld)[ok].astype(float))) / scale
    assert err_new <= max(err_ref, 1e-12)


def test_plan_structure_of_headline_config():
    g = load("cfg2")
    kwargs, f = interpolator_inputs(g)
    layout, _ = SmolyakBarycentricInterpolator(**kwargs)._assemble(f, {})
    st = Plan(layout, 1000, 1).stats
    assert st["n_terms"] == 9999 and st["n_summands"] == 8751 and st["w_raw"] == 50866 and st["w_pad"] == 234124
    assert st["n_entries"] == 1058 and st["n_rows"] == 208 and st["n_hot"] >= 40
    assert st["padded_fma"] < 17000  # lane-FMAs per point; the reference's padded contraction has 234 124


def test_plan_rejects_malformed_layouts():
    g = load("small_00")
    layout = dict(golden_layout(g))
    n = sorted(int(k.split("_")[1]) for k in layout if k.startswith("zetas_"))[0]
    bad = dict(layout)
    bad[f"dims_{n}"] = layout[f"dims_{n}"] + 100
    assert "out of range" in Plan(bad, g["x"].shape[1], int(g["d_out"])).error
    bad = dict(layout)
    nodes = layout[f"nodes_{n}"].copy()
    nodes[:, :, 1] = nodes[:, :, 0]  # duplicate node
    bad[f"nodes_{n}"] = nodes
    assert "distinct" in Plan(bad, g["x"].shape[1], int(g["d_out"])).error


WELL_CONDITIONED = [c for c in ALL_CASES if c not in ILL_CONDITIONED]


@pytest.mark.parametrize("case", WELL_CONDITIONED)
def test_dense_form_reproduces_reference_values(case):
    """The GEMM-regime form (one column of Phi per term, dense coefficient matrix in DMMA fragment order) evaluated on
    the host in the kernel's order agrees with the reference outputs like the block-sparse form does."""
    g = load(case)
    layout = _layout_of(g)
    d_in, d_out = g["x"].shape[1], int(g["d_out"])
    plan = Plan(layout, d_in, d_out, options=DENSE)
    assert plan.error is None, plan.error
    y = plan.dense(g["x"])
    assert scaled_error(y, g["y_ref"], g["cond_abs"]) < 1e-12
    assert scaled_error(y, Plan(layout, d_in, d_out)(g["x"]), g["cond_abs"]) < 1e-13


@pytest.mark.parametrize("case", [c for c in LAYOUT_CASES if c not in ILL_CONDITIONED])
def test_compact_layout_gives_the_same_plan(case):
    """smx_create_compact's description (exact shapes, node-indexed values) compiles to the same coefficients as the
    reference's padded per-group layout, and its host quadrature equals the reference integral."""
    g = load(case)
    kwargs, f = interpolator_inputs(g)
    d_in, d_out = g["x"].shape[1], int(g["d_out"])
    ip = SmolyakBarycentricInterpolator(**kwargs)
    compact, evals = ip._assemble_compact(f, {})
    # one row per function evaluation (the evaluation at the centre lives in `offset` alone for non-nested rules)
    assert compact["values"].shape in ((ip.n_f_evals, d_out), (ip.n_f_evals - 1, d_out)) and ip.n_f_evals_new == ip.n_f_evals
    padded, evals_ref = SmolyakBarycentricInterpolator(**kwargs)._assemble(f, {})
    assert set(evals) == set(evals_ref)
    a, b = Plan(compact, d_in, d_out, GRADIENT | SPARSE | DENSE), Plan(padded, d_in, d_out, GRADIENT | SPARSE | DENSE)
    assert a.error is None and b.error is None, (a.error, b.error)
    assert {k: v for k, v in a.stats.items() if k != "w_pad"} == {k: v for k, v in b.stats.items() if k != "w_pad"}
    x = g["x"][:64]
    assert np.array_equal(a(x), b(x)) and np.array_equal(a.dense(x), b.dense(x))
    assert np.array_equal(a.gradient(x), b.gradient(x))
    desc, keep = _compact_desc(compact)
    q = np.zeros(d_out)
    off = np.ascontiguousarray(np.broadcast_to(np.asarray(compact["offset"], dtype=np.float64), (d_out,)))
    assert _host.smxh_integrate_compact(d_out, off.ctypes.data, ctypes.addressof(desc), q.ctypes.data) == 0
    np.testing.assert_allclose(q, g["Q_ref"], rtol=2e-9, atol=2e-9)  # the reference's own summation noise (test_oracle.py)


def test_dense_kernel_block_dealing_covers_every_block_exactly_once():
    """smx_plan.h::deal_blocks (run by thread 0 of every CTA of the staged dense kernel): for every CTA shape, group size, ticket
    and placement of the warps on the four FP64 pipes - the regular rotations the hardware uses and arbitrary ones - every
    block of the group is multiplied exactly once (as a whole, or as two halves on two warps), no warp gets more than its
    accumulators hold, and with a regular placement the pipes differ by at most one unit (half a block for two-block warps),
    consecutive tickets putting their extra units on different pipes."""
    _host.smxh_deal_blocks.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
    _host.smxh_deal_blocks.restype = None
    rng = np.random.default_rng(5)

    def deal(nw, nb, pipe, first, blocks, ticket):
        pipe = np.ascontiguousarray(pipe, dtype=np.int32)
        out = np.full(4 * nw, -7, dtype=np.int32)
        _host.smxh_deal_blocks(nw, nb, pipe.ctypes.data, first, blocks, ticket, out.ctypes.data)
        return out.reshape(4, nw)

    for nw, nb in ((8, 2), (8, 1), (16, 1), (16, 2)):
        placements = [[(w + r) % 4 for w in range(nw)] for r in range(4)]
        placements += [rng.integers(0, 4, nw).tolist() for _ in range(12)] + [[0] * nw, [1] * (nw - 1) + [3]]
        for pi, pipe in enumerate(placements):
            regular = pi < 4
            for blocks in range(1, nw * nb):
                loads_by_ticket = []
                for ticket in (0, 1, 2, 3, 5, -1, 2**31 - 1):
                    first = 16 * (ticket & 3)
                    jb0, nbv, hblk, hsel = deal(nw, nb, pipe, first, blocks, ticket)
                    cover = np.zeros((blocks, 2), dtype=int)  # halves of every block of the group
                    load = np.zeros(4, dtype=int)             # units of half a block per pipe
                    for w in range(nw):
                        assert 0 <= nbv[w] <= nb and nbv[w] + (hblk[w] >= 0) <= nb
                        for j in range(nbv[w]):
                            cover[jb0[w] + j - first] += 1
                        load[pipe[w]] += 2 * nbv[w]
                        if hblk[w] >= 0:
                            assert nb == 2 and hsel[w] in (0, 1)
                            cover[hblk[w] - first, hsel[w]] += 1
                            load[pipe[w]] += 1
                    assert (cover == 1).all(), (nw, nb, pipe, blocks, ticket, cover.T)
                    if regular:
                        unit = 1 if nb == 2 else 2
                        assert load.max() - load.min() <= unit, (nw, nb, pipe, blocks, ticket, load)
                    loads_by_ticket.append(load)
                if regular and nb == 2 and blocks % 2 == 1:  # 26 halves: two tickets in a row level the SM
                    assert (loads_by_ticket[0] + loads_by_ticket[1]).max() - (loads_by_ticket[0] + loads_by_ticket[1]).min() == 0


def test_compact_descriptor_offsets_must_start_at_zero():
    """The compact descriptor carries no array lengths: an offset array shifted as a whole passes every per-summand difference
    check and would read past the end of val_index / dims (a GPU negative test was flaky on exactly that before the plan
    compiler checked the first offsets).  Refused deterministically, with a message that names the arrays."""
    g = load("small_02")
    kwargs, f = interpolator_inputs(g)
    d_in, d_out = g["x"].shape[1], int(g["d_out"])
    good, _ = SmolyakBarycentricInterpolator(**kwargs)._assemble_compact(f, {})
    assert Plan(good, d_in, d_out).error is None
    for key in ("val_off", "slot_off"):
        bad = dict(good)
        # (padded so that even the shifted offsets stay inside the arrays this test owns)
        bad[key] = np.ascontiguousarray(good[key] + 1)
        bad["val_index"] = np.concatenate([good["val_index"], good["val_index"][-1:]])
        bad["dims"] = np.concatenate([good["dims"], good["dims"][-1:]])
        bad["degs"] = np.concatenate([good["degs"], good["degs"][-1:]])
        bad["node_off"] = np.concatenate([good["node_off"], good["node_off"][-1:]])
        plan = Plan(bad, d_in, d_out)
        assert plan.error is not None and "must start at 0" in plan.error, plan.error


def test_compact_layout_reuses_f_evals():
    g = load("small_03")
    kwargs, f = interpolator_inputs(g)
    ip = SmolyakBarycentricInterpolator(**kwargs)
    _, evals = ip._assemble_compact(f, {})
    again = SmolyakBarycentricInterpolator(**kwargs)
    layout, _ = again._assemble_compact(lambda x: 1 / 0, evals)  # nothing new to evaluate
    assert again.n_f_evals_new == 0 and layout["values"].shape[0] in (again.n_f_evals, again.n_f_evals - 1)
