"""The plan compiler (csrc/smx_plan.cpp): reference layout -> hierarchical block-sparse layout.  Runs on the CPU
through libsmolyax_host.so; `smxh_plan_eval_host` mirrors the kernel's arithmetic and is a verification aid only."""
import ctypes

import numpy as np
import pytest

from smolyax_b200 import _build
from smolyax_b200.interpolation import SmolyakBarycentricInterpolator
from helpers import ALL_CASES, LAYOUT_CASES, golden_layout, interpolator_inputs, load, long_double, scaled_error


class _Group(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("nn", ctypes.c_int64)] + [
        (name, ctypes.c_void_p) for name in ("tau", "F", "nodes", "weights", "dims", "degs", "zetas", "quad")]


_host = ctypes.CDLL(str(_build.build_host()))
_host.smxh_plan_build.restype = ctypes.c_void_p
_host.smxh_plan_build.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
_host.smxh_plan_error.restype = ctypes.c_char_p
_host.smxh_plan_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
_host.smxh_plan_eval_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
_host.smxh_plan_gradient_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
_host.smxh_plan_free.argtypes = [ctypes.c_void_p]
STATS = ("n_terms", "n_entries", "n_rows", "n_hot", "n_chunks", "padded_fma", "n_levels", "nested", "n_summands", "w_raw", "w_pad")


class Plan:
    def __init__(self, layout, d_in, d_out):
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
        off = np.ascontiguousarray(np.broadcast_to(np.asarray(layout["offset"], dtype=np.float64), (d_out,)))
        self.d_out = d_out
        self.h = _host.smxh_plan_build(d_in, d_out, off.ctypes.data, len(ns), ctypes.addressof(arr))
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

    def gradient(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        J = np.zeros((len(x), self.d_out, x.shape[1]))
        _host.smxh_plan_gradient_host(self.h, x.ctypes.data, len(x), x.shape[1], J.ctypes.data)
        return J

    def __del__(self):
        if getattr(self, "h", None):
            _host.smxh_plan_free(self.h)


@pytest.mark.parametrize("case", ALL_CASES)
def test_plan_reproduces_reference_values(case):
    g = load(case)
    if case in LAYOUT_CASES:
        layout = golden_layout(g)
    else:
        kwargs, f = interpolator_inputs(g)
        layout, _ = SmolyakBarycentricInterpolator(**kwargs)._assemble(f, {})
    plan = Plan(layout, g["x"].shape[1], int(g["d_out"]))
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
    """d/dx_i as extra coefficient sets on the same terms (hot dimensions) + row sums (cold dimensions): equal to the
    reference's gradient wherever that one is finite, finite at the nodes, and at least as close to the 80-bit referee."""
    g = load(case)
    if case in LAYOUT_CASES:
        layout = golden_layout(g)
    else:
        kwargs, f = interpolator_inputs(g)
        layout, _ = SmolyakBarycentricInterpolator(**kwargs)._assemble(f, {})
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
