"""Pins the CPU oracle (oracle/smx_oracle.c) on outputs of the unmodified reference (tests/golden, produced by
oracle/make_golden.py on the NumPy jax stand-in): values, gradients with their NaN pattern, integrals, weights."""
import numpy as np
import pytest

from oracle import oracle
from smolyax_b200.interpolation import SmolyakBarycentricInterpolator
from helpers import BIG_CASES, LAYOUT_CASES, golden_layout, interpolator_inputs, load, long_double, scaled_error


def _check(layout, g, tol_scaled):
    y = oracle.evaluate(layout, g["x"])
    assert y.shape == g["y_ref"].shape
    assert scaled_error(y, g["y_ref"], g["cond_abs"]) < tol_scaled
    J_ref = g["J_ref"]
    J = oracle.gradient(layout, g["x"][: len(J_ref)])
    assert J.shape == J_ref.shape
    assert np.array_equal(np.isnan(J), np.isnan(J_ref))  # NaN exactly where the reference has it
    ok = ~np.isnan(J_ref)
    if ok.any():
        assert np.max(np.abs(J[ok] - J_ref[ok])) <= 1e-10 * max(1.0, np.max(np.abs(J_ref[ok])))
    Q = oracle.integral(layout)
    # signed sum of up to 16 000 summands with sum|zeta| up to 3.4e4 and an O(1) result: two fp64 summation orders
    # differ by ~1e-16 * n * sum|zeta| (measured: oracle 3e-10, reference 5e-11 away from the 80-bit referee at cfg4)
    assert np.allclose(Q, g["Q_ref"], rtol=2e-9, atol=2e-9)


@pytest.mark.parametrize("case", LAYOUT_CASES)
def test_oracle_matches_reference_outputs_on_stored_layouts(case):
    g = load(case)
    _check(golden_layout(g), g, 1e-13)


@pytest.mark.parametrize("case", BIG_CASES)
def test_oracle_matches_reference_outputs_on_baseline_configs(case):
    g = load(case)
    kwargs, f = interpolator_inputs(g)
    layout, _ = SmolyakBarycentricInterpolator(**kwargs)._assemble(f, {})  # digest-checked in test_layout.py
    _check(layout, g, 1e-13)


def test_oracle_weights():
    w = oracle.compute_weights(np.array([0.0, 1.0, -1.0]))
    assert np.array_equal(w, [-1.0, 0.5, 0.5])
    rng = np.random.default_rng(0)
    pts = rng.standard_normal(9)
    diffs = pts[:, None] - pts
    diffs[diffs == 0] = 1
    assert np.allclose(oracle.compute_weights(pts), np.prod(1 / diffs, axis=0), rtol=1e-14)


@pytest.mark.parametrize("case", LAYOUT_CASES)
def test_long_double_referee_reproduces_the_reference_in_80_bit(case):
    """oracle.evaluate_referee / gradient_referee (the C restatement carried out in long double) against the golden outputs
    of the UNMODIFIED reference run in 80-bit arithmetic on the NumPy jax stand-in (oracle/make_golden.py): equal to a few
    units of the long-double precision, same NaN pattern - so bench.py may use it as the accuracy referee at any size."""
    from oracle import oracle

    g = load(case)
    layout = golden_layout(g)
    y = oracle.evaluate_referee(layout, g["x"])
    y_gold = long_double(g, "y")
    assert float(np.max(np.abs(y - y_gold))) <= 1e-15 * max(1.0, float(np.max(np.abs(g["y_ref"]))))
    J_gold = long_double(g, "J")
    J = oracle.gradient_referee(layout, g["x"][: len(J_gold)])
    ok = ~np.isnan(g["J_ref"])
    assert np.array_equal(np.isnan(J.astype(float)), ~ok)
    if ok.any():
        assert float(np.max(np.abs((J - J_gold)[ok]))) <= 1e-14 * max(1.0, float(np.max(np.abs(g["J_ref"][ok]))))
