"""Host half of set_f: the per-group tables must equal the reference's (SHA-256 recorded by oracle/make_golden.py
from the unmodified reference), plus the f_evals reuse semantics of reference interpolation.py:119,155-163,208-228."""
import numpy as np
import pytest

from smolyax_b200 import nodes, workloads
from smolyax_b200.interpolation import SmolyakBarycentricInterpolator
from helpers import ALL_CASES, LAYOUT_CASES, golden_layout, interpolator_inputs, layout_digest, load


@pytest.mark.parametrize("case", ALL_CASES)
def test_reference_layout_is_reproduced_bit_for_bit(case):
    g = load(case)
    kwargs, f = interpolator_inputs(g)
    ip = SmolyakBarycentricInterpolator(**kwargs)
    layout, _ = ip._assemble(f, {})
    assert layout_digest(layout) == str(g["layout_sha256"])
    assert ip.n_f_evals == int(g["n_f_evals"]) == ip.n_f_evals_new
    assert ip.d_in == len(g["k"]) and ip.d_out == int(g["d_out"])


@pytest.mark.parametrize("case", LAYOUT_CASES[:6])
def test_stored_layout_arrays_match(case):
    g = load(case)
    kwargs, f = interpolator_inputs(g)
    layout, _ = SmolyakBarycentricInterpolator(**kwargs)._assemble(f, {})
    ref = golden_layout(g)
    for key, val in ref.items():
        assert np.array_equal(val, layout[key]), key  # includes the quadrature-weight tables


def test_f_evals_reuse_nested_and_non_nested():
    k = workloads.anisotropy(6)
    calls = []

    def f(x):
        calls.append(np.array(x))
        return np.array([np.sin(x.sum()), np.cos(x[0])])

    big = SmolyakBarycentricInterpolator(node_gen=nodes.Leja(dim=6), k=k, t=6.0, d_out=2)
    _, evals = big._assemble(f, {})
    assert big.n_f_evals_new == big.n_f_evals == len(evals) == len(calls)
    assert all(isinstance(key, tuple) for key in evals)  # flat {mu_tuple: value} for nested rules
    small = SmolyakBarycentricInterpolator(node_gen=nodes.Leja(dim=6), k=k, t=4.5, d_out=2)
    n_before = len(calls)
    _, evals2 = small._assemble(f, evals)
    assert small.n_f_evals_new == 0 and len(calls) == n_before and evals2 is evals

    gh = SmolyakBarycentricInterpolator(node_gen=nodes.GaussHermite(dim=4), k=workloads.anisotropy(4), t=5.0, d_out=2)
    _, ev = gh._assemble(f, {})
    assert gh.n_f_evals_new == gh.n_f_evals == sum(len(v) for v in ev.values())
    assert all(isinstance(v, dict) for v in ev.values())  # {nu: {mu_tuple: value}} for non-nested rules
    again = SmolyakBarycentricInterpolator(node_gen=nodes.GaussHermite(dim=4), k=workloads.anisotropy(4), t=5.0, d_out=2)
    again._assemble(f, ev)
    assert again.n_f_evals_new == 0


def test_batched_f_gives_the_same_tables():
    g = load("medium_00")
    kwargs, f = interpolator_inputs(g)
    one, _ = SmolyakBarycentricInterpolator(**kwargs)._assemble(f, {})
    ip = SmolyakBarycentricInterpolator(**kwargs, batched_f=True)
    many, _ = ip._assemble(f, {})
    assert ip.n_f_evals_new == ip.n_f_evals
    for key in one:
        assert np.array_equal(one[key], many[key]), key
