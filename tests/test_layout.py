"""Host half of set_f: the per-group tables must equal the reference's (SHA-256 recorded by oracle/make_golden.py
from the unmodified reference), plus the f_evals reuse semantics of reference interpolation.py:119,155-163,208-228."""
import numpy as np
import pytest

from smolyax_b200 import indices, nodes, workloads
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


def _walk_like_the_reference(layout, zero, f, nested, groups):
    """The reference's grid walk (interpolation.py:203-228) restated with plain loops over the assembled tables: the
    order of the calls of ``f`` and of the keys of ``f_evals`` that the vectorised fill has to reproduce."""
    import itertools

    calls, evals = [], {}
    for n in groups:
        if n == 0:  # the offset term of the summand without active dimensions (interpolation.py:150-165)
            store = evals if nested else evals.setdefault((), {})
            if () not in store:
                calls.append(zero.copy())
                store[()] = f(zero)
            continue
        dims, degs, tab = layout[f"dims_{n}"], layout[f"degs_{n}"], layout[f"nodes_{n}"]
        for i in range(len(dims)):
            by_dim = np.argsort(dims[i])
            nu = tuple((int(dims[i][j]), int(degs[i][j])) for j in by_dim)
            store = evals if nested else evals.setdefault(nu, {})
            for mu in itertools.product(*[range(int(v) + 1) for v in degs[i]]):
                key = tuple((int(dims[i][j]), mu[j]) for j in by_dim if mu[j] > 0)
                if key not in store:
                    x = zero.copy()
                    x[dims[i]] = [tab[i, j, mu[j]] for j in range(n)]
                    calls.append(x)
                    store[key] = f(x)
    return calls, evals


@pytest.mark.parametrize("rule, batched", [("leja", False), ("leja", True), ("gh", False), ("gh", True)])
def test_vectorised_fill_keeps_the_order_of_the_reference_walk(rule, batched):
    d_in, d_out = 7, 3
    k = workloads.anisotropy(d_in)
    gen = nodes.Leja(dim=d_in) if rule == "leja" else nodes.GaussHermite(dim=d_in)
    fam = workloads.TargetFamily(d_in, d_out)
    calls = []

    def f(x):
        x = np.asarray(x)
        calls.extend(np.atleast_2d(x).copy())
        return fam(x)

    ip = SmolyakBarycentricInterpolator(node_gen=gen, k=k, t=5.5, d_out=d_out, batched_f=batched)
    layout, evals = ip._assemble(f, {})
    zero = np.array([g(0)[0] for g in gen])
    groups = list(indices.non_zero_indices_and_zetas(k, 5.5)[0])  # bins in the order the reference meets them
    ref_calls, ref_evals = _walk_like_the_reference(layout, zero, fam, rule == "leja", groups)
    assert np.any(layout["offset"] != 0.0) and len(calls) == len(ref_calls) == ip.n_f_evals_new
    assert np.array_equal(np.array(calls), np.array(ref_calls))
    if rule == "leja":
        assert list(evals) == list(ref_evals)
        assert all(np.allclose(evals[key], ref_evals[key], rtol=1e-15) for key in ref_evals)
    else:
        assert list(evals) == list(ref_evals)
        assert all(list(evals[nu]) == list(ref_evals[nu]) for nu in ref_evals)

    # a caller's dictionary with scalar entries beside arrays (d_out = 1) is reused and broadcast
    one = SmolyakBarycentricInterpolator(node_gen=gen, k=k, t=3.0, d_out=1, batched_f=batched)
    g1 = lambda x: np.sin(np.asarray(x).sum(axis=-1))
    layout1, evals1 = one._assemble(g1, {})
    mixed = {key: (float(np.asarray(v).reshape(-1)[0]) if i % 2 else np.asarray(v).reshape(1)) for i, (key, v) in
             enumerate(evals1.items())} if rule == "leja" else None
    if mixed is not None:
        again = SmolyakBarycentricInterpolator(node_gen=gen, k=k, t=3.0, d_out=1, batched_f=batched)
        layout2, _ = again._assemble(g1, mixed)
        assert again.n_f_evals_new == 0
        assert all(np.array_equal(layout1[key], layout2[key]) for key in layout1)
