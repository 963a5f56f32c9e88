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


@pytest.mark.parametrize("rule, batched", [("leja", False), ("leja", True), ("gh", False), ("gh", True)])
def test_compact_assembly_walks_like_the_padded_one(rule, batched):
    """Same calls of ``f`` in the same order, same ``f_evals``, and ``values[val_index]`` holds exactly the entries of
    the padded tensors ``F_n`` at the grid points of each summand (C order over the sorted slots)."""
    d_in, d_out, t = 6, 2, 6.0
    k = workloads.anisotropy(d_in)
    gen = nodes.Leja(dim=d_in) if rule == "leja" else nodes.GaussHermite(dim=d_in)
    fam = workloads.TargetFamily(d_in, d_out)
    out = {}
    for mode in ("reference", "compact"):
        calls = []

        def f(x):
            calls.extend(np.atleast_2d(np.asarray(x)).copy())
            return fam(x)

        ip = SmolyakBarycentricInterpolator(node_gen=gen, k=k, t=t, d_out=d_out, batched_f=batched, layout=mode)
        layout, evals = (ip._assemble if mode == "reference" else ip._assemble_compact)(f, {})
        out[mode] = (layout, evals, np.array(calls), ip.n_f_evals_new)
    (ref, ev_r, calls_r, new_r), (cmp_, ev_c, calls_c, new_c) = out["reference"], out["compact"]
    assert new_r == new_c and np.array_equal(calls_r, calls_c)
    assert list(ev_r) == list(ev_c)
    if rule != "leja":
        assert all(list(ev_r[nu]) == list(ev_c[nu]) for nu in ev_r)
    assert np.array_equal(ref["offset"], cmp_["offset"])
    s = 0
    for n in [int(key[6:]) for key in ref if key.startswith("zetas_")]:
        F, degs, dims = ref[f"F_{n}"], ref[f"degs_{n}"], ref[f"dims_{n}"]
        for i in range(len(degs)):
            lo, hi = cmp_["slot_off"][s], cmp_["slot_off"][s + 1]
            assert np.array_equal(cmp_["dims"][lo:hi], dims[i]) and np.array_equal(cmp_["degs"][lo:hi], degs[i])
            assert cmp_["zetas"][s] == ref[f"zetas_{n}"][i]
            block = F[i][(slice(None),) + tuple(slice(0, int(v) + 1) for v in degs[i])].reshape(d_out, -1).T
            rows = cmp_["val_index"][cmp_["val_off"][s]:cmp_["val_off"][s + 1]]
            assert np.array_equal(cmp_["values"][rows], block)
            for j in range(n):
                o = cmp_["node_off"][lo + j]
                assert np.array_equal(cmp_["node_pool"][o:o + degs[i][j] + 1], ref[f"nodes_{n}"][i, j, :degs[i][j] + 1])
                assert np.array_equal(cmp_["quad_pool"][o:o + degs[i][j] + 1], ref[f"quad_{n}"][i, j, :degs[i][j] + 1])
            s += 1
    assert s == len(cmp_["zetas"]) and cmp_["values"].shape[0] == (new_c if rule == "leja" else new_c - 1)


@pytest.mark.parametrize("mode", ["reference", "compact"])
def test_layout_file_round_trip(tmp_path, mode):
    from smolyax_b200.interpolation import read_layout, write_layout

    d_in, d_out, t = 5, 3, 5.0
    k = workloads.anisotropy(d_in)
    ip = SmolyakBarycentricInterpolator(node_gen=nodes.Leja(dim=d_in), k=k, t=t, d_out=d_out, layout=mode)
    layout, _ = (ip._assemble if mode == "reference" else ip._assemble_compact)(workloads.TargetFamily(d_in, d_out), {})
    path = tmp_path / "tables.npz"
    write_layout(layout, path, d_in=d_in, d_out=d_out, k=k, t=t)
    back, meta = read_layout(path)
    assert list(back) == list(layout) and bool(back.get("compact")) == (mode == "compact")
    for key, val in layout.items():
        assert np.array_equal(np.asarray(val), back[key]) and np.asarray(val).dtype == np.asarray(back[key]).dtype, key
    assert meta["d_in"] == d_in and meta["d_out"] == d_out and meta["t"] == t and np.array_equal(meta["k"], k)
    other = SmolyakBarycentricInterpolator(node_gen=nodes.Leja(dim=d_in), k=k, t=t + 0.5, d_out=d_out)
    with pytest.raises(AssertionError, match="another index set"):
        other.load_layout(path)


@pytest.mark.reference
@pytest.mark.parametrize("rule", ["leja", "gh"])
@pytest.mark.parametrize("mode, batched", [("reference", False), ("reference", True), ("compact", False), ("compact", True)])
def test_set_f_calls_f_like_the_unmodified_reference(rule, mode, batched):
    """The reference itself (on the NumPy `jax` stand-in, build container only) and this package call ``f`` at the same
    points in the same order and return the same ``f_evals`` (keys, order, values), also when a dictionary is reused."""
    import sys
    from pathlib import Path

    sys.path[:0] = [str(Path(__file__).resolve().parent.parent / "oracle" / "jax_stub"), "/root/reference/src"]
    from smolyax import nodes as rnodes
    from smolyax.interpolation import SmolyakBarycentricInterpolator as Reference

    d_in, d_out = 6, 2
    k = workloads.anisotropy(d_in)
    fam = workloads.TargetFamily(d_in, d_out)

    def run(cls, gen, t, evals, **kw):
        calls = []

        def f(x):
            calls.extend(np.atleast_2d(np.array(x, dtype=float)))
            return fam(x)

        ip = cls(node_gen=gen, k=k, t=t, d_out=d_out, **kw)
        if cls is Reference:
            evals = ip.set_f(f=f, f_evals=evals)
        else:
            evals = (ip._assemble if mode == "reference" else ip._assemble_compact)(f, evals)[1]
        return np.array(calls), evals, ip.n_f_evals_new

    gen_r = rnodes.Leja(dim=d_in) if rule == "leja" else rnodes.GaussHermite(dim=d_in)
    gen = nodes.Leja(dim=d_in) if rule == "leja" else nodes.GaussHermite(dim=d_in)
    ev_r, ev = {}, {}
    for t in (4.0, 5.5):  # the second pass reuses the dictionary of the first
        calls_r, ev_r, new_r = run(Reference, gen_r, t, ev_r)
        calls, ev, new = run(SmolyakBarycentricInterpolator, gen, t, ev, batched_f=batched, layout=mode)
        assert new == new_r and calls.shape == calls_r.shape and np.array_equal(calls, calls_r)
        assert list(ev) == list(ev_r)
        same = np.array_equal if not batched else (lambda a, b: np.allclose(a, b, rtol=1e-15, atol=0))  # (f's own rounding)
        if rule == "leja":
            assert all(same(ev[key], ev_r[key]) for key in ev)
        else:
            assert all(list(ev[nu]) == list(ev_r[nu]) and all(same(ev[nu][key], ev_r[nu][key]) for key in ev[nu]) for nu in ev)
