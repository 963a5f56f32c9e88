"""Host combinatorics: golden cardinalities of the reference's own tests, brute-force cross-checks, and (in the
build container) exact equality with the reference's indices.py."""
import itertools as it
import sys

import numpy as np
import pytest

from smolyax_b200 import indices
from helpers import dense_indexset

# reference tests/test_indices_runtime.py:7-18
@pytest.mark.parametrize("d, t, m", [(10000, 5.3, 98), (10000, 7.78, 1000), (10000, 10.317, 10003)])
def test_golden_cardinalities(d, t, m):
    k = np.log([2 + i for i in range(d)]) / np.log(2)
    assert len(indices.indexset(k, t)) == m
    assert indices.indexset_cardinality(k, t) == m


# SURVEY.md Appendix B / BASELINE.md: thresholds and structure of the BASELINE configs
@pytest.mark.parametrize("d, n, nested, t_repr, n_nodes, n_summands", [
    (10, 1000, True, "8.831420927999998", 1000, 881),
    (1000, 10000, True, "8.224519679999998", 9999, 8751),
    (100, 10000, True, "8.749231385317653", 9986, 8808),
    (1000, 100000, False, "8.739042047999998", 99981, 15969),
])
def test_pinned_config_structure(d, n, nested, t_repr, n_nodes, n_summands):
    k = np.log((2 + np.arange(d)) / np.log(2))
    t = indices.find_approximate_threshold(k, n, nested)
    assert repr(float(t)) == t_repr
    assert indices.nodeset_cardinality(k, t, nested) == n_nodes
    n2nus, n2zetas = indices.non_zero_indices_and_zetas(k, t)
    assert sum(len(v) for n_, v in n2nus.items() if n_ > 0) == n_summands
    assert sum(sum(z) for z in n2zetas.values()) == 1  # the Smolyak coefficients sum to one


def _random_k(rng):
    d = int(rng.integers(1, 6))
    a, b = 1.1 + 2.9 * rng.random(), 0.1 + 1.9 * rng.random()
    return np.log([a + b * i for i in range(d)]) / np.log(a)


def test_membership_and_zeta_against_brute_force():
    rng = np.random.default_rng(7)
    for trial in range(12):
        k = _random_k(rng)
        t = indices.find_approximate_threshold(k, int(rng.integers(1, 100)), nested=bool(trial % 2))
        sparse = indices.indexset(k, t)
        dense = set(dense_indexset(list(k), t))
        as_dense = set()
        for nu in sparse:
            full = [0] * len(k)
            for dim, deg in nu:
                full[dim] = deg
            as_dense.add(tuple(full))
        assert as_dense == dense and len(sparse) == len(dense)
        for idx in it.product(*[range(int(np.floor(t / ki)) + 2) for ki in k]):
            assert (idx in dense) == (np.dot(idx, k) < t)
        # zeta = sum over e in {0,1}^d with nu+e in Lambda of (-1)^|e|
        for nu in dense:
            brute = sum((-1) ** sum(e) for e in it.product((0, 1), repeat=len(k)) if tuple(np.add(nu, e)) in dense)
            assert indices.smolyak_coefficient(k, len(k), t - np.dot(nu, k), 0) == brute


def test_nonzero_indices_are_consistent():
    k = np.log((2 + np.arange(30)) / np.log(2))
    t = 6.3
    n2nus, n2zetas = indices.non_zero_indices_and_zetas(k, t)
    everything = indices.indexset(k, t)
    listed = {nu for nus in n2nus.values() for nu in nus}
    assert listed <= set(everything)
    for nu in everything:
        z = indices.smolyak_coefficient(k, len(k), t - sum(k[d] * a for d, a in nu), 0)
        assert (nu in listed) == (z != 0)
    for n, nus in n2nus.items():
        assert all(len(nu) == n for nu in nus) and len(nus) == len(n2zetas[n])


def test_threshold_accuracy():  # reference tests/test_indices_runtime.py:21-34
    for d, m, nested in [(100, 1000, True), (100, 1000, False), (10000, 10000, True)]:
        accuracy = 0.01 if nested else 0.1
        k = np.log([2 + i for i in range(d)]) / np.log(2)
        t = indices.find_approximate_threshold(k, m, nested=nested, accuracy=accuracy)
        assert abs(indices.nodeset_cardinality(k, t, nested=nested) - m) / m < accuracy
    assert indices.find_approximate_threshold([1.0, 2.0], 1, True) == 1


@pytest.mark.reference
def test_identical_to_reference_indices():
    sys.path[:0] = [str(__import__("pathlib").Path(__file__).resolve().parent.parent / "oracle" / "jax_stub"), "/root/reference/src"]
    import smolyax.indices as ref

    rng = np.random.default_rng(3)
    for trial in range(60):
        k = _random_k(rng) if trial % 3 else np.sort(rng.uniform(1, 10, int(rng.integers(1, 7))))
        k = k / k[0]
        nested = bool(trial % 2)
        m = int(rng.integers(1, 300))
        t = ref.find_approximate_threshold(k, m, nested)
        assert repr(t) == repr(indices.find_approximate_threshold(k, m, nested))
        assert ref.indexset(k, t) == indices.indexset(k, t)
        assert ref.indexset_cardinality(k, t) == indices.indexset_cardinality(k, t)
        assert ref.nodeset_cardinality(k, t, nested) == indices.nodeset_cardinality(k, t, nested)
        a, za = ref.non_zero_indices_and_zetas(k, t)
        b, zb = indices.non_zero_indices_and_zetas(k, t)
        assert list(a.keys()) == list(b.keys()) and dict(a) == dict(b)
        assert {n: [int(v) for v in z] for n, z in za.items()} == dict(zb)
        # thresholds that sit exactly on sums of k (strict-inequality boundary)
        tb = float(np.dot(rng.integers(0, 3, len(k)), k)) + float(k[0])
        assert ref.indexset(k, tb) == indices.indexset(k, tb)
        assert ref.indexset_cardinality(k, tb) == indices.indexset_cardinality(k, tb)
