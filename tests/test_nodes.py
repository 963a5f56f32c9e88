"""Node families: shapes and scaling round trips (reference tests/test_nodes.py), table values, quadrature
exactness, and exact equality with the reference's tables in the build container."""
import sys

import numpy as np
import pytest

from smolyax_b200 import nodes


def _random_generators(rng, n, dmin, dmax):
    out = []
    for _ in range(n):
        d = int(rng.integers(dmin, dmax + 1))
        out.append(nodes.Leja(domains=np.sort(rng.random((d, 2)), axis=1)))
        out.append(nodes.GaussHermite(rng.standard_normal(d), rng.random(d)))
    return out


def test_random_points_and_scaling_round_trip():
    rng = np.random.default_rng(0)
    for d in range(1, 4):
        for gen in (nodes.Leja(dim=d), nodes.GaussHermite(dim=d)):
            n = int(rng.integers(1, 5))
            x = gen.get_random(n)
            assert x.shape == (n, gen.dim)
            assert np.allclose(x, gen.scale(gen.scale_back(x)))
    for gen in _random_generators(rng, 5, 2, 10):
        n = int(rng.integers(0, 5))
        x = gen.get_random(n)
        assert x.shape == ((gen.dim,) if n == 0 else (n, gen.dim))
        assert np.allclose(x, gen.scale(gen.scale_back(x)))


def test_leja_sequence_values():
    pts = nodes.Leja1D()(8)
    s = 1 / np.sqrt(2)
    assert np.array_equal(pts[:5], [0, 1, -1, s, -s])
    assert pts[5] == np.sqrt((pts[3] + 1) / 2) and pts[6] == -pts[5] and pts[7] == np.sqrt((pts[4] + 1) / 2)
    assert np.array_equal(nodes.Leja1D()(3), pts[:4])  # nested
    scaled = nodes.Leja1D([2.0, 6.0])(4)
    assert np.allclose(scaled, 4.0 + 2.0 * pts[:5])
    with pytest.raises(ValueError):
        nodes.Leja()
    with pytest.raises(ValueError):
        nodes.GaussHermite()
    with pytest.raises(AssertionError):
        nodes.Leja1D([0.0, 1.0]).scale(np.array([1.5]))  # outside the reference interval


def test_quadrature_weights_are_exact_for_polynomials():
    for n in range(0, 9):
        w = nodes.Leja1D().get_quadrature_weights(n)
        pts = nodes.Leja1D()(n)
        for p in range(n + 1):  # uniform probability measure on [-1, 1]
            assert np.isclose(np.dot(w, pts**p), (1 + (-1) ** p) / (2 * (p + 1)), atol=1e-12)
        wg = nodes.GaussHermite1D().get_quadrature_weights(n)
        xg = nodes.GaussHermite1D()(n)
        assert np.isclose(wg.sum(), 1.0) and np.isclose(np.dot(wg, xg**2), 0.5 if n >= 1 else 0.0)
    assert not nodes.GaussHermite(dim=2).is_nested and nodes.Leja(dim=2).is_nested


@pytest.mark.reference
def test_identical_to_reference_tables():
    sys.path[:0] = [str(__import__("pathlib").Path(__file__).resolve().parent.parent / "oracle" / "jax_stub"), "/root/reference/src"]
    import smolyax.nodes as ref

    for n in range(0, 41):
        assert np.array_equal(ref.Leja1D()(n), nodes.Leja1D()(n))
        assert np.array_equal(ref.Leja1D([0.2, 0.9])(n), nodes.Leja1D([0.2, 0.9])(n))
        assert np.array_equal(ref.GaussHermite1D()(n), nodes.GaussHermite1D()(n))
        assert np.array_equal(ref.GaussHermite1D(0.3, 0.7)(n), nodes.GaussHermite1D(0.3, 0.7)(n))
        assert np.array_equal(ref.GaussHermite1D().get_quadrature_weights(n), nodes.GaussHermite1D().get_quadrature_weights(n))
        if n < 25:
            assert np.array_equal(ref.Leja1D().get_quadrature_weights(n), nodes.Leja1D().get_quadrature_weights(n))
    rng = np.random.default_rng(5)
    dom = np.sort(rng.random((4, 2)), axis=1)
    a, b = ref.Leja(domains=dom), nodes.Leja(domains=dom)
    x = b.get_random(5)
    assert np.array_equal(a.scale_back(x), b.scale_back(x)) and np.array_equal(a.scale_back(x[0]), b.scale_back(x[0]))
