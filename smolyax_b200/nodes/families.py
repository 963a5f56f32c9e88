"""Leja (nested, bounded interval) and Gauss-Hermite (non-nested, real line) node families.

Tables must equal the reference's exactly, so the arithmetic follows
/root/reference/src/smolyax/nodes/leja.py:17,48-57 (recursion), :83-93 (quadrature weights), :146-148 (affine map)
and /root/reference/src/smolyax/nodes/gausshermite.py:57-62,81-87,105,122 — the same NumPy calls in the same order.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from .core import Generator, Generator1D

_SEED_LEJA = [0.0, 1.0, -1.0, 1 / np.sqrt(2), -1 / np.sqrt(2)]


class _LejaSequence:
    """The one shared, growing reference sequence on [-1, 1]: xi_j = -xi_{j-1} for even j and
    sqrt((xi_{(j+1)/2} + 1) / 2) for odd j (leja.py:48-57)."""

    values = np.array(_SEED_LEJA)

    @classmethod
    def first(cls, count: int) -> np.ndarray:
        have = cls.values.shape[0]
        if count > have:
            grown = np.empty(count)
            grown[:have] = cls.values
            for j in range(have, count):
                grown[j] = -grown[j - 1] if j % 2 == 0 else np.sqrt((grown[(j + 1) // 2] + 1) / 2)
            cls.values = grown
        return cls.values[:count]


def _affine(x, src, dst):
    """Affine map of the interval ``src`` onto ``dst`` with the reference's in-domain checks (leja.py:125-151)."""
    src, dst = np.squeeze(np.asarray(src, dtype=float)), np.squeeze(np.asarray(dst, dtype=float))
    assert src.shape == dst.shape == (2,), f"shapes {src.shape} and {dst.shape} do not match (2, )"
    assert src[0] < src[1] and dst[0] < dst[1]
    x = np.asarray(x)
    shape = x.shape
    flat = np.squeeze(x)
    assert flat.ndim <= 1
    low_ok = (flat >= src[0]) | np.isclose(flat, src[0])
    high_ok = (flat <= src[1]) | np.isclose(flat, src[1])
    assert np.all(low_ok), f"Assertion failed: Some values are below lower bounds\n{flat[~low_ok]}"
    assert np.all(high_ok), f"Assertion failed: Some values are above upper bounds\n{flat[~high_ok]}"
    unit = (flat - src[0]) / (src[1] - src[0])
    return (unit * (dst[1] - dst[0]) + dst[0]).reshape(shape)


class Leja1D(Generator1D):
    """Nested Leja nodes on ``domain`` (default ``[-1, 1]``)."""

    _REF = (-1, 1)

    def __init__(self, domain: Sequence[float] = None) -> None:
        super().__init__(is_nested=True)
        self._domain = None if domain is None else np.asarray(domain)
        self._node_cache: dict[int, np.ndarray] = {}
        self._quad_cache: dict[int, np.ndarray] = {}

    @property
    def domain(self):
        """Interval end points, or ``None`` for the reference interval."""
        return self._domain

    def __call__(self, n: int) -> np.ndarray:
        if n not in self._node_cache:
            self._node_cache[n] = self.scale(_LejaSequence.first(n + 1))
        return self._node_cache[n]

    def scale(self, x, d1=None, d2=None):
        if d1 is None:
            d1 = None if self._domain is None else self._REF
        if d2 is None:
            d2 = self._domain
        assert (d1 is None) == (d2 is None)
        if d1 is None:
            return x
        return _affine(x, d1, d2)

    def scale_back(self, x):
        if self._domain is None:
            return x
        return _affine(x, self._domain, self._REF)

    def get_random(self, n: int = 1):
        return self.scale(np.random.uniform(-1, 1, n))

    def get_quadrature_weights(self, n: int) -> np.ndarray:
        """Interpolatory weights w.r.t. the uniform probability measure: Vandermonde solve against the
        moments (1 + (-1)^i) / (2 (i + 1)) on the reference nodes (leja.py:83-93)."""
        if n not in self._quad_cache:
            pts = _LejaSequence.first(n + 1)
            vandermonde = np.vstack([pts**i for i in range(n + 1)])
            moments = np.array([(1 + (-1) ** i) / (2.0 * (i + 1)) for i in range(n + 1)])
            self._quad_cache[n] = np.linalg.solve(vandermonde, moments)
        return self._quad_cache[n]

    def __repr__(self) -> str:
        return f"Leja (domain = {self._domain})"


class Leja(Generator):
    """Leja nodes in every dimension; ``domains`` is a list of intervals, or give ``dim`` for ``[-1, 1]^dim``."""

    def __init__(self, *, domains=None, dim: int = None):
        if domains is not None:
            super().__init__([Leja1D(dom) for dom in domains])
            self._domains = np.asarray(domains)
        elif dim is not None:
            super().__init__([Leja1D()] * dim)  # one shared object, as in leja.py:242
            self._domains = None
        else:
            raise ValueError("Must specify one of 'domains' or 'dim'.")

    def __repr__(self) -> str:
        if self._domains is not None:
            return f"Leja (d = {self.dim}, domains = {self._domains.tolist()})"
        return f"Leja (d = {self.dim})"


class GaussHermite1D(Generator1D):
    """Gauss-Hermite nodes ``mean + scaling * hermgauss(n+1)``; weights w.r.t. exp(-x^2)/sqrt(pi)."""

    def __init__(self, mean: float = 0.0, scaling: float = 1.0) -> None:
        super().__init__(is_nested=False)
        self._mean = mean
        self._scaling = scaling
        self._node_cache: dict[int, np.ndarray] = {}
        self._quad_cache: dict[int, np.ndarray] = {}

    @property
    def mean(self) -> float:
        return self._mean

    @property
    def scaling(self) -> float:
        return self._scaling

    def __call__(self, n: int) -> np.ndarray:
        if n not in self._node_cache:
            self._node_cache[n] = self.scale(np.polynomial.hermite.hermgauss(n + 1)[0])
        return self._node_cache[n]

    def scale(self, x):
        return self._mean + self._scaling * x

    def scale_back(self, x):
        return (x - self._mean) / self._scaling

    def get_random(self, n: int = 1):
        return self.scale(np.random.randn(n) / np.sqrt(2))

    def get_quadrature_weights(self, n: int) -> np.ndarray:
        if n not in self._quad_cache:
            self._quad_cache[n] = np.polynomial.hermite.hermgauss(n + 1)[1] / np.sqrt(np.pi)
        return self._quad_cache[n]

    def __repr__(self) -> str:
        return f"Gauss-Hermite (mean = {self._mean}, scaling = {self._scaling})"


class GaussHermite(Generator):
    """Gauss-Hermite nodes in every dimension with optional per-dimension ``mean`` and ``scaling``."""

    def __init__(self, mean=None, scaling=None, dim: int = None):
        if dim is None:
            if scaling is not None:
                dim = len(scaling)
            elif mean is not None:
                dim = len(mean)
            else:
                raise ValueError("Must specify at least one of 'dim', 'mean', or 'scaling'.")
        self._mean = np.zeros(dim) if mean is None else np.asarray(mean)
        self._scaling = np.ones(dim) if scaling is None else np.asarray(scaling)
        if mean is None and scaling is None:
            super().__init__([GaussHermite1D()] * dim)  # one shared object, as in gausshermite.py:202
        else:
            super().__init__([GaussHermite1D(m, a) for m, a in zip(self._mean, self._scaling)])

    def scale(self, x):
        return self._mean + self._scaling * x

    def scale_back(self, x):
        return (x - self._mean) / self._scaling

    def __repr__(self) -> str:
        return f"Gauss Hermite (d = {self.dim}, mean = {self._mean.tolist()}, scaling = {self._scaling.tolist()})"
