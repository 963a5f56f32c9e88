"""Abstract 1-D node sequence and the d-dimensional container of such sequences.

Interface as in /root/reference/src/smolyax/nodes/base.py:8-129 (``Generator1D``) and :131-266 (``Generator``).
User subclasses of ``Generator1D`` keep working: the interpolator only uses ``is_nested``, ``gen(deg)`` and
``gen.get_quadrature_weights(deg)``.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Iterator, List

import numpy as np


class Generator1D(ABC):
    """A family of 1-D node sequences: ``gen(n)`` returns the ``n+1`` nodes of degree ``n``."""

    def __init__(self, is_nested: bool) -> None:
        self._nested = bool(is_nested)

    @property
    def is_nested(self) -> bool:
        """True if the nodes of degree ``n`` are the first ``n+1`` nodes of every higher degree."""
        return self._nested

    @abstractmethod
    def __call__(self, n: int) -> np.ndarray:
        """Nodes of degree ``n`` (length ``n+1``), mapped to the custom domain."""

    @abstractmethod
    def scale(self, x):
        """Map points from the reference domain to the custom domain."""

    @abstractmethod
    def scale_back(self, x):
        """Map points from the custom domain back to the reference domain."""

    @abstractmethod
    def get_random(self, n: int = 1):
        """Draw ``n`` points from the probability measure the quadrature weights integrate against."""

    @abstractmethod
    def get_quadrature_weights(self, n: int):
        """Quadrature weights belonging to the nodes of degree ``n``."""


class Generator:
    """One :class:`Generator1D` per input dimension."""

    def __init__(self, node_gens: List[Generator1D]):
        nested = {bool(g.is_nested) for g in node_gens}
        assert len(nested) == 1, "all dimensions must be nested, or none"
        self._gens = list(node_gens)
        self._nested = nested.pop()

    @property
    def dim(self) -> int:
        return len(self._gens)

    @property
    def is_nested(self) -> bool:
        return self._nested

    def __getitem__(self, i: int) -> Generator1D:
        return self._gens[i]

    def __iter__(self) -> Iterator[Generator1D]:
        return iter(self._gens)

    def __len__(self) -> int:
        return len(self._gens)

    def get_random(self, n: int = 0):
        """``n == 0``: one point of shape ``(dim,)``; otherwise ``(n, dim)`` (reference base.py:208-226)."""
        if n == 0:
            return np.squeeze([g.get_random() for g in self._gens])
        return np.array([g.get_random(n) for g in self._gens]).T

    def _per_dim(self, method: str, x):
        x = np.asarray(x)
        assert x.shape[-1] == self.dim
        if x.ndim == 1:
            return np.array([getattr(g, method)(xi) for g, xi in zip(self._gens, x)])
        return np.array([getattr(g, method)(col) for g, col in zip(self._gens, x.T)]).T

    def scale(self, x):
        return self._per_dim("scale", x)

    def scale_back(self, x):
        return self._per_dim("scale_back", x)
