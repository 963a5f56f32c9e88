"""Synthetic workloads of the BASELINE configs: anisotropy vector, target function family, evaluation points.

The shapes follow SURVEY.md §8(d):  ``k_j = log((2+j)/log 2)`` (reference README.md:68), thresholds from
``indices.find_approximate_threshold``, default-domain generators, and the target family of the reference's
benchmark, ``f_o(x) = 1 / (1 + theta_o * sum_j x_j (j+2)^(-r_o))`` with ``default_rng(0)`` perturbations of
``r = 2``, ``theta = 0.1`` by +-30 % (reference benchmarking/testfunction.py:10-27).  Used by bench.py, the golden
generator and the tests, so that all of them talk about the same inputs.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import indices, nodes

BASE_R = 2.0
BASE_THETA = 0.1
NOISE = 0.3


def anisotropy(d_in: int) -> np.ndarray:
    return np.log((2 + np.arange(d_in)) / np.log(2))


class TargetFamily:
    """Vector-valued target ``f: R^{d_in} -> R^{d_out}``; accepts one point ``(d_in,)`` or a batch ``(N, d_in)``."""

    def __init__(self, d_in: int, d_out: int):
        rng = np.random.default_rng(0)
        r = BASE_R + rng.uniform(-BASE_R * NOISE, BASE_R * NOISE, d_out)
        theta = BASE_THETA + rng.uniform(-BASE_THETA * NOISE, BASE_THETA * NOISE, d_out)
        r[0], theta[0] = BASE_R, BASE_THETA
        self.theta = theta
        self.decay = (np.arange(d_in) + 2.0) ** (-r[:, None])  # (d_out, d_in)

    def __call__(self, x):
        return 1.0 / (1.0 + self.theta * (np.asarray(x) @ self.decay.T))


@dataclass
class Workload:
    name: str
    rule: str  # "leja" | "gh"
    d_in: int
    d_out: int
    n_target: int
    n_points: int

    def generator(self):
        return nodes.Leja(dim=self.d_in) if self.rule == "leja" else nodes.GaussHermite(dim=self.d_in)

    def k(self):
        return anisotropy(self.d_in)

    def threshold(self):
        return indices.find_approximate_threshold(self.k(), self.n_target, self.rule == "leja")

    def target(self):
        return TargetFamily(self.d_in, self.d_out)

    def points(self, n: int, seed: int = 0) -> np.ndarray:
        """Host-side points from the rule's probability measure: U(-1,1)^d (Leja) or N(0,1/2)^d (Gauss-Hermite)."""
        rng = np.random.default_rng(seed)
        if self.rule == "leja":
            return rng.uniform(-1.0, 1.0, size=(n, self.d_in))
        return rng.standard_normal((n, self.d_in)) / np.sqrt(2.0)


# BASELINE.json `configs`, in order.  cfg5 re-uses the cfg2 tables with d_out = 100.
CONFIGS = {
    "cfg1": Workload("cfg1", "leja", 10, 1, 1_000, 10_000),
    "cfg2": Workload("cfg2", "leja", 1_000, 1, 10_000, 1_000_000),
    "cfg3": Workload("cfg3", "leja", 100, 10_000, 10_000, 100_000),
    "cfg4": Workload("cfg4", "gh", 1_000, 10, 100_000, 100_000),
    "cfg5": Workload("cfg5", "leja", 1_000, 100, 10_000, 100_000_000),
}
