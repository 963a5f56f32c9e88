"""Shared helpers of the test-suite: golden fixtures, generators, error measures."""
from __future__ import annotations

import hashlib
from pathlib import Path

import numpy as np

from smolyax_b200 import nodes, workloads

GOLDEN = Path(__file__).resolve().parent / "golden"
ALL_CASES = sorted(p.stem for p in GOLDEN.glob("*.npz"))
LAYOUT_CASES = [c for c in ALL_CASES if c.startswith(("small", "medium"))]
BIG_CASES = [c for c in ALL_CASES if c not in LAYOUT_CASES]
# zero-padded value tensors of the reference layout above 1 GiB: set_f assembles the compact layout (smx_create_compact), the
# per-summand kernels and the oracle on the full layout are not run (the oracle runs on sampled output columns instead)
COMPACT_CASES = {"cfg3_dout520"}
REFERENCE_LAYOUT_CASES = [c for c in ALL_CASES if c not in COMPACT_CASES]
# Non-nested rules of high degree whose hierarchical (Newton) form is refused by the plan compiler's conditioning check
# (smx_plan.cpp newton_form_error: Gauss-Hermite degree 24 loses ~1e-10 of a cardinal function in fp64): the handle then runs
# the per-summand barycentric kernels, i.e. the reference's own arithmetic.
ILL_CONDITIONED = {"medium_05"}


def load(name):
    return np.load(GOLDEN / f"{name}.npz")


def generator_of(g):
    rule, d_in = str(g["rule"]), len(g["k"])
    if "n_target" in g.files:
        return nodes.Leja(dim=d_in) if rule == "leja" else nodes.GaussHermite(dim=d_in)
    if rule == "leja":
        return nodes.Leja(domains=g["domains"])
    return nodes.GaussHermite(g["mean"], g["scaling"])


def target_of(g, gen):
    """The target the golden generator used: the benchmark family, composed with scale_back for custom domains."""
    fam = workloads.TargetFamily(len(g["k"]), int(g["d_out"]))
    if "n_target" in g.files:
        return fam
    return lambda x: fam(gen.scale_back(np.asarray(x, dtype=float)))


def golden_layout(g):
    return {k[len("layout_"):]: g[k] for k in g.files if k.startswith("layout_") and k != "layout_sha256"}


def layout_digest(layout):
    """SHA-256 over the six reference arrays + offset (same recipe as oracle/make_golden.py)."""
    h = hashlib.sha256()
    for key in sorted(k for k in layout if not k.startswith("quad")):
        a = np.ascontiguousarray(layout[key])
        h.update(key.encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def long_double(g, key):
    return g[f"{key}_ld_hi"].astype(np.longdouble) + g[f"{key}_ld_lo"].astype(np.longdouble)


def scaled_error(y, y_ref, cond_abs):
    """|y - y_ref| relative to the summand magnitude sum_nu |zeta_nu I_nu f|."""
    return float(np.max(np.abs(np.asarray(y) - y_ref) / cond_abs))


def interpolator_inputs(g):
    gen = generator_of(g)
    return dict(node_gen=gen, k=g["k"], t=float(g["t"]), d_out=int(g["d_out"])), target_of(g, gen)


def dense_indexset(k, t):
    """Brute-force Lambda(k,t) by recursion over dimensions (independent of smolyax_b200.indices)."""
    if len(k) == 0:
        return [()]
    out, j = [], 0
    while j * k[0] < t:
        out += [(j,) + rest for rest in dense_indexset(k[1:], t - j * k[0])]
        j += 1
    return out
