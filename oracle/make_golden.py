#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/src/smolyax) on the NumPy
`jax` stand-in of oracle/jax_stub.  TEST INFRASTRUCTURE — runs only in the build container (the reference does
not exist on the GPU box); the fixtures it writes are committed.

    python oracle/make_golden.py            # all cases
    python oracle/make_golden.py small cfg1 # a subset

Every fixture holds: the constructor inputs, the evaluation points, the reference's outputs in fp64
(`y_ref`, `J_ref`, `Q_ref`), the same reference code run in 80-bit long double (`*_ld_hi/lo`, the accuracy
referee of SURVEY.md §7.3), the per-point summand magnitude `cond_abs = sum_nu |zeta_nu I_nu f|` used to scale
errors, and either the reference's per-group device layout itself (small cases) or a SHA-256 of it.
"""
from __future__ import annotations

import hashlib
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.path.insert(0, str(ROOT / "oracle" / "jax_stub"))
sys.path.insert(1, "/root/reference/src")
sys.path.insert(2, str(ROOT))

from smolyax import indices as ref_indices  # noqa: E402
from smolyax import nodes as ref_nodes  # noqa: E402
from smolyax.interpolation import SmolyakBarycentricInterpolator as RefInterpolator  # noqa: E402

from smolyax_b200 import workloads  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
P = "_SmolyakBarycentricInterpolator__"
LD = np.longdouble


def split_ld(a):
    hi = np.asarray(a, dtype=np.float64)
    lo = np.asarray(np.asarray(a, dtype=LD) - hi.astype(LD), dtype=np.float64)
    return hi, lo


def layout_of(ip):
    """The six per-group arrays + offset of reference interpolation.py:230-235."""
    # the reference keeps a scalar 0 when zeta_0 == 0; store the broadcast (d_out,) form in every case
    out = {"offset": np.array(np.broadcast_to(np.asarray(getattr(ip, P + "offset"), dtype=np.float64), (ip.d_out,)))}
    for n in getattr(ip, P + "n_2_F"):
        out[f"F_{n}"] = np.asarray(getattr(ip, P + "n_2_F")[n], dtype=np.float64)
        out[f"nodes_{n}"] = np.asarray(getattr(ip, P + "n_2_nodes")[n], dtype=np.float64)
        out[f"weights_{n}"] = np.asarray(getattr(ip, P + "n_2_weights")[n], dtype=np.float64)
        out[f"dims_{n}"] = np.asarray(getattr(ip, P + "n_2_sorted_dims")[n], dtype=np.int64)
        out[f"degs_{n}"] = np.asarray(getattr(ip, P + "n_2_sorted_degs")[n], dtype=np.int64)
        out[f"zetas_{n}"] = np.asarray(getattr(ip, P + "n_2_zetas")[n], dtype=np.int64)
    return out


def quad_tables(ip, gen):
    """Zero-padded quadrature-weight tables (nn, n, taumax+1) as reference interpolation.py:361-379 builds them."""
    out = {}
    for n, degs in getattr(ip, P + "n_2_sorted_degs").items():
        degs = np.asarray(degs)
        dims = np.asarray(getattr(ip, P + "n_2_sorted_dims")[n])
        q = np.zeros(degs.shape + (int(degs.max()) + 1,))
        for s in range(degs.shape[0]):
            for j in range(n):
                w = gen[int(dims[s, j])].get_quadrature_weights(int(degs[s, j]))
                q[s, j, : len(w)] = w
        out[f"quad_{n}"] = q
    return out


def layout_digest(layout):
    h = hashlib.sha256()
    for key in sorted(layout):
        a = np.ascontiguousarray(layout[key])
        h.update(key.encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def cond_abs(ip, x):
    """sum over summands of |zeta_nu I_nu f (x)|, per point and output (plus |offset|)."""
    fn = getattr(ip, P + "compiled_tensor_product_evaluation")
    tot = np.abs(np.broadcast_to(np.asarray(getattr(ip, P + "offset"), dtype=float), (x.shape[0], ip.d_out))).copy()
    for n in getattr(ip, P + "n_2_F"):
        res = fn(x, *(getattr(ip, P + a)[n] for a in ("n_2_F", "n_2_nodes", "n_2_weights", "n_2_sorted_dims", "n_2_sorted_degs", "n_2_zetas")))
        tot += np.abs(np.asarray(res)).sum(axis=0)
    return tot


class LongDoubleTables:
    """Context manager: swap the reference's private tables for long-double copies."""

    NAMES = ("n_2_F", "n_2_nodes", "n_2_weights")

    def __init__(self, ip):
        self.ip = ip

    def __enter__(self):
        self.saved = {a: getattr(self.ip, P + a) for a in self.NAMES + ("offset",)}
        for a in self.NAMES:
            setattr(self.ip, P + a, {n: np.asarray(v).astype(LD).view(type(v)) for n, v in self.saved[a].items()})
        setattr(self.ip, P + "offset", np.asarray(self.saved["offset"]).astype(LD))
        return self.ip

    def __exit__(self, *exc):
        for a, v in self.saved.items():
            setattr(self.ip, P + a, v)


def integral_ld(ip, gen):
    """Reference integral (interpolation.py:383-389) in long double (np.einsum on object-free LD arrays)."""
    Q = np.asarray(getattr(ip, P + "offset"), dtype=LD) * np.ones(ip.d_out, dtype=LD)
    q = quad_tables(ip, gen)
    for n, F in getattr(ip, P + "n_2_F").items():
        F = np.asarray(F).astype(LD)
        z = np.asarray(getattr(ip, P + "n_2_zetas")[n]).astype(LD)
        qs = q[f"quad_{n}"].astype(LD)
        acc = F  # (nn, d_out, t1+1, ..)
        for j in range(n - 1, -1, -1):
            m = acc.shape[-1]
            acc = (acc * qs[:, j, :m].reshape((qs.shape[0],) + (1,) * (acc.ndim - 2) + (m,))).sum(axis=-1)
        Q = Q + (acc * z[:, None]).sum(axis=0)
    return Q


def build_case(name, gen, gen_args, k, t, d_out, f, x, n_grad, store_layout):
    t0 = time.time()
    ip = RefInterpolator(node_gen=gen, k=k, t=t, d_out=d_out, f=f)
    layout = layout_of(ip)
    out = dict(gen_args)
    out.update(k=np.asarray(k, dtype=np.float64), t=np.float64(t), d_out=np.int64(d_out), x=x,
               n_f_evals=np.int64(ip.n_f_evals), layout_sha256=np.array(layout_digest(layout)))
    out["y_ref"] = np.asarray(ip(x), dtype=np.float64)
    out["cond_abs"] = cond_abs(ip, x)
    xg = x[:n_grad]
    with np.errstate(all="ignore"):
        out["J_ref"] = np.asarray(ip.gradient(xg), dtype=np.float64)
    out["Q_ref"] = np.asarray(ip.integral(), dtype=np.float64)
    with LongDoubleTables(ip):
        out["y_ld_hi"], out["y_ld_lo"] = split_ld(np.asarray(ip(x.astype(LD))))
        with np.errstate(all="ignore"):
            out["J_ld_hi"], out["J_ld_lo"] = split_ld(np.asarray(ip.gradient(xg.astype(LD))))
    out["Q_ld_hi"], out["Q_ld_lo"] = split_ld(integral_ld(ip, gen))
    if store_layout:
        out.update({f"layout_{key}": v for key, v in layout.items()})
        out.update({f"layout_{key}": v for key, v in quad_tables(ip, gen).items()})
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / f"{name}.npz", **out)
    n_sum = sum(len(v) for kk, v in layout.items() if kk.startswith("zetas_"))
    print(f"{name}: d_in={len(k)} d_out={d_out} summands={n_sum} n_f_evals={ip.n_f_evals} N={len(x)} "
          f"({time.time() - t0:.1f}s, {(GOLDEN / (name + '.npz')).stat().st_size / 1024:.0f} KiB)")


def smooth_target(gen, d_in, d_out):
    """Benchmark family composed with the map to the reference domain, so random domains stay well-conditioned."""
    fam = workloads.TargetFamily(d_in, d_out)

    def f(x):
        return fam(gen.scale_back(np.asarray(x, dtype=float)))

    return f


def with_node_hits(gen, x, k, t, rng):
    """Overwrite coordinates of the last rows of x with exact interpolation nodes (one-hot / NaN branches)."""
    x = x.copy()
    d = x.shape[1]
    maxdeg = [0] * d
    for nu in ref_indices.indexset(k, t):
        for dim, deg in nu:
            maxdeg[dim] = max(maxdeg[dim], deg)
    hit_rows = min(3, len(x))
    for r in range(hit_rows):
        dim = int(rng.integers(d))
        deg = max(maxdeg[dim], 1)
        j = int(rng.integers(deg + 1))
        x[len(x) - 1 - r, dim] = gen[dim](deg)[j]
    x[len(x) - 1, :] = [g(max(m, 1))[min(1, max(m, 1))] for g, m in zip(gen, maxdeg)]  # a full grid node
    return x


def make_small():
    rng = np.random.default_rng(20240607)
    for i in range(10):
        d = int(rng.integers(1, 5))
        if i % 2 == 0:
            dom = np.sort(rng.random((d, 2)), axis=1)
            dom[:, 1] += 0.05
            gen, args = ref_nodes.Leja(domains=dom), {"rule": np.array("leja"), "domains": dom}
        else:
            mean, scaling = rng.standard_normal(d), 0.2 + rng.random(d)
            gen, args = ref_nodes.GaussHermite(mean, scaling), {"rule": np.array("gh"), "mean": mean, "scaling": scaling}
        k = np.sort(rng.uniform(1, 10, d))
        k = k / k[0]
        t = float(rng.uniform(1.5, 8))
        d_out = int(rng.integers(1, 4))
        np.random.seed(1000 + i)
        x = gen.get_random(8)
        x = with_node_hits(gen, x, k, t, rng)
        build_case(f"small_{i:02d}", gen, args, k, t, d_out, smooth_target(gen, d, d_out), x, n_grad=8, store_layout=True)
    # medium: more dimensions / summands, and high 1-D degrees (generic large-m paths)
    extra = [("leja", 5, 6.5, 2), ("gh", 5, 6.0, 3), ("leja", 6, 5.5, 1), ("gh", 4, 7.5, 2),
             ("leja", 1, 30.5, 2), ("gh", 1, 24.5, 1), ("leja", 2, 15.0, 1), ("gh", 2, 12.0, 2)]
    for i, (rule, d, t, d_out) in enumerate(extra):
        if rule == "leja":
            dom = np.sort(rng.uniform(-2, 2, (d, 2)), axis=1)
            dom[:, 1] += 0.1
            gen, args = ref_nodes.Leja(domains=dom), {"rule": np.array("leja"), "domains": dom}
        else:
            mean, scaling = rng.standard_normal(d), 0.3 + rng.random(d)
            gen, args = ref_nodes.GaussHermite(mean, scaling), {"rule": np.array("gh"), "mean": mean, "scaling": scaling}
        k = 1.0 + 0.35 * np.arange(d) + 0.1 * np.sort(rng.random(d))
        k = k / k[0]
        np.random.seed(2000 + i)
        x = with_node_hits(gen, gen.get_random(8), k, t, rng)
        build_case(f"medium_{i:02d}", gen, args, k, t, d_out, smooth_target(gen, d, d_out), x, n_grad=8, store_layout=True)


def make_cfg(name, rule, d_in, d_out, n_target, n_pts, n_grad):
    wl = workloads.Workload(name, rule, d_in, d_out, n_target, n_pts)
    gen = ref_nodes.Leja(dim=d_in) if rule == "leja" else ref_nodes.GaussHermite(dim=d_in)
    k = wl.k()
    t = ref_indices.find_approximate_threshold(k, n_target, gen.is_nested)
    x = wl.points(n_pts, seed=0)
    args = {"rule": np.array(rule), "n_target": np.int64(n_target)}
    build_case(name, gen, args, k, t, d_out, wl.target(), x, n_grad=n_grad, store_layout=False)


CASES = {
    "small": make_small,
    "cfg1": lambda: make_cfg("cfg1", "leja", 10, 1, 1_000, 64, 8),
    "cfg2": lambda: make_cfg("cfg2", "leja", 1_000, 1, 10_000, 64, 8),
    "cfg3": lambda: make_cfg("cfg3_dout16", "leja", 100, 16, 10_000, 16, 2),
    # cfg3 with enough outputs for the widest shape of the GEMM-regime kernel (16 warps x 4 output blocks = 512 columns per CTA)
    "cfg3w": lambda: make_cfg("cfg3_dout520", "leja", 100, 520, 10_000, 72, 1),
    "cfg4": lambda: make_cfg("cfg4", "gh", 1_000, 10, 100_000, 64, 8),
    # cfg5 = the cfg2 index set with 100 outputs (narrow shape of the GEMM-regime kernel); 257 points: a ragged last tile
    "cfg5": lambda: make_cfg("cfg5", "leja", 1_000, 100, 10_000, 257, 1),
    "gh100": lambda: make_cfg("gh_d100", "gh", 100, 2, 3_000, 16, 4),
}

if __name__ == "__main__":
    which = sys.argv[1:] or list(CASES)
    for c in which:
        CASES[c]()
