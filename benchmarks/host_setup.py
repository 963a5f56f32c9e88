"""Host side of the path: threshold search, index set, non-zero Smolyak coefficients and the
assembly of the tables `set_f` hands to the device — timed for this package and, where /root/reference exists (the build
container only), for the unmodified reference on the NumPy `jax` stand-in of oracle/jax_stub (its jitted calls are then
plain NumPy: the reference's numbers here are a LOWER bound of what it costs under JAX, where each `compute_weights`
call is a device dispatch and the first call of each shape compiles).  No GPU needed.

    python benchmarks/host_setup.py [cfg1 cfg2 cfg4 ...] > profiles/rNN_host_setup.txt
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REFERENCE = Path("/root/reference/src")


def best(fn, reps=3):
    out, t_best = None, float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        t_best = min(t_best, time.perf_counter() - t0)
    return out, t_best


def ours(wl, reps):
    from smolyax_b200 import indices
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    k, nested = wl.k(), wl.rule == "leja"
    rows = {}
    t, rows["find_approximate_threshold"] = best(lambda: indices.find_approximate_threshold(k, wl.n_target, nested), reps)
    lam, rows["indexset"] = best(lambda: indices.indexset(k, t), reps)
    nz, rows["non_zero_indices_and_zetas"] = best(lambda: indices.non_zero_indices_and_zetas(k, t), reps)
    f = wl.target()
    for batched in (False, True):
        def assemble():
            ip = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=k, t=t, d_out=wl.d_out, batched_f=batched,
                                                layout="reference")
            return ip._assemble(f, {})[0]
        out, rows["set_f tables" + (" (batched_f)" if batched else "")] = best(assemble, max(1, reps - 1))
        if not batched:  # (a batched f sums in another order: its values differ from per-point calls in the last bit)
            layout = out

        def assemble_compact():
            ip = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=k, t=t, d_out=wl.d_out, batched_f=batched,
                                                layout="compact")
            return ip._assemble_compact(f, {})[0]
        _, rows["set_f tables, compact layout" + (" (batched_f)" if batched else "")] = best(assemble_compact, max(1, reps - 1))
    return t, len(lam), nz, layout, rows


def reference(wl, t, reps):
    sys.path.insert(0, str(ROOT / "oracle" / "jax_stub"))
    sys.path.insert(1, str(REFERENCE))
    from smolyax import indices as rindices, nodes as rnodes
    from smolyax.interpolation import SmolyakBarycentricInterpolator as RefInterpolator

    k, nested = wl.k(), wl.rule == "leja"
    rindices.find_approximate_threshold(k, 50, nested)  # numba compilation outside the timings
    rows = {}
    t_ref, rows["find_approximate_threshold"] = best(lambda: rindices.find_approximate_threshold(k, wl.n_target, nested), reps)
    assert t_ref == t, (t_ref, t)
    lam, rows["indexset"] = best(lambda: rindices.indexset(k, t), reps)
    nz, rows["non_zero_indices_and_zetas"] = best(lambda: rindices.non_zero_indices_and_zetas(k, t), reps)
    gen = rnodes.Leja(dim=wl.d_in) if nested else rnodes.GaussHermite(dim=wl.d_in)
    f = wl.target()
    ip, rows["set_f tables"] = best(lambda: RefInterpolator(node_gen=gen, k=k, t=t, d_out=wl.d_out, f=f), 1)
    return len(lam), nz, ip, rows


def main():
    from smolyax_b200 import workloads

    names = sys.argv[1:] or ["cfg1", "cfg2", "cfg4"]
    have_ref = REFERENCE.exists()
    print(f"# host set-up times in seconds (best of 3; set_f best of 2, reference set_f once); reference = unmodified "
          f"/root/reference on the NumPy jax stand-in: {'yes' if have_ref else 'absent on this machine'}")
    for name in names:
        wl = workloads.CONFIGS[name]
        if wl.d_out > 100:  # cfg3: the padded reference tensors do not fit; time the tables at 16 outputs
            wl = workloads.Workload(name + "_dout16", wl.rule, wl.d_in, 16, wl.n_target, wl.n_points)
        t, n_lam, nz, layout, rows = ours(wl, 3)
        print(f"\n{wl.name}: {wl.rule} d_in={wl.d_in} d_out={wl.d_out} n_target={wl.n_target}  t={t!r}  |Lambda|={n_lam}")
        ref_rows = {}
        if have_ref:
            n_lam_ref, nz_ref, ip_ref, ref_rows = reference(wl, t, 3)
            assert n_lam_ref == n_lam
            same = True
            for n, F in ip_ref._SmolyakBarycentricInterpolator__n_2_F.items():
                same &= np.array_equal(np.asarray(F), layout[f"F_{n}"])
            print(f"  value tensors identical to the reference's: {same}")
        print(f"  {'step':38s} {'this package':>14s} {'reference':>12s} {'ratio':>8s}")
        for key, v in rows.items():
            r = ref_rows.get(key.replace(" (batched_f)", "").replace(", compact layout", ""))
            print(f"  {key:38s} {v:14.4f} " + (f"{r:12.4f} {r / v:8.1f}" if r is not None else f"{'-':>12s} {'-':>8s}"))


if __name__ == "__main__":
    main()
