"""Multi-index combinatorics at growing cardinality: threshold search, Lambda(k,t) and the non-zero
Smolyak coefficients for d = 10^4 (the reference's tests/test_indices_runtime.py shape), this package (C++ host library)
beside the unmodified reference (numba + Python), which is only present in the build container.  No GPU needed.

    python benchmarks/host_indices.py [n_target ...] > profiles/rNN_host_indices.txt
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REFERENCE = Path("/root/reference/src")


def timed(fn):
    t0 = time.perf_counter()
    out = fn()
    return out, time.perf_counter() - t0


def main():
    from smolyax_b200 import indices, workloads

    targets = [int(v) for v in sys.argv[1:]] or [10_000, 100_000, 1_000_000]
    rind = None
    if REFERENCE.exists():
        sys.path.insert(0, str(ROOT / "oracle" / "jax_stub"))
        sys.path.insert(1, str(REFERENCE))
        from smolyax import indices as rind
        rind.non_zero_indices_and_zetas(workloads.anisotropy(20), 3.0)  # numba compilation outside the timings
        rind.find_approximate_threshold(workloads.anisotropy(20), 30, True)
    d = 10_000
    k = workloads.anisotropy(d)
    print(f"# d = {d}, k_j = log((2+j)/log 2); seconds, one run each; reference = unmodified /root/reference: "
          f"{'yes' if rind else 'absent on this machine'}")
    print(f"{'n_target':>9s} {'|Lambda|':>9s} {'summands':>9s} | {'threshold':>10s} {'indexset':>9s} {'nonzero+zeta':>12s} | "
          f"{'ref threshold':>13s} {'ref indexset':>12s} {'ref nonzero+zeta':>16s} | same")
    for n in targets:
        t, a = timed(lambda: indices.find_approximate_threshold(k, n, True))
        lam, b = timed(lambda: indices.indexset(k, t))
        (n2nus, n2z), c = timed(lambda: indices.non_zero_indices_and_zetas(k, t))
        n_sum = sum(len(v) for v in n2nus.values())
        _, b_arr = timed(lambda: indices.indexset_arrays(k, t))  # the CSR arrays set_f consumes: no Python tuples
        _, c_arr = timed(lambda: indices.nonzero_arrays(k, t))
        ref = f"{'-':>13s} {'-':>12s} {'-':>16s} | -"
        if rind:
            t_r, ar = timed(lambda: rind.find_approximate_threshold(k, n, True))
            lam_r, br = timed(lambda: rind.indexset(k, t))
            (r2nus, r2z), cr = timed(lambda: rind.non_zero_indices_and_zetas(k, t))
            same = t_r == t and lam_r == lam and dict(r2nus) == dict(n2nus) and all(list(r2z[key]) == list(n2z[key]) for key in r2z)
            ref = f"{ar:13.3f} {br:12.3f} {cr:16.3f} | {same}"
        print(f"{n:9d} {len(lam):9d} {n_sum:9d} | {a:10.3f} {b:9.3f} {c:12.3f} | {ref} | arrays only: indexset {b_arr:.3f}, "
              f"nonzero+zeta {c_arr:.3f}", flush=True)


if __name__ == "__main__":
    main()
