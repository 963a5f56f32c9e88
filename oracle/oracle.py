"""ctypes front-end of the CPU oracle (oracle/smx_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module; nothing under
smolyax_b200/ does.  The functions take the reference's per-group layout (a dict with keys ``offset``, ``F_n``,
``nodes_n``, ``weights_n``, ``dims_n``, ``degs_n``, ``zetas_n`` and optionally ``quad_n``), i.e. exactly what
reference interpolation.py:230-235 keeps on the device.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> Path:
    so, src = HERE / "libsmx_oracle.so", HERE / "smx_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        # the image's default $CC (/opt/gcc) ships without libgomp.spec; /usr/bin/gcc has OpenMP
        last = None
        for cc, omp in (("/usr/bin/gcc", "-fopenmp"), ("gcc", "-fopenmp"), ("gcc", "")):
            flags = f"-O2 -fPIC -shared {omp} -ffp-contract=off -fno-fast-math -std=c11"
            last = subprocess.run(["make", "-C", str(HERE), "-B", "libsmx_oracle.so", f"CC={cc}", f"CFLAGS={flags}"],
                                  capture_output=True, text=True)
            if last.returncode == 0:
                break
        else:
            raise RuntimeError("could not build the CPU oracle:\n" + last.stdout + last.stderr)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build()))
        _LIB.smo_max_threads.restype = ctypes.c_int
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_ip)


def group_sizes(layout):
    return sorted(int(key.split("_")[1]) for key in layout if key.startswith("zetas_"))


def _tau(layout, n):
    return np.asarray(layout[f"F_{n}"].shape[2:], dtype=np.int64) - 1


def max_threads() -> int:
    return int(lib().smo_max_threads())


def use_all_cores() -> int:
    """Let OpenMP use every core this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    import os

    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().smo_set_num_threads(ctypes.c_int(n))
    return max_threads()


def compute_weights(nodes):
    nodes_, pn = _d(nodes)
    w = np.empty_like(nodes_)
    lib().smo_compute_weights(pn, ctypes.c_int64(len(nodes_)), w.ctypes.data_as(_dp))
    return w


def evaluate(layout, x):
    """Restates reference interpolation.py:264-304 (``__call__``)."""
    x_, px = _d(np.atleast_2d(x))
    N, d_in = x_.shape
    d_out = int(np.asarray(layout["F_%d" % group_sizes(layout)[0]]).shape[1]) if group_sizes(layout) else len(layout["offset"])
    y = np.ascontiguousarray(np.broadcast_to(np.asarray(layout["offset"], dtype=np.float64), (N, d_out))).copy()
    for n in group_sizes(layout):
        keep = [_d(layout[f"F_{n}"]), _d(layout[f"nodes_{n}"]), _d(layout[f"weights_{n}"]), _i(layout[f"dims_{n}"]),
                _i(layout[f"degs_{n}"]), _i(layout[f"zetas_{n}"]), _i(_tau(layout, n))]
        lib().smo_group_eval(px, ctypes.c_int64(N), ctypes.c_int64(d_in), *(k[1] for k in keep[:6]),
                             ctypes.c_int64(len(keep[5][0])), ctypes.c_int(n), keep[6][1], ctypes.c_int64(d_out),
                             y.ctypes.data_as(_dp))
    return y


def gradient(layout, x):
    """Restates reference interpolation.py:306-345 (``gradient``)."""
    x_, px = _d(np.atleast_2d(x))
    N, d_in = x_.shape
    d_out = int(np.asarray(layout["F_%d" % group_sizes(layout)[0]]).shape[1]) if group_sizes(layout) else len(layout["offset"])
    J = np.zeros((N, d_out, d_in))
    for n in group_sizes(layout):
        keep = [_d(layout[f"F_{n}"]), _d(layout[f"nodes_{n}"]), _d(layout[f"weights_{n}"]), _i(layout[f"dims_{n}"]),
                _i(layout[f"degs_{n}"]), _i(layout[f"zetas_{n}"]), _i(_tau(layout, n))]
        lib().smo_group_gradient(px, ctypes.c_int64(N), ctypes.c_int64(d_in), ctypes.c_int64(d_in),
                                 *(k[1] for k in keep[:6]), ctypes.c_int64(len(keep[5][0])), ctypes.c_int(n),
                                 keep[6][1], ctypes.c_int64(d_out), J.ctypes.data_as(_dp))
    return J


def integral(layout):
    """Restates reference interpolation.py:347-390 (``integral``); needs the ``quad_n`` tables."""
    ns = group_sizes(layout)
    d_out = int(np.asarray(layout["F_%d" % ns[0]]).shape[1]) if ns else len(layout["offset"])
    Q = np.ascontiguousarray(np.broadcast_to(np.asarray(layout["offset"], dtype=np.float64), (d_out,))).copy()
    for n in ns:
        keep = [_d(layout[f"F_{n}"]), _d(layout[f"quad_{n}"]), _i(layout[f"zetas_{n}"]), _i(_tau(layout, n))]
        lib().smo_group_integral(keep[0][1], keep[1][1], keep[2][1], ctypes.c_int64(len(keep[2][0])), ctypes.c_int(n),
                                 keep[3][1], ctypes.c_int64(d_out), Q.ctypes.data_as(_dp))
    return Q


def _referee(layout, x, fn, shape_tail):
    x_, px = _d(np.atleast_2d(x))
    N, d_in = x_.shape
    ns = group_sizes(layout)
    d_out = int(np.asarray(layout["F_%d" % ns[0]]).shape[1]) if ns else len(layout["offset"])
    assert np.dtype(np.longdouble).itemsize == 16 and np.finfo(np.longdouble).nmant == 63, "needs x87 80-bit long double"
    out = np.zeros((N, d_out) + ((d_in,) if shape_tail else ()), dtype=np.longdouble)
    if not shape_tail:
        out += np.asarray(layout["offset"], dtype=np.float64).astype(np.longdouble)
    for n in ns:
        keep = [_d(layout[f"F_{n}"]), _d(layout[f"nodes_{n}"]), _d(layout[f"weights_{n}"]), _i(layout[f"dims_{n}"]),
                _i(layout[f"degs_{n}"]), _i(layout[f"zetas_{n}"]), _i(_tau(layout, n))]
        args = [px, ctypes.c_int64(N), ctypes.c_int64(d_in)] + ([ctypes.c_int64(d_in)] if shape_tail else []) + [k[1] for k in keep[:6]] + \
               [ctypes.c_int64(len(keep[5][0])), ctypes.c_int(n), keep[6][1], ctypes.c_int64(d_out), ctypes.c_void_p(out.ctypes.data)]
        getattr(lib(), fn)(*args)
    return out


def evaluate_referee(layout, x):
    """``evaluate`` carried out in 80-bit long double on the same fp64 tables (accuracy referee)."""
    return _referee(layout, x, "smo_group_eval_ld", False)


def gradient_referee(layout, x):
    """``gradient`` carried out in 80-bit long double (NaN at node hits, like the fp64 version)."""
    return _referee(layout, x, "smo_group_gradient_ld", True)
