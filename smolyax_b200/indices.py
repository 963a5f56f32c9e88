r"""
Anisotropic multi-index sets :math:`\Lambda_{k,t} = \{\nu \in \mathbb N_0^d : \sum_j k_j \nu_j < t\}`,
Smolyak coefficients and node-set cardinalities.

Same public functions and results as the reference's ``smolyax.indices``
(/root/reference/src/smolyax/indices.py:20,72,121,214,262,292); the searches run in the C++ host
library (csrc/smx_host.cpp) and reproduce the reference's visiting order and floating-point path, so the
outputs are *identical*, not merely equivalent.  ``k`` must be increasing (reference README.md:12-15).
"""
from __future__ import annotations

import ctypes
from collections import defaultdict
from typing import Sequence

import numpy as np

from . import _build

_c_dp = ctypes.POINTER(ctypes.c_double)
_c_i64p = ctypes.POINTER(ctypes.c_int64)
_c_i32p = ctypes.POINTER(ctypes.c_int32)


def _load():
    lib = ctypes.CDLL(str(_build.build_host()))
    lib.smxh_indexset.restype = ctypes.c_void_p
    lib.smxh_indexset.argtypes = [_c_dp, ctypes.c_int64, ctypes.c_double]
    lib.smxh_nonzero_indices_and_zetas.restype = ctypes.c_void_p
    lib.smxh_nonzero_indices_and_zetas.argtypes = [_c_dp, ctypes.c_int64, ctypes.c_double]
    lib.smxh_indexset_cardinality.restype = ctypes.c_int64
    lib.smxh_indexset_cardinality.argtypes = [_c_dp, ctypes.c_int64, ctypes.c_double]
    lib.smxh_nodeset_cardinality_non_nested.restype = ctypes.c_int64
    lib.smxh_nodeset_cardinality_non_nested.argtypes = [_c_dp, ctypes.c_int64, ctypes.c_double]
    lib.smxh_smolyak_coefficient.restype = ctypes.c_int64
    lib.smxh_smolyak_coefficient.argtypes = [_c_dp, ctypes.c_int64, ctypes.c_double, ctypes.c_int64]
    lib.smxh_result_count.restype = ctypes.c_int64
    lib.smxh_result_count.argtypes = [ctypes.c_void_p]
    lib.smxh_result_nnz.restype = ctypes.c_int64
    lib.smxh_result_nnz.argtypes = [ctypes.c_void_p]
    lib.smxh_result_copy.restype = None
    lib.smxh_result_copy.argtypes = [ctypes.c_void_p, _c_i64p, _c_i32p, _c_i32p, _c_i64p]
    lib.smxh_result_free.restype = None
    lib.smxh_result_free.argtypes = [ctypes.c_void_p]
    return lib


_lib = _load()


def _kvec(k) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(k, dtype=np.float64).ravel())


def _fetch(handle, with_zetas: bool):
    """Copy a C++ result (CSR of (dim,deg) pairs) into NumPy arrays and release it."""
    try:
        n, nnz = _lib.smxh_result_count(handle), _lib.smxh_result_nnz(handle)
        offsets = np.empty(n + 1, dtype=np.int64)
        dims = np.empty(max(nnz, 1), dtype=np.int32)
        degs = np.empty(max(nnz, 1), dtype=np.int32)
        zetas = np.empty(max(n, 1), dtype=np.int64) if with_zetas else None
        _lib.smxh_result_copy(
            handle,
            offsets.ctypes.data_as(_c_i64p),
            dims.ctypes.data_as(_c_i32p),
            degs.ctypes.data_as(_c_i32p),
            zetas.ctypes.data_as(_c_i64p) if with_zetas else None,
        )
    finally:
        _lib.smxh_result_free(handle)
    return offsets, dims[:nnz], degs[:nnz], (zetas[:n] if with_zetas else None)


def indexset_arrays(k: Sequence[float], t: float):
    """Array form of :func:`indexset`: ``(offsets, dims, degs)`` — multi-index ``i`` is the slice
    ``offsets[i]:offsets[i+1]`` of the (dimension, degree) pairs, dimensions ascending."""
    kv = _kvec(k)
    h = _lib.smxh_indexset(kv.ctypes.data_as(_c_dp), len(kv), float(t))
    return _fetch(h, False)[:3]


def nonzero_arrays(k: Sequence[float], t: float):
    """Array form of :func:`non_zero_indices_and_zetas`: ``(offsets, dims, degs, zetas)`` in the order the
    reference's walk emits the multi-indices (not yet binned by number of active dimensions)."""
    kv = _kvec(k)
    h = _lib.smxh_nonzero_indices_and_zetas(kv.ctypes.data_as(_c_dp), len(kv), float(t))
    return _fetch(h, True)


def _tuples(offsets, dims, degs):
    pairs = list(zip(dims.tolist(), degs.tolist()))
    off = offsets.tolist()
    return [tuple(pairs[off[i] : off[i + 1]]) for i in range(len(off) - 1)]


def indexset(k: Sequence[float], t: float) -> list:
    r"""The multi-index set :math:`\Lambda_{k,t}` as a list of sparse tuples ``((j, nu_j), ...)`` over the
    dimensions with ``nu_j > 0`` (reference indices.py:20-69, same order)."""
    return _tuples(*indexset_arrays(k, t))


def indexset_cardinality(k: Sequence[float], t: float) -> int:
    r"""``len(indexset(k, t))`` without building the set (reference indices.py:72-118)."""
    kv = _kvec(k)
    return int(_lib.smxh_indexset_cardinality(kv.ctypes.data_as(_c_dp), len(kv), float(t)))


def smolyak_coefficient(k: Sequence[float], d: int, rem_t: float, parity: int) -> int:
    r"""Smolyak coefficient :math:`\zeta_{\Lambda,\nu} = \sum_{e \in \{0,1\}^d,\ \nu+e \in \Lambda} (-1)^{|e|}`
    from the remaining budget ``rem_t = t - sum_j nu_j k_j`` (reference indices.py:121-168)."""
    kv = _kvec(k)
    return int(_lib.smxh_smolyak_coefficient(kv.ctypes.data_as(_c_dp), int(d), float(rem_t), int(parity)))


def non_zero_indices_and_zetas(k: Sequence[float], t: float):
    """Multi-indices with non-zero Smolyak coefficient and those coefficients, binned by the number ``n`` of
    active dimensions: ``(n2nus, n2zetas)`` (reference indices.py:214-259; same order within each bin)."""
    offsets, dims, degs, zetas = nonzero_arrays(k, t)
    n2nus, n2zetas = defaultdict(list), defaultdict(list)
    for nu, z in zip(_tuples(offsets, dims, degs), zetas.tolist()):
        n2nus[len(nu)].append(nu)
        n2zetas[len(nu)].append(z)
    return n2nus, n2zetas


def nodeset_cardinality(k: Sequence[float], t: float, nested: bool = False) -> int:
    """Number of distinct interpolation nodes of the Smolyak operator (reference indices.py:262-289)."""
    if nested:
        return indexset_cardinality(k, t)
    kv = _kvec(k)
    return int(_lib.smxh_nodeset_cardinality_non_nested(kv.ctypes.data_as(_c_dp), len(kv), float(t)))


def find_approximate_threshold(
    k: Sequence[float], m: int, nested: bool, max_iter: int = 32, accuracy: float = 0.001
) -> float:
    """Threshold ``t`` for which the node set has approximately ``m`` nodes: geometric bracketing by a factor
    1.2, then bisection (reference indices.py:292-355; identical arithmetic, hence identical ``t``)."""
    assert m > 0
    if m == 1:
        return 1
    kv = _kvec(k)

    def card(t):
        return nodeset_cardinality(kv, t, nested)

    lo, hi = 1.0, 2.0
    while card(lo) > m:
        lo, hi = lo / 1.2, lo
    while card(hi) < m:
        lo, hi = hi, hi * 1.2

    t_cand = lo + (hi - lo) / 2.0
    m_cand = card(t_cand)
    for _ in range(max_iter):
        if m_cand > m:
            hi = t_cand
        else:
            lo = t_cand
        t_cand = lo + (hi - lo) / 2.0
        m_cand = card(t_cand)
        if np.abs(m_cand - m) / m < accuracy:
            break
    return t_cand
