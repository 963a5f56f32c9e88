"""ctypes binding of libsmolyax_b200.so (the C-ABI of include/smolyax_b200.h).

The library is the only compute path of this package.  If it cannot be loaded (and cannot be built because nvcc is
absent) importing this module raises: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_uint32, c_void_p

import numpy as np

from . import _build

SMX_KEEP_GROUPS = 1
SMX_NO_FAST_PATH = 2
SMX_GRAD_FINITE_AT_NODES = 4
SMX_DENSE_PATH = 8
SMX_NO_DENSE_PATH = 16

SMX_ERR_UNSUPPORTED = 4

_STATUS = {1: "invalid argument", 2: "CUDA error", 3: "out of device memory", 4: "unsupported shape", 5: "no sm_100 device"}


class SmolyaxCudaError(RuntimeError):
    pass


class GroupDesc(ctypes.Structure):
    _fields_ = [
        ("n", c_int32),
        ("nn", c_int64),
        ("tau", c_void_p),
        ("F", c_void_p),
        ("nodes", c_void_p),
        ("weights", c_void_p),
        ("dims", c_void_p),
        ("degs", c_void_p),
        ("zetas", c_void_p),
        ("quad", c_void_p),
    ]


class InterpDesc(ctypes.Structure):
    _fields_ = [
        ("d_in", c_int64),
        ("d_out", c_int64),
        ("offset", c_void_p),
        ("n_groups", c_int32),
        ("groups", POINTER(GroupDesc)),
        ("flags", c_uint32),
    ]


COMPACT_FIELDS = (("n_active", np.int32), ("slot_off", np.int64), ("dims", np.int64), ("degs", np.int64),
                  ("node_off", np.int64), ("node_pool", np.float64), ("quad_pool", np.float64), ("zetas", np.int64),
                  ("val_off", np.int64), ("val_index", np.int64), ("values", np.float64))


class CompactDesc(ctypes.Structure):
    _fields_ = [("n_summands", c_int64)] + [(name, c_void_p) for name, _ in COMPACT_FIELDS] + [("n_values", c_int64)]


class Info(ctypes.Structure):
    _fields_ = [(name, c_int64) for name in (
        "d_in", "d_out", "n_summands", "w_raw", "w_pad", "n_terms", "n_entries", "n_rows", "n_chunks", "padded_fma",
        "device_bytes")] + [(name, c_int32) for name in ("has_fast_path", "has_groups", "nested", "has_dense_path")] + [
        ("dense_terms", c_int64), ("grad_jobs", c_int64), ("grad_items", c_int64)]


EXPORTS = {
    "smx_create": (c_int, [POINTER(InterpDesc), c_int, POINTER(c_void_p)]),
    "smx_create_compact": (c_int, [c_int64, c_int64, c_void_p, POINTER(CompactDesc), c_uint32, c_int, POINTER(c_void_p)]),
    "smx_destroy": (c_int, [c_void_p]),
    "smx_eval": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "smx_gradient": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "smx_integral": (c_int, [c_void_p, c_void_p, c_void_p]),
    "smx_eval_host": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64]),
    "smx_prepare": (c_int, [c_void_p, c_int64]),
    "smx_group_eval": (c_int, [c_void_p, c_int64, c_int64, c_int64, POINTER(GroupDesc), c_int64, c_void_p, c_int, c_void_p]),
    "smx_group_gradient": (c_int, [c_void_p, c_int64, c_int64, c_int64, POINTER(GroupDesc), c_int64, c_void_p, c_int, c_void_p]),
    "smx_group_integral": (c_int, [POINTER(GroupDesc), c_int64, c_void_p, c_int, c_void_p]),
    "smx_compute_weights": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "smx_basis": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "smx_get_info": (c_int, [c_void_p, POINTER(Info)]),
    "smx_launch_count": (c_int64, []),
    "smx_last_kernel": (c_char_p, []),
    "smx_last_error": (c_char_p, []),
    "smx_version": (c_int, []),
    "smx_arch": (c_char_p, []),
    "smx_build_info": (c_char_p, []),
}


ABI_VERSION = 200  # smx_version() of the library these ctypes struct layouts were written for


def _load():
    # build_cuda() returns at once when the binary's stamp matches the sources in the tree and rebuilds it otherwise; a box
    # without nvcc (never the case in this image) keeps whatever binary travelled with the tree.  No CPU fallback either way.
    if _build.nvcc_path() is not None or not _build.CUDA_LIB.exists():
        path = _build.build_cuda()
    else:
        path = _build.CUDA_LIB
    lib = ctypes.CDLL(str(path))
    for name, (restype, argtypes) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export what the header declares
        fn.restype, fn.argtypes = restype, argtypes
    if lib.smx_version() != ABI_VERSION:
        raise ImportError(f"{path} reports ABI {lib.smx_version()}, this binding is written for {ABI_VERSION}: stale binary")
    return lib


lib = _load()


def check(status: int, where: str = ""):
    if status != 0:
        msg = lib.smx_last_error().decode(errors="replace")
        text = f"{where}: {_STATUS.get(status, status)}: {msg}"
        if status == 1:
            raise AssertionError(text)  # the reference signals bad inputs with assert
        raise SmolyaxCudaError(text)


def group_sizes(layout):
    return sorted(int(key.split("_")[1]) for key in layout if key.startswith("zetas_"))


def pack_groups(layout, ptr_of):
    """Build the smx_group_desc array for a reference-layout dict.  ``ptr_of(array, dtype)`` returns
    ``(keepalive, address)`` — host arrays for smx_create, device tensors for the seam twins."""
    ns = group_sizes(layout)
    arr = (GroupDesc * max(len(ns), 1))()
    keep = []
    for i, n in enumerate(ns):
        F = layout[f"F_{n}"]
        tau = np.ascontiguousarray(np.asarray(F.shape[2:], dtype=np.int64) - 1)
        keep.append(tau)
        arr[i].n, arr[i].nn, arr[i].tau = n, F.shape[0], tau.ctypes.data
        for field, key, dt in (("F", "F", np.float64), ("nodes", "nodes", np.float64), ("weights", "weights", np.float64),
                               ("dims", "dims", np.int64), ("degs", "degs", np.int64), ("zetas", "zetas", np.int64),
                               ("quad", "quad", np.float64)):
            a = layout.get(f"{key}_{n}")
            if a is None:
                setattr(arr[i], field, None)
                continue
            k, addr = ptr_of(a, dt)
            keep.append(k)
            setattr(arr[i], field, addr)
    return arr, len(ns), keep


def host_ptr(a, dt):
    a = np.ascontiguousarray(a, dtype=dt)
    return a, a.ctypes.data


def create(layout, d_in: int, d_out: int, flags: int, device: int = -1):
    arr, n_groups, keep = pack_groups(layout, host_ptr)
    off = np.ascontiguousarray(np.broadcast_to(np.asarray(layout["offset"], dtype=np.float64), (d_out,)))
    desc = InterpDesc(d_in, d_out, off.ctypes.data, n_groups, arr, flags)
    handle = c_void_p()
    check(lib.smx_create(ctypes.byref(desc), device, ctypes.byref(handle)), "smx_create")
    del keep
    return handle


def pack_compact(layout):
    """smx_compact_desc for a compact layout dict (interpolation._assemble_compact); returns (desc, keepalive)."""
    desc = CompactDesc()
    keep = []
    desc.n_summands = len(layout["zetas"])
    for name, dt in COMPACT_FIELDS:
        a = layout.get(name)
        if a is None:
            setattr(desc, name, None)
            continue
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        setattr(desc, name, a.ctypes.data)
    desc.n_values = layout["values"].shape[0]
    return desc, keep


def create_compact(layout, d_in: int, d_out: int, flags: int, device: int = -1):
    desc, keep = pack_compact(layout)
    off = np.ascontiguousarray(np.broadcast_to(np.asarray(layout["offset"], dtype=np.float64), (d_out,)))
    handle = c_void_p()
    check(lib.smx_create_compact(d_in, d_out, off.ctypes.data, ctypes.byref(desc), flags, device, ctypes.byref(handle)),
          "smx_create_compact")
    del keep
    return handle


def info(handle) -> dict:
    out = Info()
    check(lib.smx_get_info(handle, ctypes.byref(out)), "smx_get_info")
    return {name: int(getattr(out, name)) for name, _ in Info._fields_}


def last_kernel() -> str:
    """Name and template arguments of the last kernel this thread launched (smx_last_kernel)."""
    return lib.smx_last_kernel().decode()

