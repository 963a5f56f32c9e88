"""jax.ffi registration of the seam twins (reference interpolation.py:243-248, 293-302, 334-343, 389).

For a caller that keeps the reference's Python and stays inside ``jax.jit``: ``group_eval`` / ``group_gradient`` take the
reference's own device arrays of one group (or of a slice of its summands) and return the SUM over the summands - what
``jnp.sum(res, axis=0)`` yields after the reference's ``jit(vmap(..))`` call - computed by ``smx_group_eval`` /
``smx_group_gradient`` of ``libsmolyax_b200.so`` on XLA's stream.

The handlers live in ``csrc/smx_xla_ffi.cc`` and are compiled on first use into ``libsmolyax_xla_ffi.so`` IF jaxlib's
headers are there (``jax.ffi.include_dir()``).  Without jax this module still imports; every entry point then raises
``RuntimeError`` naming what is missing (there is no fallback path).
"""
import ctypes
import subprocess
from pathlib import Path

from . import _build

SOURCE = _build.CSRC / "smx_xla_ffi.cc"
LIBRARY = _build.PKG / "libsmolyax_xla_ffi.so"
TARGETS = {"smx_group_eval": "SmxGroupEval", "smx_group_gradient": "SmxGroupGradient", "smx_group_integral": "SmxGroupIntegral"}
_registered = False


def include_dir():
    """jaxlib's FFI header directory, or None when jax (or its ffi module) is not installed."""
    try:
        import jax
        ffi = getattr(jax, "ffi", None) or __import__("jax.extend.ffi", fromlist=["ffi"])
        return Path(ffi.include_dir())
    except Exception:
        return None


def compile_command(include: Path, output: Path = LIBRARY) -> list:
    """g++ line for the handler library: C++17, jaxlib's headers, the C-ABI header, linked against libsmolyax_b200.so."""
    cuda_inc = Path(_build.nvcc_path()).parent.parent / "include" if _build.nvcc_path() else Path("/usr/local/cuda/include")
    return ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", f"-I{include}", f"-I{_build.ROOT / 'include'}", f"-I{cuda_inc}",
            str(SOURCE), "-o", str(output), f"-L{_build.PKG}", "-lsmolyax_b200", f"-Wl,-rpath,{_build.PKG}"]


def build(force: bool = False) -> Path:
    inc = include_dir()
    if inc is None:
        raise RuntimeError("jax.ffi is not available: install jax/jaxlib to use the XLA FFI handlers (csrc/smx_xla_ffi.cc)")
    _build.build_cuda()
    if force or not LIBRARY.exists() or LIBRARY.stat().st_mtime < SOURCE.stat().st_mtime:
        subprocess.run(compile_command(inc), check=True)
    return LIBRARY


def register() -> None:
    """Compile (if needed) and register the three FFI targets for platform "CUDA"."""
    global _registered
    if _registered:
        return
    lib = ctypes.CDLL(str(build()))
    import jax
    ffi = getattr(jax, "ffi", None) or __import__("jax.extend.ffi", fromlist=["ffi"])
    for target, symbol in TARGETS.items():
        ffi.register_ffi_target(target, ffi.pycapsule(getattr(lib, symbol)), platform="CUDA")
    _registered = True


def _call(target, out_shape, *operands):
    register()
    import jax
    import jax.numpy as jnp
    ffi = getattr(jax, "ffi", None) or __import__("jax.extend.ffi", fromlist=["ffi"])
    return ffi.ffi_call(target, jax.ShapeDtypeStruct(out_shape, jnp.float64))(*operands)


def group_eval(x, F, nodes, weights, dims, degs, zetas):
    """sum_s zeta_s I_s(x), shape (N, d_out): drop-in for ``jnp.sum(compiled_tensor_product_evaluation(..), axis=0)``."""
    return _call("smx_group_eval", (x.shape[0], F.shape[1]), x, F, nodes, weights, dims, degs, zetas)


def group_gradient(x, F, nodes, weights, dims, degs, zetas):
    """sum_s zeta_s grad I_s(x), shape (N, d_out, d_in): drop-in for the gradient twin (interpolation.py:334-343)."""
    return _call("smx_group_gradient", (x.shape[0], F.shape[1], x.shape[1]), x, F, nodes, weights, dims, degs, zetas)


def group_integral(F, quad, zetas):
    """sum_s zeta_s <F_s, quad_s>, shape (d_out,): the einsum of interpolation.py:389."""
    return _call("smx_group_integral", (F.shape[1],), F, quad, zetas)
