"""XLA FFI handlers (csrc/smx_xla_ffi.cc, smolyax_b200/xla_ffi.py): this image has no jaxlib, so
 * the translation unit is compiled (syntax and types only) against a stub of the FFI declarations it uses - the stub's
   handler macro fails to compile if an implementation does not match the argument list of its binding;
 * the module imports without jax and refuses loudly;
 * with jaxlib present (not here) the library is built against the real headers and its exports are checked.
"""
import shutil
import subprocess
from pathlib import Path

import pytest

from smolyax_b200 import _build, xla_ffi

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_translation_unit_is_well_formed_against_the_stub_headers():
    cuda_inc = "/usr/local/cuda/include"
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", f"-I{ROOT / 'tests' / 'xla_ffi_stub'}", f"-I{ROOT / 'include'}",
           f"-I{cuda_inc}", str(xla_ffi.SOURCE)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    src = xla_ffi.SOURCE.read_text()
    for symbol in xla_ffi.TARGETS.values():  # every registered target has its handler symbol
        assert f"XLA_FFI_DEFINE_HANDLER_SYMBOL({symbol}," in src
    for entry in ("smx_group_eval(", "smx_group_gradient(", "smx_group_integral("):  # .. and goes through the C-ABI
        assert entry in src
    assert "#include <torch" not in src and "oracle" not in src


def test_a_binding_that_does_not_match_its_implementation_is_rejected(tmp_path):
    """The stub is strict enough to be worth compiling against: dropping one argument of a binding must not compile."""
    src = xla_ffi.SOURCE.read_text()
    marker = ".Arg<ffi::Buffer<ffi::S64>>()   // zetas    (nn)\n"
    assert marker in src
    broken = tmp_path / "broken.cc"
    broken.write_text(src.replace(marker, "", 1))
    cmd = ["g++", "-std=c++17", "-fsyntax-only", f"-I{ROOT / 'tests' / 'xla_ffi_stub'}", f"-I{ROOT / 'include'}", "-I/usr/local/cuda/include",
           str(broken)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode != 0 and "does not match its binding" in res.stderr


def test_without_jax_the_module_imports_and_refuses_loudly():
    if xla_ffi.include_dir() is not None:
        pytest.skip("jaxlib is installed here")
    with pytest.raises(RuntimeError, match="jax.ffi is not available"):
        xla_ffi.build()
    with pytest.raises(RuntimeError, match="jax.ffi is not available"):
        xla_ffi.register()
    cmd = xla_ffi.compile_command(Path("/nonexistent/include"))
    assert "-lsmolyax_b200" in cmd and str(xla_ffi.SOURCE) in cmd


def test_with_jaxlib_the_library_builds_and_exports_the_handlers():
    if xla_ffi.include_dir() is None:
        pytest.skip("no jaxlib in this image (the GPU box has none either): built and tested where jax is installed")
    import ctypes

    lib = ctypes.CDLL(str(xla_ffi.build(force=True)))
    for symbol in xla_ffi.TARGETS.values():
        assert hasattr(lib, symbol)
