"""In-tree build of the two native libraries.

* ``libsmolyax_host.so``  — g++ only (multi-index combinatorics); loadable without a GPU.
* ``libsmolyax_b200.so`` — nvcc, sm_100a only: the CUDA kernels and the C-ABI of include/smolyax_b200.h.

Both are written next to this file so that they travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
ROOT = PKG.parent
HOST_LIB = PKG / "libsmolyax_host.so"
CUDA_LIB = PKG / "libsmolyax_b200.so"

HOST_SOURCES = ["smx_host.cpp", "smx_plan.cpp"]
CUDA_SOURCES = ["smx_api.cu", "smx_seam.cu", "smx_fast.cu", "smx_fast_kernel.cu", "smx_fast_multi.cu", "smx_dense_kernel.cu", "smx_plan.cpp"]
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def _run(cmd):
    proc = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("build failed:\n$ " + " ".join(map(str, cmd)) + "\n" + proc.stdout + proc.stderr)
    return proc.stdout + proc.stderr


def build_host(force: bool = False) -> Path:
    srcs = [CSRC / s for s in HOST_SOURCES]
    if not force and _newer(HOST_LIB, srcs + sorted(CSRC.glob("*.h"))):
        return HOST_LIB
    cxx = os.environ.get("CXX", "g++")
    tmp = HOST_LIB.with_suffix(f".so.tmp{os.getpid()}")
    _run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-pthread", *srcs, "-o", tmp])
    os.replace(tmp, HOST_LIB)
    return HOST_LIB


def nvcc_path():
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return cand if Path(cand).exists() else None


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    """Compile every translation unit for sm_100a (objects in csrc/.build, in parallel) and link the shared library."""
    from concurrent.futures import ThreadPoolExecutor

    srcs = [CSRC / s for s in CUDA_SOURCES]
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "smolyax_b200.h"]
    if not force and _newer(CUDA_LIB, srcs + headers):
        return CUDA_LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libsmolyax_b200.so (sm_100a); there is no CPU fallback")
    objdir = CSRC / ".build"
    objdir.mkdir(exist_ok=True)
    common = [nvcc, *NVCC_ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math",
              "--fmad=true", "-I", ROOT / "include", "-I", CSRC]
    if verbose:
        common += ["-Xptxas", "-v"]

    def compile_one(src: Path):
        obj = objdir / (src.name + ".o")
        if not force and _newer(obj, [src] + headers):
            return obj, ""
        return obj, _run([*common, "-c", src, "-o", obj])

    with ThreadPoolExecutor(max_workers=len(srcs)) as pool:
        results = list(pool.map(compile_one, srcs))
    tmp = CUDA_LIB.with_suffix(f".so.tmp{os.getpid()}")
    _run([nvcc, *NVCC_ARCH, "-shared", *[obj for obj, _ in results], "-o", tmp])
    os.replace(tmp, CUDA_LIB)
    if verbose:
        print("".join(out for _, out in results))
    return CUDA_LIB


def build_all(force: bool = False, verbose: bool = False):
    return build_host(force), build_cuda(force, verbose)


if __name__ == "__main__":
    import sys

    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", HOST_LIB.name, CUDA_LIB.name)
