"""In-tree build of the two native libraries.

* ``libsmolyax_host.so``  — g++ only (multi-index combinatorics); loadable without a GPU.
* ``libsmolyax_b200.so`` — nvcc, sm_100a only: the CUDA kernels and the C-ABI of include/smolyax_b200.h.

Both are written next to this file so that they travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import time
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
ROOT = PKG.parent
HOST_LIB = PKG / "libsmolyax_host.so"
CUDA_LIB = PKG / "libsmolyax_b200.so"

HOST_SOURCES = ["smx_host.cpp", "smx_plan.cpp"]
CUDA_SOURCES = ["smx_api.cu", "smx_seam.cu", "smx_fast.cu", "smx_fast_kernel.cu", "smx_fast_pipe.cu", "smx_grad_kernel.cu", "smx_fast_multi.cu", "smx_dense_kernel.cu", "smx_plan.cpp"]
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


def source_hash(files, extra=()) -> str:
    """SHA-256 over the given source files (names and contents) and the build flags: what a binary is stamped with."""
    h = hashlib.sha256()
    for f in sorted(Path(f) for f in files):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    for e in extra:
        h.update(str(e).encode())
    return h.hexdigest()[:16]


def _stamp_path(lib: Path) -> Path:
    return lib.with_suffix(".so.stamp")


def read_stamp(lib: Path) -> dict:
    """Stamp written next to a built library: hash of the sources it was built from, compiler, flags, time."""
    try:
        return json.loads(_stamp_path(lib).read_text())
    except Exception:
        return {}


def build_host(force: bool = False) -> Path:
    srcs = [CSRC / s for s in HOST_SOURCES]
    flags = ["-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-pthread"]
    digest = source_hash(srcs + sorted(CSRC.glob("*.h")), flags)
    if not force and HOST_LIB.exists() and read_stamp(HOST_LIB).get("source_hash") == digest:
        return HOST_LIB
    cxx = os.environ.get("CXX", "g++")
    tmp = HOST_LIB.with_suffix(f".so.tmp{os.getpid()}")
    _run([cxx, *flags, *srcs, "-o", tmp])
    os.replace(tmp, HOST_LIB)
    _stamp_path(HOST_LIB).write_text(json.dumps({"source_hash": digest, "compiler": cxx, "flags": flags, "built": time.strftime("%Y-%m-%dT%H:%M:%S")}))
    return HOST_LIB


def nvcc_path():
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return cand if Path(cand).exists() else None


LAST_BUILD = {}  # library name -> "rebuilt" | "reused (source hash matches)", for the build report of __graft_entry__.build()


def nvcc_version(nvcc) -> str:
    try:
        out = subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout
        return next((ln.strip() for ln in out.splitlines() if "release" in ln), "nvcc ?")
    except Exception:
        return "nvcc ?"


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    """Compile every translation unit for sm_100a (objects in csrc/.build, in parallel) and link the shared library.

    The library carries a stamp (smx_build_info(), and a `.stamp` file beside it): hash of its sources and flags, compiler
    version.  Without `force` the existing binary is reused only when that hash matches the sources in the tree."""
    from concurrent.futures import ThreadPoolExecutor

    srcs = [CSRC / s for s in CUDA_SOURCES]
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "smolyax_b200.h"]
    tuning = bool(os.environ.get("SMX_TUNING"))
    flags = [*NVCC_ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math", "--fmad=true"]
    if tuning:  # A/B timing builds: the kernels' tuning knobs read the environment (smx_plan.h tune_int)
        flags += ["-DSMX_TUNING", *os.environ.get("SMX_EXTRA_NVCC_FLAGS", "").split()]
    digest = source_hash(srcs + headers, flags)
    stamp = read_stamp(CUDA_LIB)
    if not force and CUDA_LIB.exists() and stamp.get("source_hash") == digest:
        LAST_BUILD.setdefault(CUDA_LIB.name, "reused (source hash matches)")  # (a rebuild earlier in this process stays on record)
        return CUDA_LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libsmolyax_b200.so (sm_100a); there is no CPU fallback")
    objdir = CSRC / ".build"
    objdir.mkdir(exist_ok=True)
    common = [nvcc, *flags, "-I", ROOT / "include", "-I", CSRC]
    if verbose:
        common += ["-Xptxas", "-v"]
    version = nvcc_version(nvcc)
    text = f"{version}; sources {digest}; built {time.strftime('%Y-%m-%dT%H:%M:%S')}"

    def compile_one(src: Path):
        obj = objdir / (src.name + (".tuning.o" if tuning else ".o"))
        extra = [f'-DSMX_BUILD_STAMP="{text}"'] if src.name == "smx_api.cu" else []
        if not force and not extra and _newer(obj, [src] + headers):
            return obj, ""
        return obj, _run([*common, *extra, "-c", src, "-o", obj])

    with ThreadPoolExecutor(max_workers=len(srcs)) as pool:
        results = list(pool.map(compile_one, srcs))
    tmp = CUDA_LIB.with_suffix(f".so.tmp{os.getpid()}")
    _run([nvcc, *NVCC_ARCH, "-shared", *[obj for obj, _ in results], "-o", tmp])
    os.replace(tmp, CUDA_LIB)
    _stamp_path(CUDA_LIB).write_text(json.dumps({"source_hash": digest, "compiler": version, "flags": flags, "tuning": tuning,
                                                 "built": time.strftime("%Y-%m-%dT%H:%M:%S")}))
    LAST_BUILD[CUDA_LIB.name] = "rebuilt"
    if verbose:
        print("".join(out for _, out in results))
    return CUDA_LIB


def build_all(force: bool = False, verbose: bool = False):
    return build_host(force), build_cuda(force, verbose)


if __name__ == "__main__":
    import sys

    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", HOST_LIB.name, CUDA_LIB.name)
