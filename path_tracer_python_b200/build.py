"""Builds libb200pt.so (strict numerics, the parity build) and libb200pt_fast.so (production numerics)
in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m path_tracer_python_b200.build [--force] [--verbose] [--only=strict|fast|debug|wide]

Flags that matter: ``--fmad=false`` (PT-F32 contract: no FFMA contraction; the fast variant is the same
sources with ``-DPT_FAST=1 --fmad=true``), ``-lineinfo`` (ncu source page),
``-gencode arch=compute_100a,code=sm_100a`` (B200 only; no other architectures, no PTX JIT).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libb200pt.so")
FAST_LIB_PATH = os.path.join(HERE, "libb200pt_fast.so")
FAST_DEFINES = ("PT_FAST=1",)
WIDE_LIB_PATH = os.path.join(HERE, "libb200pt_wide.so")
WIDE_DEFINES = ("PT_WIDE=1",)
DEBUG_LIB_PATH = os.path.join(HERE, "libb200pt_dbg.so")
DEBUG_DEFINES = ("PT_DEBUG_BOUNDS=1",)
SOURCES = ["b200pt.cu", "pt_host.cpp"]
HEADERS = ["pt_device.cuh", "pt_math.cuh", os.path.join("..", "..", "include", "b200pt.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--fmad=false", "-ftz=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math", "-Xcompiler", "-ffp-contract=off",
    "-shared", "-cudart", "shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build(lib_path: str = LIB_PATH) -> bool:
    if not os.path.exists(lib_path):
        return True
    t = os.path.getmtime(lib_path)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """`defines` / `out` build an experimental variant next to the product library (A/B
    measurements through the PT_LIB_PATH override of _lib.py); the default build takes neither."""
    if out is None and not defines:
        if not force and not needs_build():
            return LIB_PATH
    out = out or LIB_PATH
    flags = list(NVCC_FLAGS)
    if any(d.split("=")[0] == "PT_FAST" and not d.endswith("=0") for d in defines):
        flags[flags.index("--fmad=false")] = "--fmad=true"
    cmd = [_nvcc()] + flags + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else [])
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    cmd += ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return out


def build_fast(force: bool = False, verbose: bool = False) -> str:
    """The production-numerics variant (PT_FAST: FMA contraction and approximate reciprocals around the strict
    intersection tests) next to the strict library."""
    if not force and not needs_build(FAST_LIB_PATH):
        return FAST_LIB_PATH
    return build_native(force=True, verbose=verbose, defines=FAST_DEFINES, out=FAST_LIB_PATH)


def build_wide(force: bool = False, verbose: bool = False) -> str:
    """The 4-wide BVH variant (PT_WIDE: the packer collapses the binary tree into 128-byte 4-child nodes).  Same
    results; measured 5 % slower than the binary layout on config 2, so not the default."""
    if not force and not needs_build(WIDE_LIB_PATH):
        return WIDE_LIB_PATH
    return build_native(force=True, verbose=verbose, defines=WIDE_DEFINES, out=WIDE_LIB_PATH)


def build_debug(force: bool = False, verbose: bool = False) -> str:
    """Strict numerics with every node / primitive / queue / bin / slab / stack index checked in the kernels
    (PT_DEBUG_BOUNDS); tests/test_gpu_debug.py runs the golden suite through it."""
    if not force and not needs_build(DEBUG_LIB_PATH):
        return DEBUG_LIB_PATH
    return build_native(force=True, verbose=verbose, defines=DEBUG_DEFINES, out=DEBUG_LIB_PATH)


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    only = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--only=")]
    force, verbose = "--force" in sys.argv, "--verbose" in sys.argv
    if defs or outs:
        print(build_native(force=force, verbose=verbose, defines=defs, out=outs[0] if outs else None))
    else:
        if not only or only[0] == "strict":
            print(build_native(force=force, verbose=verbose))
        if not only or only[0] == "fast":
            print(build_fast(force=force, verbose=verbose))
        if not only or only[0] == "debug":
            print(build_debug(force=force, verbose=verbose))
        if not only or only[0] == "wide":
            print(build_wide(force=force, verbose=verbose))
