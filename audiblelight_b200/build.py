"""Builds the C-ABI CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc", "alrender.cu")
_DEPS = [_SRC, os.path.join(_HERE, "csrc", "alr_kernels.cuh"), os.path.join(_HERE, "csrc", "alr_fft.cuh"),
         os.path.join(os.path.dirname(_HERE), "include", "alrender.h")]
_LIB = os.path.join(_HERE, "libalrender.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def library_path() -> str:
    return _LIB


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libalrender.so")
    return exe


def is_stale() -> bool:
    if not os.path.exists(_LIB):
        return True
    t = os.path.getmtime(_LIB)
    return any(os.path.getmtime(d) > t for d in _DEPS if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> audiblelight_b200/libalrender.so"""
    if not force and not is_stale():
        return _LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", _LIB + ".tmp", _SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(_LIB + ".tmp", _LIB)
    if verbose:
        print(res.stderr)
    return _LIB
