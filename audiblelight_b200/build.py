"""Builds the C-ABI CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc", "alrender.cu")
_DEPS = [_SRC, os.path.join(_HERE, "csrc", "alr_kernels.cuh"), os.path.join(_HERE, "csrc", "alr_fft.cuh"),
         os.path.join(_HERE, "csrc", "alr_fused.cuh"), os.path.join(_HERE, "csrc", "alr_sweep.cuh"),
         os.path.join(os.path.dirname(_HERE), "include", "alrender.h")]
_LIB = os.path.join(_HERE, "libalrender.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def library_path() -> str:
    """The library to load; ALR_LIBRARY overrides it (used to A/B kernel variants built with extra -D flags)."""
    return os.environ.get("ALR_LIBRARY", _LIB)


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


def build_library(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> audiblelight_b200/libalrender.so
    (`defines` / `out` build a kernel variant next to it, e.g. defines=["ALR_CMAC_BINS=1"], out="variant.so")."""
    lib = _LIB if out is None else os.path.join(_HERE, out)
    if out is None and not force and not is_stale():
        return lib
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in defines]
    cmd += ["-o", lib + ".tmp", _SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(lib + ".tmp", lib)
    if verbose:
        print(res.stderr)
    return lib
