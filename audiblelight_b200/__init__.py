"""audiblelight_b200 — B200-native (sm_100a) renderer for AudibleLight's synthesis hot path.

Host layer in Python, compute in hand-written CUDA kernels behind a C-ABI shared library
(`include/alrender.h`, built to `audiblelight_b200/libalrender.so`). There is no CPU fallback: importing
`audiblelight_b200.synthesize` works anywhere, but every render call needs the library and a GPU.
"""
__version__ = "0.1.0"

from .build import build_library, library_path  # noqa: F401
