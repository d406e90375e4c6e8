"""CPU oracle for the PCM_16 packing of a scene mix (TEST INFRASTRUCTURE, scope rows f2/f4).

`Scene.generate` stores `sf.write(path, mix.T, sr)` (core.py:1840-1847): soundfile / libsndfile 1.2 pick PCM_16 for a
.wav path and convert float -> short in src/pcm.c (f2les_array) as `psf_lrintf(x * 0x7FFF)` assigned to a short, with
clipping disabled by default (values beyond +-1 wrap). soundfile is neither vendored nor installable offline ->
PARITY UNPINNED against the real library; this file restates the published conversion.
"""
import numpy as np


def pcm16_from_float(mix_ct: np.ndarray) -> np.ndarray:
    """(C, T) float32 -> (T, C) int16."""
    x = np.asarray(mix_ct, dtype=np.float32).T * np.float32(32767.0)   # float arithmetic, as in the C loop
    return np.rint(x).astype(np.int64).astype(np.int16)                  # lrintf (half to even), then 16-bit truncation
