"""f4 — a consumer of the mix: the STFT / visibility front-end of the reference's acoustic imaging
(`audiblelight/imaging.py`: `extract_visibilities` :455-492, `form_visibility` :697-719, the per-band loop of
`get_visibility_matrix` :775-853), computed on the GPU from `scene.audio[mic]`.

Only this front-end is rebuilt: the APGD solver, the spherical tesselation and the contour extraction behind it are a
different subsystem with their own dependencies (pyunlocbox, astropy, scikit-image) and stay in the reference.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import ALR_MEM_DEVICE, ALR_MEM_HOST
from .renderer import Renderer

# config.py:79-84
AIMG_FMIN, AIMG_FMAX, AIMG_NBANDS, AIMG_BANDWIDTH, AIMG_TSTI = 1500, 4500, 9, 50.0, 10e-3


def _renderer(renderer: Optional[Renderer]) -> Renderer:
    if renderer is not None:
        return renderer
    from .synthesize import get_renderer
    return get_renderer()


def visibility_bands(mix, rate: float, freqs: Sequence[float], bw: float = AIMG_BANDWIDTH, t_sti: float = AIMG_TSTI,
                     n_sti_per_block: int = 10, alpha: float = 1.0, renderer: Optional[Renderer] = None, stream: int = 0):
    """Visibility matrices of every band in one GPU call.

    mix: (C, T) float32 — numpy array (host) or CUDA torch tensor; returns complex128 (n_bands, n_blocks, C, C) as a
    numpy array (host input) or a torch tensor (device input). Band b, block k equals
    `form_visibility(mix.T, rate, freqs[b], bw, t_sti, n_sti_per_block * t_sti)[k]` of the reference."""
    rnd = _renderer(renderer)
    is_dev = hasattr(mix, "data_ptr")
    if tuple(mix.shape).__len__() != 2:
        raise ValueError("mix must have shape (channels, samples)")
    if "float32" not in str(mix.dtype):
        raise TypeError(f"mix must be float32, got {mix.dtype}")
    if not (mix.is_contiguous() if is_dev else mix.flags.c_contiguous):
        raise ValueError("mix must be C-contiguous")
    n_ch, n_samples = int(mix.shape[0]), int(mix.shape[1])
    n_stft = int(rate * t_sti)
    if n_stft == 0:
        raise ValueError("Not enough samples per time frame.")  # imaging.py:469
    n_blocks = (n_samples // n_stft) // int(n_sti_per_block)
    fc = np.ascontiguousarray(freqs, dtype=np.float64)
    nb = int(fc.shape[0])
    if is_dev:
        import torch
        out = torch.zeros((max(n_blocks, 0), nb, n_ch, n_ch, 2), dtype=torch.float64, device=mix.device)
        optr, mptr = out.data_ptr(), mix.data_ptr()
    else:
        out = np.zeros((max(n_blocks, 0), nb, n_ch, n_ch, 2), dtype=np.float64)
        optr, mptr = out.ctypes.data, mix.ctypes.data
    if n_blocks > 0:
        _lib.check(rnd._lib.alr_visibilities(rnd._h, C.c_void_p(mptr), n_ch, n_samples, float(rate), float(t_sti),
                                             C.c_void_p(fc.ctypes.data), nb, float(bw), int(n_sti_per_block), float(alpha),
                                             C.c_void_p(optr), ALR_MEM_DEVICE if is_dev else ALR_MEM_HOST,
                                             C.c_void_p(stream)))
    if is_dev:
        import torch
        return torch.view_as_complex(out).permute(1, 0, 2, 3)
    return out.view(np.complex128)[..., 0].transpose(1, 0, 2, 3)


def form_visibility(data: np.ndarray, rate: float, fc: float, bw: float, t_sti: float, t_stationarity: float,
                    renderer: Optional[Renderer] = None) -> np.ndarray:
    """imaging.py:697-719 — data (samples, channels) -> (n_blocks, channels, channels) complex128."""
    mix = np.ascontiguousarray(np.asarray(data).T, dtype=np.float32)
    return np.ascontiguousarray(visibility_bands(mix, rate, [fc], bw, t_sti, int(t_stationarity / t_sti), 1.0, renderer)[0])


def band_frequencies(fmin: float = AIMG_FMIN, fmax: float = AIMG_FMAX, nbands: int = AIMG_NBANDS) -> np.ndarray:
    """The "linear" scale of get_visibility_matrix (imaging.py:818-819)."""
    if fmin >= fmax:
        raise ValueError(f"Minimum frequency must be smaller than maximum frequency (current minimum: {fmin}, maximum: {fmax}).")
    return np.linspace(fmin, fmax, nbands)
