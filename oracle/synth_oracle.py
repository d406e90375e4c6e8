"""CPU oracle for AudibleLight's synthesis hot path (TEST INFRASTRUCTURE — not product code).

A float64 numpy restatement of the arithmetic in the reference's `audiblelight/synthesize.py:40-608`
(+ `utils.py:667-706`).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this module; the product (`audiblelight_b200/`) never does.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the unmodified reference from
/root/reference (stub modules for the uninstalled, arithmetic-free imports) and stores its outputs for
seeded inputs under `tests/golden/*.npz`; `tests/test_oracle.py` checks every function here against them
and against the reference's own known-answer tests (`tests/test_synthesize.py:42-57,307-377`).

The FFTs themselves live in scipy (`scipy.signal.fftconvolve`, `scipy.fft.rfft/irfft`; reference pins
scipy 1.13-1.16 in `pyproject.toml`, un-vendored; this image has scipy 1.18).  `fftconvolve`'s published
recipe is restated in `linear_convolve` on top of `scipy.fft` (pocketfft), including its per-operand precision.

The functions are array-in / array-out (no Event / Scene objects) so the same oracle serves the C-ABI
tests, the Python drop-in tests and the benchmark baseline.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
from scipy import fft as sp_fft

FFT_SIZE, WIN_SIZE, HOP_SIZE = 512, 256, 128  # config.py:9-11


# --------------------------------------------------------------------------------------------------
# small helpers (utils.py)
# --------------------------------------------------------------------------------------------------
def tiny(x) -> float:
    """utils.py:691-706 — smallest normal of x's float dtype; float32's for non-float input."""
    dt = np.asarray(x).dtype
    if not (np.issubdtype(dt, np.floating) or np.issubdtype(dt, np.complexfloating)):
        dt = np.dtype(np.float32)
    return float(np.finfo(dt).tiny)


def pad_or_truncate(audio: np.ndarray, n: int) -> np.ndarray:
    """utils.py:667-688 — zero-pad or cut a (channels, samples) array to n samples."""
    have = audio.shape[1]
    if have == n:
        return audio
    if have > n:
        return audio[:, :n]
    out = np.zeros((audio.shape[0], n), dtype=audio.dtype)
    out[:, :have] = audio
    return out


def apply_snr(x: np.ndarray, snr: float) -> np.ndarray:
    """synthesize.py:40-49 — scale so the absolute peak over ALL channels equals snr (peak floor 1e-15)."""
    peak = max(1e-15, float(np.max(np.abs(x)))) if x.size else 1e-15
    return x * snr / peak


def db_to_multiplier(db: float, x) -> float:
    """synthesize.py:52-68 — 10^(db/20) / (x + tiny(x))."""
    return 10.0 ** (db / 20.0) / (x + tiny(x))


def normalize_irs(irs: np.ndarray) -> np.ndarray:
    """synthesize.py:404-428 — divide by the mean (over axis -2) of the L2 norms taken over axis -1."""
    energy = np.sqrt(np.sum(np.abs(irs) ** 2, axis=-1, keepdims=True))
    energy = energy + tiny(energy)
    return irs / energy.mean(axis=-2, keepdims=True)


def ir_scales(irs_cnl: np.ndarray) -> np.ndarray:
    """Per-IR scalar a_l that `normalize_irs` applies when called as at synthesize.py:560.

    irs_cnl is (C, N, Lh); the reference normalises the (N, C, Lh) view, i.e. the mean runs over capsules.
    """
    e = np.sqrt(np.sum(np.asarray(irs_cnl, dtype=np.float64) ** 2, axis=-1))  # (C, N)
    e = e + tiny(e)
    return 1.0 / e.mean(axis=0)  # (N,)


# --------------------------------------------------------------------------------------------------
# static (time-invariant) convolution
# --------------------------------------------------------------------------------------------------
def linear_convolve(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Full linear convolution along the last axis, following scipy.signal.fftconvolve's recipe
    (scipy/signal/_signaltools.py `_freq_domain_conv`): real FFTs of length next_fast_len(la+lb-1), each
    operand transformed IN ITS OWN PRECISION — the reference passes float32 audio and float64 IRs
    (synthesize.py:103,490), so the audio spectrum is complex64 and carries ~1e-7 relative noise — then a
    complex128 product and a float64 inverse."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    if b.dtype not in (np.float32, np.float64):
        b = b.astype(np.float64)
    n = a.shape[-1] + b.shape[-1] - 1
    m = sp_fft.next_fast_len(n, True)
    spec = sp_fft.rfft(a, m, axis=-1) * sp_fft.rfft(b, m, axis=-1)
    return sp_fft.irfft(spec, m, axis=-1)[..., :n]


def time_invariant_convolution(audio: np.ndarray, ir: np.ndarray) -> np.ndarray:
    """synthesize.py:71-106 — mono audio (Lx,) with IR (Lh, C) -> (C, Lx+Lh-1); same error strings."""
    if audio.ndim != 1:
        raise ValueError(f"Only mono input is supported, but got {audio.ndim} dimensions!")
    if ir.ndim != 2:
        raise ValueError(
            f"Expected shape of IR should be (n_samples, n_channels), but got ({ir.shape}) instead"
        )
    return linear_convolve(audio[None, :], ir.T)


# --------------------------------------------------------------------------------------------------
# moving (time-variant) convolution — literal STFT-domain form
# --------------------------------------------------------------------------------------------------
def n_stft_frames(n_samples: int, hop: int = HOP_SIZE) -> int:
    """synthesize.py:123."""
    return 2 * int(math.ceil(n_samples / (2.0 * hop))) + 1


def stft(y: np.ndarray, fft_size=FFT_SIZE, win_size=WIN_SIZE, hop_size=HOP_SIZE) -> np.ndarray:
    """synthesize.py:109-145 — sin^2 window, (win-hop) left pad, frames on the last axis then moved to the
    front: output (frames, freq, *leading dims of y)."""
    y = np.asarray(y)
    window = np.sin(np.pi / win_size * np.arange(win_size)) ** 2
    frames = n_stft_frames(y.shape[-1], hop_size)
    left = win_size - hop_size
    right = frames * hop_size - y.shape[-1]
    padded = np.zeros(y.shape[:-1] + (left + y.shape[-1] + right,), dtype=np.float64)
    padded[..., left:left + y.shape[-1]] = y
    idx = np.arange(win_size)[:, None] + hop_size * np.arange(frames)[None, :]  # (win, frames)
    seg = padded[..., idx] * window[:, None]  # (..., win, frames)
    spec = np.fft.rfft(seg, fft_size, axis=-2)  # (..., freq, frames)
    spec = np.moveaxis(np.moveaxis(spec, -2, 0), -1, 0)  # (frames, freq, ...)
    return np.ascontiguousarray(spec)


def ir_start_frames(ir_times: np.ndarray, sr: float, hop_size: int = HOP_SIZE) -> np.ndarray:
    """synthesize.py:169 — round-half-to-even frame index at which each IR starts (float array)."""
    return np.round((np.asarray(ir_times) * sr + hop_size) / hop_size)


def interpolation_matrix(ir_times: np.ndarray, sr: float, hop_size: int = HOP_SIZE,
                         n_frames: Optional[int] = None) -> np.ndarray:
    """synthesize.py:148-181 — (n_frames, n_irs) linear cross-fade weights, assignment semantics."""
    frames = ir_start_frames(ir_times, sr, hop_size)
    if n_frames is None:
        n_frames = int(frames[-1])
    w = np.zeros((n_frames, len(frames)))
    for l in range(len(frames) - 1):
        rows = np.arange(frames[l], frames[l + 1] + 1, dtype=int) - 1
        ramp = np.linspace(0, 1, len(rows))
        w[rows, l] = 1 - ramp
        w[rows, l + 1] = ramp
    return w


def ctf_convolve(s_audio: np.ndarray, s_ir: np.ndarray, w_ir: np.ndarray) -> np.ndarray:
    """synthesize.py:184-252 — Y[i] = sum_{k<=i, k<n_frames_ir} sum_l H[k,:,:,l] * w[i-k,l] * X[i-k].

    Same frame-serial structure and per-frame active-IR sub-selection as the reference (so its cost on a
    CPU is representative), written as one contraction per output frame.
    """
    n_fr_ir, n_freq, n_ch, n_irs = s_ir.shape
    n_frames = min(s_audio.shape[0], w_ir.shape[0])
    out = np.empty((n_frames, n_freq, n_ch), dtype=complex)
    w_c = w_ir.astype(complex)
    for i in range(n_frames):
        depth = min(i + 1, n_fr_ir)
        src = np.arange(i, i - depth, -1)  # source frames i, i-1, ... paired with IR frames 0, 1, ...
        w_win = w_c[src]  # (depth, n_irs)
        h_win = s_ir[:depth]
        active = np.any(w_win != 0, axis=0)
        if active.mean() < 0.5:  # synthesize.py:236
            h_win = h_win[:, :, :, active]
            w_win = w_win[:, active]
        ctf = np.einsum("kfcl,kl->kfc", h_win, w_win)
        out[i] = np.einsum("kfc,kf->fc", ctf, s_audio[src])
    return out


def istft_ola(spec: np.ndarray, fft_size=FFT_SIZE, win_size=WIN_SIZE, hop_size=HOP_SIZE) -> np.ndarray:
    """synthesize.py:255-274 — un-normalised inverse (norm="forward" => x fft_size), OLA, crop
    [win : n_frames*hop] -> (n_frames*hop - win, channels)."""
    n_frames, _, n_ch = spec.shape
    frames = np.fft.irfft(spec, n=fft_size, axis=1) * fft_size
    buf = np.zeros(((n_frames + 1) * hop_size + win_size, n_ch))
    for i in range(n_frames):
        buf[i * hop_size:i * hop_size + fft_size] += frames[i]
    return buf[win_size:n_frames * hop_size]


def time_variant_convolution(irs: np.ndarray, audio: np.ndarray, duration: float, sample_rate: float,
                             fft_size=FFT_SIZE, win_size=WIN_SIZE, hop_size=HOP_SIZE) -> np.ndarray:
    """synthesize.py:277-310 — irs (C, N, Lh), audio (Lx,) -> (C, n_frames*hop - win)."""
    ir_spec = stft(irs, fft_size, win_size, hop_size)
    au_spec = stft(audio, fft_size, win_size, hop_size)
    ir_times = np.linspace(0, duration, irs.shape[1])
    w = interpolation_matrix(ir_times, sample_rate, hop_size)
    return istft_ola(ctf_convolve(au_spec, ir_spec, w), fft_size, win_size, hop_size).T


# --------------------------------------------------------------------------------------------------
# moving convolution — closed form; seconds instead of minutes.
# Cross-checked against the literal form above in tests/test_oracle.py.
# --------------------------------------------------------------------------------------------------
def crossfade_gains(w: np.ndarray, n_samples: int, hop_size: int = HOP_SIZE) -> np.ndarray:
    """g[l, t] = sum_j w[j, l] * win(t + hop - hop*j): Hann-smoothed source-side weight of IR l."""
    win_size = 2 * hop_size
    window = np.sin(np.pi / win_size * np.arange(win_size)) ** 2
    n_frames, n_irs = w.shape
    g = np.zeros((n_irs, (n_frames + 1) * hop_size))
    for j in range(n_frames):
        g[:, j * hop_size:j * hop_size + win_size] += w[j][:, None] * window[None, :]
    g = g[:, hop_size:]  # undo the (win-hop) left pad
    out = np.zeros((n_irs, n_samples))
    m = min(n_samples, g.shape[1])
    out[:, :m] = g[:, :m]
    return out


def time_variant_convolution_closed(irs: np.ndarray, audio: np.ndarray, duration: float,
                                    sample_rate: float) -> np.ndarray:
    """y_c = 512 * sum_l h_{l,c} * (g_l x), cut to n_frames*128-256 samples (same output as the STFT form)."""
    irs = np.asarray(irs, dtype=np.float64)
    audio = np.asarray(audio, dtype=np.float64)
    n_ch, n_irs, _ = irs.shape
    ir_times = np.linspace(0, duration, n_irs)
    w = interpolation_matrix(ir_times, sample_rate, HOP_SIZE)
    n_frames = min(n_stft_frames(len(audio)), w.shape[0])
    n_out = n_frames * HOP_SIZE - WIN_SIZE
    g = crossfade_gains(w[:n_frames], len(audio))
    out = np.zeros((n_ch, max(n_out, 0)))
    for l in range(n_irs):
        nz = np.flatnonzero(g[l])
        if nz.size == 0:
            continue
        lo, hi = nz[0], nz[-1] + 1
        if lo >= n_out:
            continue
        seg = g[l, lo:hi] * audio[lo:hi]
        part = linear_convolve(seg[None, :], irs[:, l, :])
        take = min(part.shape[1], n_out - lo)
        out[:, lo:lo + take] += part[:, :take]
    return FFT_SIZE * out


# --------------------------------------------------------------------------------------------------
# event render / dry audio / mixdown
# --------------------------------------------------------------------------------------------------
@dataclass
class EventResult:
    spatial: np.ndarray                 # (C, Lx) float64  -> event.spatial_audio[mic]
    event_scale: float                  # db_to_multiplier(ref_db + snr, mean|apply_snr(y)|)
    dry: Optional[np.ndarray] = None    # (Lx+Lh-1,) float64 -> event._spatial_audio_dry[mic]


def valid_audio(y: np.ndarray) -> None:
    """The part of librosa.util.valid_audio the path relies on (synthesize.py:552,603,398)."""
    if not np.issubdtype(np.asarray(y).dtype, np.floating):
        raise ValueError("Audio data must be floating-point")
    if not np.isfinite(y).all():
        raise ValueError("Audio buffer is not finite everywhere")


def dry_audio(audio: np.ndarray, irs_norm: np.ndarray, event_scale: float, sample_rate: float,
              ref_ir_channel: int, direct_path_time_ms: Tuple[float, float]) -> np.ndarray:
    """synthesize.py:468-496 — window the (normalised) first IR of the reference capsule around its signed
    peak, convolve, scale by event_scale."""
    if ref_ir_channel > irs_norm.shape[0]:  # sic: off-by-one guard of the reference (:470)
        raise ValueError(
            f"Reference channel index out of range for IRs with {irs_norm.shape[0]} channels"
        )
    low, high = direct_path_time_ms
    low_sp = int(low * sample_rate / 1000)
    high_sp = int(high * sample_rate / 1000)
    h = np.array(irs_norm[ref_ir_channel, 0, :], dtype=np.float64)
    peak = int(np.argmax(h))
    if peak + high_sp < h.shape[0]:
        h[peak + high_sp:] = 0
    if peak - low_sp > 0:
        h[:peak - low_sp] = 0
    return linear_convolve(audio, h) * event_scale


def render_event(audio: np.ndarray, irs: np.ndarray, snr: float, ref_db: float, *, is_moving: bool,
                 duration: Optional[float] = None, sample_rate: Optional[float] = None,
                 ref_ir_channel: Optional[int] = None,
                 direct_path_time_ms: Optional[Tuple[float, float]] = None,
                 literal: bool = True) -> EventResult:
    """synthesize.py:507-608 — audio (Lx,) float32 peak-normalised, irs (C, N, Lh) float64.

    literal=True uses the STFT-domain loop exactly as the reference does; False the closed form.
    """
    irs = np.array(irs, dtype=np.float64, copy=True)
    n_ch, n_emitters, _ = irs.shape
    valid_audio(audio)
    n_audio = audio.shape[0]
    irs_n = normalize_irs(irs.transpose(1, 0, 2)).transpose(1, 0, 2)
    if n_emitters == 1:
        if is_moving:
            raise ValueError("Moving Event has only one emitter!")
        spatial = time_invariant_convolution(audio, irs_n[:, 0].T)
    elif n_emitters == 0:
        spatial = np.repeat(audio[:, None], n_ch, 1).T
    else:
        if not is_moving:
            raise ValueError("Expected a moving event!")
        fn = time_variant_convolution if literal else time_variant_convolution_closed
        spatial = fn(irs_n, audio, duration, sample_rate)
    spatial = pad_or_truncate(spatial, n_audio)
    spatial = apply_snr(spatial, snr)
    event_scale = db_to_multiplier(ref_db + snr, np.mean(np.abs(spatial)))
    spatial = event_scale * spatial
    if spatial.shape != (n_ch, n_audio):
        raise ValueError(f"Incompatible shapes: {spatial.shape} vs {(n_ch, n_audio)}")
    valid_audio(spatial)
    dry = None
    if ref_ir_channel is not None and direct_path_time_ms is not None:
        dry = dry_audio(audio, irs_n, event_scale, sample_rate, ref_ir_channel, direct_path_time_ms)
    return EventResult(spatial=spatial, event_scale=float(event_scale), dry=dry)


def event_slice(scene_start: float, scene_end: float, sample_rate: float, total: int) -> Tuple[int, int]:
    """synthesize.py:361-362 — Python round (half-to-even) of seconds*sr, clipped to the scene."""
    return max(0, round(scene_start * sample_rate)), min(round(scene_end * sample_rate), total)


@dataclass
class MixResult:
    scene: np.ndarray                                        # (C, T) float32 -> scene.audio[mic]
    padded: List[Optional[np.ndarray]] = field(default_factory=list)      # per event (C, T) float32
    dry_padded: List[Optional[np.ndarray]] = field(default_factory=list)  # per event (T,) float32
    slices: List[Tuple[int, int]] = field(default_factory=list)


def mix_scene(duration: float, sample_rate: float,
              spatial: Sequence[np.ndarray], starts: Sequence[float], ends: Sequence[float],
              ambiences: Sequence[Tuple[np.ndarray, float]] = (),
              dry: Optional[Sequence[Optional[np.ndarray]]] = None) -> MixResult:
    """synthesize.py:314-401 for one microphone: float32 scene buffer, ambience first (scaled to ref_db by
    its mean |.|), then every event added into [start, end) in order; per-event padded copies."""
    channels = max(s.shape[0] for s in spatial)
    total = round(duration * sample_rate)
    scene = np.zeros((channels, total), dtype=np.float32)
    for noise, amb_ref_db in ambiences:
        if noise.shape != scene.shape:
            raise ValueError(
                f"Scene ambient noise does not match expected shape. "
                f"Expected {scene.shape}, but got {noise.shape}."
            )
        scene += db_to_multiplier(amb_ref_db, np.mean(np.abs(noise))) * noise
    res = MixResult(scene=scene)
    for i, (sp, s0, s1) in enumerate(zip(spatial, starts, ends)):
        a, b = event_slice(s0, s1, sample_rate, total)
        res.slices.append((a, b))
        if b <= a:
            res.padded.append(None)
            res.dry_padded.append(None)
            continue
        piece = pad_or_truncate(sp, b - a)
        scene[:, a:b] += piece
        pad = np.zeros_like(scene)
        pad[:, a:b] += piece
        res.padded.append(pad)
        d = dry[i] if dry is not None else None
        if d is not None:
            dp = np.zeros(total, dtype=scene.dtype)
            dp[a:b] += pad_or_truncate(d[None, :], b - a)[0]
            res.dry_padded.append(dp)
        else:
            res.dry_padded.append(None)
    valid_audio(scene)
    return res
