"""CPU oracle for the linear event augmentations (TEST INFRASTRUCTURE, scope row f1).

fade / invert / reverse restate audiblelight/augmentation.py:1490-1601 and are pinned by tests/golden/augment.npz
(produced by the unmodified reference). The IIR filters and pre/de-emphasis restate the published algorithms of the
un-vendored dependencies (pedalboard 0.9.17 / JUCE dsp::IIR, librosa 0.11.0) on top of scipy.signal.lfilter:
PARITY UNPINNED for those (the reference's own tests only assert "output differs, same shape",
tests/test_augmentation.py:233-266).
"""
import math

import numpy as np
from scipy import signal

FADE_SHAPES = ["linear", "exponential", "logarithmic", "quarter_sine", "half_sine", "none"]


def _fade_in(length, fade_len, shape):
    if fade_len == 0 or shape == "none":
        return np.ones(length)
    f = np.linspace(0, 1, fade_len)
    if shape == "exponential":
        f = np.power(2, (f - 1)) * f
    elif shape == "logarithmic":
        f = np.log10(0.1 + f) + 1
    elif shape == "quarter_sine":
        f = np.sin(f * math.pi / 2)
    elif shape == "half_sine":
        f = np.sin(f * math.pi - math.pi / 2) / 2 + 0.5
    return np.clip(np.concatenate((f, np.ones(length - fade_len))), 0, 1)


def _fade_out(length, fade_len, shape):
    if fade_len == 0 or shape == "none":
        return np.ones(length)
    f = np.linspace(0, 1, fade_len)
    if shape == "linear":
        f = -f + 1
    elif shape == "exponential":
        f = np.power(2, -f) * (1 - f)
    elif shape == "logarithmic":
        f = np.log10(1.1 - f) + 1
    elif shape == "quarter_sine":
        f = np.sin(f * math.pi / 2 + math.pi / 2)
    elif shape == "half_sine":
        f = np.sin(f * math.pi + math.pi / 2) / 2 + 0.5
    return np.clip(np.concatenate((np.ones(length - fade_len), f)), 0, 1)


def fade(x, sample_rate, fade_in_len, fade_out_len, fade_in_shape, fade_out_shape):
    """augmentation.py:1532-1554."""
    n = x.shape[-1]
    fi = min(int(round(fade_in_len * sample_rate)), n)
    fo = min(int(round(fade_out_len * sample_rate)), n)
    return x * (_fade_in(n, fi, fade_in_shape) * _fade_out(n, fo, fade_out_shape))


def invert(x):
    return np.negative(x)


def reverse(x):
    return np.flip(x, axis=-1)


def gain_db(x, db):
    return x * 10.0 ** (db / 20.0)


def biquad(x, b, a):
    """Zero-initial-state IIR section (pedalboard plugins are called with reset=True, augmentation.py:107-112)."""
    return signal.lfilter(np.asarray(b, dtype=np.float64), np.asarray(a, dtype=np.float64), np.asarray(x, dtype=np.float64))


def preemphasis(x, coef):
    """librosa.effects.preemphasis: lfilter([1, -coef], [1], x, zi = 2 x[0] - x[1])."""
    x = np.asarray(x, dtype=np.float64)
    zi = np.atleast_1d(2 * x[0] - x[1]) if x.shape[-1] > 1 else np.atleast_1d(x[0])
    y, _ = signal.lfilter([1.0, -coef], [1.0], x, zi=zi)
    return y


def deemphasis(x, coef):
    """librosa.effects.deemphasis: lfilter([1], [1, -coef], x) from a zero state, minus the linear-extrapolation term
    ((2 - coef) x[0] - x[1]) / (3 - coef) * coef^n."""
    x = np.asarray(x, dtype=np.float64)
    y, _ = signal.lfilter([1.0], [1.0, -coef], x, zi=np.zeros(1))
    if x.shape[-1] > 1:
        y = y - ((2 - coef) * x[0] - x[1]) / (3 - coef) * (coef ** np.arange(x.shape[-1]))
    return y


def peak_normalize(x):
    """event.py:535-536."""
    x = np.asarray(x)
    tiny = np.finfo(x.dtype if np.issubdtype(x.dtype, np.floating) else np.float32).tiny
    return x / np.max(np.abs(x) + tiny)


def delay(x, sample_rate, delay_seconds, feedback, mix):
    """pedalboard 0.9.17 Delay (Delay.h process loop, restated; JUCE DelayLine without interpolation): per sample
    d = pop(); push(x + feedback * d); y = (1 - mix) * x + mix * d, with an integer delay of int(delay_seconds * sr)."""
    x = np.asarray(x, dtype=np.float64)
    D = int(delay_seconds * sample_rate)
    if D == 0:
        return x.copy()
    n = x.shape[-1]
    line = np.zeros(n + D)  # line[k] = value pushed at sample k
    y = np.empty(n)
    for i in range(n):
        d = line[i - D] if i >= D else 0.0
        line[i] = x[i] + feedback * d
        y[i] = (1.0 - mix) * x[i] + mix * d
    return y
