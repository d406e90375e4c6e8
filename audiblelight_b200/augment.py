"""Host-side description of the linear event augmentations the device can apply to the dry audio (scope row f1).

Each constructor returns an `AugOp` for `EventJob.aug_ops`; `Renderer.pack` turns them into `alr_aug_op` records.

Parity status:
  * fade / invert / reverse restate the reference's own numpy code (augmentation.py:1490-1601) -> PINNED by golden
    vectors produced by the unmodified reference (tests/golden/augment.npz).
  * gain, the first-order low/high-pass, the shelves, the peak filter (pedalboard 0.9.17 -> JUCE dsp::IIR) and
    pre/de-emphasis (librosa 0.11) live in dependencies that are neither vendored nor installable offline. Their
    coefficient formulas below are the published JUCE `IIR::ArrayCoefficients` / RBJ-cookbook and librosa ones;
    the device result is checked against scipy.signal.lfilter with the same coefficients -> parity UNPINNED.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

(ALR_AUG_GAIN, ALR_AUG_INVERT, ALR_AUG_REVERSE, ALR_AUG_FADE, ALR_AUG_BIQUAD, ALR_AUG_PREEMPHASIS, ALR_AUG_DEEMPHASIS,
 ALR_AUG_DELAY) = range(8)
FADE_SHAPES = ["linear", "exponential", "logarithmic", "quarter_sine", "half_sine", "none"]  # augmentation.py FADE_SHAPES


@dataclass
class AugOp:
    type: int
    p: Tuple[float, ...] = ()
    fade_in_len: float = 0.0      # seconds (Fade)
    fade_out_len: float = 0.0
    fade_in_shape: int = 5
    fade_out_shape: int = 5
    sample_rate: float = 0.0


def gain_db(db: float) -> AugOp:
    """pedalboard.Gain(gain_db) (augmentation.py:1105-1136): y = x * 10^(dB/20)."""
    return AugOp(ALR_AUG_GAIN, (10.0 ** (db / 20.0),))


def invert() -> AugOp:
    """augmentation.py:1557-1580."""
    return AugOp(ALR_AUG_INVERT)


def reverse() -> AugOp:
    """augmentation.py:1583-1601."""
    return AugOp(ALR_AUG_REVERSE)


def fade(sample_rate: float, fade_in_len: float, fade_out_len: float, fade_in_shape: str, fade_out_shape: str) -> AugOp:
    """augmentation.py:1403-1554. The sample counts min(int(round(len * sr)), L) are fixed when the job is packed."""
    for sh in (fade_in_shape, fade_out_shape):
        if sh not in FADE_SHAPES:
            raise ValueError(f"Expected `shape` to be one of {', '.join(FADE_SHAPES)} but got {sh}")
    return AugOp(ALR_AUG_FADE, (), float(fade_in_len), float(fade_out_len), FADE_SHAPES.index(fade_in_shape),
                 FADE_SHAPES.index(fade_out_shape), float(sample_rate))


def fade_samples(op: AugOp, n_audio: int) -> Tuple[int, int]:
    return (min(int(round(op.fade_in_len * op.sample_rate)), n_audio),
            min(int(round(op.fade_out_len * op.sample_rate)), n_audio))


def biquad(b: Sequence[float], a: Sequence[float]) -> AugOp:
    """Generic second-order section with zero initial state; coefficients are normalised by a[0]."""
    b = list(b) + [0.0] * (3 - len(b))
    a = list(a) + [0.0] * (3 - len(a))
    return AugOp(ALR_AUG_BIQUAD, (b[0] / a[0], b[1] / a[0], b[2] / a[0], a[1] / a[0], a[2] / a[0]))


def lowpass_coeffs(sample_rate: float, cutoff_hz: float):
    """JUCE makeFirstOrderLowPass (pedalboard.LowpassFilter, augmentation.py:303-345)."""
    n = math.tan(math.pi * cutoff_hz / sample_rate)
    return [n, n], [n + 1.0, n - 1.0]


def highpass_coeffs(sample_rate: float, cutoff_hz: float):
    """JUCE makeFirstOrderHighPass (pedalboard.HighpassFilter, augmentation.py:406-446)."""
    n = math.tan(math.pi * cutoff_hz / sample_rate)
    return [1.0, -1.0], [n + 1.0, n - 1.0]


def low_shelf_coeffs(sample_rate: float, cutoff_hz: float, gain_db_: float, q: float):
    """JUCE makeLowShelf (pedalboard.LowShelfFilter, augmentation.py:449-504); gainFactor = 10^(dB/20)."""
    A = math.sqrt(10.0 ** (gain_db_ / 20.0))
    am1, ap1 = A - 1.0, A + 1.0
    omega = 2.0 * math.pi * max(cutoff_hz, 2.0) / sample_rate
    coso = math.cos(omega)
    beta = math.sin(omega) * math.sqrt(A) / q
    amc = am1 * coso
    return ([A * (ap1 - amc + beta), A * 2.0 * (am1 - ap1 * coso), A * (ap1 - amc - beta)],
            [ap1 + amc + beta, -2.0 * (am1 + ap1 * coso), ap1 + amc - beta])


def high_shelf_coeffs(sample_rate: float, cutoff_hz: float, gain_db_: float, q: float):
    """JUCE makeHighShelf (pedalboard.HighShelfFilter, augmentation.py:348-403)."""
    A = math.sqrt(10.0 ** (gain_db_ / 20.0))
    am1, ap1 = A - 1.0, A + 1.0
    omega = 2.0 * math.pi * max(cutoff_hz, 2.0) / sample_rate
    coso = math.cos(omega)
    beta = math.sin(omega) * math.sqrt(A) / q
    amc = am1 * coso
    return ([A * (ap1 + amc + beta), A * -2.0 * (am1 + ap1 * coso), A * (ap1 + amc - beta)],
            [ap1 - amc + beta, 2.0 * (am1 - ap1 * coso), ap1 - amc - beta])


def peak_coeffs(sample_rate: float, cutoff_hz: float, gain_db_: float, q: float):
    """JUCE makePeakFilter (pedalboard.PeakFilter; MultibandEqualizer is a cascade of these, augmentation.py:507-660)."""
    A = math.sqrt(10.0 ** (gain_db_ / 20.0))
    omega = 2.0 * math.pi * max(cutoff_hz, 2.0) / sample_rate
    alpha = math.sin(omega) / (2.0 * q)
    c2 = -2.0 * math.cos(omega)
    return [1.0 + alpha * A, c2, 1.0 - alpha * A], [1.0 + alpha / A, c2, 1.0 - alpha / A]


def lowpass(sample_rate: float, cutoff_hz: float) -> AugOp:
    return biquad(*lowpass_coeffs(sample_rate, cutoff_hz))


def highpass(sample_rate: float, cutoff_hz: float) -> AugOp:
    return biquad(*highpass_coeffs(sample_rate, cutoff_hz))


def low_shelf(sample_rate: float, cutoff_hz: float, gain_db_: float, q: float) -> AugOp:
    return biquad(*low_shelf_coeffs(sample_rate, cutoff_hz, gain_db_, q))


def high_shelf(sample_rate: float, cutoff_hz: float, gain_db_: float, q: float) -> AugOp:
    return biquad(*high_shelf_coeffs(sample_rate, cutoff_hz, gain_db_, q))


def peak(sample_rate: float, cutoff_hz: float, gain_db_: float, q: float) -> AugOp:
    return biquad(*peak_coeffs(sample_rate, cutoff_hz, gain_db_, q))


def multiband_equalizer(sample_rate: float, bands: Sequence[Tuple[float, float, float]]) -> List[AugOp]:
    """bands: (cutoff_hz, gain_db, q) per peak filter, applied in order (augmentation.py:643-660)."""
    return [peak(sample_rate, f, g, q) for f, g, q in bands]


def preemphasis(coef: float) -> AugOp:
    """librosa.effects.preemphasis (augmentation.py:1350-1385)."""
    return AugOp(ALR_AUG_PREEMPHASIS, (float(coef),))


def deemphasis(coef: float) -> AugOp:
    """librosa.effects.deemphasis (augmentation.py:1388-1400)."""
    return AugOp(ALR_AUG_DEEMPHASIS, (float(coef),))


def delay(sample_rate: float, delay_seconds: float, feedback: float, mix: float) -> AugOp:
    """Delay (augmentation.py:1046-1102 -> pedalboard.Delay): integer delay line of int(delay_seconds * sample_rate)
    samples with feedback and a linear dry/wet mix. Parity UNPINNED (pedalboard is not available offline)."""
    if not 0.0 <= feedback <= 1.0 or not 0.0 <= mix <= 1.0:
        raise ValueError("feedback and mix must lie in [0, 1]")
    return AugOp(ALR_AUG_DELAY, (float(int(delay_seconds * sample_rate)), float(feedback), float(mix)))


# ---- mapping from the reference's Augmentation objects ---------------------------------------------------------------------
def from_reference(aug) -> "List[AugOp] | None":
    """Device ops equivalent to one `audiblelight.augmentation.EventAugmentation` instance, or None when the effect is
    not a linear filter the device implements (compressors, pitch shift, time warps, codecs ... stay on the host).
    Dispatch is by class name and the instance's own `params` dict / `sample_rate` (augmentation.py: `self.params`
    of every class), so nothing from the reference package needs importing."""
    name = type(aug).__name__
    p = dict(getattr(aug, "params", {}) or {})
    sr = float(getattr(aug, "sample_rate", 0.0) or 0.0)
    try:
        if name == "LowpassFilter":
            return [lowpass(sr, p["cutoff_frequency_hz"])]
        if name == "HighpassFilter":
            return [highpass(sr, p["cutoff_frequency_hz"])]
        if name == "LowShelfFilter":
            return [low_shelf(sr, p["cutoff_frequency_hz"], p["gain_db"], p["q"])]
        if name == "HighShelfFilter":
            return [high_shelf(sr, p["cutoff_frequency_hz"], p["gain_db"], p["q"])]
        if name == "MultibandEqualizer":
            return multiband_equalizer(sr, list(zip(p["cutoff_frequency_hz"], p["gain_db"], p["q"])))
        if name == "Gain":
            return [gain_db(p["gain_db"])]
        if name == "Invert":
            return [invert()]
        if name == "Reverse":
            return [reverse()]
        if name == "Preemphasis":
            return [preemphasis(p["coef"])]
        if name == "Deemphasis":
            return [deemphasis(p["coef"])]
        if name == "Fade":
            return [fade(sr, p["fade_in_len"], p["fade_out_len"], p["fade_in_shape"], p["fade_out_shape"])]
        if name == "Delay":
            return [delay(sr, p["delay_seconds"], p["feedback"], p["mix"])]
    except (KeyError, TypeError, ValueError):
        return None
    return None


def chain_from_reference(augmentations, max_ops: int = 8) -> "List[AugOp] | None":
    """Op list for `Event.augmentations` (applied in order, event.py:530-532), or None if any entry is unsupported or
    the chain is longer than the device's limit."""
    ops: List[AugOp] = []
    for aug in augmentations:
        o = from_reference(aug)
        if o is None:
            return None
        ops += o
    return ops if len(ops) <= max_ops else None
