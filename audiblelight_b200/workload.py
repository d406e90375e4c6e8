"""Synthetic workloads of the named shapes (BASELINE.json configs / SURVEY.md §8d), seeded per scene.

A scene's *structure* (event durations, SNRs, start times, IR counts) comes from numpy's
`default_rng(1000 + scene_idx)` on the host, so every implementation (GPU, oracle, reference) sees the same
scene; the bulk samples (dry audio, RIR taps, ambience) are drawn either with numpy from the same generator
(host path, used by tests / CPU baseline) or with a seeded torch CUDA generator (device path, used for the
180 GB-scale benchmark where drawing 18 GB per GPU on the host would take minutes).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import augment as _aug
from .renderer import EventJob, SceneJob, event_slice, moving_frames, scene_samples


@dataclass
class EventSpec:
    n_audio: int
    n_irs: int
    snr: float
    start: float  # seconds
    aug: Optional[tuple] = None  # one linear augmentation: (kind, cutoff_hz[, gain_db, q]); None = no augmentation


@dataclass
class SceneSpec:
    index: int
    sr: int
    duration: float
    channels: int
    n_ir_samples: int
    ref_db: float
    events: List[EventSpec]
    ambience: bool


def c3_scene_spec(scene_idx: int, sr: int = 24000, duration: float = 60.0, channels: int = 4, lh: int = 24000,
                  n_static: int = 6, n_moving: int = 3, ir_rate: float = 10.0, ambience: bool = True,
                  augment: bool = False) -> SceneSpec:
    """DCASE-style SELD scene with moving events (configs[2] / the unit of configs[4]): 60 s @ 24 kHz, 4 channels,
    1 s RIRs, 6 static + 3 moving events of U(2, 10) s, moving events with one RIR per 100 ms, SNR U(5, 30)."""
    rng = np.random.default_rng(1000 + scene_idx)
    events = []
    for k in range(n_static + n_moving):
        dur = float(rng.uniform(2.0, 10.0))
        n_audio = int(round(dur * sr))
        dur = n_audio / sr
        moving = k >= n_static
        n_irs = int(round(ir_rate * dur)) + 1 if moving else 1
        events.append(EventSpec(n_audio=n_audio, n_irs=n_irs, snr=float(rng.uniform(5.0, 30.0)),
                                start=float(rng.uniform(0.0, duration - dur))))
    if augment:
        # every event is preceded by ONE linear augmentation with seeded parameters (ranges of
        # augmentation.py scaled to the 24 kHz Nyquist) and Event.load_audio's peak normalisation. A separate
        # generator keeps the scene structure identical to the un-augmented workload.
        arng = np.random.default_rng(9000 + scene_idx)
        for e in events:
            kind = ["lowpass", "highpass", "low_shelf", "high_shelf", "peak"][int(arng.integers(0, 5))]
            if kind == "lowpass":
                e.aug = (kind, float(arng.uniform(0.125, 0.45) * sr))
            elif kind == "highpass":
                e.aug = (kind, float(arng.uniform(32.0, 1024.0)))
            else:
                e.aug = (kind, float(arng.uniform(100.0, 0.4 * sr)), float(arng.uniform(-20.0, 10.0)), float(arng.uniform(0.1, 1.0)))
    return SceneSpec(index=scene_idx, sr=sr, duration=duration, channels=channels, n_ir_samples=lh, ref_db=-65.0,
                     events=events, ambience=ambience)


def c2_scene_spec(scene_idx: int) -> SceneSpec:
    """configs[1]: 60 s @ 24 kHz, 9 static events with ambience (one 4-channel array; FOA + MIC = two of these)."""
    return c3_scene_spec(scene_idx, n_static=9, n_moving=0)


def c1_scene_spec(scene_idx: int = 0) -> SceneSpec:
    """configs[0]: one static 10 s event at 24 kHz with a 4-channel 1 s RIR, no ambience."""
    return SceneSpec(index=scene_idx, sr=24000, duration=10.0, channels=4, n_ir_samples=24000, ref_db=-65.0,
                     events=[EventSpec(n_audio=240000, n_irs=1, snr=10.0, start=0.0)], ambience=False)


def c4_scene_spec(scene_idx: int = 0) -> SceneSpec:
    """configs[3]: Eigenmike em64, 48 kHz, 2 s RIRs, 5 overlapping static events of 10 s in a 30 s scene."""
    rng = np.random.default_rng(1000 + scene_idx)
    events = [EventSpec(n_audio=480000, n_irs=1, snr=float(rng.uniform(5.0, 30.0)), start=float(rng.uniform(0, 20.0)))
              for _ in range(5)]
    return SceneSpec(index=scene_idx, sr=48000, duration=30.0, channels=64, n_ir_samples=96000, ref_db=-65.0,
                     events=events, ambience=False)


def aug_coeffs(aug: tuple, sr: float):
    """(b, a) of an EventSpec.aug entry (formulas of audiblelight_b200.augment; shared with the CPU baseline)."""
    kind = aug[0]
    if kind == "lowpass":
        return _aug.lowpass_coeffs(sr, aug[1])
    if kind == "highpass":
        return _aug.highpass_coeffs(sr, aug[1])
    fn = {"low_shelf": _aug.low_shelf_coeffs, "high_shelf": _aug.high_shelf_coeffs, "peak": _aug.peak_coeffs}[kind]
    return fn(sr, aug[1], aug[2], aug[3])


def algorithmic_bytes(spec: SceneSpec) -> int:
    """B_alg of SURVEY.md §8(d): fp32, every compulsory array touched once — RIRs in, dry audio in, ambience in,
    event.spatial_audio out, scene mix out."""
    C, T = spec.channels, scene_samples(spec.duration, spec.sr)
    b = 0
    for e in spec.events:
        b += 4 * C * e.n_irs * spec.n_ir_samples + 4 * C * e.n_audio + 4 * e.n_audio
    b += 4 * C * T * (2 if spec.ambience else 1)
    return b


def ir_bytes(spec: SceneSpec) -> int:
    return sum(4 * spec.channels * e.n_irs * spec.n_ir_samples for e in spec.events)


def _decay(lh: int):
    return np.exp(-np.arange(lh, dtype=np.float64) / (lh / 6.0))


def host_scene_arrays(spec: SceneSpec, dtype=np.float32):
    """numpy inputs: [(audio f32 (Lx,), irs (C, N, Lh))...], ambience (C, T) or None. IRs are returned in `dtype`
    (float64 for the oracle — the reference's backends deliver float64 — float32 for the GPU)."""
    rng = np.random.default_rng(5000 + spec.index)
    decay = _decay(spec.n_ir_samples)
    out = []
    for e in spec.events:
        x = rng.standard_normal(e.n_audio).astype(np.float32)
        x = (x / np.max(np.abs(x) + np.finfo(np.float32).tiny)).astype(np.float32)
        h = (rng.standard_normal((spec.channels, e.n_irs, spec.n_ir_samples)) * decay).astype(dtype)
        out.append((x, h))
    amb = None
    if spec.ambience:
        T = scene_samples(spec.duration, spec.sr)
        a = rng.standard_normal((spec.channels, T))
        amb = (a / np.max(np.abs(a), axis=1, keepdims=True)).astype(np.float32)
    return out, amb


def device_scene_arrays(spec: SceneSpec, device):
    """Same shapes drawn on the GPU with a seeded torch generator (float32)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(5000 + spec.index)
    decay = torch.exp(-torch.arange(spec.n_ir_samples, device=device, dtype=torch.float32) / (spec.n_ir_samples / 6.0))
    out = []
    for e in spec.events:
        x = torch.randn(e.n_audio, device=device, generator=g)
        x = x / x.abs().max()
        h = torch.randn((spec.channels, e.n_irs, spec.n_ir_samples), device=device, generator=g) * decay
        out.append((x, h))
    amb = None
    if spec.ambience:
        T = scene_samples(spec.duration, spec.sr)
        a = torch.randn((spec.channels, T), device=device, generator=g)
        amb = a / a.abs().amax(dim=1, keepdim=True)
    return out, amb


def scene_jobs(spec: SceneSpec, arrays, amb, scene_index: int):
    """EventJobs + SceneJob for one scene (one microphone) from its arrays (numpy or torch)."""
    T = scene_samples(spec.duration, spec.sr)
    jobs = []
    for e, (x, h) in zip(spec.events, arrays):
        dur = e.n_audio / float(spec.sr)
        j = EventJob(audio=x, irs=h, n_channels=spec.channels, snr=e.snr, ref_db=spec.ref_db, scene=scene_index)
        if e.aug is not None:
            j.aug_ops = [_aug.biquad(*aug_coeffs(e.aug, float(spec.sr)))]
            j.normalize_audio = True
        if e.n_irs > 1:
            j.ir_frames, j.n_frames = moving_frames(dur, float(spec.sr), e.n_irs, e.n_audio)
        j.scene_start, j.scene_end = event_slice(e.start, e.start + dur, spec.sr, T)
        jobs.append(j)
    sj = SceneJob(n_channels=spec.channels, n_samples=T, ambience=[amb] if amb is not None else [],
                  ambience_ref_db=[spec.ref_db] if amb is not None else [])
    return jobs, sj


# ---- duck-typed Scene / Event / Ambience objects of the same workload ---------------------------------------------------
# Minimal stand-ins with the attributes the drop-in reads from the reference's classes (core.py / event.py /
# ambience.py); RIRs are float64 (C, N, Lh) arrays in ordinary (pageable) memory, as the reference's backends deliver
# them. Used by bench.py (`e2e.objects_mode`) and tools/dataset_throughput.py.
class SynEmitter:
    def __init__(self, polar):
        self.coordinates_relative_polar = polar


class SynEvent:
    def __init__(self, alias, audio, sr, n_irs, snr, start, rng):
        from collections import OrderedDict
        self.alias, self.audio, self.sample_rate, self.snr = alias, audio, float(sr), snr
        self.duration = len(audio) / float(sr)
        self.scene_start = round(start, 1)  # the DCASE metadata grid is 100 ms
        self.scene_end = self.scene_start + round(self.duration, 1)
        self.is_moving = n_irs > 1
        self.n = n_irs
        self.class_id, self.filename = int(rng.integers(0, 13)), f"{alias}.wav"
        self.emitters = [SynEmitter({"mic000": np.array([[rng.uniform(-180, 180), rng.uniform(-40, 40), rng.uniform(0.5, 5)]])})
                         for _ in range(n_irs)]
        self.spatial_audio, self._spatial_audio_padded = OrderedDict(), OrderedDict()
        self._spatial_audio_dry, self._spatial_audio_dry_padded = OrderedDict(), OrderedDict()

    def load_audio(self, ignore_cache=False, normalize=True):
        return self.audio

    def __len__(self):
        return self.n


class SynAmbience:
    def __init__(self, noise, ref_db):
        self.noise, self.ref_db = noise, ref_db

    def load_ambience(self, normalize=True):
        return self.noise


class SynScene:
    def __init__(self, idx: int, spec: Optional[SceneSpec] = None):
        import types
        from collections import OrderedDict
        spec = spec if spec is not None else c3_scene_spec(idx)
        arrays, amb = host_scene_arrays(spec, dtype=np.float64)
        rng = np.random.default_rng(idx)
        self.duration, self.sample_rate, self.ref_db = spec.duration, spec.sr, spec.ref_db
        evs = [SynEvent(f"event{k:03d}", x, spec.sr, e.n_irs, e.snr, min(e.start, spec.duration - len(x) / spec.sr - 0.2), rng)
               for k, (e, (x, h)) in enumerate(zip(spec.events, arrays))]
        self.events = OrderedDict((e.alias, e) for e in evs)
        self.ambience = OrderedDict(amb=SynAmbience(amb, spec.ref_db)) if amb is not None else OrderedDict()
        self.audio = OrderedDict()
        irs = np.concatenate([h for _, h in arrays], axis=1)
        self.state = types.SimpleNamespace(name="synthetic", microphones=OrderedDict(mic000=None), num_emitters=irs.shape[1],
                                           simulate=lambda: None, get_irs=lambda: OrderedDict(mic000=irs))
        self.index = idx

    def get_events(self):
        return list(self.events.values())

    def to_dict(self):
        return dict(index=self.index, duration=self.duration, events=list(self.events))
