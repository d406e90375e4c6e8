"""Low-level host API: flat event / scene jobs -> one `alr_render` call of the C-ABI library.

Arrays may be numpy arrays (host path, ALR_MEM_HOST) or CUDA torch tensors (device-resident path,
ALR_MEM_DEVICE); torch is used only for allocation and `data_ptr()` hand-off.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from . import augment as _augment
from ._lib import (ALR_GAIN_EVENT, ALR_GAIN_NONE, ALR_MEM_DEVICE, ALR_MEM_HOST, AlrAugOp, AlrEvent, AlrEventStats,
                   AlrProfile, AlrScene)

FFT_SIZE, WIN_SIZE, HOP_SIZE = 512, 256, 128  # the only STFT geometry the kernels implement (config.py:9-11)


# ---- bit-exact host arithmetic shared by every entry point ------------------------------------------------------
def n_stft_frames(n_samples: int, hop_size: int = HOP_SIZE) -> int:
    """Number of STFT frames the reference computes for a signal (synthesize.py:123)."""
    return 2 * int(np.ceil(n_samples / (2.0 * hop_size))) + 1


def moving_frames(duration: float, sample_rate: float, n_irs: int, n_audio: int,
                  hop_size: int = HOP_SIZE) -> Tuple[np.ndarray, int]:
    """IR start frames and frame count of a moving event, with the reference's own expressions:
    ir_times = linspace(0, duration, N) (synthesize.py:302); frames = np.round((t*sr + hop)/hop) (:169, half to
    even); n_frames = min(audio frames, int(frames[-1])) (:170, :208-210)."""
    ir_times = np.linspace(0, duration, n_irs)
    frames = np.round((ir_times * sample_rate + hop_size) / hop_size)
    n_w = int(frames[-1])
    return frames.astype(np.int32), min(n_stft_frames(n_audio, hop_size), n_w)


def event_slice(scene_start: float, scene_end: float, sample_rate, total: int) -> Tuple[int, int]:
    """Sample slice of an event in the scene buffer (synthesize.py:361-362): Python round(), half to even."""
    return max(0, round(scene_start * sample_rate)), min(round(scene_end * sample_rate), total)


def scene_samples(duration, sample_rate) -> int:
    """T = round(scene.duration * scene.sample_rate) (synthesize.py:331)."""
    return round(duration * sample_rate)


# ---- jobs ---------------------------------------------------------------------------------------------------------
@dataclass
class EventJob:
    """One (event, microphone) render; mirrors `alr_event` of include/alrender.h."""
    audio: object = None                  # (Lx,) float32
    irs: object = None                    # (C, N, Lh) float32 (any strides with a contiguous last axis) or None
    n_channels: int = 0
    snr: float = 0.0
    ref_db: float = -65.0
    ir_frames: Optional[np.ndarray] = None  # int32 (N,) for moving events
    n_frames: int = 0
    normalize_irs: bool = True
    gain_mode: int = ALR_GAIN_EVENT
    n_out: Optional[int] = None           # default: len(audio)
    dry: Optional[Tuple[int, int, int]] = None  # (ref channel, low samples, high samples)
    scene: int = -1
    scene_start: int = 0
    scene_end: int = 0
    spatial: object = None                # (C, n_out) float32 out (allocated when None); INPUT when prerendered
    dry_out: object = None                # (Lx+Lh-1,) float32 out
    prerendered: bool = False             # spatial is an input that is only mixed
    stats: Optional[dict] = None
    # f1: linear augmentations applied to the dry audio on the device before the convolution (audiblelight_b200.augment)
    aug_ops: Sequence[object] = ()        # list of augment.AugOp, applied in order
    normalize_audio: bool = False         # then x / max(|x| + tiny) as Event.load_audio(normalize=True)
    audio_out: object = None              # optional (Lx,) float32 out: the augmented / normalised dry audio
    keep_spatial: bool = True             # False (host arrays, mixed events only): render + mix on the device, no copy back


@dataclass
class SceneJob:
    """One (scene, microphone) mixdown; mirrors `alr_scene`."""
    n_channels: int
    n_samples: int
    ambience: Sequence[object] = ()       # each (C, T) float32, or None: Gaussian noise generated on the device (f3)
    ambience_ref_db: Sequence[float] = ()
    ambience_seed: Sequence[int] = ()     # one per layer when any layer is None (Philox key of the generated noise)
    mix: object = None                    # (C, T) float32 out
    pcm16: object = None                  # optional (T, C) int16 out: the PCM_16 samples sf.write(path, mix.T, sr) stores
    keep_mix: bool = True                 # False (host arrays, pcm16 given): only the PCM copy is downloaded


def _is_torch(a) -> bool:
    return hasattr(a, "data_ptr")


def _ptr(a) -> int:
    if a is None:
        return 0
    return a.data_ptr() if _is_torch(a) else a.ctypes.data


_NP_F32 = np.dtype(np.float32)


def _require_f32(a, name: str, contiguous: bool = True) -> None:
    """The C-ABI takes raw float32 pointers: anything else (float64 arrays above all) would be reinterpreted silently."""
    if a is None:
        return
    if isinstance(a, np.ndarray):  # (fast path: this runs for every buffer of every event)
        if a.dtype != _NP_F32:
            raise TypeError(f"{name} must be float32, got {a.dtype}")
        if contiguous and not a.flags.c_contiguous:
            raise ValueError(f"{name} must be C-contiguous")
        return
    if "float32" not in str(a.dtype):
        raise TypeError(f"{name} must be float32, got {a.dtype}")
    if contiguous and not (a.is_contiguous() if _is_torch(a) else a.flags.c_contiguous):
        raise ValueError(f"{name} must be C-contiguous")


def _strides_elems(a):
    if _is_torch(a):
        return tuple(a.stride())
    return tuple(s // a.itemsize for s in a.strides)


class PinnedPool:
    """Page-locked host blocks (alr_pinned_alloc) handed out as numpy arrays and recycled between calls. The drop-in
    converts the reference's float64 RIRs to float32 straight into these blocks, so the upload runs at the full PCIe
    rate instead of the pageable-memory rate. Blocks are kept by size class until `close()`."""

    def __init__(self, lib, handle):
        self._lib, self._h = lib, handle
        self._free = {}   # size class -> [address]
        self._used = []   # (size class, address)

    def take(self, shape, dtype=np.float32) -> np.ndarray:
        dtype = np.dtype(dtype)
        count = int(np.prod(shape, dtype=np.int64))
        nbytes = count * dtype.itemsize
        # best fit among the free blocks; new blocks come in quarter-octave size classes so that batches with
        # slightly different array sizes reuse them instead of calling cudaHostAlloc again (slow: it pins pages)
        best = None
        for cls, bucket in self._free.items():
            if bucket and cls >= nbytes and (best is None or cls < best):
                best = cls
        if best is not None and best <= max(4 * nbytes, 1 << 22):
            cls, addr = best, self._free[best].pop()
        else:
            cls = 1 << 20
            while cls < nbytes:
                cls *= 2
            for frac in (5, 6, 7):  # 5/8, 6/8, 7/8 of the power of two
                if cls // 8 * frac >= nbytes:
                    cls = cls // 8 * frac
                    break
            p = C.c_void_p()
            _lib.check(self._lib.alr_pinned_alloc(self._h, cls, C.byref(p)))
            addr = p.value
        self._used.append((cls, addr))
        buf = (C.c_byte * max(nbytes, 1)).from_address(addr)
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def recycle(self) -> None:
        """All blocks handed out so far may be reused (their arrays must no longer be in use)."""
        for cls, addr in self._used:
            self._free.setdefault(cls, []).append(addr)
        self._used = []

    def close(self) -> None:
        self.recycle()
        for bucket in self._free.values():
            for addr in bucket:
                self._lib.alr_pinned_free(self._h, C.c_void_p(addr))
        self._free = {}


class Renderer:
    """One context per GPU (alr_create). Not thread-safe; calls block until results are ready."""

    def __init__(self, device: int = -1, workspace_limit: Optional[int] = None, profiling: bool = False, **options):
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.alr_create(int(device), C.byref(h)))
        self._h = h
        self.device = device
        if workspace_limit is not None:
            _lib.check(self._lib.alr_set_workspace_limit(self._h, int(workspace_limit)))
        if profiling:
            _lib.check(self._lib.alr_set_profiling(self._h, 1))
        for name, value in options.items():  # alr_set_option switches: fused=1, ring_bytes=..., lookahead=..., mix_group=...
            self.set_option(name, value)
        self.pool = PinnedPool(self._lib, self._h)

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self, "pool", None) is not None:
                self.pool.close()
            self._lib.alr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name: str, value: int):
        _lib.check(self._lib.alr_set_option(self._h, name.encode(), int(value)))

    def set_profiling(self, enable: bool):
        _lib.check(self._lib.alr_set_profiling(self._h, 1 if enable else 0))

    def profile(self) -> dict:
        p = AlrProfile()
        _lib.check(self._lib.alr_get_profile(self._h, C.byref(p)))
        return {k: getattr(p, k) for k, _ in AlrProfile._fields_}

    # -- marshalling ---------------------------------------------------------------------------------------------
    @staticmethod
    def _alloc_like(ref, shape):
        if _is_torch(ref):
            import torch
            return torch.empty(shape, dtype=torch.float32, device=ref.device)
        return np.empty(shape, dtype=np.float32)

    def pack(self, events: Sequence[EventJob], scenes: Sequence[SceneJob]):
        """Builds the ctypes descriptor arrays once (reusable across repeated render calls on the same buffers)."""
        keep = []
        n_ev, n_sc = len(events), len(scenes)
        ev_arr = (AlrEvent * max(n_ev, 1))()
        device_mode = None
        for i, e in enumerate(events):
            a = ev_arr[i]
            ref = e.spatial if e.prerendered else e.audio
            mode = _is_torch(ref)
            if device_mode is None:
                device_mode = mode
            elif device_mode != mode:
                raise ValueError("cannot mix host (numpy) and device (torch) buffers in one call")
            C_ = int(e.n_channels)
            _require_f32(e.audio, "audio")
            _require_f32(e.irs, "irs", contiguous=False)
            _require_f32(e.spatial, "spatial")
            _require_f32(e.dry_out, "dry_out")
            _require_f32(e.audio_out, "audio_out")
            if e.prerendered:
                n_out = int(e.spatial.shape[1])
                a.n_irs = -1
                a.n_channels = C_
                a.n_out = n_out
                a.spatial = _ptr(e.spatial)
                a.gain_mode = ALR_GAIN_NONE
            else:
                lx = int(e.audio.shape[0])
                n_out = int(e.n_out) if e.n_out is not None else lx
                a.audio = _ptr(e.audio)
                a.n_audio = lx
                n_irs = 0 if e.irs is None else int(e.irs.shape[1])
                a.n_irs = n_irs
                a.n_channels = C_
                if n_irs > 0:
                    sc, sn, st = _strides_elems(e.irs)
                    if st != 1:
                        raise ValueError("IR taps must be contiguous along the last axis")
                    if int(e.irs.shape[0]) != C_:
                        raise ValueError("irs.shape[0] != n_channels")
                    a.irs = _ptr(e.irs)
                    a.ir_stride_c, a.ir_stride_n = int(sc), int(sn)
                    a.n_ir_samples = int(e.irs.shape[2])
                if n_irs > 1:
                    fr = np.ascontiguousarray(e.ir_frames, dtype=np.int32)
                    if fr.shape != (n_irs,):
                        raise ValueError("ir_frames must have one entry per IR")
                    keep.append(fr)
                    a.ir_frames = fr.ctypes.data
                    a.n_frames = int(e.n_frames)
                if e.aug_ops:
                    ops = (AlrAugOp * len(e.aug_ops))()
                    for k, op in enumerate(e.aug_ops):
                        ops[k].type = int(op.type)
                        ops[k].fade_in_shape, ops[k].fade_out_shape = int(op.fade_in_shape), int(op.fade_out_shape)
                        if op.type == _augment.ALR_AUG_FADE:
                            ops[k].fade_in_samples, ops[k].fade_out_samples = _augment.fade_samples(op, lx)
                        for q, v in enumerate(op.p):
                            ops[k].p[q] = float(v)
                    keep.append(ops)
                    a.aug_ops = C.cast(ops, C.c_void_p)
                    a.n_aug_ops = len(e.aug_ops)
                a.normalize_audio = 1 if e.normalize_audio else 0
                if (e.aug_ops or e.normalize_audio) and e.audio_out is not None:
                    if tuple(e.audio_out.shape) != (lx,):
                        raise ValueError("audio_out must have shape (n_audio,)")
                    a.audio_out = _ptr(e.audio_out)
                a.normalize_irs = 1 if e.normalize_irs else 0
                a.gain_mode = int(e.gain_mode)
                a.snr = float(e.snr)
                a.ref_db = float(e.ref_db)
                if not e.keep_spatial:
                    if _is_torch(e.audio) or e.scene < 0:
                        raise ValueError("keep_spatial=False needs host arrays and an event that is mixed into a scene")
                    e.spatial = None
                elif e.spatial is None:
                    e.spatial = self._alloc_like(e.audio, (C_, n_out))
                if e.spatial is not None and tuple(e.spatial.shape) != (C_, n_out):
                    raise ValueError(f"spatial buffer has shape {tuple(e.spatial.shape)}, expected {(C_, n_out)}")
                a.spatial = _ptr(e.spatial)
                a.n_out = n_out
                if e.dry is not None:
                    ch, lo, hi = e.dry
                    a.dry_channel, a.dry_low, a.dry_high = int(ch), int(lo), int(hi)
                    n_dry = lx + int(e.irs.shape[2]) - 1
                    if e.dry_out is None:
                        e.dry_out = self._alloc_like(e.audio, (n_dry,))
                    a.dry = _ptr(e.dry_out)
            a.scene = int(e.scene)
            a.scene_start = int(e.scene_start)
            a.scene_end = int(e.scene_end)
        sc_arr = (AlrScene * max(n_sc, 1))()
        for i, s in enumerate(scenes):
            b = sc_arr[i]
            b.n_channels = int(s.n_channels)
            b.n_samples = int(s.n_samples)
            n_amb = len(s.ambience)
            b.n_ambience = n_amb
            if n_amb:
                if any(amb is None for amb in s.ambience) and len(s.ambience_seed) != n_amb:
                    raise ValueError("ambience_seed needs one entry per ambience layer when a layer is generated (None)")
                for amb in s.ambience:
                    if amb is None:
                        continue
                    if tuple(amb.shape) != (s.n_channels, s.n_samples):
                        raise ValueError(
                            f"Scene ambient noise does not match expected shape. "
                            f"Expected {(s.n_channels, s.n_samples)}, but got {tuple(amb.shape)}.")
                    _require_f32(amb, "ambience")
                    if device_mode is None:
                        device_mode = _is_torch(amb)
                ptrs = (C.c_void_p * n_amb)(*[_ptr(amb) for amb in s.ambience])
                dbs = (C.c_double * n_amb)(*[float(d) for d in s.ambience_ref_db])
                keep += [ptrs, dbs]
                b.ambience = C.cast(ptrs, C.c_void_p)
                b.ambience_ref_db = C.cast(dbs, C.c_void_p)
                if len(s.ambience_seed) == n_amb:
                    seeds = (C.c_uint64 * n_amb)(*[int(v) & 0xFFFFFFFFFFFFFFFF for v in s.ambience_seed])
                    keep.append(seeds)
                    b.ambience_seed = C.cast(seeds, C.c_void_p)
            if s.pcm16 is not None:
                if tuple(s.pcm16.shape) != (s.n_samples, s.n_channels) or "int16" not in str(s.pcm16.dtype):
                    raise ValueError("pcm16 must be an int16 array of shape (n_samples, n_channels)")
                if not (s.pcm16.is_contiguous() if _is_torch(s.pcm16) else s.pcm16.flags.c_contiguous):
                    raise ValueError("pcm16 must be contiguous")
                b.pcm16 = _ptr(s.pcm16)
            if not s.keep_mix:
                if s.pcm16 is None or _is_torch(s.pcm16):
                    raise ValueError("keep_mix=False needs a host pcm16 buffer")
                b.mix = 0
                continue
            if s.mix is None:
                ref = next((a for a in s.ambience if a is not None), None)
                if ref is None:
                    ref = next((e.spatial for e in events if e.spatial is not None), None)
                if ref is None:
                    ref = np.empty(0, dtype=np.float32)
                s.mix = self._alloc_like(ref, (s.n_channels, s.n_samples))
            _require_f32(s.mix, "mix")
            b.mix = _ptr(s.mix)
        stats = (AlrEventStats * max(n_ev, 1))()
        return dict(ev=ev_arr, sc=sc_arr, n_ev=n_ev, n_sc=n_sc, stats=stats, keep=keep,
                    mode=ALR_MEM_DEVICE if device_mode else ALR_MEM_HOST, events=events, scenes=scenes)

    def run(self, packed, stream: int = 0):
        """One alr_render call on pre-packed descriptors. `stream` is a raw cudaStream_t (0 = default stream)."""
        rc = self._lib.alr_render(self._h, packed["ev"], packed["n_ev"], packed["sc"], packed["n_sc"], packed["mode"],
                                  packed["stats"], C.c_void_p(stream))
        _lib.check(rc)
        return packed["stats"]

    def render(self, events: Sequence[EventJob], scenes: Sequence[SceneJob] = (), stream: int = 0):
        """Renders the events (+ optional scene mixdowns) and fills `spatial` / `dry_out` / `mix` and `stats`."""
        packed = self.pack(events, scenes)
        st = self.run(packed, stream)
        for i, e in enumerate(events):
            s = st[i]
            e.stats = dict(peak=s.peak, mean_abs=s.mean_abs, gain=s.gain, event_scale=s.event_scale,
                           nonfinite=bool(s.nonfinite), dry_peak=s.dry_peak)
        return events, scenes

    # -- FFT unit-test hooks ----------------------------------------------------------------------------------------
    def debug_rfft(self, blocks):
        """blocks: CUDA float32 tensor (n_blocks, n_valid<=P) -> packed spectra (n_blocks, P, 2)."""
        import torch
        P = self._lib.alr_partition_size()
        n_blocks, n_valid = blocks.shape
        out = torch.empty((n_blocks, P, 2), dtype=torch.float32, device=blocks.device)
        _lib.check(self._lib.alr_debug_rfft(self._h, blocks.data_ptr(), n_blocks, blocks.stride(0), n_valid,
                                            out.data_ptr(), None))
        return out

    def debug_irfft(self, spec):
        import torch
        P = self._lib.alr_partition_size()
        n_blocks = spec.shape[0]
        out = torch.empty((n_blocks, 2 * P), dtype=torch.float32, device=spec.device)
        _lib.check(self._lib.alr_debug_irfft(self._h, spec.data_ptr(), n_blocks, out.data_ptr(), None))
        return out


def debug_plan(job: EventJob) -> dict:
    """Host-only: the partition plan the C++ planner derives for one event (no GPU needed)."""
    lib = _lib.load()
    dummy = np.zeros(1, dtype=np.float32)
    lx = int(job.audio.shape[0])
    a = AlrEvent()
    a.audio = _ptr(job.audio)
    a.n_audio = lx
    n_irs = int(job.irs.shape[1])
    a.n_irs = n_irs
    a.n_channels = int(job.n_channels)
    sc, sn, _ = _strides_elems(job.irs)
    a.irs = _ptr(job.irs)
    a.ir_stride_c, a.ir_stride_n = int(sc), int(sn)
    a.n_ir_samples = int(job.irs.shape[2])
    fr = None
    if n_irs > 1:
        fr = np.ascontiguousarray(job.ir_frames, dtype=np.int32)
        a.ir_frames = fr.ctypes.data
        a.n_frames = int(job.n_frames)
    a.n_out = int(job.n_out) if job.n_out is not None else lx
    a.spatial = dummy.ctypes.data
    a.scene = -1
    header = np.zeros(8, dtype=np.int32)
    cap_ir = 6 * max(n_irs, 1)
    irs = np.zeros(cap_ir, dtype=np.int32)
    cap_w = 4 * (int(job.n_frames) + 2 * n_irs + 16)
    wband = np.zeros(cap_w, dtype=np.float32)
    cap_l = 2 * (a.n_out // 64 + 16)
    lrange = np.zeros(cap_l, dtype=np.int32)
    _lib.check(lib.alr_debug_plan(C.byref(a), header.ctypes.data, irs.ctypes.data, cap_ir, wband.ctypes.data, cap_w,
                                  lrange.ctypes.data, cap_l))
    K, B_valid, B_out, n_valid, xlimit, n_ir, n_w, n_l = [int(v) for v in header]
    return dict(K=K, B_valid=B_valid, B_out=B_out, n_valid=n_valid, xlimit=xlimit,
                irs=irs[:6 * n_ir].reshape(n_ir, 6).copy(), wband=wband[:n_w].copy(),
                lrange=lrange[:2 * n_l].reshape(n_l, 2).copy(), P=lib.alr_partition_size())
