"""Drop-in replacements for the synthesis hot path of `audiblelight/synthesize.py` (reference lines cited per
function). Same names, signatures, side effects and error strings; the arithmetic runs in the CUDA library.

    import audiblelight_b200.synthesize as syn
    syn.install()           # rebinds the three functions Scene.generate imports at call time (core.py:1828-1838)
    scene.generate(...)     # now renders on the GPU

or call `render_event_audio`, `render_audio_for_all_scene_events`, `generate_scene_audio_from_events`,
`time_invariant_convolution`, `time_variant_convolution` directly, or `render_scenes([...])` for whole batches.

The objects passed in are the reference's own `Scene` / `Event` / `Ambience` (duck-typed here: nothing from the
reference package is imported unless `install()` is called). There is no CPU fallback.
"""
from __future__ import annotations

import threading
from collections import OrderedDict
from time import time
from typing import Iterable, List, Optional, Sequence

import numpy as np

try:  # same logger as the reference when available
    from loguru import logger
except Exception:  # pragma: no cover
    import logging

    logger = logging.getLogger("audiblelight_b200")

from . import augment as _augment
from .renderer import (ALR_GAIN_EVENT, ALR_GAIN_NONE, FFT_SIZE, HOP_SIZE, WIN_SIZE, EventJob, Renderer, SceneJob,
                       event_slice, moving_frames, scene_samples)

DEFAULT_REF_DB = -65  # config.py:23

try:  # the reference raises librosa's ParameterError from librosa.util.valid_audio (synthesize.py:552,603,398)
    from librosa.util.exceptions import ParameterError  # type: ignore
except Exception:
    class ParameterError(ValueError):
        """Stand-in for librosa.util.exceptions.ParameterError when librosa is not installed."""


_renderers = {}
_lock = threading.Lock()

# f1 (opt-in): run Event.augmentations on the device when every entry is a linear filter the library implements.
# Off by default because the IIR effects wrap pedalboard / librosa, which are not available offline: their device
# arithmetic follows the published formulas but is not pinned against the real dependencies (see augment.py).
DEVICE_AUGMENTATIONS = False

# f3 (opt-in): Gaussian ambience (Ambience(noise="gaussian"), ambience.py:155-163) that has not been loaded yet is drawn on
# the device inside the mixdown call instead of on the host and uploaded (23 MB per one-minute 4-channel layer). The
# reference draws it from numpy's unseeded global generator, so only the distribution is defined; the device stream is
# keyed by a seed taken from that same global generator. `ambience.audio` stays unset (nothing comes back to the host).
DEVICE_AMBIENCE = False


def _load_raw_audio(event) -> np.ndarray:
    """The first half of Event.load_audio (event.py:519-527): the resampled mono clip BEFORE augmentation and
    normalisation, with the reference's own loader call."""
    import librosa  # the reference's dependency; present wherever Event objects exist
    audio_raw, _ = librosa.load(event.filepath, sr=event.sample_rate, mono=True, offset=event.event_start,
                                duration=event.duration, dtype=np.float32)
    return audio_raw


def get_renderer(device: int = -1, slot: int = 0) -> Renderer:
    """Per-process, per-device renderer context (alr_create is done once). `slot` > 0 gives further persistent
    contexts of the same device for callers that keep several batches in flight (audiblelight_b200.dataset)."""
    with _lock:
        r = _renderers.get((device, slot))
        if r is None:
            r = Renderer(device)
            _renderers[(device, slot)] = r
        return r


# ---- small host helpers (utils.py / validation), exact reference semantics ---------------------------------------
def _valid_audio(y: np.ndarray) -> None:
    """librosa.util.valid_audio as the path uses it: floating dtype and finite everywhere."""
    y = np.asarray(y)
    if not np.issubdtype(y.dtype, np.floating):
        raise ParameterError("Audio data must be floating-point")
    # one pass, no temporary: a NaN or an infinity anywhere makes the float64 sum non-finite (inf - inf = nan), and a
    # float64 sum of finite float32 / float64 audio cannot overflow for any buffer that fits in memory
    if y.size and not np.isfinite(np.sum(y, dtype=np.float64)):
        if not np.isfinite(y).all():  # (float64 input with values near 1e308 could overflow the sum: settle it exactly)
            raise ParameterError("Audio buffer is not finite everywhere")


def _validate_shape(shape_a, shape_b) -> None:
    """utils.validate_shape (utils.py:483-503)."""
    n = max(len(shape_a), len(shape_b))
    pa = tuple(shape_a) + (None,) * (n - len(shape_a))
    pb = tuple(shape_b) + (None,) * (n - len(shape_b))
    for i, (a, b) in enumerate(zip(pa, pb)):
        if a is not None and b is not None and a != b:
            raise ValueError(f"Incompatible shapes at index {i}: {a} != {b} (full shapes: {pa} vs {pb})")


def _check_stft_geometry(fft_size, win_size, hop_size) -> None:
    """The kernels are specialised for the geometry the reference always uses (render_audio_for_all_scene_events
    never passes anything else, synthesize.py:666-672)."""
    for name, v in (("fft_size", fft_size), ("win_size", win_size), ("hop_size", hop_size)):
        if isinstance(v, bool) or not isinstance(v, (int, float, np.integer, np.floating)):
            raise TypeError("Expected a positive numeric input, but got {}".format(type(v)))
        if v < 0:
            raise ValueError(f"Expected a positive numeric input, but got {v}")
    if (int(fft_size), int(win_size), int(hop_size)) != (FFT_SIZE, WIN_SIZE, HOP_SIZE):
        raise ValueError(
            f"audiblelight_b200 implements fft_size/win_size/hop_size = {FFT_SIZE}/{WIN_SIZE}/{HOP_SIZE} only, "
            f"got {fft_size}/{win_size}/{hop_size}")


_CONVERT_POOL = None


def _convert_pool():
    """Small thread pool for the big dtype conversions of a batch (numpy releases the GIL while it copies)."""
    global _CONVERT_POOL
    if _CONVERT_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        import os
        _CONVERT_POOL = ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 2) - 1)))
    return _CONVERT_POOL


def _to_f64_many(arrays):
    """float32 -> float64 copies of a batch of result arrays (what the reference stores in event.spatial_audio), in parallel."""
    arrays = list(arrays)
    if sum(a.size for a in arrays) < (1 << 22):
        return [a.astype(np.float64) for a in arrays]
    return list(_convert_pool().map(lambda a: a.astype(np.float64), arrays))


def _as_f32(a, pool=None) -> np.ndarray:
    """C-contiguous float32 copy (no copy if it already is one). The reference's backends deliver float64 RIRs
    (worldstate.py:2210-2212): for a batch of scenes this conversion is the largest host cost of the drop-in, so
    big arrays are converted in slices by a small thread pool (numpy releases the GIL while it copies). With `pool`
    (a renderer's PinnedPool) the result lives in page-locked memory, which the library uploads at the full PCIe rate;
    such arrays are only valid until the pool is recycled."""
    a = np.asarray(a)
    if pool is not None and a.ndim > 0 and a.size >= (1 << 14):
        out = pool.take(a.shape)
        if a.size < (1 << 21):
            np.copyto(out, a, casting="same_kind")
            return out
    else:
        if a.dtype == np.float32 and a.flags.c_contiguous:
            return a
        if a.ndim == 0 or a.size < (1 << 21):
            return np.ascontiguousarray(a, dtype=np.float32)
        out = np.empty(a.shape, dtype=np.float32)
    axis = int(np.argmax(a.shape))
    n = a.shape[axis]
    parts = min(16, n)
    bounds = [n * k // parts for k in range(parts + 1)]

    def conv(k):
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(bounds[k], bounds[k + 1])
        np.copyto(out[tuple(sl)], a[tuple(sl)], casting="same_kind")
    list(_convert_pool().map(conv, range(parts)))
    return out


# ---- raw convolutions ------------------------------------------------------------------------------------------------
def time_invariant_convolution(audio: np.ndarray, ir: np.ndarray) -> np.ndarray:
    """synthesize.py:71-106 — mono audio (n_samples,) * IR (n_ir_samples, n_channels) -> (n_channels, n+m-1)."""
    if audio.ndim != 1:
        raise ValueError(f"Only mono input is supported, but got {audio.ndim} dimensions!")
    if ir.ndim != 2:
        raise ValueError(
            f"Expected shape of IR should be (n_samples, n_channels), but got ({ir.shape}) instead"
        )
    lh, n_ch = ir.shape
    irs = _as_f32(np.asarray(ir).T)[:, None, :]
    job = EventJob(audio=_as_f32(audio), irs=irs, n_channels=n_ch, normalize_irs=False, gain_mode=ALR_GAIN_NONE,
                   n_out=audio.shape[0] + lh - 1)
    get_renderer().render([job])
    return job.spatial.astype(np.float64)


def time_variant_convolution(irs: np.ndarray, event, fft_size=FFT_SIZE, win_size=WIN_SIZE, hop_size=HOP_SIZE) -> np.ndarray:
    """synthesize.py:277-310 — irs (n_channels, n_irs, n_ir_samples), audio from event.load_audio() ->
    (n_channels, n_frames*hop - win)."""
    _check_stft_geometry(fft_size, win_size, hop_size)
    audio = event.load_audio()
    n_ch, n_irs, _ = irs.shape
    frames, n_frames = moving_frames(event.duration, event.sample_rate, len(event), audio.shape[0])
    if len(event) != n_irs:
        raise ValueError(f"Event has {len(event)} emitters but {n_irs} IRs were given")
    n_out = max(n_frames * HOP_SIZE - WIN_SIZE, 0)
    if n_out == 0:
        return np.zeros((n_ch, 0))
    job = EventJob(audio=_as_f32(audio), irs=_as_f32(irs), n_channels=n_ch, ir_frames=frames, n_frames=n_frames,
                   normalize_irs=False, gain_mode=ALR_GAIN_NONE, n_out=n_out)
    get_renderer().render([job])
    return job.spatial.astype(np.float64)


# ---- event rendering ----------------------------------------------------------------------------------------------------
def _event_job(event, irs: np.ndarray, ref_db, ignore_cache: bool, pool=None) -> EventJob:
    """Everything render_event_audio does on the host before the arithmetic (synthesize.py:544-587)."""
    n_ch, n_emitters, n_ir_samples = irs.shape
    ops = None
    augs = list(getattr(event, "augmentations", None) or [])
    cached = bool(getattr(event, "is_audio_loaded", False)) and not ignore_cache
    if DEVICE_AUGMENTATIONS and augs and not cached:
        ops = _augment.chain_from_reference(augs)
    if ops is not None:
        # the device applies the chain and the peak normalisation of Event.load_audio (event.py:530-536)
        audio = _load_raw_audio(event)
        _valid_audio(audio)
        n_audio = audio.shape[0]
        job = EventJob(audio=_as_f32(audio, pool), n_channels=n_ch, snr=float(event.snr), ref_db=float(ref_db), n_out=n_audio,
                       aug_ops=ops, normalize_audio=True, audio_out=np.empty(n_audio, dtype=np.float32))
    else:
        audio = event.load_audio(ignore_cache=ignore_cache, normalize=True)
        _valid_audio(audio)
        n_audio = audio.shape[0]
        job = EventJob(audio=_as_f32(audio, pool), n_channels=n_ch, snr=float(event.snr), ref_db=float(ref_db), n_out=n_audio)
    if n_emitters == 1:
        if event.is_moving:
            raise ValueError("Moving Event has only one emitter!")
        job.irs = _as_f32(irs, pool)
    elif n_emitters == 0:
        logger.warning(
            f"No IRs were found for Event with alias {event.alias}. Audio is being tiled along the "
            f"channel dimension to match the expected shape {n_ch, n_audio}."
        )
        job.irs = None
    else:
        if not event.is_moving:
            raise ValueError("Expected a moving event!")
        if len(event) != n_emitters:
            raise ValueError(f"Event has {len(event)} emitters but {n_emitters} IRs were given")
        job.irs = _as_f32(irs, pool)
        job.ir_frames, job.n_frames = moving_frames(event.duration, event.sample_rate, len(event), n_audio)
    # dry / direct-path audio (compute_dry_audio, synthesize.py:432-504)
    ref_ch = getattr(event, "ref_ir_channel", None)
    dp = getattr(event, "direct_path_time_ms", None)
    if ref_ch is not None and dp is not None:
        if ref_ch > n_ch:  # sic (synthesize.py:470)
            raise ValueError(f"Reference channel index out of range for IRs with {n_ch} channels")
        if n_emitters > 0:
            if ref_ch >= n_ch:
                raise IndexError(f"index {ref_ch} is out of bounds for axis 0 with size {n_ch}")
            low, high = dp
            job.dry = (int(ref_ch), int(low * event.sample_rate / 1000), int(high * event.sample_rate / 1000))
    elif ref_ch is not None or dp is not None:
        logger.warning(
            "Only one of `ref_ir_channel` or `direct_path_time` were specified when creating the Event. "
            "Dry audio will not be computed for this Event. Pass both variables to compute dry audio."
        )
    return job


def _store_event_result(event, job: EventJob, mic_alias: str, spatial64=None) -> None:
    n_ch, n_audio = job.n_channels, job.audio.shape[0]
    if job.stats["nonfinite"]:
        raise ParameterError("Audio buffer is not finite everywhere")
    if job.audio_out is not None:  # device-side augmentation: what Event.load_audio would have cached (event.py:538)
        _valid_audio(job.audio_out)
        event.audio = job.audio_out
    if job.irs is None:
        spatial = job.spatial  # N == 0 stays float32 (:577)
    else:
        spatial = spatial64 if spatial64 is not None else job.spatial.astype(np.float64)
    _validate_shape(spatial.shape, (n_ch, n_audio))
    event.spatial_audio[mic_alias] = spatial
    if job.dry is not None:
        event._spatial_audio_dry[mic_alias] = job.dry_out.astype(np.float64)


def render_event_audio(event, irs: np.ndarray, mic_alias: str, ref_db=DEFAULT_REF_DB, ignore_cache: Optional[bool] = True,
                       fft_size=FFT_SIZE, win_size=WIN_SIZE, hop_size=HOP_SIZE) -> None:
    """synthesize.py:507-608 — renders `event.spatial_audio[mic_alias]` (and the dry audio when requested)."""
    if mic_alias in event.spatial_audio.keys() and not ignore_cache:
        return
    _check_stft_geometry(fft_size, win_size, hop_size)
    job = _event_job(event, np.asarray(irs), ref_db, ignore_cache)
    get_renderer().render([job])
    _store_event_result(event, job, mic_alias)


def validate_scene(scene) -> None:
    """synthesize.py:681-739 — same checks, same messages."""
    if scene.state.num_emitters == 0:
        raise ValueError("WorldState has no emitters!")
    if len(scene.state.microphones) == 0:
        raise ValueError("WorldState has no microphones!")
    if len(scene.events) == 0:
        raise ValueError("Scene has no events!")
    total_ems = 0
    for alias, ev in scene.events.items():
        try:
            total_ems += len(ev)
        except ValueError:
            raise ValueError(f"Event with alias '{alias}' has no emitters registered. Has it been orphaned?")
    if not scene.state.name.upper() == "RLR":
        return
    if scene.state.ctx.get_listener_count() == 0:
        raise ValueError("Ray-tracing engine has no listeners!")
    if scene.state.ctx.get_source_count() == 0:
        raise ValueError("Ray-tracing engine has no sources!")
    vals = (total_ems, scene.state.num_emitters, scene.state.ctx.get_source_count())
    if not all(v == vals[0] for v in vals):
        raise ValueError(
            f"Mismatching number of emitters, events, and sources! "
            f"Got {len(scene.events)} events, {scene.state.num_emitters} emitters, "
            f"{scene.state.ctx.get_source_count()} sources. "
            f"Have any been orphaned?"
        )
    capsules = sum(m.n_listeners for m in scene.state.microphones.values())
    if capsules != scene.state.ctx.get_listener_count():
        raise ValueError(
            f"Mismatching number of microphones and listeners! "
            f"Got {capsules} capsules, {scene.state.ctx.get_listener_count()} listeners. "
            f"Have any been orphaned?"
        )


def _scene_event_jobs(scene, ignore_cache: bool, pool=None):
    """(mic, event) jobs of one scene in the reference's loop order (synthesize.py:653-675), honouring the cache."""
    if ignore_cache:
        scene.state.simulate()
    else:
        try:
            _ = scene.state.irs
        except AttributeError:
            scene.state.simulate()
    validate_scene(scene)
    irs = scene.state.get_irs()
    jobs = []
    for mic_alias, mic_ir in irs.items():
        emitter_counter = 0
        for _, event in scene.events.items():
            n = len(event)
            event_irs = mic_ir[:, emitter_counter:n + emitter_counter, :]
            if not (mic_alias in event.spatial_audio.keys() and not ignore_cache):
                jobs.append((mic_alias, event, _event_job(event, np.asarray(event_irs), scene.ref_db, ignore_cache, pool)))
            emitter_counter += n
    return jobs


def render_audio_for_all_scene_events(scene, ignore_cache: Optional[bool] = False) -> None:
    """synthesize.py:613-677 — all (microphone, event) renders of the scene, as ONE batched GPU call."""
    rnd = get_renderer()
    rnd.pool.recycle()  # inputs are staged in page-locked memory for the duration of this call only
    jobs = _scene_event_jobs(scene, bool(ignore_cache), rnd.pool)
    start = time()
    if jobs:
        rnd.render([j for _, _, j in jobs])
        for mic_alias, event, job in jobs:
            _store_event_result(event, job, mic_alias)
    logger.info(f"Rendered scene audio in {(time() - start):.2f} seconds!")


# ---- mixdown ----------------------------------------------------------------------------------------------------------------
def _is_ambience(obj) -> bool:
    return hasattr(obj, "load_ambience") and hasattr(obj, "ref_db")


def _mix_jobs_for_mic(scene, mic_alias: str, scene_index: int, prerendered: bool, event_jobs=None, pool=None):
    """SceneJob + per-event placement for one microphone (synthesize.py:327-378)."""
    channels = max(ev.spatial_audio[mic_alias].shape[0] for ev in scene.events.values()) if prerendered else \
        max(j.n_channels for j in event_jobs)
    total = scene_samples(scene.duration, scene.sample_rate)
    ambs, dbs, seeds = [], [], []
    if len(scene.ambience) > 0:
        for ambience in scene.ambience.values():
            if not _is_ambience(ambience):
                raise TypeError(f"Expected scene ambient noise to be of type Ambience, but got {type(ambience)}!")
            if (DEVICE_AMBIENCE and getattr(ambience, "beta", None) == "gaussian"
                    and getattr(ambience, "audio", None) is None and int(getattr(ambience, "channels", channels)) == channels):
                ambs.append(None)
                dbs.append(float(ambience.ref_db))
                seeds.append(int(np.random.randint(0, 2 ** 62)))
                continue
            seeds.append(0)
            noise = ambience.load_ambience(normalize=True)
            if noise.shape != (channels, total):
                raise ValueError(
                    f"Scene ambient noise does not match expected shape. "
                    f"Expected {(channels, total)}, but got {noise.shape}."
                )
            ambs.append(_as_f32(noise, pool))
            dbs.append(float(ambience.ref_db))
    sjob = SceneJob(n_channels=channels, n_samples=total, ambience=ambs, ambience_ref_db=dbs,
                    ambience_seed=seeds if any(a is None for a in ambs) else ())
    placements = []
    for event in scene.events.values():
        s0, s1 = event_slice(event.scene_start, event.scene_end, scene.sample_rate, total)
        if s1 <= s0:
            logger.warning(f"Skipping event due to invalid slice: start={s0}, end={s1}")
        placements.append((event, s0, s1))
    return sjob, placements


class _Thunk:
    __slots__ = ("fn",)

    def __init__(self, fn):
        self.fn = fn


class LazyPaddedDict(OrderedDict):
    """`event._spatial_audio_padded` / `_spatial_audio_dry_padded` whose arrays are built on first access.

    The reference materialises a zero-padded (C, T) float32 copy of every event for every microphone inside
    generate_scene_audio_from_events (synthesize.py:381-395) — 23 MB per event of a one-minute 4-channel scene, most
    of the 0.49 s that function takes — although only scripts/ssseg/generate_dataset.py ever
    reads them. Here the dict stores what is needed to build the copy (a reference to the rendered event audio, which
    render calls replace rather than mutate, and the scene slice) and builds it when somebody asks: same keys, same
    values, same dtype."""

    def set_lazy(self, key, fn) -> None:
        OrderedDict.__setitem__(self, key, _Thunk(fn))

    def __getitem__(self, key):
        v = OrderedDict.__getitem__(self, key)
        if isinstance(v, _Thunk):
            v = v.fn()
            OrderedDict.__setitem__(self, key, v)
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def materialize(self) -> "OrderedDict":
        return OrderedDict(self.items())

    def __reduce__(self):  # pickling / deepcopy see plain arrays
        return (OrderedDict, (self.items(),))


def _lazy_dict(event, name: str) -> LazyPaddedDict:
    d = getattr(event, name, None)
    if not isinstance(d, LazyPaddedDict):
        d = LazyPaddedDict(d or ())
        setattr(event, name, d)
    return d


def _store_padded(event, mic_alias, spatial, s0, s1, channels, total, dry) -> None:
    """Per-event zero-padded copies (synthesize.py:381-395), built lazily (LazyPaddedDict)."""
    n = s1 - s0

    def build_spatial():
        padded = np.zeros((channels, total), dtype=np.float32)
        take = min(n, spatial.shape[1])
        padded[:, s0:s0 + take] += spatial[:, :take]
        return padded

    _lazy_dict(event, "_spatial_audio_padded").set_lazy(mic_alias, build_spatial)
    if dry is not None:
        def build_dry():
            dpad = np.zeros(total, dtype=np.float32)
            take = min(n, dry.shape[0])
            dpad[s0:s0 + take] += dry[:take]
            return dpad

        _lazy_dict(event, "_spatial_audio_dry_padded").set_lazy(mic_alias, build_dry)


def generate_scene_audio_from_events(scene) -> None:
    """synthesize.py:314-401 — mixes the already rendered `event.spatial_audio` of every microphone with the
    ambience into `scene.audio[mic]` (float32), and stores the per-event padded copies."""
    mics = list(scene.state.microphones.keys())
    jobs, sjobs, per_mic = [], [], []
    for mi, mic_alias in enumerate(mics):
        sjob, placements = _mix_jobs_for_mic(scene, mic_alias, mi, True)
        sjobs.append(sjob)
        per_mic.append(placements)
        for event, s0, s1 in placements:
            if s1 <= s0:
                continue
            sp = event.spatial_audio[mic_alias]
            if sp.shape[0] != sjob.n_channels:
                raise ValueError(
                    f"operands could not be broadcast together with shapes {(sjob.n_channels, s1 - s0)} {sp.shape}")
            jobs.append(EventJob(spatial=_as_f32(sp), n_channels=sp.shape[0], prerendered=True, scene=mi,
                                 scene_start=s0, scene_end=s1))
    get_renderer().render(jobs, sjobs)
    for mic_alias, sjob, placements in zip(mics, sjobs, per_mic):
        for event, s0, s1 in placements:
            if s1 <= s0:
                continue
            _store_padded(event, mic_alias, event.spatial_audio[mic_alias], s0, s1, sjob.n_channels, sjob.n_samples,
                          event._spatial_audio_dry.get(mic_alias))
        _valid_audio(sjob.mix)
        _validate_shape(sjob.mix.shape, (sjob.n_channels, sjob.n_samples))
        scene.audio[mic_alias] = sjob.mix


# ---- batch entry: many scenes, render + mix in one GPU call ----------------------------------------------------------------
def render_scenes(scenes: Sequence, ignore_cache: bool = True, device: int = -1, store_padded: bool = True,
                  pcm16: bool = False, keep_event_audio: bool = True, keep_mix: bool = True,
                  renderer: Optional[Renderer] = None, pinned: bool = False):
    """Renders and mixes a whole batch of Scene objects with one `alr_render` call (events are rendered and mixed
    on the device without a host round trip). Equivalent to calling `render_audio_for_all_scene_events(scene,
    ignore_cache)` and `generate_scene_audio_from_events(scene)` on every scene.

    Dataset generation (audiblelight_b200.dataset) only needs the mix as the 16-bit PCM that `Scene.generate` writes:
    `pcm16=True` also returns, per scene, `{mic_alias: (T, C) int16}` packed on the device; `keep_event_audio=False`
    leaves `event.spatial_audio` (and the padded copies) unset and `keep_mix=False` leaves `scene.audio` unset, so
    that neither is copied back from the GPU. `renderer` selects an explicit context (one per concurrent caller; the
    default is the per-device singleton). `pinned=True` stages the converted inputs and the PCM output in the
    renderer's page-locked pool: faster transfers, but the returned PCM arrays are only valid until the next call on
    the same renderer."""
    rnd = renderer if renderer is not None else get_renderer(device)
    pool = rnd.pool if pinned else None
    if pool is not None:
        pool.recycle()
    if not keep_mix and not pcm16:
        raise ValueError("keep_mix=False needs pcm16=True")
    if not keep_event_audio:
        store_padded = False
    all_jobs: List[EventJob] = []
    all_scenes: List[SceneJob] = []
    book = []
    for scene in scenes:
        ev_jobs = _scene_event_jobs(scene, ignore_cache, pool)
        by_key = {(m, id(e)): j for m, e, j in ev_jobs}
        for mic_alias in scene.state.microphones.keys():
            mic_jobs = []
            for event in scene.events.values():
                j = by_key.get((mic_alias, id(event)))
                if j is None:  # cached: mix the stored result
                    sp = event.spatial_audio[mic_alias]
                    j = EventJob(spatial=_as_f32(sp), n_channels=sp.shape[0], prerendered=True)
                mic_jobs.append(j)
            sjob, placements = _mix_jobs_for_mic(scene, mic_alias, len(all_scenes), False, mic_jobs, pool)
            for j, (event, s0, s1) in zip(mic_jobs, placements):
                j.scene, j.scene_start, j.scene_end = len(all_scenes), s0, s1
                if j.n_channels != sjob.n_channels and s1 > s0:
                    raise ValueError("all events of a microphone must have the same number of channels")
                if not keep_event_audio and not j.prerendered and j.dry is None:
                    j.keep_spatial = False
            if pcm16:
                sjob.pcm16 = (pool.take((sjob.n_samples, sjob.n_channels), np.int16) if pool is not None else
                              np.empty((sjob.n_samples, sjob.n_channels), dtype=np.int16))
                sjob.keep_mix = keep_mix
            all_jobs += mic_jobs
            all_scenes.append(sjob)
            book.append((scene, mic_alias, sjob, mic_jobs, placements))
    start = time()
    rnd.render(all_jobs, all_scenes)
    packed = {id(scene): OrderedDict() for scene in scenes}
    # float64 copies of every rendered event in one parallel pass (the single largest host cost after the RIR conversion)
    to64 = [j for _, _, _, mic_jobs, _ in book for j in mic_jobs if not j.prerendered and j.keep_spatial and j.irs is not None]
    as64 = dict(zip((id(j) for j in to64), _to_f64_many(j.spatial for j in to64)))
    for scene, mic_alias, sjob, mic_jobs, placements in book:
        for j, (event, s0, s1) in zip(mic_jobs, placements):
            if j.stats is not None and j.stats["nonfinite"]:
                raise ParameterError("Audio buffer is not finite everywhere")
            if not j.prerendered and j.keep_spatial:
                _store_event_result(event, j, mic_alias, as64.get(id(j)))
            if store_padded and s1 > s0:
                _store_padded(event, mic_alias, event.spatial_audio[mic_alias], s0, s1, sjob.n_channels,
                              sjob.n_samples, event._spatial_audio_dry.get(mic_alias))
        if sjob.keep_mix:
            _valid_audio(sjob.mix)
            scene.audio[mic_alias] = sjob.mix
        if pcm16:
            packed[id(scene)][mic_alias] = sjob.pcm16
    logger.info(f"Rendered scene audio in {(time() - start):.2f} seconds!")
    return [packed[id(scene)] for scene in scenes] if pcm16 else None


# ---- installation into the reference package -----------------------------------------------------------------------------------
_PATCHED = ("render_event_audio", "render_audio_for_all_scene_events", "generate_scene_audio_from_events",
            "time_invariant_convolution", "time_variant_convolution")
_originals = {}


def install(device_augmentations: bool = False, device_ambience: bool = False) -> None:
    """Rebinds the hot-path functions on `audiblelight.synthesize`. `Scene.generate` imports them at call time
    (core.py:1828-1831), so every existing caller (tests, scripts/seld/generate_dataset.py) picks them up.
    `device_augmentations=True` additionally moves linear `Event.augmentations` chains onto the GPU (f1),
    `device_ambience=True` draws not-yet-loaded Gaussian ambience on the GPU (f3, see DEVICE_AMBIENCE)."""
    import audiblelight.synthesize as ref  # noqa
    global DEVICE_AUGMENTATIONS, DEVICE_AMBIENCE
    DEVICE_AUGMENTATIONS = bool(device_augmentations)
    DEVICE_AMBIENCE = bool(device_ambience)
    g = globals()
    for name in _PATCHED:
        if name not in _originals:
            _originals[name] = getattr(ref, name)
        setattr(ref, name, g[name])


def uninstall() -> None:
    if not _originals:
        return
    import audiblelight.synthesize as ref  # noqa
    for name, fn in _originals.items():
        setattr(ref, name, fn)
    _originals.clear()
