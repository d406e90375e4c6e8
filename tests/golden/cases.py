"""Seeded input definitions shared by `make_golden.py` (runs the reference) and the tests (run oracle / GPU).

Inputs are regenerated from the seed (numpy's default_rng stream is stable); only reference OUTPUTS are stored.
"""
import numpy as np

# name -> spec of a single `render_event_audio` call (synthesize.py:507)
EVENT_CASES = {
    # static, the quickstart shape scaled down
    "static_4ch": dict(seed=11, sr=24000, lx=6000, lh=2000, c=4, n=1, snr=10.0, ref_db=-65),
    "static_ir_longer_than_audio": dict(seed=12, sr=16000, lx=1500, lh=4000, c=2, n=1, snr=5.0, ref_db=-50),
    "static_neg_snr": dict(seed=13, sr=24000, lx=3000, lh=700, c=4, n=1, snr=-7.5, ref_db=-65),
    "static_mono_odd": dict(seed=14, sr=44100, lx=4411, lh=1023, c=1, n=1, snr=30.0, ref_db=-80),
    "static_silent": dict(seed=15, sr=24000, lx=2000, lh=500, c=4, n=1, snr=10.0, ref_db=-65, silent=True),
    # no emitters -> dry audio tiled over channels (synthesize.py:572-577)
    "no_ir_tiled": dict(seed=16, sr=24000, lx=2500, lh=100, c=4, n=0, snr=12.0, ref_db=-65),
    # moving
    "moving_2ir": dict(seed=21, sr=24000, lx=6000, lh=1500, c=4, n=2, snr=10.0, ref_db=-65),
    "moving_5ir": dict(seed=22, sr=24000, lx=9000, lh=2500, c=4, n=5, snr=20.0, ref_db=-65),
    # Lx = 24000 @ 24 kHz: last IR frame is 188.5 -> round-half-even gives 188
    "moving_half_even": dict(seed=23, sr=24000, lx=24000, lh=3000, c=2, n=11, snr=15.0, ref_db=-65),
    "moving_ir_longer_than_audio": dict(seed=24, sr=16000, lx=3000, lh=8000, c=2, n=3, snr=8.0, ref_db=-60),
    "moving_44k_odd": dict(seed=25, sr=44100, lx=13001, lh=999, c=3, n=4, snr=6.0, ref_db=-65),
    # dry / direct-path audio (synthesize.py:432-504)
    "static_dry": dict(seed=31, sr=24000, lx=5000, lh=2400, c=4, n=1, snr=10.0, ref_db=-65,
                       ref_ir_channel=1, direct_path_time_ms=(6.0, 50.0)),
    "moving_dry": dict(seed=32, sr=24000, lx=5000, lh=1200, c=4, n=3, snr=10.0, ref_db=-65,
                       ref_ir_channel=0, direct_path_time_ms=(2.0, 10.0)),
}

# scenes for `generate_scene_audio_from_events` (synthesize.py:314): a list of events with start times
SCENE_CASES = {
    "scene_static_ambience": dict(
        seed=41, sr=8000, duration=2.0, c=4, ref_db=-65, ambience_ref_db=[-65],
        events=[dict(lx=4000, lh=900, n=1, snr=10.0, start=0.1),
                dict(lx=6000, lh=900, n=1, snr=25.0, start=1.0),  # runs past the scene end -> truncated
                dict(lx=3000, lh=900, n=1, snr=5.0, start=0.50006)]),
    "scene_moving_two_ambiences": dict(
        seed=42, sr=16000, duration=1.5, c=2, ref_db=-55, ambience_ref_db=[-55, -70],
        events=[dict(lx=8000, lh=1600, n=4, snr=12.0, start=0.25),
                dict(lx=5000, lh=1600, n=1, snr=18.0, start=0.0),
                dict(lx=4000, lh=1600, n=1, snr=9.0, start=1.6)]),  # starts after the end -> skipped
    "scene_no_ambience_dry": dict(
        seed=43, sr=24000, duration=1.0, c=4, ref_db=-65, ambience_ref_db=[],
        events=[dict(lx=12000, lh=2000, n=1, snr=10.0, start=0.2, ref_ir_channel=2,
                     direct_path_time_ms=(6.0, 50.0)),
                dict(lx=7000, lh=2000, n=3, snr=14.0, start=0.6)]),
}


def make_audio(rng, lx, silent=False):
    """Peak-normalised float32 mono, like Event.load_audio (event.py:520-536)."""
    if silent:
        return np.zeros(lx, dtype=np.float32)
    x = rng.standard_normal(lx).astype(np.float32)
    x = x / np.max(np.abs(x) + np.finfo(np.float32).tiny)
    return x.astype(np.float32)


def make_irs(rng, c, n, lh):
    """Exponentially decaying Gaussian RIRs, float64 as the backends deliver them (worldstate.py:2210)."""
    t = np.arange(lh)
    decay = np.exp(-t / max(lh / 6.0, 1.0))
    irs = rng.standard_normal((c, n, lh)) * decay
    # a direct-path spike so the dry-audio peak search has something to find
    if n > 0:
        irs[:, :, min(lh - 1, 20)] += 4.0
    return irs


def make_ambience(rng, c, total):
    """Per-channel peak-normalised Gaussian noise (ambience.py:160-165, 211-214)."""
    a = rng.standard_normal((c, total))
    return a / np.max(np.abs(a), axis=1, keepdims=True)


def event_inputs(spec):
    rng = np.random.default_rng(spec["seed"])
    audio = make_audio(rng, spec["lx"], spec.get("silent", False))
    irs = make_irs(rng, spec["c"], spec["n"], spec["lh"])
    return audio, irs


def scene_inputs(spec):
    rng = np.random.default_rng(spec["seed"])
    total = round(spec["duration"] * spec["sr"])
    evs = []
    for e in spec["events"]:
        audio = make_audio(rng, e["lx"])
        irs = make_irs(rng, spec["c"], e["n"], e["lh"])
        evs.append((audio, irs))
    ambs = [make_ambience(rng, spec["c"], total) for _ in spec["ambience_ref_db"]]
    return evs, ambs


# ---- DCASE 2024 metadata (synthesize.py:742-878) -------------------------------------------------------------------------
class DcaseEmitter:
    def __init__(self, polar_by_mic):
        self.coordinates_relative_polar = polar_by_mic  # {mic: (1, 3) array: azimuth deg, elevation deg, distance m}


class DcaseEvent:
    def __init__(self, scene_start, duration, class_id, filename, emitters):
        self.scene_start = scene_start
        self.scene_end = scene_start + duration
        self.class_id = class_id
        self.filename = filename
        self.emitters = emitters
        self.is_moving = len(emitters) > 1


class DcaseScene:
    """The attributes generate_dcase2024_metadata reads from a Scene."""

    def __init__(self, duration, mics, events):
        import types
        from collections import OrderedDict
        self.duration = duration
        self.state = types.SimpleNamespace(microphones=OrderedDict((m, None) for m in mics))
        self._events = events

    def get_events(self):
        return list(self._events)


def dcase_static_scene(duration, events, mic="poltest"):
    """events: [(az, el, dist_m, scene_start, duration, class_id, filename)] — the form of the reference's own
    expected-table tests (tests/test_dcase_metadata.py:247-352)."""
    evs = [DcaseEvent(st, du, cls, fn, [DcaseEmitter({mic: np.array([[az, el, dist]], dtype=np.float64)})])
           for az, el, dist, st, du, cls, fn in events]
    return DcaseScene(duration, [mic], evs)


def dcase_random_scene(seed, duration=60.0, n_events=9, mics=("mic000", "mic001"), moving_fraction=0.4):
    """Seeded scene with static and moving events on a 0.1 s grid (Scene.add_event rounds nothing, but the metadata
    function requires starts/ends that land on the frame grid after round(., 1)), repeated files and ties."""
    rng = np.random.default_rng(seed)
    files = [f"file{k}.wav" for k in range(max(2, n_events - 2))]
    evs = []
    for _ in range(n_events):
        dur = float(rng.uniform(0.3, 10.0))
        start = float(rng.uniform(-0.5, duration - 0.2))  # some start before 0 / end after the scene
        n_em = int(rng.integers(2, 40)) if rng.random() < moving_fraction else 1
        ems = []
        for _e in range(n_em):
            ems.append(DcaseEmitter({m: np.array([[rng.uniform(-180, 180), rng.uniform(-90, 90), rng.uniform(0.2, 9.0)]])
                                     for m in mics}))
        fn = files[int(rng.integers(0, len(files)))]
        evs.append(DcaseEvent(start, dur, int(rng.integers(0, 13)), fn, ems))
    # exact half-way values exercise round-half-even
    evs.append(DcaseEvent(1.0, 0.5, 3, "half.wav", [DcaseEmitter({m: np.array([[2.5, -3.5, 0.125]]) for m in mics})]))
    return DcaseScene(duration, list(mics), evs)


DCASE_RANDOM_SEEDS = [101, 102, 103, 104, 105, 106]
