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
