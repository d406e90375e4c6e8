"""bench.py's CPU arm: the UNMODIFIED reference functions on the host cores.

TEST/BENCH INFRASTRUCTURE — imported only by bench.py (`--impl reference` and the `cpu_baseline` leg). Nothing here
touches the renderer: `audiblelight_b200.workload` is imported for the scene SPECS (pure numpy), the CUDA library is
never loaded by this module.

Every worker process (one per host core, single-threaded BLAS/FFT) takes one scene of the benchmark workload and calls
the reference's own `render_event_audio` (synthesize.py:507-608) for each event and `generate_scene_audio_from_events`
(:314-401) for the mixdown, on float64 RIRs as the reference's backends deliver them, through duck-typed Event / Scene
stand-ins (baseline/ref_loader.py; the reference's Event/Scene classes need its whole non-installable dependency set).

What is timed:
  * EVERYTHING of the scene, in full, by default (`moving="all"`): every static event, all three moving events with
    all of their RIRs through the frame-serial STFT-domain loop (perform_time_variant_convolution, :184-252; ~7 CPU-
    seconds per audio-second once an event has more than ~25 RIRs) and the complete mixdown with ambience. One sample =
    one scene per core = 2-3.5 minutes of wall time; nothing is extrapolated.
  * `moving="shortest"` (quick checks only, ALR_REFERENCE_MOVING=shortest) renders ONE whole moving event per scene and
    scales it to the other two with the loop's work model W(n) = sum_{i<n} min(i + 1, n_ir_frames). That model was
    checked against whole events with the unmodified reference (profiles/r02_reference_extrapolation.txt): +7 % and
    +13 % when predicting 9 s and 6 s events from a 3.5 s one, but -60 % from a 2 s event, because below ~25 RIRs the
    reference skips its per-frame boolean sub-select copy (:231-240), which costs ~4x the contraction itself. Round 1
    extrapolated from the first 3 s of one event; the default no longer scales anything.
  * the one linear augmentation per event (C5) runs through scipy.signal.lfilter (oracle/augment_oracle.py):
    pedalboard / librosa, which the reference wraps for it, are not installable offline. < 1 % of the scene time.

When baseline/_ref is missing (reference never installed), `run()` falls back to the oracle port and says so
(`kind: "port"`).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# measured with tools/ref_extrapolation_check.py in the build container (unmodified reference, 1 core):
EXTRAPOLATION_CHECK = ("work model vs whole events, profiles/r02_reference_extrapolation.txt: +7..13 % inside the sub-select "
                       "regime, -60 % from a 2 s event")


def _work(n_frames: int, n_ir_frames: int) -> float:
    n = min(n_frames, n_ir_frames)
    return n * (n + 1) / 2.0 + max(0, n_frames - n_ir_frames) * float(n_ir_frames)


def _n_stft_frames(n: int, hop: int = 128) -> int:
    return 2 * int(np.ceil(n / (2.0 * hop))) + 1


def reference_installed() -> bool:
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    from baseline import install_reference, ref_loader
    if os.path.isdir(os.path.join("/root/reference", "audiblelight")):
        return True
    return ref_loader.reference_available() and install_reference.verify()


def _worker(args):
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    scene_idx, workload, moving_mode = args
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from audiblelight_b200 import workload as wl  # scene specs only (numpy); no CUDA library is loaded
    from baseline import ref_loader
    from oracle import augment_oracle as ao
    syn = ref_loader.load_reference_synthesize()
    spec = {"c5": lambda i: wl.c3_scene_spec(i, augment=True), "c3": wl.c3_scene_spec, "c2": wl.c2_scene_spec,
            "c1": wl.c1_scene_spec, "c4": wl.c4_scene_spec}[workload](scene_idx)
    rng = np.random.default_rng(5000 + scene_idx)
    decay = np.exp(-np.arange(spec.n_ir_samples) / (spec.n_ir_samples / 6.0))
    T = round(spec.duration * spec.sr)
    events, ir_list = [], []
    t_aug = 0.0
    for k, e in enumerate(spec.events):
        x = rng.standard_normal(e.n_audio).astype(np.float32)
        x = (x / np.max(np.abs(x) + np.finfo(np.float32).tiny)).astype(np.float32)
        if e.aug is not None:  # Event.load_audio: augmentation, then peak normalisation (event.py:530-536)
            t0 = time.perf_counter()
            x = ao.peak_normalize(ao.biquad(x, *wl.aug_coeffs(e.aug, float(spec.sr)))).astype(np.float32)
            t_aug += time.perf_counter() - t0
        h = rng.standard_normal((spec.channels, e.n_irs, spec.n_ir_samples)) * decay  # float64, (C, N, Lh)
        ev = ref_loader.RefEvent(x, spec.sr, e.n_irs, e.snr, scene_start=e.start, alias=f"event{k:03d}")
        events.append(ev)
        ir_list.append(h)
    moving = [k for k, e in enumerate(spec.events) if e.n_irs > 1]
    if moving_mode == "all":
        measured = list(moving)
    else:
        measured = sorted(moving, key=lambda k: spec.events[k].n_audio)[:1]
    n_ir_frames = _n_stft_frames(spec.n_ir_samples)
    t_static = t_moving_meas = 0.0
    w_meas = 0.0
    for k, (ev, h) in enumerate(zip(events, ir_list)):
        if k in moving and k not in measured:
            # not rendered in a bounded run: zeros of the right shape so that the mixdown does the same work
            ev.spatial_audio["mic000"] = np.zeros((spec.channels, len(ev.audio)))
            continue
        t0 = time.perf_counter()
        syn.render_event_audio(ev, h, "mic000", ref_db=spec.ref_db)
        dt = time.perf_counter() - t0
        if k in moving:
            t_moving_meas += dt
            w_meas += _work(_n_stft_frames(len(ev.audio)), n_ir_frames)
        else:
            t_static += dt
    amb = None
    if spec.ambience:
        a = rng.standard_normal((spec.channels, T))
        amb = {"amb0": ref_loader.RefAmbience((a / np.max(np.abs(a), axis=1, keepdims=True)).astype(np.float32), spec.ref_db)}
    scene = ref_loader.RefScene(spec.duration, spec.sr, spec.ref_db, events, amb)
    t0 = time.perf_counter()
    syn.generate_scene_audio_from_events(scene)
    t_mix = time.perf_counter() - t0
    w_all = sum(_work(_n_stft_frames(spec.events[k].n_audio), n_ir_frames) for k in moving)
    t_moving = t_moving_meas * (w_all / w_meas) if w_meas > 0 else 0.0
    return dict(scene=scene_idx, duration=spec.duration, t_static=t_static + t_aug, t_mix=t_mix,
                t_moving_measured=t_moving_meas, t_moving=t_moving, scaled=len(measured) < len(moving),
                measured_audio_s=sum(spec.events[k].n_audio for k in measured) / float(spec.sr),
                moving_audio_s=sum(spec.events[k].n_audio for k in moving) / float(spec.sr))


def host_cores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def pick_scenes(workload: str, n: int, first_scene: int, pool: int = 64, key: str = "shortest_event"):
    """Bounded runs take `n` of `pool` consecutive scenes: with key="shortest_event" those whose shortest moving event is
    shortest (the `moving="shortest"` quick mode), with key="least_moving_audio" those with the least moving audio in
    total (whole scenes, nothing scaled — the cheapest scenes for the CPU, so the CPU figure errs on the high side)."""
    if workload not in ("c5", "c3"):
        return [first_scene + i for i in range(n)]
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from audiblelight_b200 import workload as wl
    cand = []
    for i in range(first_scene, first_scene + max(pool, n)):
        mov = [e.n_audio for e in wl.c3_scene_spec(i).events if e.n_irs > 1]
        cand.append((min(mov) if key == "shortest_event" else sum(mov), i))
    return [i for _, i in sorted(cand)[:n]]


def run(n_workers=None, workload: str = "c5", moving: str = None, first_scene: int = 0, select: str = "first"):
    """One bounded sample: `n_workers` scenes, one per core, concurrently. Returns the bench fields."""
    avail = host_cores()
    if moving is None:
        moving = os.environ.get("ALR_REFERENCE_MOVING", "all")
    if n_workers is None:
        n_workers = max(1, min(avail, 64))
    if not reference_installed():
        from oracle import cpu_baseline  # the port: same algorithm, restated (oracle/synth_oracle.py)
        out = cpu_baseline.run(n_workers=n_workers, first_scene=first_scene)
        out["kind"] = "port"
        out["sample"] = "baseline/_ref missing -> oracle port; " + out["sample"]
        return out
    if moving != "all":
        scenes = pick_scenes(workload, n_workers, first_scene)
    elif select == "cheapest":
        scenes = pick_scenes(workload, n_workers, first_scene, key="least_moving_audio")
    else:
        scenes = [first_scene + i for i in range(n_workers)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(n_workers) as pool:
        res = pool.map(_worker, [(i, workload, moving) for i in scenes], chunksize=1)
    wall = time.perf_counter() - t0
    per_scene = [r["t_static"] + r["t_mix"] + r["t_moving"] for r in res]
    value = sum(r["duration"] / t for r, t in zip(res, per_scene))  # all workers concurrently, one scene each
    scaled = any(r["scaled"] for r in res)
    sample = (f"{n_workers} scenes of the workload, one per core, unmodified reference render_event_audio + "
              f"generate_scene_audio_from_events: static events and the mixdown with ambience in full; moving events: ")
    if scaled:
        sample += (f"ONE WHOLE event per scene (the shortest, mean {np.mean([r['measured_audio_s'] for r in res]):.1f} s of "
                   f"{np.mean([r['moving_audio_s'] for r in res]):.1f} s of moving audio per scene), scaled to the scene's "
                   f"other moving events with the loop's work model (SCALED; {EXTRAPOLATION_CHECK})")
    else:
        sample += "all of them, in full (no scaling)"
        if select == "cheapest" and workload in ("c5", "c3"):
            sample += ("; scenes = those with the least moving audio among 64 consecutive ones (bounds the leg's run time; "
                       "biases the CPU figure upwards)")
    return dict(value=value, unit="scene-seconds/s", cores=n_workers, kind="reference", sample=sample,
                per_core=value / n_workers, wall_s=wall, mean_scene_cpu_s=float(np.mean(per_scene)),
                mean_static_s=float(np.mean([r["t_static"] for r in res])),
                mean_mix_s=float(np.mean([r["t_mix"] for r in res])),
                mean_moving_s=float(np.mean([r["t_moving"] for r in res])),
                mean_moving_measured_s=float(np.mean([r["t_moving_measured"] for r in res])),
                host_cores_available=avail, scenes=scenes)


if __name__ == "__main__":
    import json
    print(json.dumps(run(n_workers=int(sys.argv[1]) if len(sys.argv) > 1 else None)))
