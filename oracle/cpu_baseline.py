"""CPU baseline for bench.py: the oracle port of the reference's own algorithm, timed on the host cores.

TEST/BENCH INFRASTRUCTURE (only bench.py's `cpu_baseline` / `--impl reference` legs import this).

The reference's CPU path needs ~10 CPU-seconds per audio-second of a moving event (the
frame-serial STFT-domain loop of perform_time_variant_convolution, synthesize.py:184-252), i.e. 3-5 minutes
per C3 scene and core, so a full scene cannot be timed inside a benchmark run. The bounded sample is:

  * every worker (one per host core, one scene each) renders ALL static events of its scene and the complete
    mixdown with ambience in full (oracle `render_event` / `mix_scene`, the literal port), and
  * renders the first `moving_seconds` of ONE moving event of its scene with the literal STFT-domain port
    (`time_variant_convolution`), whose measured time is extrapolated to the scene's moving events with the
    loop's exact work model W(n) = sum_{i<n} min(i+1, n_ir_frames) (contraction depth per output frame).

`moving_seconds` defaults to 3 s (31 RIRs at 10 per second) on purpose: from ~25 RIRs on, fewer than half of the
RIRs are active inside the 1-s convolution window, which switches the reference (and this port) into its
per-frame boolean sub-select copy (synthesize.py:231-240) — the regime every 2.4-10 s event of the workload
runs in and the reason it costs ~10 CPU-seconds per audio-second. A shorter sample would flatter the CPU.

The result is labelled as extrapolated in `sample`.
"""
from __future__ import annotations

import math
import multiprocessing as mp
import os
import time

import numpy as np


def _work(n_frames: int, n_ir_frames: int) -> float:
    n = min(n_frames, n_ir_frames)
    return n * (n + 1) / 2.0 + max(0, n_frames - n_ir_frames) * float(n_ir_frames)


def _worker(args):
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    scene_idx, moving_seconds = args
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from audiblelight_b200 import workload as wl
    from oracle import synth_oracle as orc
    from oracle import augment_oracle as ao
    spec = wl.c3_scene_spec(scene_idx, augment=True)
    # inputs (float64 IRs as the reference's backends deliver them); generation is not timed
    rng = np.random.default_rng(5000 + scene_idx)
    decay = np.exp(-np.arange(spec.n_ir_samples) / (spec.n_ir_samples / 6.0))
    T = round(spec.duration * spec.sr)
    t_static = 0.0
    spatial, starts, ends = [], [], []
    moving = []
    for e in spec.events:
        x = rng.standard_normal(e.n_audio).astype(np.float32)
        x = (x / np.max(np.abs(x) + np.finfo(np.float32).tiny)).astype(np.float32)
        if e.aug is not None:  # Event.load_audio: augmentation, then peak normalisation (event.py:530-536)
            t0 = time.perf_counter()
            x = ao.peak_normalize(ao.biquad(x, *wl.aug_coeffs(e.aug, float(spec.sr)))).astype(np.float32)
            t_static += time.perf_counter() - t0
        dur = e.n_audio / float(spec.sr)
        if e.n_irs == 1:
            h = rng.standard_normal((spec.channels, 1, spec.n_ir_samples)) * decay
            t0 = time.perf_counter()
            res = orc.render_event(x, h, e.snr, spec.ref_db, is_moving=False)
            t_static += time.perf_counter() - t0
            spatial.append(res.spatial)
        else:
            moving.append((e, x))
            spatial.append(np.zeros((spec.channels, e.n_audio)))  # placeholder with the right shape for the mix
        starts.append(e.start)
        ends.append(e.start + dur)
    amb = rng.standard_normal((spec.channels, T))
    amb = amb / np.max(np.abs(amb), axis=1, keepdims=True)
    t0 = time.perf_counter()
    orc.mix_scene(spec.duration, spec.sr, spatial, starts, ends, [(amb, spec.ref_db)])
    t_mix = time.perf_counter() - t0
    # one truncated moving event, literal STFT-domain path
    t_moving_full = 0.0
    t_sub = 0.0
    if moving:
        e, x = moving[0]
        n_sub = min(e.n_audio, int(round(moving_seconds * spec.sr)))
        n_irs_sub = max(2, int(round(10.0 * n_sub / spec.sr)) + 1)
        h = rng.standard_normal((spec.channels, n_irs_sub, spec.n_ir_samples)) * decay
        t0 = time.perf_counter()
        orc.render_event(x[:n_sub], h, e.snr, spec.ref_db, is_moving=True, duration=n_sub / float(spec.sr),
                         sample_rate=float(spec.sr), literal=True)
        t_sub = time.perf_counter() - t0
        n_ir_frames = orc.n_stft_frames(spec.n_ir_samples)
        w_sub = _work(orc.n_stft_frames(n_sub), n_ir_frames)
        for ev, _ in moving:
            t_moving_full += t_sub * _work(orc.n_stft_frames(ev.n_audio), n_ir_frames) / w_sub
    return dict(scene=scene_idx, t_static=t_static, t_mix=t_mix, t_sub=t_sub, t_moving_full=t_moving_full,
                duration=spec.duration)


def run(n_workers=None, moving_seconds: float = 3.0, first_scene: int = 0):
    """Returns dict(value=scene-seconds/s over all workers, cores, sample, per_core, wall_s)."""
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if n_workers is None:
        n_workers = max(1, min(avail, 64))
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(n_workers) as pool:
        res = pool.map(_worker, [(first_scene + i, moving_seconds) for i in range(n_workers)])
    wall = time.perf_counter() - t0
    per_scene = [r["t_static"] + r["t_mix"] + r["t_moving_full"] for r in res]
    value = sum(r["duration"] / t for r, t in zip(res, per_scene))  # all workers concurrently, one scene each
    return dict(
        value=value, unit="scene-seconds/s", cores=n_workers, kind="port",
        sample=(f"{n_workers} C3-style scenes, one per core: static events + ambience mixdown timed in full; moving "
                f"events: first {moving_seconds:g} s of one event per scene with the literal STFT-domain port, "
                f"extrapolated to the scene's 3 moving events with the loop's work model (EXTRAPOLATED)"),
        per_core=value / n_workers, wall_s=wall,
        mean_scene_cpu_s=float(np.mean(per_scene)),
        mean_static_s=float(np.mean([r["t_static"] for r in res])),
        mean_mix_s=float(np.mean([r["t_mix"] for r in res])),
        mean_moving_s=float(np.mean([r["t_moving_full"] for r in res])),
        host_cores_available=avail,
    )


if __name__ == "__main__":
    import json
    print(json.dumps(run(n_workers=None)))
