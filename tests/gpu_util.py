"""Shared helpers for the -m gpu tests (all compute goes through the C-ABI library)."""
import numpy as np

import cases
from audiblelight_b200.renderer import EventJob, SceneJob, moving_frames, event_slice, scene_samples

TOL = 1e-5  # BASELINE.json north_star: max-abs error <= 1e-5 of full scale (1.0), fp32


def event_job(spec, audio, irs, **kw):
    n = irs.shape[1]
    job = EventJob(audio=np.ascontiguousarray(audio, dtype=np.float32),
                   irs=np.ascontiguousarray(irs, dtype=np.float32) if n > 0 else None,
                   n_channels=irs.shape[0], snr=spec["snr"], ref_db=spec["ref_db"], **kw)
    if n > 1:
        job.ir_frames, job.n_frames = moving_frames(len(audio) / float(spec["sr"]), float(spec["sr"]), n, len(audio))
    if spec.get("ref_ir_channel") is not None and spec.get("direct_path_time_ms") is not None:
        low, high = spec["direct_path_time_ms"]
        job.dry = (spec["ref_ir_channel"], int(low * float(spec["sr"]) / 1000), int(high * float(spec["sr"]) / 1000))
    return job


def scene_jobs(spec):
    evs_in, ambs = cases.scene_inputs(spec)
    total = scene_samples(spec["duration"], spec["sr"])
    jobs = []
    for e, (audio, irs) in zip(spec["events"], evs_in):
        es = dict(sr=spec["sr"], snr=e["snr"], ref_db=spec["ref_db"], ref_ir_channel=e.get("ref_ir_channel"),
                  direct_path_time_ms=e.get("direct_path_time_ms"))
        j = event_job(es, audio, irs)
        dur = len(audio) / float(spec["sr"])
        j.scene = 0
        j.scene_start, j.scene_end = event_slice(float(e["start"]), float(e["start"]) + dur, spec["sr"], total)
        jobs.append(j)
    scene = SceneJob(n_channels=spec["c"], n_samples=total,
                     ambience=[np.ascontiguousarray(a, dtype=np.float32) for a in ambs],
                     ambience_ref_db=list(spec["ambience_ref_db"]))
    return jobs, scene
