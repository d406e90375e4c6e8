#!/usr/bin/env python
"""bench.py — scene-seconds rendered per second on the synthesis hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scenes-per-gpu S]

Workload (config.workload): configs[4] of BASELINE.json — a batch of one-minute C3-style SELD scenes with moving
events (60 s @ 24 kHz, 4 channels, 1 s RIRs, 6 static + 3 moving events of 2-10 s, one RIR per 100 ms, Gaussian
ambience), scene-sharded: every GPU renders `--scenes-per-gpu` scenes per step (128 by default = 1024 scenes on
8 GPUs), weak scaling, no collective on the data path. A step = one pass of the whole hot path (RIR spectra,
cross-fade, partitioned convolution, gains, mixdown) over the GPU's batch.

Printed JSON (rank 0): `value` = whole-job scene-seconds/s with inputs resident in HBM; `e2e` = the same metric
through the C-ABI with HOST (pinned) buffers, host<->device copies inside the timed region; `roofline` for the
dominant kernel from CUDA events recorded on the render stream inside the timed region (plus the fp32 side of the
roofline: the path sits at the fp32 ridge); `cpu_baseline` = the UNMODIFIED reference functions
(baseline/_ref, see baseline/reference_arm.py) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "scene_seconds_per_second"
UNIT = "scene-seconds/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes-per-gpu", type=int, default=None,
                    help="scenes rendered per step and GPU; default by workload: c5 / c2 128, c1 1024 (10 s scenes: fewer are "
                         "launch-bound), c4 16 (em64 scenes: 64 channels, 2 s RIRs at 48 kHz)")
    ap.add_argument("--e2e-scenes", type=int, default=64, help="scenes per e2e step and GPU (host buffers); 16 / 32 / 64 give 27.1 / 27.8 / 28.4 K")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="c5", choices=["c5", "c2", "c1", "c4"])
    ap.add_argument("--cpu-workers", type=int, default=None)
    ap.add_argument("--workspace-mb", type=int, default=None, help="override the renderer's workspace limit (experiments)")
    args = ap.parse_args()
    if args.scenes_per_gpu is None:
        args.scenes_per_gpu = {"c1": 1024, "c4": 16}.get(args.workload, 128)
    return args


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _loop(self):
        # NVML (nvidia_ml_py) answers in well under a millisecond, so even a 0.4 s timed region gets dozens of samples;
        # nvidia-smi (one process per sample, ~50-100 ms) is the fallback.
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            while not self._stop.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = int(get_reasons(h))
                row = [str(self.idx), str(sm), str(mx), "", hex(r)]
                row += ["Active" if r & bits[n] else "Not Active"
                        for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")]
                self.rows.append(row)
                self._stop.wait(0.005)
            return
        except Exception:
            pass
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.idx)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def scene_spec(workload: str, idx: int):
    from audiblelight_b200 import workload as wl
    if workload == "c5":  # C3-style scene + one linear augmentation per event
        return wl.c3_scene_spec(idx, augment=True)
    return {"c2": wl.c2_scene_spec, "c1": wl.c1_scene_spec, "c4": wl.c4_scene_spec}[workload](idx)


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation (unmodified functions from baseline/_ref, driven by
    baseline/reference_arm.py) on the host cores, same metric / config; rank 0 only. Each step is one bounded sample:
    one scene per core. Steps stop early when the arm's time budget (~4 minutes) is used up; `steps` is what ran."""
    if rank != 0:
        return
    from baseline import reference_arm
    vals, last = [], None
    spent = 0.0
    budget = float(os.environ.get("ALR_REFERENCE_BUDGET_S", "240"))
    for i in range(max(1, args.steps)):
        last = reference_arm.run(n_workers=args.cpu_workers, workload=args.workload, first_scene=64 * i)
        vals.append(last["value"])
        spent += last["wall_s"]
        if spent + last["wall_s"] > budget:
            break
    value = sum(vals) / len(vals)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": len(vals), "steps_requested": args.steps, "warmup": 0, "ms_per_step": 1000.0 * last["wall_s"],
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": last["sample"],
                         "per_core": last["per_core"], "mean_scene_cpu_s": last["mean_scene_cpu_s"],
                         "values_per_step": vals},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def partition_size():
    """Partition length P the library is compiled with, read from the source (the reference arm must not load the
    CUDA library just to fill in `config`)."""
    import re
    try:
        src = open(os.path.join(ROOT, "audiblelight_b200", "csrc", "alr_fft.cuh")).read()
        return int(re.search(r"#define\s+ALR_P\s+(\d+)", src).group(1))
    except Exception:
        return None


def config_dict(args, world):
    names = {"c5": "configs[4]: batch of one-minute C3-style SELD scenes with moving events (60 s @ 24 kHz, 4 ch, "
                   "1 s RIRs, 6 static + 3 moving events, 10 RIR/s, one linear augmentation + peak normalisation per "
                   "event, Gaussian ambience), scene-sharded",
             "c2": "configs[1]: 60 s @ 24 kHz, 4 ch, 9 static events + ambience",
             "c1": "configs[0]: one static 10 s event, 4-ch 1 s RIR",
             "c4": "configs[3]: em64 64 ch, 48 kHz, 2 s RIRs, 5 static events"}
    return {"workload": names[args.workload], "scenes_per_gpu": args.scenes_per_gpu,
            "scenes_total": args.scenes_per_gpu * world, "parallelism": f"scene-sharded x{world}, no collective",
            "cache": "inputs per step (>= 18 GB per GPU at 128 scenes) are far larger than the 126 MB L2",
            "partition": partition_size()}


FP32_PEAK_TFLOPS = 69.8  # measured: plain FFMA, tools/micro/ffma2_bench.cu on B200 @ 1965 MHz (profiles/r01_ffma2.txt)


def event_flops(job, P):
    """fp32 flops of one event in the partitioned closed form the kernels implement: RIR partition
    FFTs, source block FFTs, spectral multiply-accumulates (8 flops per complex MAC) and the inverse FFTs, with the
    planner's own partition plan (alr_debug_plan; host only). A P-point complex FFT counts 5 P log2 P."""
    import math
    from audiblelight_b200.renderer import debug_plan
    if job.irs is None:
        return 0.0
    C, N = int(job.n_channels), int(job.irs.shape[1])
    plan = debug_plan(job)
    K, B = plan["K"], plan["B_valid"]
    f_fft = 5.0 * P * math.log2(P)
    xb0, xnb = plan["irs"][:, 0], plan["irs"][:, 1]
    pairs = 0
    for l in range(N):
        for k in range(K):
            # source blocks j of IR l with output block xb0 + j + k inside the valid range
            pairs += max(0, min(int(xnb[l]), B - int(xb0[l]) - k))
    return (N * C * K + int(xnb.sum()) + B * C) * f_fft + 8.0 * P * C * pairs


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from audiblelight_b200 import workload as wl
    from audiblelight_b200.renderer import Renderer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the renderer)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line (no NCCL version banner)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident inputs: scene i of the job goes to rank i % world ----------------------------------------
    S = args.scenes_per_gpu
    from audiblelight_b200 import sharding
    specs = [scene_spec(args.workload, i) for i in sharding.shard_scene_indices(S, rank, world)]
    jobs, scenes = [], []
    for si, sp in enumerate(specs):
        arrays, amb = wl.device_scene_arrays(sp, dev)
        j, sj = wl.scene_jobs(sp, arrays, amb, si)
        jobs += j
        scenes.append(sj)
    b_alg = sum(wl.algorithmic_bytes(sp) for sp in specs)
    b_ir = sum(wl.ir_bytes(sp) for sp in specs)
    scene_seconds = sum(sp.duration for sp in specs)
    rnd = Renderer(local_rank, profiling=True,
                   workspace_limit=None if args.workspace_mb is None else args.workspace_mb << 20)
    packed = rnd.pack(jobs, scenes)
    stream = torch.cuda.current_stream().cuda_stream
    flops_step = sum(event_flops(j, partition_size() or 2048) for j in jobs) if rank == 0 else 0.0

    for _ in range(max(args.warmup, 3)):
        rnd.run(packed, stream)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_acc = {}
    ev0.record()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        rnd.run(packed, stream)
        p = rnd.profile()
        for k, v in p.items():
            prof_acc[k] = prof_acc.get(k, 0.0) + v
    ev1.record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    barrier()
    ms_per_step = elapsed_ms / args.steps
    value = scene_seconds * world / (ms_per_step / 1000.0)

    # ---- end-to-end through the C-ABI with pinned HOST buffers ------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        Se = min(args.e2e_scenes, S)
        h_jobs, h_scenes = [], []
        for si in range(Se):
            sp = specs[si]
            arrays, amb = wl.device_scene_arrays(sp, dev)

            def pin(t):
                return t.cpu().pin_memory().numpy()

            arrays = [(pin(x), pin(h)) for x, h in arrays]
            amb = pin(amb) if amb is not None else None
            j, sj = wl.scene_jobs(sp, arrays, amb, si)
            for e in j:
                e.spatial = torch.empty((e.n_channels, e.audio.shape[0]), dtype=torch.float32).pin_memory().numpy()
            sj.mix = torch.empty((sj.n_channels, sj.n_samples), dtype=torch.float32).pin_memory().numpy()
            h_jobs += j
            h_scenes.append(sj)
        e2e_ss = sum(specs[i].duration for i in range(Se)) * world

        def time_calls(fn, steps):
            for _ in range(2):
                fn()
            barrier()
            t0 = time.perf_counter()
            last = None
            for _ in range(steps):
                last = fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return dt, last

        def one_call():
            rnd.render(h_jobs, h_scenes, stream)  # pack + plan + H2D + kernels + D2H, synchronous
            return rnd.profile()

        def entry(dt, p, note):
            return {"value": e2e_ss * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(p["h2d_bytes"]),
                    "d2h_bytes_per_step": int(p["d2h_bytes"]),
                    "pcie_gbs": (p["h2d_bytes"] + p["d2h_bytes"]) * args.e2e_steps / dt / 1e9, "note": note}

        # (1) every output of the C-ABI call comes back: all event.spatial_audio arrays and the mixes
        dt, p = time_calls(one_call, args.e2e_steps)
        all_outputs = entry(dt, p, "every event.spatial_audio (C, Lx) and every scene mix downloaded")
        # (2) the step's RESULT is the scene mix (what Scene.generate writes, core.py:1840-1847): events are rendered and
        # mixed on the device, scene.audio (float32) is the download. This is the headline e2e.
        for e in h_jobs:
            e.keep_spatial = False
        dt, p = time_calls(one_call, args.e2e_steps)
        e2e = entry(dt, p, "inputs (dry audio, RIRs, ambience) uploaded from pinned host memory, every scene mix (C, T) float32 "
                           "downloaded; event audio is rendered and mixed on the device")
        e2e.update({"scenes_per_step_per_gpu": Se, "steps": args.e2e_steps,
                    "timing": "wall clock around Renderer.render (descriptor packing, planning, pinned H2D, kernels, D2H)",
                    "all_outputs": all_outputs})
        # (3) as dataset generation runs it (audiblelight_b200.dataset): only the 16-bit PCM of each mix is copied back
        for sj in h_scenes:
            sj.pcm16 = torch.empty((sj.n_samples, sj.n_channels), dtype=torch.int16).pin_memory().numpy()
            sj.keep_mix = False
        dt, p = time_calls(one_call, args.e2e_steps)
        e2e["dataset_mode"] = entry(dt, p, "mix-only: PCM_16 (T, C) of every scene mix is the only download")
        # (3b) informational: as (3), and the Gaussian ambience of every scene is DRAWN ON THE DEVICE inside the call (f3,
        # alr_scene.ambience_seed) instead of being uploaded: 22 % fewer host->device bytes. The workload's ambience is
        # Gaussian noise, which the reference generates inside this very call (Ambience.load_ambience, synthesize.py:350).
        saved_amb = [(sj.ambience, sj.ambience_seed) for sj in h_scenes]
        if all(len(sj.ambience) == 1 for sj in h_scenes):
            for k, sj in enumerate(h_scenes):
                sj.ambience, sj.ambience_seed = [None], [4242 + k]
            dt, p = time_calls(one_call, args.e2e_steps)
            e2e["dataset_mode_device_ambience"] = entry(dt, p, "PCM_16 mixes down, Gaussian ambience generated on the device "
                                                               "(not uploaded)")
            for sj, (a_, s_) in zip(h_scenes, saved_amb):
                sj.ambience, sj.ambience_seed = a_, s_
        # (4) two contexts, two host threads, alternating steps: the upload of step i + 1 overlaps the kernels and the
        # download of step i (how audiblelight_b200.dataset drives batches). Same per-step work and copies as (2).
        try:
            import copy
            import threading as _th
            for sj in h_scenes:
                sj.pcm16, sj.keep_mix = None, True
            rnd_b = Renderer(local_rank)
            jobs_b, scenes_b = copy.copy(h_jobs), []
            jobs_b = [copy.copy(e) for e in h_jobs]
            for sj in h_scenes:
                sb = copy.copy(sj)
                sb.mix = torch.empty((sj.n_channels, sj.n_samples), dtype=torch.float32).pin_memory().numpy()
                scenes_b.append(sb)
            streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
            work = [(rnd, h_jobs, h_scenes, streams[0]), (rnd_b, jobs_b, scenes_b, streams[1])]

            def loop(k, n):
                r_, j_, s_, st_ = work[k]
                for _ in range(n):
                    r_.render(j_, s_, st_.cuda_stream)

            def both(n):
                ts = [_th.Thread(target=loop, args=(k, n)) for k in range(2)]
                for t_ in ts:
                    t_.start()
                for t_ in ts:
                    t_.join()

            both(1)
            barrier()
            t0 = time.perf_counter()
            both(args.e2e_steps)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            e2e["two_contexts"] = {"value": e2e_ss * 2 * args.e2e_steps / dt, "unit": UNIT,
                                   "note": "two renderer contexts driven by two host threads, steps alternate; per-step "
                                           "copies as the headline e2e"}
            rnd_b.close()
            del jobs_b, scenes_b
        except Exception as exc:
            e2e["two_contexts"] = {"error": repr(exc)}
        del h_jobs, h_scenes
        # What a user of install() gets: the reference's own objects in, results on the objects out. float64 RIRs in
        # pageable numpy memory (as the reference's backends deliver them), converted to float32 by the host layer,
        # every event's spatial audio and every mix copied back and stored on the Event / Scene objects
        # (audiblelight_b200.synthesize.render_scenes == render_audio_for_all_scene_events + generate_scene_audio_from_events
        # per scene). Informational; building the synthetic objects is not timed.
        try:
            from audiblelight_b200 import synthesize as syn
            So = min(8, Se)
            objs = [wl.SynScene(specs[i].index) for i in range(So)]
            syn.render_scenes(objs, renderer=rnd, pinned=True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                syn.render_scenes(objs, renderer=rnd, pinned=True)
                p = rnd.profile()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            e2e["objects_mode"] = {"value": sum(o.duration for o in objs) * world * 2 / dt, "unit": UNIT,
                                   "scenes_per_step_per_gpu": So, "h2d_bytes_per_step": int(p["h2d_bytes"]),
                                   "d2h_bytes_per_step": int(p["d2h_bytes"]),
                                   "note": "duck-typed Scene objects with float64 pageable RIRs -> render_scenes -> "
                                           "event.spatial_audio (float64) + scene.audio on the objects"}
            del objs
        except Exception as exc:  # informational leg: never lose the headline over it
            e2e["objects_mode"] = {"error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA events recorded on the render stream between the launches) -------------
    peak, peak_src = load_peaks()
    kern_ms = {k[3:]: prof_acc[k] / args.steps for k in ("ms_ir_fft", "ms_x_fft", "ms_cmac", "ms_cmac_static", "ms_ifft",
                                                         "ms_mix", "ms_other", "ms_fused")}
    dom = max(kern_ms, key=kern_ms.get)
    # algorithmic bytes each kernel class is responsible for (DESIGN.md "Algorithmic bytes"):
    out_bytes = sum(4 * sp.channels * e.n_audio for sp in specs for e in sp.events)
    mix_bytes = sum(4 * sp.channels * round(sp.duration * sp.sr) * (2 if sp.ambience else 1) for sp in specs)
    x_bytes = sum(4 * e.n_audio for sp in specs for e in sp.events)
    b_ir_static = sum(4 * sp.channels * e.n_irs * sp.n_ir_samples for sp in specs for e in sp.events if e.n_irs == 1)
    alg = {"ir_fft": b_ir, "x_fft": x_bytes, "cmac": b_ir - b_ir_static, "cmac_static": b_ir_static, "ifft": out_bytes,
           "mix": out_bytes + mix_bytes, "other": 0, "fused": b_ir - b_ir_static}
    n_launch_dom = {"ir_fft": 1, "x_fft": 1, "cmac": 1, "cmac_static": 1, "ifft": 1, "fused": 1}.get(dom, None)
    chunks = max(1, int(prof_acc["n_chunks"] / args.steps))
    launches_dom = chunks if n_launch_dom else None
    dom_ms = kern_ms[dom]
    achieved = alg[dom] / (dom_ms / 1000.0) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        name = {"ir_fft": "k_ir_fft", "cmac": "k_cmac", "fused": "k_mov_fused"}.get(dom)
        if name in tr and args.workload == "c5" and args.scenes_per_gpu == 128:
            traffic = tr[name]["dram_read_bytes"] + tr[name]["dram_write_bytes"]
    except Exception:
        pass
    per_launch = (lambda v: v / launches_dom) if launches_dom else (lambda v: v)
    roofline = {
        "bound": "hbm", "kernel": {"ir_fft": "k_ir_fft", "x_fft": "k_x_fft", "cmac": "k_cmac", "cmac_static": "k_cmac_static",
                                   "ifft": "k_ifft_ola", "fused": "k_mov_fused",
                                   "mix": "k_mix+k_apply_gain+k_amb_*", "other": "misc"}[dom],
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "traffic_note": "DRAM bytes of one launch under ncu (profiles/ncu_traffic.json); compare with algorithmic_bytes_per_launch",
        "algorithmic_bytes_per_launch": per_launch(alg[dom]), "kernel_us_per_launch": per_launch(dom_ms * 1000.0),
        "peak_source": peak_src, "algorithmic_bytes_per_step": alg[dom], "kernel_ms_per_step": dom_ms,
        "launches_per_step": launches_dom,
        "kernel_share_of_step": dom_ms / ms_per_step,
        "kernel_ms": kern_ms,
        "pipeline": {"algorithmic_bytes_per_step": b_alg, "achieved": b_alg / (ms_per_step / 1000.0) / 1e9,
                     "frac": b_alg / (ms_per_step / 1000.0) / 1e9 / peak,
                     "note": "whole hot path: B_alg of SURVEY.md 8(d) / step time"},
    }
    # the fp32 side (report max(t_HBM, t_fp32) as the honest bound; no tensor cores on this path)
    t_hbm_ms = b_alg / (peak * 1e9) * 1e3
    t_fp32_ms = flops_step / (FP32_PEAK_TFLOPS * 1e12) * 1e3
    roofline["fp32"] = {
        "flops_per_step": flops_step, "achieved": flops_step / (ms_per_step / 1000.0) / 1e12, "peak": FP32_PEAK_TFLOPS,
        "unit": "TFLOP/s", "frac": flops_step / (ms_per_step / 1000.0) / 1e12 / FP32_PEAK_TFLOPS,
        "peak_source": "measured FFMA rate, tools/micro/ffma2_bench.cu (profiles/r01_ffma2.txt)",
        "model": "partitioned closed form: (N C K + sum xnb + B C) 5 P log2 P + 8 P C (active X.H pairs), planner's plan"}
    roofline["bound_times_ms"] = {"hbm": t_hbm_ms, "fp32": t_fp32_ms}
    roofline["pipeline_bound"] = "fp32" if t_fp32_ms > t_hbm_ms else "hbm"
    roofline["pipeline"]["frac_of_bound"] = max(t_hbm_ms, t_fp32_ms) / ms_per_step
    cpu = None
    if not args.no_cpu_baseline:
        from baseline import reference_arm
        c = reference_arm.run(n_workers=args.cpu_workers, workload=args.workload, select="cheapest")
        cpu = {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": c["kind"], "sample": c["sample"],
               "per_core": c["per_core"], "mean_scene_cpu_s": c["mean_scene_cpu_s"], "wall_s": c["wall_s"]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded torch CUDA generator; structure from numpy "
        "default_rng(1000+scene))",
        "config": config_dict(args, world),
        "clocks": sampler.summary(),
        "e2e": e2e,
        "gpu_launches": int(prof_acc["kernel_launches"]),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "host_ms_per_step_wall": 1000.0 * t_wall / args.steps,
        "workspace_bytes": int(prof_acc["workspace_bytes"] / args.steps),
        "chunks_per_step": chunks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
