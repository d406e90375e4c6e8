"""k_small_rir vs the general partitioned pipeline on the renders it was built for (device-resident inputs):
  * dry / direct-path sub-events (compute_dry_audio): 1 704-tap window of a 1 s RIR, 5 s events;
  * short static RIRs (512 taps, 4 capsules), 5 s events.
    python tools/small_rir_bench.py > profiles/r02_small_rir.txt
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiblelight_b200.renderer import EventJob, Renderer  # noqa: E402


def run(rnd, jobs, reps=20):
    packed = rnd.pack(jobs, [])
    for _ in range(3):
        rnd.run(packed)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        rnd.run(packed)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, rnd.profile()["kernel_launches"]


def main():
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    n_ev, lx = 256, 120000
    decay = torch.exp(-torch.arange(24000, device=dev) / 4000.0)
    xs = [torch.randn(lx, device=dev, generator=g) for _ in range(n_ev)]
    h_long = [torch.randn((4, 1, 24000), device=dev, generator=g) * decay for _ in range(n_ev)]
    h_short = [torch.randn((4, 1, 512), device=dev, generator=g) for _ in range(n_ev)]
    print(f"# {n_ev} events of {lx / 24000:.0f} s, B200, device-resident, ms per alr_render call (mean of 20)")
    for name, mk in (
        ("static 1 s RIR + dry window (6 / 65 ms)", lambda i: EventJob(audio=xs[i], irs=h_long[i], n_channels=4, snr=10.0, dry=(0, 144, 1560))),
        ("static 1 s RIR, no dry audio (reference point)", lambda i: EventJob(audio=xs[i], irs=h_long[i], n_channels=4, snr=10.0)),
        ("static 512-tap RIR, 4 capsules", lambda i: EventJob(audio=xs[i], irs=h_short[i], n_channels=4, snr=10.0)),
    ):
        res = []
        for small in (1, 0):
            r = Renderer(0, small_rir=small)
            ms, launches = run(r, [mk(i) for i in range(n_ev)])
            res.append((ms, launches))
            r.close()
        print(f"{name:50s} k_small_rir {res[0][0]:7.3f} ms ({res[0][1]} launches)   general pipeline {res[1][0]:7.3f} ms ({res[1][1]} launches)")


if __name__ == "__main__":
    main()
