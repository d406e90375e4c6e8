"""Two render calls in flight (two contexts, two host threads) vs one: does overlapping the fill/drain of consecutive
host-buffer calls raise end-to-end throughput?   python tools/e2e_pipelined_probe.py [scenes per call] [calls]"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from audiblelight_b200 import workload as wl  # noqa: E402
from audiblelight_b200.renderer import Renderer  # noqa: E402

dev = torch.device("cuda", 0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
K = int(sys.argv[2]) if len(sys.argv) > 2 else 6
mix_only = len(sys.argv) > 3 and sys.argv[3] == "mix"


def make_batch(first):
    jobs, scenes = [], []
    for si in range(first, first + S):
        sp = wl.c3_scene_spec(si, augment=True)
        arrays, amb = wl.device_scene_arrays(sp, dev)
        pin = lambda t: t.cpu().pin_memory().numpy()  # noqa: E731
        arrays = [(pin(x), pin(h)) for x, h in arrays]
        j, sj = wl.scene_jobs(sp, arrays, pin(amb), si - first)
        for e in j:
            if mix_only:
                e.keep_spatial = False
            else:
                e.spatial = torch.empty((e.n_channels, e.audio.shape[0]), dtype=torch.float32).pin_memory().numpy()
        if mix_only:
            sj.pcm16 = torch.empty((sj.n_samples, sj.n_channels), dtype=torch.int16).pin_memory().numpy()
            sj.keep_mix = False
        else:
            sj.mix = torch.empty((sj.n_channels, sj.n_samples), dtype=torch.float32).pin_memory().numpy()
        jobs += j
        scenes.append(sj)
    return jobs, scenes


batches = [make_batch(0), make_batch(S)]
rnds = [Renderer(0), Renderer(0)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for r, b, s in zip(rnds, batches, streams):
    for _ in range(2):
        r.render(*b, stream=s.cuda_stream)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(K):
    rnds[0].render(*batches[0], stream=streams[0].cuda_stream)
torch.cuda.synchronize()
t_serial = (time.perf_counter() - t0) / K
print(f"serial: {1e3 * t_serial:.1f} ms per {S}-scene call -> {S * 60 / t_serial:.0f} scene-s/s")


def worker(w):
    torch.cuda.set_device(0)
    for _ in range(K):
        rnds[w].render(*batches[w], stream=streams[w].cuda_stream)


torch.cuda.synchronize()
t0 = time.perf_counter()
ths = [threading.Thread(target=worker, args=(w,)) for w in range(2)]
[t.start() for t in ths]
[t.join() for t in ths]
torch.cuda.synchronize()
t_pipe = (time.perf_counter() - t0) / (2 * K)
print(f"2 in flight: {1e3 * t_pipe:.1f} ms per {S}-scene call -> {S * 60 / t_pipe:.0f} scene-s/s")
