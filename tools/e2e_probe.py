"""Where does the end-to-end time go? Wall time of Renderer.render vs pack vs the library call vs device span."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiblelight_b200 import workload as wl
from audiblelight_b200.renderer import Renderer
dev = torch.device("cuda", 0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rnd = Renderer(0, profiling=False)
jobs, scenes = [], []
for si in range(S):
    sp = wl.c3_scene_spec(si)
    arrays, amb = wl.device_scene_arrays(sp, dev)
    pin = lambda t: t.cpu().pin_memory().numpy()
    arrays = [(pin(x), pin(h)) for x, h in arrays]
    amb = pin(amb)
    j, sj = wl.scene_jobs(sp, arrays, amb, si)
    for e in j:
        e.spatial = torch.empty((e.n_channels, e.audio.shape[0]), dtype=torch.float32).pin_memory().numpy()
    sj.mix = torch.empty((sj.n_channels, sj.n_samples), dtype=torch.float32).pin_memory().numpy()
    jobs += j; scenes.append(sj)
for _ in range(2): rnd.render(jobs, scenes)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    packed = rnd.pack(jobs, scenes); t1 = time.perf_counter()
    rnd.run(packed); t2 = time.perf_counter()
    p = rnd.profile()
    print(f"pack {1e3*(t1-t0):.1f} ms, alr_render wall {1e3*(t2-t1):.1f} ms, device span {p['ms_total']:.1f} ms, "
          f"host plan {p['ms_host_plan']:.1f} ms, chunks {p['n_chunks']}, h2d {p['h2d_bytes']/1e9:.2f} GB d2h {p['d2h_bytes']/1e9:.2f} GB")
