"""Where does the host time of the drop-in's batch entry go?  (render_scenes on duck-typed Scene objects with float64 RIRs)
    python tools/objects_mode_profile.py [scenes]
"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiblelight_b200 import synthesize as syn, workload as wl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scenes = [wl.SynScene(i) for i in range(n)]
rnd = syn.get_renderer(0)
syn.render_scenes(scenes, renderer=rnd, pinned=True)
t0 = time.perf_counter()
syn.render_scenes(scenes, renderer=rnd, pinned=True)
print(f"{n} scenes: {time.perf_counter() - t0:.3f} s  ->  {n * 60 / (time.perf_counter() - t0):.0f} scene-s/s; GPU part {rnd.profile()['ms_total']:.1f} ms")
pr = cProfile.Profile()
pr.enable()
syn.render_scenes(scenes, renderer=rnd, pinned=True)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
