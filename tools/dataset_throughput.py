"""End-to-end throughput of the batch dataset driver on synthetic C3-style scenes (60 s, 6 static + 3 moving events):
host objects -> audiblelight_b200.dataset.generate_scenes -> WAV (PCM_16) + JSON + DCASE CSV files on disk.

    python tools/dataset_throughput.py [scenes] [batch] [outdir]

The Scene / Event / Ambience objects are minimal stand-ins with the attributes the drop-in reads from the reference's
classes (core.py / event.py / ambience.py); RIRs are float64 (C, N, Lh) arrays as the reference's backends deliver them,
so the float64 -> float32 conversion the host layer has to do is part of the measured time.
"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiblelight_b200 import dataset, workload as wl  # noqa: E402


Scene = wl.SynScene


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    out = sys.argv[3] if len(sys.argv) > 3 else tempfile.mkdtemp(prefix="alr_dataset_")
    os.makedirs(out, exist_ok=True)
    t0 = time.perf_counter()
    scenes = [Scene(i) for i in range(n)]
    t_build = time.perf_counter() - t0
    dataset.generate_scenes(scenes[:3 * batch], out, batch_scenes=batch)  # warm-up (both contexts, buffers, pinned pools)
    for pipeline in (False, True):
        t0 = time.perf_counter()
        written = dataset.generate_scenes(scenes, out, batch_scenes=batch, pipeline=pipeline)
        dt = time.perf_counter() - t0
        nbytes = sum(os.path.getsize(p) for w in written for k in w for p in w[k])
        print(f"pipeline={pipeline}: {n} scenes ({n * 60} scene-seconds) in {dt:.2f} s -> {n * 60 / dt:.0f} scene-s/s, "
              f"{nbytes / 1e6:.0f} MB written to {out}")
    print(f"(building the {n} synthetic scene objects took {t_build:.1f} s; not part of the timed region)")


if __name__ == "__main__":
    main()
