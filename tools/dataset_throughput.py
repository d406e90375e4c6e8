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
import types
from collections import OrderedDict

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiblelight_b200 import dataset, workload as wl  # noqa: E402


class Emitter:
    def __init__(self, polar):
        self.coordinates_relative_polar = polar


class Event:
    def __init__(self, alias, audio, sr, n_irs, snr, start, rng):
        self.alias, self.audio, self.sample_rate, self.snr = alias, audio, float(sr), snr
        self.duration = len(audio) / float(sr)
        self.scene_start = round(start, 1)  # the metadata grid is 100 ms
        self.scene_end = self.scene_start + round(self.duration, 1)
        self.is_moving = n_irs > 1
        self.n = n_irs
        self.class_id, self.filename = int(rng.integers(0, 13)), f"{alias}.wav"
        self.emitters = [Emitter({"mic000": np.array([[rng.uniform(-180, 180), rng.uniform(-40, 40), rng.uniform(0.5, 5)]])})
                         for _ in range(n_irs)]
        self.spatial_audio, self._spatial_audio_padded = OrderedDict(), OrderedDict()
        self._spatial_audio_dry, self._spatial_audio_dry_padded = OrderedDict(), OrderedDict()

    def load_audio(self, ignore_cache=False, normalize=True):
        return self.audio

    def __len__(self):
        return self.n


class Ambience:
    def __init__(self, noise, ref_db):
        self.noise, self.ref_db = noise, ref_db

    def load_ambience(self, normalize=True):
        return self.noise


class Scene:
    def __init__(self, idx):
        spec = wl.c3_scene_spec(idx)
        arrays, amb = wl.host_scene_arrays(spec, dtype=np.float64)
        rng = np.random.default_rng(idx)
        self.duration, self.sample_rate, self.ref_db = spec.duration, spec.sr, spec.ref_db
        evs = [Event(f"event{k:03d}", x, spec.sr, e.n_irs, e.snr, min(e.start, spec.duration - len(x) / spec.sr - 0.2), rng)
               for k, (e, (x, h)) in enumerate(zip(spec.events, arrays))]
        self.events = OrderedDict((e.alias, e) for e in evs)
        self.ambience = OrderedDict(amb=Ambience(amb, spec.ref_db))
        self.audio = OrderedDict()
        irs = np.concatenate([h for _, h in arrays], axis=1)
        self.state = types.SimpleNamespace(name="synthetic", microphones=OrderedDict(mic000=None), num_emitters=irs.shape[1],
                                           simulate=lambda: None, get_irs=lambda: OrderedDict(mic000=irs))
        self.index = idx

    def get_events(self):
        return list(self.events.values())

    def to_dict(self):
        return dict(index=self.index, duration=self.duration, events=list(self.events))


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
