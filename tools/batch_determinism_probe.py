"""Renders a 40-event batch 12 times and compares every event bit by bit with its individual render (prints the events that\ndiffer, with a per-IR scale fit for moving events). Written while hunting a sporadic race in k_ir_fft (round 1)."""
import numpy as np, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "tests/golden")
import cases
from audiblelight_b200.renderer import Renderer, EventJob, moving_frames
from oracle import synth_oracle as orc
def build():
    rng = np.random.default_rng(11)
    jobs, meta = [], []
    for i in range(40):
        lx = int(rng.integers(500, 9000)); lh = int(rng.integers(50, 4000)); n = int(rng.choice([1, 1, 2, 5]))
        x = cases.make_audio(rng, lx); h = cases.make_irs(rng, 4, n, lh).astype(np.float32)
        j = EventJob(audio=x, irs=h, n_channels=4, snr=10.0 + i % 7, ref_db=-65.0)
        if n > 1: j.ir_frames, j.n_frames = moving_frames(lx / 24000.0, 24000.0, n, lx)
        jobs.append(j); meta.append((lx, lh, n))
    return jobs, meta
rnd = Renderer(0)
ref, meta = build()
for s in ref: rnd.render([s])
for rep in range(12):
    jobs, _ = build()
    rnd.render(jobs)
    for i, (a, b) in enumerate(zip(jobs, ref)):
        d = np.abs(a.spatial - b.spatial).max()
        if d > 0:
            lx, lh, n = meta[i]
            print("rep", rep, "event", i, meta[i], "rel", d / np.abs(b.spatial).max(), "stats batch", a.stats["peak"], a.stats["mean_abs"], a.stats["gain"], "single", b.stats["peak"], b.stats["mean_abs"], b.stats["gain"])
            if n > 1:
                irs = a.irs.astype(np.float64)
                irs_n = orc.normalize_irs(irs.transpose(1, 0, 2)).transpose(1, 0, 2)
                ys = []
                for l in range(n):
                    z = np.zeros_like(irs_n); z[:, l] = irs_n[:, l]
                    y = orc.time_variant_convolution_closed(z, a.audio, lx / 24000.0, 24000.0)
                    ys.append(orc.pad_or_truncate(y, lx).ravel())
                A = np.stack(ys, 1)
                beta_b = np.linalg.lstsq(A, a.spatial.astype(np.float64).ravel(), rcond=None)[0]
                beta_s = np.linalg.lstsq(A, b.spatial.astype(np.float64).ravel(), rcond=None)[0]
                print("   per-IR scale ratio batch/single:", beta_b / beta_s)
print("done")
