"""How well does the work model of baseline/reference_arm.py predict the unmodified reference's time for a WHOLE
moving event from another whole moving event?  (run in the build container: needs /root/reference or baseline/_ref)

    python tools/ref_extrapolation_check.py 2.0 3.5 6.0 > profiles/r02_reference_extrapolation.txt
"""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(dur):
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    import numpy as np
    from baseline import ref_loader
    from baseline.reference_arm import _n_stft_frames, _work
    syn = ref_loader.load_reference_synthesize()
    sr, C, Lh = 24000, 4, 24000
    n = int(round(dur * sr))
    N = int(round(10.0 * n / sr)) + 1
    rng = np.random.default_rng(int(dur * 1000))
    x = rng.standard_normal(n).astype(np.float32)
    x /= np.abs(x).max()
    h = rng.standard_normal((C, N, Lh)) * np.exp(-np.arange(Lh) / (Lh / 6.0))
    ev = ref_loader.RefEvent(x, sr, N, 10.0)
    t0 = time.perf_counter()
    syn.render_event_audio(ev, h, "mic000", ref_db=-65.0)
    return dur, N, time.perf_counter() - t0, _work(_n_stft_frames(n), _n_stft_frames(Lh))


if __name__ == "__main__":
    durs = [float(a) for a in sys.argv[1:]] or [2.0, 3.5, 6.0]
    with mp.get_context("fork").Pool(len(durs)) as pool:
        res = pool.map(one, durs)
    print("# unmodified reference render_event_audio, whole moving events (24 kHz, 4 ch, 1 s RIRs, 10 RIR/s), 1 core each")
    print("# duration_s n_irs seconds work_model")
    for d, N, t, w in res:
        print(f"{d:.1f} {N} {t:.2f} {w:.0f}")
    d0, _, t0, w0 = res[0]
    print(f"# prediction from the {d0:.1f} s event (time x work ratio) vs measured:")
    for d, N, t, w in res[1:]:
        pred = t0 * w / w0
        print(f"{d:.1f} s: predicted {pred:.2f} s, measured {t:.2f} s, error {100.0 * (pred - t) / t:+.1f} %")
