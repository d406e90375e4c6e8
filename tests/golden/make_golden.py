"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) on seeded inputs.

    python tests/golden/make_golden.py          # writes tests/golden/*.npz

Only runs where /root/reference exists (the build container). The .npz files are committed; tests and the
GPU box never need the reference itself.
"""
import os
import sys
from collections import OrderedDict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402
import ref_loader  # noqa: E402


def run_event(syn, spec, audio, irs, alias="ev", start=0.0):
    ev = ref_loader.RefEvent(audio, spec["sr"], irs.shape[1], spec["snr"], scene_start=start, alias=alias,
                             ref_ir_channel=spec.get("ref_ir_channel"),
                             direct_path_time_ms=spec.get("direct_path_time_ms"))
    syn.render_event_audio(ev, irs, "mic000", ref_db=spec["ref_db"])
    return ev


def main():
    syn = ref_loader.load_reference_synthesize()
    out = {}

    # --- known-answer bits of the small helpers -------------------------------------------------
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 50))
    out["kat_apply_snr_in"] = x
    out["kat_apply_snr_out"] = syn.apply_snr(x, 7.0)
    out["kat_db_mult"] = np.array([syn.db_to_multiplier(db, v) for db, v in
                                   [(0, 1.0), (6.0206, 1.0), (-6.0206, 1.0), (20, 0.1), (-20, 10.0),
                                    (-65 + 12.5, 0.0371)]])
    irs = rng.standard_normal((5, 4, 300))
    out["kat_norm_irs_in"] = irs
    out["kat_norm_irs_out"] = syn.normalize_irs(irs)
    # interpolation matrices for awkward rates / durations (exact integer structure)
    for i, (dur, n, sr) in enumerate([(1.0, 11, 24000), (0.25, 2, 24000), (0.2948, 4, 44100),
                                      (0.5625, 3, 16000), (2.0, 21, 24000), (0.375, 6, 48000)]):
        t = np.linspace(0, dur, n)
        out[f"kat_interp_{i}_args"] = np.array([dur, n, sr])
        out[f"kat_interp_{i}"] = syn.generate_interpolation_matrix(t, sr, 128)
    y = rng.standard_normal((2, 3, 700))
    out["kat_stft_in"] = y
    out["kat_stft_out"] = syn.stft(y, 512, 256, 128)
    out["kat_tiny"] = np.array([syn.utils.tiny(np.float32(1)), syn.utils.tiny(np.float64(1)), syn.utils.tiny(3)])
    np.savez_compressed(os.path.join(HERE, "kat_helpers.npz"), **out)

    # --- conv primitives -----------------------------------------------------------------------
    out = {}
    a = cases.make_audio(np.random.default_rng(2), 3000)
    h = cases.make_irs(np.random.default_rng(3), 4, 1, 801)[:, 0].T  # (Lh, C)
    out["tic_out"] = syn.time_invariant_convolution(a, h)
    spec = cases.EVENT_CASES["moving_5ir"]
    audio, irs = cases.event_inputs(spec)
    ev = ref_loader.RefEvent(audio, spec["sr"], irs.shape[1], spec["snr"])
    out["tvc_out"] = syn.time_variant_convolution(irs, ev, 512, 256, 128)
    np.savez_compressed(os.path.join(HERE, "conv_primitives.npz"), **out)

    # --- render_event_audio --------------------------------------------------------------------
    out = {}
    for name, spec in cases.EVENT_CASES.items():
        audio, irs = cases.event_inputs(spec)
        ev = run_event(syn, spec, audio, irs)
        out[f"{name}__spatial"] = ev.spatial_audio["mic000"]
        if "mic000" in ev._spatial_audio_dry:
            out[f"{name}__dry"] = ev._spatial_audio_dry["mic000"]
        out[f"{name}__audio_sha"] = np.array([float(np.abs(audio).sum()), float(np.abs(irs).sum())])
        print(name, ev.spatial_audio["mic000"].shape, float(np.abs(ev.spatial_audio["mic000"]).max()))
    np.savez_compressed(os.path.join(HERE, "events.npz"), **out)

    # --- generate_scene_audio_from_events --------------------------------------------------------
    out = {}
    for name, spec in cases.SCENE_CASES.items():
        evs_in, ambs = cases.scene_inputs(spec)
        events = []
        for i, (e, (audio, irs)) in enumerate(zip(spec["events"], evs_in)):
            espec = dict(sr=spec["sr"], snr=e["snr"], ref_db=spec["ref_db"],
                         ref_ir_channel=e.get("ref_ir_channel"),
                         direct_path_time_ms=e.get("direct_path_time_ms"))
            events.append(run_event(syn, espec, audio, irs, alias=f"event{i:03d}", start=e["start"]))
        amb = OrderedDict((f"amb{i}", ref_loader.RefAmbience(a, db))
                          for i, (a, db) in enumerate(zip(ambs, spec["ambience_ref_db"])))
        scene = ref_loader.RefScene(spec["duration"], spec["sr"], spec["ref_db"], events, amb)
        syn.generate_scene_audio_from_events(scene)
        out[f"{name}__scene"] = scene.audio["mic000"]
        sl = []
        for i, ev in enumerate(events):
            a = max(0, round(ev.scene_start * scene.sample_rate))
            b = min(round(ev.scene_end * scene.sample_rate), scene.audio["mic000"].shape[1])
            sl.append((a, b))
            if "mic000" in ev._spatial_audio_padded:
                # store only the non-zero support of the padded copy to keep the fixture small
                out[f"{name}__padded{i}_sum"] = np.array(
                    [float(np.abs(ev._spatial_audio_padded["mic000"]).sum()),
                     float(np.abs(ev._spatial_audio_padded["mic000"][:, a:b]).sum())])
            if "mic000" in ev._spatial_audio_dry_padded:
                out[f"{name}__drypadded{i}"] = ev._spatial_audio_dry_padded["mic000"]
        out[f"{name}__slices"] = np.array(sl, dtype=np.int64)
        print(name, scene.audio["mic000"].shape, scene.audio["mic000"].dtype, sl)
    # event-slice rounding KATs (synthesize.py:361-362): python round() half-to-even on float products
    kat = []
    for s, d, sr, tot in [(0.50006, 0.375, 8000, 16000), (0.1, 0.5, 8000, 16000), (1.00003125, 2.0, 16000, 40000),
                          (0.0000312, 1.0, 16000, 16000), (2.5 / 24000, 0.5, 24000, 24000),
                          (3.5 / 24000, 0.5, 24000, 24000), (59.99, 5.0, 24000, 1440000)]:
        a = max(0, round(s * sr))
        b = min(round((s + d) * sr), tot)
        kat.append((s, d, sr, tot, a, b))
    out["slice_kat"] = np.array(kat, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "scenes.npz"), **out)


def make_augment_golden():
    """Fade / Invert / Reverse of the unmodified reference (augmentation.py:1403-1601) -> tests/golden/augment.npz."""
    ref_loader.load_reference_synthesize()
    import audiblelight.augmentation as aug
    out = {}
    x = np.random.default_rng(77).standard_normal(6000).astype(np.float32)
    out["x"] = x
    fade_cases = [(24000, 0.05, 0.1, "half_sine", "logarithmic"), (24000, 0.0, 0.2, "none", "exponential"),
                  (16000, 0.125, 0.0, "quarter_sine", "none"), (44100, 0.01, 0.01, "linear", "linear"),
                  (24000, 1.0, 1.0, "exponential", "quarter_sine"), (24000, 0.1, 0.05, "logarithmic", "half_sine")]
    out["fade_cases"] = np.array([[c[0], c[1], c[2], aug.Fade.FADE_SHAPES.index(c[3]), aug.Fade.FADE_SHAPES.index(c[4])]
                                  for c in fade_cases])
    for i, (sr, fi, fo, si, so) in enumerate(fade_cases):
        out[f"fade_{i}"] = aug.Fade(sample_rate=sr, fade_in_len=fi, fade_out_len=fo, fade_in_shape=si, fade_out_shape=so)(x)
    out["invert"] = aug.Invert(24000)(x)
    out["reverse"] = aug.Reverse(24000)(x)
    np.savez_compressed(os.path.join(HERE, "augment.npz"), **out)


def make_dcase_golden():
    """Rows of the unmodified reference generate_dcase2024_metadata (synthesize.py:742-878) on the seeded duck-typed
    scenes of cases.dcase_random_scene -> tests/golden/dcase.npz."""
    syn = ref_loader.load_reference_synthesize()
    out = {}
    for seed in cases.DCASE_RANDOM_SEEDS:
        scene = cases.dcase_random_scene(seed)
        res = syn.generate_dcase2024_metadata(scene)
        for mic, df in res.items():
            out[f"s{seed}_{mic}"] = df.reset_index(drop=False).to_numpy().astype(np.int64)
            out[f"s{seed}_{mic}_csv"] = np.frombuffer(df.to_csv(sep=",", encoding="utf-8", header=None).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "dcase.npz"), **out)
    print("dcase.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
    make_augment_golden()
    make_dcase_golden()
