"""Pins the CPU oracle (oracle/synth_oracle.py) to the reference: golden vectors produced by the unmodified
reference (tests/golden/make_golden.py) and the reference's own known-answer tests
(/root/reference/tests/test_synthesize.py:42-57, 307-337, 340-377)."""
import os

import numpy as np
import pytest

import cases
from oracle import synth_oracle as orc

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def kat():
    return np.load(os.path.join(G, "kat_helpers.npz"))


@pytest.fixture(scope="module")
def gev():
    return np.load(os.path.join(G, "events.npz"))


@pytest.fixture(scope="module")
def gsc():
    return np.load(os.path.join(G, "scenes.npz"))


# ---- reference's own KATs -------------------------------------------------------------------------
@pytest.mark.parametrize("x,snr,peak", [
    (np.array([0.5, -1.0, 0.25]), 5, 5), (np.array([[0.1, 0.2], [-0.4, 0.3]]), 10, 10),
    (np.zeros(10), 5, 0), (np.array([1e-20, -1e-20]), 3, 3e-5), (np.array([2.0, -4.0]), -2, 2),
])
def test_apply_snr_peak(x, snr, peak):
    # tests/test_synthesize.py:42-57 — peak equals snr, all-zero input stays zero
    out = orc.apply_snr(x, snr)
    assert np.isclose(np.abs(out).max(), peak, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("db,x,expected", [(0, 1.0, 1.0), (6.0206, 1.0, 2.0), (-6.0206, 1.0, 0.5),
                                           (20, 0.1, 100.0), (-20, 10.0, 0.01)])
def test_db_to_multiplier_kat(db, x, expected):
    # tests/test_synthesize.py:307-337
    assert np.isclose(orc.db_to_multiplier(db, x), expected, atol=1e-4)


def test_normalize_irs_unit_mean_energy():
    # tests/test_synthesize.py:340-377 — mean capsule energy becomes 1
    rng = np.random.default_rng(0)
    irs = rng.standard_normal((3, 4, 1000))
    out = orc.normalize_irs(irs)
    e = np.sqrt((out ** 2).sum(-1))
    assert np.allclose(e.mean(-1), 1.0)


def test_time_invariant_convolution_errors():
    # tests/test_synthesize.py:25-39
    with pytest.raises(ValueError, match="Only mono input is supported"):
        orc.time_invariant_convolution(np.zeros((2, 10)), np.zeros((10, 4)))
    with pytest.raises(ValueError, match="Expected shape of IR should be"):
        orc.time_invariant_convolution(np.zeros(10), np.zeros(10))


# ---- golden: helpers --------------------------------------------------------------------------------
def test_helpers_golden(kat):
    assert np.array_equal(orc.apply_snr(kat["kat_apply_snr_in"], 7.0), kat["kat_apply_snr_out"])
    got = np.array([orc.db_to_multiplier(db, v) for db, v in
                    [(0, 1.0), (6.0206, 1.0), (-6.0206, 1.0), (20, 0.1), (-20, 10.0), (-65 + 12.5, 0.0371)]])
    assert np.array_equal(got, kat["kat_db_mult"])
    assert np.allclose(orc.normalize_irs(kat["kat_norm_irs_in"]), kat["kat_norm_irs_out"], rtol=1e-14, atol=0)
    assert np.array_equal(np.array([orc.tiny(np.float32(1)), orc.tiny(np.float64(1)), orc.tiny(3)]), kat["kat_tiny"])


@pytest.mark.parametrize("i", range(6))
def test_interpolation_matrix_bit_exact(kat, i):
    dur, n, sr = kat[f"kat_interp_{i}_args"]
    w = orc.interpolation_matrix(np.linspace(0, dur, int(n)), sr, 128)
    assert w.shape == kat[f"kat_interp_{i}"].shape
    assert np.array_equal(w, kat[f"kat_interp_{i}"])


def test_stft_golden(kat):
    got = orc.stft(kat["kat_stft_in"])
    ref = kat["kat_stft_out"]
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 1e-11


def test_conv_primitives_golden():
    g = np.load(os.path.join(G, "conv_primitives.npz"))
    a = cases.make_audio(np.random.default_rng(2), 3000)
    h = cases.make_irs(np.random.default_rng(3), 4, 1, 801)[:, 0].T
    got = orc.time_invariant_convolution(a, h)
    assert got.shape == g["tic_out"].shape == (4, 3800)
    assert np.abs(got - g["tic_out"]).max() < 1e-11
    spec = cases.EVENT_CASES["moving_5ir"]
    audio, irs = cases.event_inputs(spec)
    dur = len(audio) / float(spec["sr"])
    lit = orc.time_variant_convolution(irs, audio, dur, float(spec["sr"]))
    assert lit.shape == g["tvc_out"].shape
    scale = np.abs(g["tvc_out"]).max()
    assert np.abs(lit - g["tvc_out"]).max() < 1e-12 * scale
    closed = orc.time_variant_convolution_closed(irs, audio, dur, float(spec["sr"]))
    assert closed.shape == g["tvc_out"].shape
    assert np.abs(closed - g["tvc_out"]).max() < 1e-11 * scale


# ---- golden: render_event_audio -------------------------------------------------------------------------
def _render(spec, literal):
    audio, irs = cases.event_inputs(spec)
    return orc.render_event(audio, irs, spec["snr"], spec["ref_db"], is_moving=spec["n"] > 1,
                            duration=len(audio) / float(spec["sr"]), sample_rate=float(spec["sr"]),
                            ref_ir_channel=spec.get("ref_ir_channel"),
                            direct_path_time_ms=spec.get("direct_path_time_ms"), literal=literal)


@pytest.mark.parametrize("literal", [True, False])
@pytest.mark.parametrize("name", list(cases.EVENT_CASES))
def test_render_event_golden(gev, name, literal):
    spec = cases.EVENT_CASES[name]
    audio, irs = cases.event_inputs(spec)
    assert np.allclose(gev[f"{name}__audio_sha"], [np.abs(audio).sum(), np.abs(irs).sum()], rtol=0, atol=0)
    res = _render(spec, literal)
    ref = gev[f"{name}__spatial"]
    # float64 everywhere except the no-IR case, where the reference keeps the float32 dry audio (synthesize.py:577)
    assert res.spatial.shape == ref.shape and res.spatial.dtype == ref.dtype
    assert ref.dtype == (np.float32 if spec["n"] == 0 else np.float64)
    # full scale is 1.0; the oracle must sit far inside the 1e-5 product tolerance
    assert np.abs(res.spatial - ref).max() < (1e-12 if spec["n"] else 1e-9)
    if f"{name}__dry" in gev.files:
        assert res.dry.shape == gev[f"{name}__dry"].shape
        assert np.abs(res.dry - gev[f"{name}__dry"]).max() < 1e-12
    else:
        assert res.dry is None


def test_render_event_errors():
    audio, irs = cases.event_inputs(cases.EVENT_CASES["static_4ch"])
    with pytest.raises(ValueError, match="Moving Event has only one emitter!"):
        orc.render_event(audio, irs, 10.0, -65, is_moving=True)
    audio, irs = cases.event_inputs(cases.EVENT_CASES["moving_2ir"])
    with pytest.raises(ValueError, match="Expected a moving event!"):
        orc.render_event(audio, irs, 10.0, -65, is_moving=False)
    bad = audio.copy()
    bad[3] = np.nan
    with pytest.raises(ValueError):
        orc.render_event(bad, irs, 10.0, -65, is_moving=True, duration=0.25, sample_rate=24000.0)


# ---- golden: scene mix ----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.SCENE_CASES))
def test_mix_scene_golden(gsc, name):
    spec = cases.SCENE_CASES[name]
    evs_in, ambs = cases.scene_inputs(spec)
    spatial, dry, starts, ends = [], [], [], []
    for e, (audio, irs) in zip(spec["events"], evs_in):
        dur = len(audio) / float(spec["sr"])
        r = orc.render_event(audio, irs, e["snr"], spec["ref_db"], is_moving=e["n"] > 1, duration=dur,
                             sample_rate=float(spec["sr"]), ref_ir_channel=e.get("ref_ir_channel"),
                             direct_path_time_ms=e.get("direct_path_time_ms"), literal=False)
        spatial.append(r.spatial)
        dry.append(r.dry)
        starts.append(float(e["start"]))
        ends.append(float(e["start"]) + dur)
    mix = orc.mix_scene(spec["duration"], spec["sr"], spatial, starts, ends,
                        list(zip(ambs, spec["ambience_ref_db"])), dry)
    ref = gsc[f"{name}__scene"]
    assert mix.scene.dtype == ref.dtype == np.float32 and mix.scene.shape == ref.shape
    assert np.abs(mix.scene.astype(np.float64) - ref).max() < 1e-9
    assert np.array_equal(np.array(mix.slices, dtype=np.int64), gsc[f"{name}__slices"])  # timings bit-exact
    for i, p in enumerate(mix.padded):
        key = f"{name}__padded{i}_sum"
        if p is None:
            assert key not in gsc.files
            continue
        a, b = mix.slices[i]
        assert np.allclose([np.abs(p).sum(), np.abs(p[:, a:b]).sum()], gsc[key], rtol=1e-6)
        dk = f"{name}__drypadded{i}"
        if dk in gsc.files:
            assert np.abs(mix.dry_padded[i].astype(np.float64) - gsc[dk]).max() < 1e-9
        else:
            assert mix.dry_padded[i] is None


def test_event_slice_kat(gsc):
    for s, d, sr, tot, a, b in gsc["slice_kat"]:
        assert orc.event_slice(s, s + d, sr, int(tot)) == (int(a), int(b))


def test_mix_scene_ambience_shape_error():
    with pytest.raises(ValueError, match="Scene ambient noise does not match expected shape"):
        orc.mix_scene(1.0, 1000, [np.zeros((2, 10))], [0.0], [0.01], [(np.zeros((2, 999)), -65)])
