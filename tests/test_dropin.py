"""Drop-in host layer (audiblelight_b200.synthesize): same names, signatures, side effects and error strings as
audiblelight/synthesize.py. CPU tests cover everything that happens before the GPU call; -m gpu tests drive
duck-typed Scene / Event / Ambience objects (tests/golden/ref_loader.py stand-ins) end to end and compare with
the reference's golden outputs."""
import inspect
import os
import types
from collections import OrderedDict

import numpy as np
import pytest

import cases
import ref_loader
from ref_loader import RefAmbience, RefEvent, RefScene

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def syn():
    import audiblelight_b200.synthesize as s
    return s


# ---- CPU: signatures and validation ---------------------------------------------------------------------------
def test_signatures_match_reference(syn):
    # synthesize.py:507-516, :613-615, :314, :71, :277-283
    assert list(inspect.signature(syn.render_event_audio).parameters) == [
        "event", "irs", "mic_alias", "ref_db", "ignore_cache", "fft_size", "win_size", "hop_size"]
    p = inspect.signature(syn.render_event_audio).parameters
    assert (p["ref_db"].default, p["ignore_cache"].default, p["fft_size"].default, p["win_size"].default,
            p["hop_size"].default) == (-65, True, 512, 256, 128)
    assert list(inspect.signature(syn.render_audio_for_all_scene_events).parameters) == ["scene", "ignore_cache"]
    assert inspect.signature(syn.render_audio_for_all_scene_events).parameters["ignore_cache"].default is False
    assert list(inspect.signature(syn.generate_scene_audio_from_events).parameters) == ["scene"]
    assert list(inspect.signature(syn.time_invariant_convolution).parameters) == ["audio", "ir"]
    assert list(inspect.signature(syn.time_variant_convolution).parameters) == [
        "irs", "event", "fft_size", "win_size", "hop_size"]


def test_time_invariant_convolution_invalid(syn):
    # tests/test_synthesize.py:25-39 of the reference
    with pytest.raises(ValueError, match="Only mono input is supported"):
        syn.time_invariant_convolution(np.zeros((2, 100)), np.zeros((100, 4)))
    with pytest.raises(ValueError, match="Expected shape of IR should be"):
        syn.time_invariant_convolution(np.zeros(100), np.zeros(100))


def test_render_event_audio_errors_before_compute(syn):
    audio, irs = cases.event_inputs(cases.EVENT_CASES["static_4ch"])
    ev = RefEvent(audio, 24000, 1, 10.0)
    ev.is_moving = True
    with pytest.raises(ValueError, match="Moving Event has only one emitter!"):
        syn.render_event_audio(ev, irs, "mic000")
    audio, irs = cases.event_inputs(cases.EVENT_CASES["moving_2ir"])
    ev = RefEvent(audio, 24000, 2, 10.0)
    ev.is_moving = False
    with pytest.raises(ValueError, match="Expected a moving event!"):
        syn.render_event_audio(ev, irs, "mic000")
    bad = audio.copy()
    bad[5] = np.inf
    with pytest.raises(Exception, match="finite"):
        syn.render_event_audio(RefEvent(bad, 24000, 2, 10.0), irs, "mic000")
    with pytest.raises(ValueError, match="512/256/128"):
        syn.render_event_audio(RefEvent(audio, 24000, 2, 10.0), irs, "mic000", fft_size=1024)
    ev = RefEvent(audio, 24000, 2, 10.0, ref_ir_channel=7, direct_path_time_ms=(6, 50))
    with pytest.raises(ValueError, match="Reference channel index out of range"):
        syn.render_event_audio(ev, irs, "mic000")


def test_cached_event_is_skipped(syn):
    audio, irs = cases.event_inputs(cases.EVENT_CASES["static_4ch"])
    ev = RefEvent(audio, 24000, 1, 10.0)
    marker = np.ones((4, 3))
    ev.spatial_audio["mic000"] = marker
    syn.render_event_audio(ev, irs, "mic000", ignore_cache=False)  # returns before touching the GPU
    assert ev.spatial_audio["mic000"] is marker


class FakeState:
    name = "fake"

    def __init__(self, irs_by_mic, n_emitters):
        self._irs = irs_by_mic
        self.microphones = OrderedDict((m, types.SimpleNamespace(n_listeners=1)) for m in irs_by_mic)
        self.num_emitters = n_emitters
        self.simulated = 0

    def simulate(self):
        self.simulated += 1
        self.irs = self._irs

    def get_irs(self):
        return self._irs


def test_validate_scene_messages(syn):
    # tests/test_synthesize.py:229-288 of the reference
    sc = RefScene(1.0, 8000, -65, [])
    sc.state = FakeState(OrderedDict(mic000=np.zeros((4, 0, 10))), 0)
    with pytest.raises(ValueError, match="WorldState has no emitters!"):
        syn.validate_scene(sc)
    sc.state.num_emitters = 1
    sc.state.microphones = OrderedDict()
    with pytest.raises(ValueError, match="WorldState has no microphones!"):
        syn.validate_scene(sc)
    sc.state.microphones = OrderedDict(mic000=None)
    with pytest.raises(ValueError, match="Scene has no events!"):
        syn.validate_scene(sc)


def test_ambience_type_and_shape_errors(syn):
    audio, irs = cases.event_inputs(cases.EVENT_CASES["static_4ch"])
    ev = RefEvent(audio, 24000, 1, 10.0, alias="e0")
    ev.spatial_audio["mic000"] = np.zeros((4, len(audio)))
    sc = RefScene(1.0, 24000, -65, [ev], OrderedDict(a=object()))
    with pytest.raises(TypeError, match="Expected scene ambient noise to be of type Ambience"):
        syn.generate_scene_audio_from_events(sc)
    sc = RefScene(1.0, 24000, -65, [ev], OrderedDict(a=RefAmbience(np.zeros((4, 100)), -65)))
    with pytest.raises(ValueError, match="Scene ambient noise does not match expected shape"):
        syn.generate_scene_audio_from_events(sc)


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference checkout not present")
def test_install_rebinds_reference_functions(syn):
    ref = ref_loader.load_reference_synthesize()
    orig = ref.render_event_audio
    syn.install()
    try:
        assert ref.render_event_audio is syn.render_event_audio
        assert ref.generate_scene_audio_from_events is syn.generate_scene_audio_from_events
        assert ref.render_audio_for_all_scene_events is syn.render_audio_for_all_scene_events
    finally:
        syn.uninstall()
    assert ref.render_event_audio is orig


# ---- GPU: the full object-level flow ---------------------------------------------------------------------------------
def _scene_objects(spec):
    evs_in, ambs = cases.scene_inputs(spec)
    events, ir_list = [], []
    for i, (e, (audio, irs)) in enumerate(zip(spec["events"], evs_in)):
        events.append(RefEvent(audio, spec["sr"], irs.shape[1], e["snr"], scene_start=e["start"], alias=f"event{i:03d}",
                               ref_ir_channel=e.get("ref_ir_channel"), direct_path_time_ms=e.get("direct_path_time_ms")))
        ir_list.append(irs)
    amb = OrderedDict((f"amb{i}", RefAmbience(a, db)) for i, (a, db) in enumerate(zip(ambs, spec["ambience_ref_db"])))
    scene = RefScene(spec["duration"], spec["sr"], spec["ref_db"], events, amb)
    all_irs = np.concatenate(ir_list, axis=1)  # (C, sum N, Lh) like WorldState.get_irs
    scene.state = FakeState(OrderedDict(mic000=all_irs), all_irs.shape[1])
    return scene


@pytest.mark.gpu
@pytest.mark.parametrize("batched", [False, True])
@pytest.mark.parametrize("name", list(cases.SCENE_CASES))
def test_scene_flow_golden(syn, name, batched):
    gsc = np.load(os.path.join(G, "scenes.npz"))
    spec = cases.SCENE_CASES[name]
    scene = _scene_objects(spec)
    if batched:
        syn.render_scenes([scene], ignore_cache=False)
    else:
        syn.render_audio_for_all_scene_events(scene)
        syn.generate_scene_audio_from_events(scene)
    assert scene.state.simulated == 1
    ref = gsc[f"{name}__scene"]
    got = scene.audio["mic000"]
    assert got.dtype == np.float32 and got.shape == ref.shape
    assert np.abs(got.astype(np.float64) - ref).max() <= 1e-5
    slices = gsc[f"{name}__slices"]
    for i, ev in enumerate(scene.events.values()):
        a, b = slices[i]
        sp = ev.spatial_audio["mic000"]
        assert sp.shape == (spec["c"], len(ev.audio)) and sp.dtype == np.float64
        if b > a:
            pad = ev._spatial_audio_padded["mic000"]
            assert pad.shape == ref.shape and pad.dtype == np.float32
            assert np.allclose([np.abs(pad).sum(), np.abs(pad[:, a:b]).sum()], gsc[f"{name}__padded{i}_sum"], rtol=1e-4)
            dk = f"{name}__drypadded{i}"
            if dk in gsc.files:
                assert np.abs(ev._spatial_audio_dry_padded["mic000"].astype(np.float64) - gsc[dk]).max() <= 1e-5
        else:
            assert "mic000" not in ev._spatial_audio_padded


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(cases.EVENT_CASES))
def test_render_event_audio_object_golden(syn, name):
    gev = np.load(os.path.join(G, "events.npz"))
    spec = cases.EVENT_CASES[name]
    audio, irs = cases.event_inputs(spec)
    ev = RefEvent(audio, spec["sr"], irs.shape[1], spec["snr"], ref_ir_channel=spec.get("ref_ir_channel"),
                  direct_path_time_ms=spec.get("direct_path_time_ms"))
    syn.render_event_audio(ev, irs, "mic000", ref_db=spec["ref_db"])
    ref = gev[f"{name}__spatial"]
    got = ev.spatial_audio["mic000"]
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert np.abs(got.astype(np.float64) - ref).max() <= 1e-5
    if f"{name}__dry" in gev.files:
        assert np.abs(ev._spatial_audio_dry["mic000"] - gev[f"{name}__dry"]).max() <= 1e-5
    else:
        assert "mic000" not in ev._spatial_audio_dry


@pytest.mark.gpu
def test_raw_convolution_functions(syn):
    g = np.load(os.path.join(G, "conv_primitives.npz"))
    a = cases.make_audio(np.random.default_rng(2), 3000)
    h = cases.make_irs(np.random.default_rng(3), 4, 1, 801)[:, 0].T
    out = syn.time_invariant_convolution(a, h)
    assert out.shape == g["tic_out"].shape and out.dtype == np.float64
    assert np.abs(out - g["tic_out"]).max() < 3e-6 * np.abs(g["tic_out"]).max()
    spec = cases.EVENT_CASES["moving_5ir"]
    audio, irs = cases.event_inputs(spec)
    ev = RefEvent(audio, spec["sr"], irs.shape[1], spec["snr"])
    out = syn.time_variant_convolution(irs, ev, 512, 256, 128)
    assert out.shape == g["tvc_out"].shape
    assert np.abs(out - g["tvc_out"]).max() < 3e-6 * np.abs(g["tvc_out"]).max()


@pytest.mark.gpu
def test_two_microphones_and_cache(syn):
    spec = cases.SCENE_CASES["scene_static_ambience"]
    scene = _scene_objects(spec)
    irs0 = scene.state._irs["mic000"]
    scene.state._irs = OrderedDict(mic000=irs0, mic001=irs0[:, :, ::-1].copy())
    scene.state.microphones = OrderedDict(mic000=None, mic001=None)
    syn.render_audio_for_all_scene_events(scene)
    first = {k: ev.spatial_audio["mic000"] for k, ev in scene.events.items()}
    assert all(set(ev.spatial_audio) == {"mic000", "mic001"} for ev in scene.events.values())
    syn.render_audio_for_all_scene_events(scene, ignore_cache=False)  # cached: untouched, no re-simulation
    assert scene.state.simulated == 1
    assert all(scene.events[k].spatial_audio["mic000"] is v for k, v in first.items())
    syn.generate_scene_audio_from_events(scene)
    assert set(scene.audio) == {"mic000", "mic001"}
    gsc = np.load(os.path.join(G, "scenes.npz"))
    assert np.abs(scene.audio["mic000"].astype(np.float64) - gsc["scene_static_ambience__scene"]).max() <= 1e-5
