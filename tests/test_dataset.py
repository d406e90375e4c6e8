"""Scope row f2 (+ the PCM packing of f4): DCASE 2024 metadata, PCM_16 WAV output and the batch driver
(audiblelight_b200.dataset). The metadata is integer host code: BIT-EXACT against the reference's own expected tables
and against rows produced by the unmodified reference function (tests/golden/dcase.npz)."""
import json
import os
import types
from collections import OrderedDict

import numpy as np
import pytest

import cases
from audiblelight_b200 import augment as A
from audiblelight_b200 import dataset
from oracle import wav_oracle
from ref_loader import RefAmbience, RefEvent, RefScene

G = os.path.join(os.path.dirname(__file__), "golden")

# (azimuth, elevation, distance m, scene_start, duration, class id, file) and the expected rows, verbatim from the
# reference's tests/test_dcase_metadata.py:247-352 (class ids of event.DCASE_SOUND_EVENT_CLASSES)
REFERENCE_TABLES = [
    ([(-50, 30, 1.81, 1.0, 0.1, 1, "93853.wav"), (10, -20, 2.43, 1.1, 0.2, 1, "93856.wav"), (-40, 0, 0.80, 1.3, 0.04, 8, "music.wav")],
     [[10, 1, 0, -50, 30, 181], [11, 1, 0, -50, 30, 181], [11, 1, 1, 10, -20, 243], [12, 1, 1, 10, -20, 243],
      [13, 1, 1, 10, -20, 243], [13, 8, 0, -40, 0, 80]]),
    ([(95.0, 5.0, 1.0, 10.0, 0.5, 9, "3471.wav"), (129, -18, 0.5, 10.2, 0.3, 4, "9547.wav")],
     [[100, 9, 0, 95, 5, 100], [101, 9, 0, 95, 5, 100], [102, 4, 0, 129, -18, 50], [102, 9, 0, 95, 5, 100],
      [103, 4, 0, 129, -18, 50], [103, 9, 0, 95, 5, 100], [104, 4, 0, 129, -18, 50], [104, 9, 0, 95, 5, 100],
      [105, 4, 0, 129, -18, 50], [105, 9, 0, 95, 5, 100]]),
    ([(-55.0, 9.0, 2.64, 25.5, 0.4, 7, "35632.wav"), (-61.0, -6.0, 2.18, 27.5, 0.5, 10, "95709.wav")],
     [[255, 7, 0, -55, 9, 264], [256, 7, 0, -55, 9, 264], [257, 7, 0, -55, 9, 264], [258, 7, 0, -55, 9, 264],
      [259, 7, 0, -55, 9, 264], [275, 10, 0, -61, -6, 218], [276, 10, 0, -61, -6, 218], [277, 10, 0, -61, -6, 218],
      [278, 10, 0, -61, -6, 218], [279, 10, 0, -61, -6, 218], [280, 10, 0, -61, -6, 218]]),
]


@pytest.mark.parametrize("events,expected", REFERENCE_TABLES)
def test_dcase_rows_reference_expected_tables(events, expected):
    scene = cases.dcase_static_scene(30, events)
    assert np.array_equal(dataset.dcase2024_rows(scene)["poltest"], np.array(expected))


@pytest.mark.parametrize("seed", cases.DCASE_RANDOM_SEEDS)
def test_dcase_rows_and_csv_match_reference_golden(seed):
    gold = np.load(os.path.join(G, "dcase.npz"))
    rows = dataset.dcase2024_rows(cases.dcase_random_scene(seed))
    assert list(rows) == ["mic000", "mic001"]
    for mic, got in rows.items():
        assert got.dtype == np.int64 and np.array_equal(got, gold[f"s{seed}_{mic}"])
        assert dataset.dcase_csv(got).encode() == gold[f"s{seed}_{mic}_csv"].tobytes()


def test_dcase_dataframe_form():
    pd = pytest.importorskip("pandas")
    gold = np.load(os.path.join(G, "dcase.npz"))
    seed = cases.DCASE_RANDOM_SEEDS[0]
    res = dataset.generate_dcase2024_metadata(cases.dcase_random_scene(seed))
    df = res["mic001"]
    assert isinstance(df, pd.DataFrame) and df.index.name == "frame_number"
    assert list(df.columns) == dataset.DCASE_2024_COLUMNS[1:]
    assert np.array_equal(df.reset_index(drop=False).to_numpy(), gold[f"s{seed}_mic001"])
    assert df.to_csv(sep=",", encoding="utf-8", header=None).encode() == gold[f"s{seed}_mic001_csv"].tobytes()


@pytest.mark.parametrize("starts", [[10, 5, 0], [0, 5, 10], [5, 0, 10]])
def test_dcase_source_ids_follow_start_order(starts):
    # the property the reference checks in tests/test_dcase_metadata.py:378-420: same class, different files ->
    # source ids 0, 1, 2 in order of scene_start; another class starts again at 0
    evs = [(0, 0, 1.0, float(st), 1.0, 7, f"door{k}.wav") for k, st in enumerate(starts)] + [(0, 0, 1.0, 20.0, 1.0, 8, "m.wav")]
    rows = dataset.dcase2024_rows(cases.dcase_static_scene(60, evs))["poltest"]
    for st, want in zip(sorted(starts), [0, 1, 2]):
        assert set(rows[rows[:, 0] == int(st * 10) + 1][:, 2]) == {want}
    assert set(rows[rows[:, 1] == 8][:, 2]) == {0}
    # a repeated file keeps its id
    evs = [(0, 0, 1.0, 0.0, 1.0, 7, "a.wav"), (0, 0, 1.0, 5.0, 1.0, 7, "b.wav"), (0, 0, 1.0, 9.0, 1.0, 7, "a.wav")]
    rows = dataset.dcase2024_rows(cases.dcase_static_scene(60, evs))["poltest"]
    assert set(rows[rows[:, 0] >= 90][:, 2]) == {0} and set(rows[(rows[:, 0] >= 50) & (rows[:, 0] < 70)][:, 2]) == {1}


def test_dcase_invalid_class_and_off_grid():
    scene = cases.dcase_static_scene(30, [(0, 0, 1.0, 1.0, 1.0, "asdf", "a.wav")])
    with pytest.raises(ValueError, match="valid DCASE class indices"):
        dataset.dcase2024_rows(scene)
    with pytest.raises(IndexError):  # the reference's np.where(...)[0][0] on a grid that does not hold the time
        dataset.dcase2024_rows(cases.dcase_static_scene(30, [(0, 0, 1.0, 1.3, 1.0, 1, "a.wav")]), temporal_resolution=0.5)
    assert dataset.dcase2024_rows(cases.DcaseScene(10, ["m"], []))["m"].shape == (0, 6)


def test_pcm16_oracle_known_values():
    x = np.array([[0.0, 0.5, -0.5, 1.0, -1.0, 1.5, 1e-5, 3.0517578125e-05 * 1.5]], dtype=np.float32)
    # 0.5*32767 = 16383.5 -> 16384 (half to even); 1.5*32767 = 49150.5 -> 49150 -> wraps to -16386 in 16 bits
    assert wav_oracle.pcm16_from_float(x)[:, 0].tolist() == [0, 16384, -16384, 32767, -32767, -16386, 0, 1]


def test_wav_roundtrip(tmp_path):
    from scipy.io import wavfile
    pcm = np.random.default_rng(0).integers(-32768, 32767, size=(1000, 4)).astype(np.int16)
    p = tmp_path / "a.wav"
    dataset.write_wav_pcm16(p, pcm, 24000)
    assert os.path.getsize(p) == 44 + pcm.nbytes
    sr, back = dataset.read_wav_pcm16(p)
    assert sr == 24000 and np.array_equal(back, pcm)
    sr2, back2 = wavfile.read(p)
    assert sr2 == 24000 and np.array_equal(back2, pcm)
    dataset.write_wav_pcm16(tmp_path / "m.wav", pcm[:, 0], 8000)
    assert dataset.read_wav_pcm16(tmp_path / "m.wav")[1].shape == (1000, 1)


def test_reference_augmentation_objects_map_to_device_ops():
    def _mk(name, sr, params):
        o = type(name, (), {})()
        o.sample_rate, o.params = sr, params
        return o
    ops = A.chain_from_reference([
        _mk("LowpassFilter", 24000, dict(cutoff_frequency_hz=5000.0)),
        _mk("HighShelfFilter", 24000, dict(cutoff_frequency_hz=3000.0, gain_db=-6.0, q=0.7)),
        _mk("MultibandEqualizer", 24000, dict(n_bands=2, gain_db=[3.0, -2.0], cutoff_frequency_hz=[500.0, 4000.0], q=[1.0, 0.5])),
        _mk("Gain", 24000, dict(gain_db=-3.0)), _mk("Invert", 24000, {}), _mk("Deemphasis", 24000, dict(coef=0.5)),
        _mk("Fade", 24000, dict(fade_in_len=0.1, fade_out_len=0.2, fade_in_shape="linear", fade_out_shape="half_sine")),
    ])
    assert [o.type for o in ops] == [A.ALR_AUG_BIQUAD, A.ALR_AUG_BIQUAD, A.ALR_AUG_BIQUAD, A.ALR_AUG_BIQUAD, A.ALR_AUG_GAIN,
                                     A.ALR_AUG_INVERT, A.ALR_AUG_DEEMPHASIS, A.ALR_AUG_FADE]
    assert ops[0].p == A.lowpass(24000, 5000.0).p and ops[3].p == A.peak(24000, 4000.0, -2.0, 0.5).p
    assert A.from_reference(_mk("Delay", 24000, dict(delay_seconds=0.05, feedback=0.3, mix=0.2)))[0].p == (1200.0, 0.3, 0.2)
    # non-linear / unsupported effects stay on the host
    assert A.chain_from_reference([_mk("Compressor", 24000, dict(threshold_db=-10))]) is None
    assert A.chain_from_reference([_mk("Gain", 24000, dict(gain_db=1.0))] * 9) is None
    assert A.chain_from_reference([]) == []


# ---- GPU ---------------------------------------------------------------------------------------------------------------
class _State:
    name = "fake"

    def __init__(self, irs_by_mic, n_emitters):
        self._irs = irs_by_mic
        self.microphones = OrderedDict((m, types.SimpleNamespace(n_listeners=1)) for m in irs_by_mic)
        self.num_emitters = n_emitters

    def simulate(self):
        self.irs = self._irs

    def get_irs(self):
        return self._irs


def _scene(spec, idx=0, ref_db=None, mics=("mic000",)):
    evs_in, ambs = cases.scene_inputs(spec)
    events, ir_list = [], []
    for i, (e, (audio, irs)) in enumerate(zip(spec["events"], evs_in)):
        ev = RefEvent(audio, spec["sr"], irs.shape[1], e["snr"], scene_start=e["start"], alias=f"event{i:03d}")
        ev.class_id, ev.filename = i % 13, f"file{i}.wav"
        ev.emitters = [cases.DcaseEmitter({m: np.array([[10.0 * i + k, -5.0 * i, 1.0 + 0.01 * k]]) for m in mics})
                       for k in range(irs.shape[1])]
        events.append(ev)
        ir_list.append(irs)
    db = spec["ref_db"] if ref_db is None else ref_db
    amb = OrderedDict((f"amb{i}", RefAmbience(a, d if ref_db is None else ref_db))
                      for i, (a, d) in enumerate(zip(ambs, spec["ambience_ref_db"])))
    scene = RefScene(spec["duration"], spec["sr"], db, events, amb, mics=mics)
    all_irs = np.concatenate(ir_list, axis=1)
    scene.state = _State(OrderedDict((m, all_irs * (1.0 + 0.1 * k)) for k, m in enumerate(mics)), all_irs.shape[1])
    # (one golden scene holds an event that starts after the scene end to exercise the mixer's skip rule; the
    # reference's metadata function raises IndexError for such an event, so it is left out of the metadata here)
    scene.get_events = lambda: [e for e in scene.events.values() if e.scene_start < scene.duration]
    scene.to_dict = lambda: dict(duration=scene.duration, index=idx, events=list(scene.events))
    return scene


@pytest.mark.gpu
@pytest.mark.parametrize("ref_db", [None, -3])  # -3 dB: |mix| exceeds 1.0, exercising the 16-bit wrap of libsndfile
@pytest.mark.parametrize("name", list(cases.SCENE_CASES))
def test_gpu_pcm16_is_bit_exact_conversion_of_the_mix(name, ref_db):
    import audiblelight_b200.synthesize as syn
    scene = _scene(cases.SCENE_CASES[name], ref_db=ref_db)
    pcm = syn.render_scenes([scene], pcm16=True)
    got = pcm[0]["mic000"]
    assert got.dtype == np.int16 and got.shape == scene.audio["mic000"].T.shape
    assert np.array_equal(got, wav_oracle.pcm16_from_float(scene.audio["mic000"]))
    if ref_db is not None:
        assert np.abs(scene.audio["mic000"]).max() > 1.0
    # mix-only mode: nothing but the PCM comes back, and it is the same PCM
    scene2 = _scene(cases.SCENE_CASES[name], ref_db=ref_db)
    r = syn.get_renderer()
    pcm2 = syn.render_scenes([scene2], pcm16=True, keep_event_audio=False, keep_mix=False)
    assert np.array_equal(pcm2[0]["mic000"], got)
    assert "mic000" not in scene2.audio and all("mic000" not in e.spatial_audio for e in scene2.events.values())
    assert got.nbytes <= r.profile()["d2h_bytes"] < got.nbytes + 4096  # the PCM plus the per-event statistics


@pytest.mark.gpu
def test_gpu_pcm16_device_buffers():
    torch = pytest.importorskip("torch")
    import gpu_util
    from audiblelight_b200.renderer import Renderer
    spec = cases.SCENE_CASES["scene_static_ambience"]
    jobs, sjob = gpu_util.scene_jobs(spec)
    for j in jobs:
        j.audio = torch.from_numpy(j.audio).cuda()
        j.irs = torch.from_numpy(j.irs).cuda() if j.irs is not None else None
    sjob.ambience = [torch.from_numpy(a).cuda() for a in sjob.ambience]
    sjob.pcm16 = torch.zeros((sjob.n_samples, sjob.n_channels), dtype=torch.int16, device="cuda")
    r = Renderer(0)
    r.render(jobs, [sjob])
    torch.cuda.synchronize()
    assert np.array_equal(sjob.pcm16.cpu().numpy(), wav_oracle.pcm16_from_float(sjob.mix.cpu().numpy()))
    r.close()


@pytest.mark.gpu
def test_gpu_generate_scenes_writes_the_dataset(tmp_path):
    import audiblelight_b200.synthesize as syn
    names = list(cases.SCENE_CASES)
    scenes = [_scene(cases.SCENE_CASES[n], idx=i, mics=("mic000", "mic001") if i == 0 else ("mic000",))
              for i, n in enumerate(names)]
    written = dataset.generate_scenes(scenes, tmp_path, batch_scenes=2,
                                      audio_fnames=[f"mix_{i}.wav" for i in range(len(scenes))],
                                      metadata_fnames=[f"meta_{i}" for i in range(len(scenes))])
    assert sorted(p.name for p in written[0]["audio"]) == ["mix_0_mic000.wav", "mix_0_mic001.wav"]
    assert [p.name for p in written[0]["csv"]] == ["meta_0_mic000.csv", "meta_0_mic001.csv"]
    assert [p.name for p in written[1]["json"]] == ["meta_1.json"]
    for i, n in enumerate(names):
        ref_scene = _scene(cases.SCENE_CASES[n], idx=i, mics=("mic000", "mic001") if i == 0 else ("mic000",))
        syn.render_scenes([ref_scene])
        for p in written[i]["audio"]:
            mic = p.stem.split("_")[-1]
            sr, pcm = dataset.read_wav_pcm16(p)
            assert sr == int(ref_scene.sample_rate)
            assert np.array_equal(pcm, wav_oracle.pcm16_from_float(ref_scene.audio[mic]))
        assert json.load(open(written[i]["json"][0]))["index"] == i
        rows = dataset.dcase2024_rows(ref_scene)
        for p in written[i]["csv"]:
            assert open(p, newline="").read() == dataset.dcase_csv(rows[p.stem.split("_")[-1]])
    # the pipelined driver (default; here 3 scenes in batches of 2 and of 1) writes the same bytes as the sequential one
    for bs in (1, 2):
        seq_dir = tmp_path / f"seq{bs}"
        seq_dir.mkdir()
        again = [_scene(cases.SCENE_CASES[n], idx=i, mics=("mic000", "mic001") if i == 0 else ("mic000",))
                 for i, n in enumerate(names)]
        w2 = dataset.generate_scenes(again, seq_dir, batch_scenes=bs, pipeline=(bs == 1),
                                     audio_fnames=[f"mix_{i}.wav" for i in range(len(scenes))],
                                     metadata_fnames=[f"meta_{i}" for i in range(len(scenes))])
        for a, b in zip(written, w2):
            for kind in ("audio", "json", "csv"):
                assert [p.name for p in a[kind]] == [p.name for p in b[kind]]
                for pa, pb in zip(a[kind], b[kind]):
                    assert open(pa, "rb").read() == open(pb, "rb").read()
    # keep_audio=True behaves like Scene.generate: results stay on the objects
    dataset.generate_scenes(scenes[:1], tmp_path, keep_audio=True, metadata_json=False, metadata_dcase=False)
    assert scenes[0].audio["mic001"].dtype == np.float32 and "mic000" in next(iter(scenes[0].events.values())).spatial_audio


@pytest.mark.gpu
def test_gpu_device_augmentation_of_event_objects(monkeypatch):
    """f1 drop-in wiring: Event.augmentations made of linear effects run on the device when enabled."""
    import audiblelight_b200.synthesize as syn
    from oracle import augment_oracle as ao

    def mk(name, sr, params):
        o = type(name, (), {})()
        o.sample_rate, o.params = sr, params
        return o
    spec = cases.EVENT_CASES["static_4ch"]
    raw, irs = cases.event_inputs(spec)
    raw = (0.3 * raw).astype(np.float32)
    augs = [mk("Fade", spec["sr"], dict(fade_in_len=0.05, fade_out_len=0.02, fade_in_shape="quarter_sine", fade_out_shape="linear")),
            mk("Invert", spec["sr"], {})]
    y = ao.peak_normalize(ao.invert(ao.fade(raw.astype(np.float64), spec["sr"], 0.05, 0.02, "quarter_sine", "linear")).astype(np.float32))

    def make_event(audio):
        ev = RefEvent(audio, spec["sr"], 1, spec["snr"])
        ev.is_audio_loaded = False
        return ev
    # host path: the event's own load_audio delivers the augmented audio
    host_ev = make_event(y.astype(np.float32))
    syn.render_event_audio(host_ev, irs, "mic000", ref_db=spec["ref_db"])
    # device path
    dev_ev = make_event(np.zeros(1, np.float32))
    dev_ev.augmentations = augs
    monkeypatch.setattr(syn, "_load_raw_audio", lambda e: raw)
    monkeypatch.setattr(syn, "DEVICE_AUGMENTATIONS", True)
    syn.render_event_audio(dev_ev, irs, "mic000", ref_db=spec["ref_db"])
    assert np.abs(dev_ev.audio - y).max() < 2e-6
    scale = np.abs(host_ev.spatial_audio["mic000"]).max()
    assert np.abs(dev_ev.spatial_audio["mic000"] - host_ev.spatial_audio["mic000"]).max() < 1e-5 * max(scale, 1.0)
    assert np.abs(dev_ev.spatial_audio["mic000"] - host_ev.spatial_audio["mic000"]).max() < 2e-5 * scale
    # an unsupported effect keeps the host path (load_audio is used, the raw loader is not called)
    monkeypatch.setattr(syn, "_load_raw_audio", lambda e: (_ for _ in ()).throw(AssertionError("raw loader used")))
    mixed_ev = make_event(y.astype(np.float32))
    mixed_ev.augmentations = augs + [mk("Compressor", spec["sr"], {})]
    syn.render_event_audio(mixed_ev, irs, "mic000", ref_db=spec["ref_db"])
    assert np.array_equal(mixed_ev.spatial_audio["mic000"], host_ev.spatial_audio["mic000"])
