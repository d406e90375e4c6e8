"""f3: Gaussian ambience generated on the device (alr_scene.ambience_seed). The reference draws this noise from numpy's
unseeded global generator (ambience.py:155-163), so parity is distributional: i.i.d. N(0, 1) samples, independent
channels, per-channel peak normalisation (:210-214), and the mixdown scale of generate_scene_audio_from_events
(synthesize.py:350-352) applied to it exactly as to an uploaded layer."""
import numpy as np
import pytest

from audiblelight_b200.renderer import EventJob, Renderer, SceneJob
from oracle import synth_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rnd():
    r = Renderer(0)
    yield r
    r.close()


def _gen(rnd, C, T, seed, ref_db=-50.0):
    sj = SceneJob(n_channels=C, n_samples=T, ambience=[None], ambience_ref_db=[ref_db], ambience_seed=[seed],
                  mix=np.zeros((C, T), np.float32))
    rnd.render([], [sj])
    return sj.mix


def test_generated_ambience_statistics(rnd):
    C, T = 4, 1440000
    mix = _gen(rnd, C, T, seed=1234).astype(np.float64)
    # the mix of a scene with one ambience layer is scale * noise with mean|.| == 10^(ref_db / 20)
    assert np.isclose(np.abs(mix).mean(), 10 ** (-50.0 / 20.0), rtol=1e-5)
    x = mix / mix.std(axis=1, keepdims=True)
    assert np.all(np.abs(x.mean(axis=1)) < 4.0 / np.sqrt(T))
    kurt = (x ** 4).mean(axis=1)
    assert np.all(np.abs(kurt - 3.0) < 0.05)                       # Gaussian: E x^4 = 3
    assert np.all(np.abs((x ** 3).mean(axis=1)) < 0.02)            # symmetric
    lag1 = (x[:, 1:] * x[:, :-1]).mean(axis=1)
    assert np.all(np.abs(lag1) < 5.0 / np.sqrt(T))                 # white
    cc = np.corrcoef(x)
    assert np.all(np.abs(cc - np.eye(C)) < 5.0 / np.sqrt(T))       # channels independent
    # same peak / mean ratio as numpy's generator gives after the per-channel peak normalisation
    ref = np.random.default_rng(0).standard_normal((C, T))
    ref /= np.abs(ref).max(axis=1, keepdims=True)
    r_ref = (np.abs(ref).max(axis=1) / np.abs(ref).mean(axis=1)).mean()
    r_dev = (np.abs(mix).max(axis=1) / np.abs(mix).mean(axis=1)).mean()
    assert abs(r_dev / r_ref - 1.0) < 0.08
    # flat spectrum: band powers within a few percent of each other
    spec = np.abs(np.fft.rfft(x[0])) ** 2
    bands = spec[1:1 + 16 * (len(spec) // 16)].reshape(16, -1).mean(axis=1)
    assert bands.max() / bands.min() < 1.03


def test_generated_ambience_is_seeded_and_layout_independent(rnd):
    a = _gen(rnd, 3, 100003, seed=7)
    b = _gen(rnd, 3, 100003, seed=7)
    c = _gen(rnd, 3, 100003, seed=8)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    # a longer layer with the same seed starts with the same raw stream (normalisation differs by the channel peak)
    d = _gen(rnd, 3, 200000, seed=7)
    ratio = a[:, :1000] / d[:, :1000]
    assert np.allclose(ratio, ratio[:, :1], rtol=1e-5)


def test_generated_layer_mixes_like_an_uploaded_one(rnd):
    """Generate a layer, read it back through a mix-only scene, then feed the SAME samples as an ordinary input layer
    next to an event: identical mixdown to the scene that generates the layer in place; device buffers == host buffers."""
    import torch
    C, T = 4, 48000
    rng = np.random.default_rng(5)
    noise = _gen(rnd, C, T, seed=99, ref_db=0.0)           # scale * normalised noise
    noise = noise / np.abs(noise).max(axis=1, keepdims=True)  # back to the peak-normalised layer (up to float rounding)
    x = rng.standard_normal(20000).astype(np.float32)
    x /= np.abs(x).max()
    h = (rng.standard_normal((C, 1, 3000)) * np.exp(-np.arange(3000) / 500.0)).astype(np.float32)

    def scene(amb, seeds):
        ev = EventJob(audio=x, irs=h, n_channels=C, snr=10.0, ref_db=-65.0, scene=0, scene_start=5000, scene_end=25000)
        sj = SceneJob(n_channels=C, n_samples=T, ambience=amb, ambience_ref_db=[-65.0], ambience_seed=seeds)
        rnd.render([ev], [sj])
        return sj.mix
    m_gen = scene([None], [99])
    m_in = scene([noise], ())
    assert np.abs(m_gen - m_in).max() <= 2e-7
    # device-resident call
    ev = EventJob(audio=torch.from_numpy(x).cuda(), irs=torch.from_numpy(h).cuda(), n_channels=C, snr=10.0, ref_db=-65.0,
                  scene=0, scene_start=5000, scene_end=25000)
    sj = SceneJob(n_channels=C, n_samples=T, ambience=[None], ambience_ref_db=[-65.0], ambience_seed=[99])
    rnd.render([ev], [sj])
    assert np.array_equal(sj.mix.cpu().numpy(), m_gen)
    # oracle on the read-back layer
    res = orc.render_event(x, h.astype(np.float64), 10.0, -65.0, is_moving=False)
    mix = orc.mix_scene(T / 24000.0, 24000, [res.spatial], [5000 / 24000.0], [25000 / 24000.0], [(noise.astype(np.float64), -65.0)])
    assert np.abs(m_gen.astype(np.float64) - mix.scene).max() <= 1e-5


def test_missing_seed_is_an_error(rnd):
    sj = SceneJob(n_channels=2, n_samples=1000, ambience=[None], ambience_ref_db=[-60.0])
    with pytest.raises(Exception):
        rnd.render([], [sj])


def test_dropin_device_ambience_opt_in():
    """audiblelight_b200.synthesize with DEVICE_AMBIENCE: a Gaussian Ambience object that has not been loaded is drawn on
    the device; a loaded one (or any other kind) is uploaded as before."""
    from collections import OrderedDict
    from audiblelight_b200 import synthesize as syn
    from audiblelight_b200 import workload as wl

    class GaussAmbience:
        def __init__(self, channels, ref_db):
            self.beta, self.audio, self.channels, self.ref_db, self.loaded = "gaussian", None, channels, ref_db, 0

        def load_ambience(self, ignore_cache=False, normalize=True):
            self.loaded += 1
            raise AssertionError("the host generator must not run for a device-generated layer")

    spec = wl.c3_scene_spec(2, duration=20.0, n_static=2, n_moving=1)
    scene = wl.SynScene(2, spec)
    amb = GaussAmbience(spec.channels, -60.0)
    scene.ambience = OrderedDict(amb=amb)
    old = syn.DEVICE_AMBIENCE
    syn.DEVICE_AMBIENCE = True
    try:
        syn.render_scenes([scene])
    finally:
        syn.DEVICE_AMBIENCE = old
    mix = scene.audio["mic000"].astype(np.float64)
    assert mix.shape == (spec.channels, round(20.0 * spec.sr)) and amb.loaded == 0 and amb.audio is None
    # ambience floor dominates outside the events: mean|mix| there == 10^(ref_db / 20)
    quiet = np.ones(mix.shape[1], bool)
    for ev in scene.events.values():
        quiet[max(0, round(ev.scene_start * spec.sr)):round(ev.scene_end * spec.sr)] = False
    if quiet.sum() > 20000:
        assert np.isclose(np.abs(mix[:, quiet]).mean(), 10 ** (-60.0 / 20.0), rtol=0.02)
