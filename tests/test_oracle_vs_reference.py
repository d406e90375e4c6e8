"""Live comparison of the oracle with the UNMODIFIED reference (only where /root/reference exists, i.e. the build
container; skipped on the GPU box). Complements the committed golden vectors with randomised shapes."""
import numpy as np
import pytest

import cases
import ref_loader
from oracle import synth_oracle as orc

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def syn():
    return ref_loader.load_reference_synthesize()


@pytest.mark.parametrize("seed", range(6))
def test_render_event_fuzz(syn, seed):
    rng = np.random.default_rng(100 + seed)
    sr = int(rng.choice([16000, 24000, 44100, 48000]))
    lx = int(rng.integers(800, 7000))
    lh = int(rng.integers(50, 3000))
    c = int(rng.integers(1, 5))
    n = int(rng.choice([1, 1, 2, 3, 6]))
    snr = float(rng.uniform(-10, 30))
    ref_db = float(rng.integers(-80, -50))
    audio = cases.make_audio(rng, lx)
    irs = cases.make_irs(rng, c, n, lh)
    ev = ref_loader.RefEvent(audio, sr, n, snr)
    syn.render_event_audio(ev, irs, "m", ref_db=ref_db)
    want = ev.spatial_audio["m"]
    for literal in (True, False):
        got = orc.render_event(audio, irs, snr, ref_db, is_moving=n > 1, duration=lx / float(sr), sample_rate=float(sr),
                               literal=literal).spatial
        assert got.shape == want.shape
        assert np.abs(got - want).max() < 1e-11


def test_literal_port_matches_reference_loop_bitwise(syn):
    rng = np.random.default_rng(5)
    x = cases.make_audio(rng, 5000)
    irs = cases.make_irs(rng, 3, 7, 2000)
    S, X = syn.stft(irs, 512, 256, 128), syn.stft(x, 512, 256, 128)
    w = syn.generate_interpolation_matrix(np.linspace(0, 5000 / 24000, 7), 24000.0, 128)
    a = syn.perform_time_variant_convolution(X, S, w)
    b = orc.ctf_convolve(X, S, w)
    assert np.abs(a - b).max() <= 1e-12 * np.abs(a).max()
    assert np.array_equal(syn.istft_overlap_synthesis(a, 512, 256, 128), orc.istft_ola(a))


def test_interpolation_matrix_fuzz(syn):
    rng = np.random.default_rng(9)
    for _ in range(40):
        sr = float(rng.choice([16000, 24000, 44100, 48000]))
        dur = float(rng.uniform(0.05, 4.0))
        n = int(rng.integers(2, 60))
        t = np.linspace(0, dur, n)
        assert np.array_equal(syn.generate_interpolation_matrix(t, sr, 128), orc.interpolation_matrix(t, sr, 128))
