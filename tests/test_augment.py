"""Scope row f1 — linear event augmentations (audiblelight/augmentation.py) on the device.

CPU: the oracle restatement of Fade / Invert / Reverse against golden vectors of the unmodified reference, and
sanity of the (unpinned) filter formulas. GPU: every op, chains, peak normalisation and an end-to-end render through
the C-ABI against the oracle."""
import os

import numpy as np
import pytest

TOL = 1e-5  # BASELINE.json north_star: max-abs error <= 1e-5 of full scale, also for event.audio after the device-side chain

import cases
from audiblelight_b200 import augment as A
from oracle import augment_oracle as ao
from oracle import synth_oracle as orc

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(G, "augment.npz"))


# ---- CPU -------------------------------------------------------------------------------------------------------------
def test_oracle_fade_invert_reverse_match_reference(gold):
    x = gold["x"]
    for i, (sr, fi, fo, si, so) in enumerate(gold["fade_cases"]):
        got = ao.fade(x, sr, fi, fo, ao.FADE_SHAPES[int(si)], ao.FADE_SHAPES[int(so)])
        assert np.array_equal(got, gold[f"fade_{i}"])          # bit-exact restatement
    assert np.array_equal(ao.invert(x), gold["invert"])
    assert np.array_equal(ao.reverse(x), gold["reverse"])


def test_fade_endpoints():
    # tests/test_augmentation.py:303-330 of the reference: a full-length linear fade-in starts at 0 and ends at 1
    x = np.ones(1000)
    y = ao.fade(x, 1000, 1.0, 0.0, "linear", "none")
    assert y[0] == 0.0 and y[-1] == 1.0
    y = ao.fade(x, 1000, 0.0, 1.0, "none", "linear")
    assert y[0] == 1.0 and y[-1] == 0.0


def test_filter_formulas_sanity():
    sr = 24000.0
    w0 = 0.0
    def resp(b, a, w):
        z = np.exp(-1j * w * np.arange(3))
        b = np.array(list(b) + [0] * (3 - len(b))); a = np.array(list(a) + [0] * (3 - len(a)))
        return abs((b * z).sum() / (a * z).sum())
    b, a = A.lowpass_coeffs(sr, 3000.0)
    assert np.isclose(resp(b, a, 0.0), 1.0) and resp(b, a, np.pi) < 1e-9
    assert np.isclose(resp(b, a, 2 * np.pi * 3000 / sr), 2 ** -0.5, rtol=1e-6)      # -3 dB at the cutoff
    b, a = A.highpass_coeffs(sr, 500.0)
    assert resp(b, a, 0.0) < 1e-12 and np.isclose(resp(b, a, np.pi), 1.0)
    b, a = A.low_shelf_coeffs(sr, 400.0, -12.0, 0.7)
    assert np.isclose(resp(b, a, 0.0), 10 ** (-12 / 20), rtol=1e-9) and np.isclose(resp(b, a, np.pi), 1.0)
    b, a = A.high_shelf_coeffs(sr, 4000.0, 6.0, 0.7)
    assert np.isclose(resp(b, a, np.pi), 10 ** (6 / 20), rtol=1e-9) and np.isclose(resp(b, a, 0.0), 1.0)
    b, a = A.peak_coeffs(sr, 2000.0, 9.0, 1.0)
    assert np.isclose(resp(b, a, 2 * np.pi * 2000 / sr), 10 ** (9 / 20), rtol=1e-9)
    assert np.isclose(resp(b, a, 0.0), 1.0) and np.isclose(resp(b, a, np.pi), 1.0)
    assert w0 == 0.0


def test_deemphasis_inverts_preemphasis():
    x = np.random.default_rng(0).standard_normal(3000)
    for coef in (0.0, 0.3, 0.97):
        assert np.abs(ao.deemphasis(ao.preemphasis(x, coef), coef) - x).max() < 1e-9


def test_delay_oracle_is_a_feedback_comb():
    # impulse in -> dry tap at 0, wet taps mix * feedback^(k-1) at k * D
    x = np.zeros(50)
    x[0] = 1.0
    y = ao.delay(x, 10.0, 1.2, 0.5, 0.25)  # D = 12
    want = np.zeros(50)
    want[0] = 0.75
    for k in range(1, 5):
        want[12 * k] = 0.25 * 0.5 ** (k - 1)
    assert np.allclose(y, want, atol=1e-15)
    assert np.array_equal(ao.delay(x, 10.0, 0.0, 0.0, 0.3), x)
    op = A.delay(24000.0, 0.0123, 0.4, 0.3)
    assert op.type == A.ALR_AUG_DELAY and op.p[0] == float(int(0.0123 * 24000.0))
    with pytest.raises(ValueError):
        A.delay(24000.0, 0.1, 1.5, 0.3)


# ---- GPU -------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def rnd():
    from audiblelight_b200.renderer import Renderer
    r = Renderer(0)
    yield r
    r.close()


def _augment_only(rnd, x, ops, normalize=False):
    """Runs the device augmentation and returns the dry audio it produced (event rendered without IRs)."""
    from audiblelight_b200.renderer import EventJob
    out = np.zeros(len(x), np.float32)
    job = EventJob(audio=np.ascontiguousarray(x, np.float32), irs=None, n_channels=1, snr=1.0, ref_db=0.0, aug_ops=ops,
                   normalize_audio=normalize, audio_out=out)
    rnd.render([job])
    return out


@pytest.mark.gpu
def test_gpu_fade_invert_reverse_golden(rnd, gold):
    x = gold["x"]
    for i, (sr, fi, fo, si, so) in enumerate(gold["fade_cases"]):
        got = _augment_only(rnd, x, [A.fade(sr, fi, fo, A.FADE_SHAPES[int(si)], A.FADE_SHAPES[int(so)])])
        assert np.abs(got - gold[f"fade_{i}"]).max() < 2e-6 * np.abs(x).max()
    assert np.array_equal(_augment_only(rnd, x, [A.invert()]), gold["invert"])
    assert np.array_equal(_augment_only(rnd, x, [A.reverse()]), gold["reverse"])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [100, 511, 512, 513, 6000, 100001])
def test_gpu_iir_filters_vs_lfilter(rnd, n):
    sr = 24000.0
    x = np.random.default_rng(n).standard_normal(n).astype(np.float32)
    specs = [A.lowpass_coeffs(sr, 6000.0), A.highpass_coeffs(sr, 100.0), A.low_shelf_coeffs(sr, 300.0, -15.0, 0.4),
             A.high_shelf_coeffs(sr, 5000.0, 8.0, 0.9), A.peak_coeffs(sr, 1500.0, -9.0, 2.0)]
    for b, a in specs:
        got = _augment_only(rnd, x, [A.biquad(b, a)])
        want = ao.biquad(x, b, a)
        assert np.abs(got - want).max() <= TOL * max(1.0, np.abs(want).max())
    got = _augment_only(rnd, x, [A.gain_db(-7.5)])
    assert np.abs(got - ao.gain_db(x.astype(np.float64), -7.5)).max() < 1e-6 * np.abs(x).max()
    for coef in (0.2, 0.97):
        assert np.abs(_augment_only(rnd, x, [A.preemphasis(coef)]) - ao.preemphasis(x, coef)).max() < 1e-5
        assert np.abs(_augment_only(rnd, x, [A.deemphasis(coef)]) - ao.deemphasis(x, coef)).max() <= TOL * max(1.0, np.abs(ao.deemphasis(x, coef)).max())


@pytest.mark.gpu
@pytest.mark.parametrize("n,delay_s", [(100, 0.001), (5000, 0.01), (5000, 0.5), (100001, 0.0371), (3000, 0.0)])
def test_gpu_delay_vs_oracle(rnd, n, delay_s):
    sr = 24000.0
    x = np.random.default_rng(n).standard_normal(n).astype(np.float32)
    got = _augment_only(rnd, x, [A.delay(sr, delay_s, 0.45, 0.35)])
    want = ao.delay(x, sr, delay_s, 0.45, 0.35)
    assert np.abs(got - want).max() < 2e-6 * np.abs(want).max()


@pytest.mark.gpu
def test_gpu_chain_and_normalise(rnd):
    sr = 24000.0
    x = np.random.default_rng(5).standard_normal(30000).astype(np.float32)
    ops = [A.fade(sr, 0.1, 0.2, "half_sine", "exponential"), A.highpass(sr, 200.0)] + \
        A.multiband_equalizer(sr, [(1200.0, 6.0, 1.0), (5000.0, -4.0, 0.7)]) + [A.invert(), A.gain_db(3.0)]
    got = _augment_only(rnd, x, ops, normalize=True)
    y = ao.fade(x.astype(np.float64), sr, 0.1, 0.2, "half_sine", "exponential")
    y = ao.biquad(y, *A.highpass_coeffs(sr, 200.0))
    y = ao.biquad(y, *A.peak_coeffs(sr, 1200.0, 6.0, 1.0))
    y = ao.biquad(y, *A.peak_coeffs(sr, 5000.0, -4.0, 0.7))
    y = ao.peak_normalize(ao.gain_db(ao.invert(y), 3.0))
    assert np.isclose(np.abs(got).max(), 1.0, atol=1e-6)
    assert np.abs(got - y).max() <= TOL


@pytest.mark.gpu
def test_gpu_render_with_augmentation_equals_render_of_augmented_audio(rnd):
    """Event.load_audio semantics: augment -> peak-normalise -> render_event_audio (event.py:530-536 + synthesize.py:551)."""
    from audiblelight_b200.renderer import EventJob, moving_frames
    sr = 24000.0
    rng = np.random.default_rng(6)
    raw = (0.3 * rng.standard_normal(20000)).astype(np.float32)
    irs = cases.make_irs(rng, 4, 5, 3000)
    ops = [A.lowpass(sr, 7000.0), A.fade(sr, 0.05, 0.05, "linear", "linear")]
    job = EventJob(audio=raw, irs=irs.astype(np.float32), n_channels=4, snr=14.0, ref_db=-65.0, aug_ops=ops, normalize_audio=True)
    job.ir_frames, job.n_frames = moving_frames(20000 / sr, sr, 5, 20000)
    rnd.render([job])
    y = ao.fade(ao.biquad(raw, *A.lowpass_coeffs(sr, 7000.0)), sr, 0.05, 0.05, "linear", "linear")
    y = ao.peak_normalize(y).astype(np.float32)
    res = orc.render_event(y, irs, 14.0, -65.0, is_moving=True, duration=20000 / sr, sample_rate=sr, literal=False)
    assert np.abs(job.spatial - res.spatial).max() <= 1e-5


@pytest.mark.gpu
def test_gpu_normalise_only_and_silent_audio(rnd):
    x = np.random.default_rng(7).standard_normal(5000).astype(np.float32) * 0.01
    got = _augment_only(rnd, x, [], normalize=True)
    assert np.abs(got - ao.peak_normalize(x)).max() < 1e-6
    assert np.all(_augment_only(rnd, np.zeros(3000, np.float32), [A.gain_db(6.0)], normalize=True) == 0)   # silent file -> no NaN


def _random_chain(rng, sr, n):
    """(device ops, oracle function) for a random chain of 1-6 linear effects."""
    ops, fns = [], []
    for _ in range(int(rng.integers(1, 7))):
        kind = rng.choice(["gain", "invert", "reverse", "fade", "lowpass", "highpass", "low_shelf", "high_shelf", "peak",
                           "preemphasis", "deemphasis", "delay"])
        if kind == "gain":
            db = float(rng.uniform(-12, 6))
            ops.append(A.gain_db(db)); fns.append(lambda x, db=db: ao.gain_db(x, db))
        elif kind == "invert":
            ops.append(A.invert()); fns.append(ao.invert)
        elif kind == "reverse":
            ops.append(A.reverse()); fns.append(ao.reverse)
        elif kind == "fade":
            fi, fo = float(rng.uniform(0, 1.2 * n / sr)), float(rng.uniform(0, 1.2 * n / sr))
            si, so = (A.FADE_SHAPES[int(rng.integers(0, 6))] for _ in range(2))
            ops.append(A.fade(sr, fi, fo, si, so)); fns.append(lambda x, a=(fi, fo, si, so): ao.fade(x, sr, *a))
        elif kind in ("lowpass", "highpass"):
            fc = float(rng.uniform(50.0, 0.45 * sr))
            b, a = (A.lowpass_coeffs if kind == "lowpass" else A.highpass_coeffs)(sr, fc)
            ops.append(A.biquad(b, a)); fns.append(lambda x, b=b, a=a: ao.biquad(x, b, a))
        elif kind in ("low_shelf", "high_shelf", "peak"):
            fc, g, q = float(rng.uniform(200.0, 0.4 * sr)), float(rng.uniform(-15, 9)), float(rng.uniform(0.3, 2.0))
            b, a = {"low_shelf": A.low_shelf_coeffs, "high_shelf": A.high_shelf_coeffs, "peak": A.peak_coeffs}[kind](sr, fc, g, q)
            ops.append(A.biquad(b, a)); fns.append(lambda x, b=b, a=a: ao.biquad(x, b, a))
        elif kind == "preemphasis":
            c = float(rng.uniform(0.0, 0.97))
            ops.append(A.preemphasis(c)); fns.append(lambda x, c=c: ao.preemphasis(x, c))
        elif kind == "deemphasis":
            c = float(rng.uniform(0.0, 0.9))
            ops.append(A.deemphasis(c)); fns.append(lambda x, c=c: ao.deemphasis(x, c))
        else:
            d, fb, mix = float(rng.uniform(0.0, 1.5 * n / sr)), float(rng.uniform(0.0, 0.6)), float(rng.uniform(0.0, 1.0))
            ops.append(A.delay(sr, d, fb, mix)); fns.append(lambda x, a=(d, fb, mix): ao.delay(x, sr, *a))

    def run(x):
        y = x.astype(np.float64)
        for f in fns:
            y = np.asarray(f(y), dtype=np.float64)
        return y
    return ops, run


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(8))
def test_gpu_random_augmentation_chains(rnd, seed):
    """Several events with random chains and ragged lengths in ONE call (the ops of one level share launches)."""
    from audiblelight_b200.renderer import EventJob
    rng = np.random.default_rng(300 + seed)
    sr = float(rng.choice([8000.0, 24000.0, 44100.0]))
    jobs, wants = [], []
    for _ in range(int(rng.integers(2, 9))):
        n = int(rng.choice([2, 3, 31, 32, 33, 511, 512, 513, 1024, 1500, 5000, 20001]))
        x = rng.standard_normal(n).astype(np.float32)
        ops, run = _random_chain(rng, sr, n)
        norm = bool(rng.random() < 0.5)
        y = run(x)
        wants.append(ao.peak_normalize(y) if norm else y)
        jobs.append(EventJob(audio=x, irs=None, n_channels=1, snr=1.0, ref_db=0.0, aug_ops=ops, normalize_audio=norm,
                             audio_out=np.zeros(n, np.float32)))
    rnd.render(jobs)
    for j, want in zip(jobs, wants):
        scale = max(1.0, np.abs(want).max())
        assert np.abs(j.audio_out - want).max() <= TOL * scale, (seed, [o.type for o in j.aug_ops], len(want))
