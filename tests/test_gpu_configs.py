"""GPU parity at BASELINE.json's real sizes and on awkward shapes (pytest -m gpu).

Full-size cases are compared with the oracle's closed form (cross-checked against the reference's literal STFT loop in
tests/test_oracle.py); properties that hold at any size (superposition, equal-IR moving == static, impulse RIR ==
delay, scale invariance of the event gain) are checked on the raw convolution outputs."""
import numpy as np
import pytest

import cases
from audiblelight_b200 import workload as wl
from audiblelight_b200.renderer import ALR_GAIN_NONE, EventJob, Renderer, SceneJob, event_slice, moving_frames, scene_samples
from oracle import synth_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def rnd():
    r = Renderer(0)
    yield r
    r.close()


def _oracle_scene(spec, arrays64, amb):
    spatial, starts, ends = [], [], []
    for e, (x, h) in zip(spec.events, arrays64):
        dur = e.n_audio / float(spec.sr)
        if e.aug is not None:
            from oracle import augment_oracle as ao
            x = ao.peak_normalize(ao.biquad(x, *wl.aug_coeffs(e.aug, float(spec.sr)))).astype(np.float32)
        res = orc.render_event(x, h, e.snr, spec.ref_db, is_moving=e.n_irs > 1, duration=dur, sample_rate=float(spec.sr),
                               literal=False)
        spatial.append(res.spatial)
        starts.append(e.start)
        ends.append(e.start + dur)
    ambs = [(amb.astype(np.float64), spec.ref_db)] if amb is not None else []
    return spatial, orc.mix_scene(spec.duration, spec.sr, spatial, starts, ends, ambs)


def _run_scene(rnd, spec):
    arrays, amb = wl.host_scene_arrays(spec, dtype=np.float32)
    jobs, sj = wl.scene_jobs(spec, arrays, amb, 0)
    rnd.render(jobs, [sj])
    arrays64 = [(x, h.astype(np.float64)) for x, h in arrays]
    spatial, mix = _oracle_scene(spec, arrays64, amb)
    return jobs, sj, spatial, mix


def test_config1_quickstart_static_10s(rnd):
    """configs[0]: one static 10 s event at 24 kHz, 4-channel 1 s RIR."""
    jobs, sj, spatial, mix = _run_scene(rnd, wl.c1_scene_spec())
    assert jobs[0].spatial.shape == (4, 240000)
    assert np.abs(jobs[0].spatial - spatial[0]).max() <= TOL
    assert np.abs(sj.mix.astype(np.float64) - mix.scene).max() <= TOL


def test_config2_static_scene_60s(rnd):
    """configs[1]: 60 s @ 24 kHz, 9 static events + ambience."""
    jobs, sj, spatial, mix = _run_scene(rnd, wl.c2_scene_spec(3))
    assert sj.mix.shape == (4, 1440000)
    for j, s in zip(jobs, spatial):
        assert np.abs(j.spatial - s).max() <= TOL
    err = np.abs(sj.mix.astype(np.float64) - mix.scene).max()
    assert err <= TOL and err <= 2e-5 * np.abs(mix.scene).max()


def test_config3_moving_scene_60s(rnd):
    """configs[2] / the unit of configs[4]: 6 static + 3 moving events (one RIR per 100 ms, 1 s RIRs), 60 s scene."""
    spec = wl.c3_scene_spec(5)
    jobs, sj, spatial, mix = _run_scene(rnd, spec)
    assert sum(e.n_irs > 1 for e in spec.events) == 3
    for j, s in zip(jobs, spatial):
        err = np.abs(j.spatial - s).max()
        assert err <= TOL and err <= 3e-5 * np.abs(s).max()
    assert np.abs(sj.mix.astype(np.float64) - mix.scene).max() <= TOL


def test_config5_unit_scene_with_augmentations(rnd):
    """The unit of configs[4] as bench.py runs it: C3 scene + one seeded linear augmentation and peak normalisation per
    event on the device (f1)."""
    spec = wl.c3_scene_spec(9, augment=True)
    assert all(e.aug is not None for e in spec.events)
    jobs, sj, spatial, mix = _run_scene(rnd, spec)
    for j, s_ in zip(jobs, spatial):
        assert np.abs(j.spatial - s_).max() <= TOL
    assert np.abs(sj.mix.astype(np.float64) - mix.scene).max() <= TOL


def test_config4_em64_single_event(rnd):
    """configs[3] shape: 64 channels, 48 kHz, 2 s RIR, 10 s static event (one of the five, to bound the CPU time)."""
    rng = np.random.default_rng(64)
    x = cases.make_audio(rng, 480000)
    h = (rng.standard_normal((64, 1, 96000)) * np.exp(-np.arange(96000) / 16000.0)).astype(np.float32)
    job = EventJob(audio=x, irs=h, n_channels=64, snr=12.0, ref_db=-65.0)
    rnd.render([job])
    res = orc.render_event(x, h.astype(np.float64), 12.0, -65.0, is_moving=False)
    err = np.abs(job.spatial - res.spatial).max()
    assert job.spatial.shape == (64, 480000)
    assert err <= TOL and err <= 3e-5 * np.abs(res.spatial).max()


# ---- size-independent properties -----------------------------------------------------------------------------------------
def _raw(rnd, x, h, sr=24000, n_out=None, **kw):
    n = h.shape[1]
    job = EventJob(audio=np.ascontiguousarray(x, np.float32), irs=np.ascontiguousarray(h, np.float32), n_channels=h.shape[0],
                   normalize_irs=False, gain_mode=ALR_GAIN_NONE, n_out=n_out, **kw)
    if n > 1:
        job.ir_frames, job.n_frames = moving_frames(len(x) / float(sr), float(sr), n, len(x))
    rnd.render([job])
    return job.spatial


def test_superposition_full_size(rnd):
    rng = np.random.default_rng(1)
    x1, x2 = cases.make_audio(rng, 150001), cases.make_audio(rng, 150001)
    h = cases.make_irs(rng, 4, 31, 24000)
    y1, y2, y12 = _raw(rnd, x1, h), _raw(rnd, x2, h), _raw(rnd, x1 + x2, h)
    scale = np.abs(y12).max()
    assert np.abs(y12 - (y1 + y2)).max() < 5e-6 * scale


def test_moving_with_identical_irs_equals_static_times_512(rnd):
    rng = np.random.default_rng(2)
    x = cases.make_audio(rng, 100000)
    h1 = cases.make_irs(rng, 3, 1, 12000)
    hN = np.repeat(h1, 17, axis=1)
    ym = _raw(rnd, x, hN)
    ys = _raw(rnd, x, h1)
    fr, n_frames = moving_frames(100000 / 24000.0, 24000.0, 17, 100000)
    n_valid = n_frames * 128 - 256
    assert np.abs(ym[:, :n_valid] - 512.0 * ys[:, :n_valid]).max() < 5e-6 * np.abs(ym).max()
    assert np.all(ym[:, n_valid:] == 0)


def test_impulse_rir_is_a_delay(rnd):
    rng = np.random.default_rng(3)
    x = cases.make_audio(rng, 50000)
    h = np.zeros((2, 1, 5000), np.float32)
    h[0, 0, 1234] = 1.0
    h[1, 0, 4999] = -0.5
    y = _raw(rnd, x, h, n_out=50000 + 5000 - 1)
    ref0 = np.zeros(54999); ref0[1234:1234 + 50000] = x
    ref1 = np.zeros(54999); ref1[4999:4999 + 50000] = -0.5 * x
    assert np.abs(y[0] - ref0).max() < 2e-6 and np.abs(y[1] - ref1).max() < 2e-6


def test_event_gain_is_scale_invariant(rnd):
    """render_event_audio's result does not depend on the overall scale of the RIRs (apply_snr + db_to_multiplier)."""
    rng = np.random.default_rng(4)
    x = cases.make_audio(rng, 30000)
    h = cases.make_irs(rng, 4, 1, 8000).astype(np.float32)
    a = EventJob(audio=x, irs=h, n_channels=4, snr=9.0, ref_db=-60.0)
    b = EventJob(audio=x, irs=(h * 37.5).astype(np.float32), n_channels=4, snr=9.0, ref_db=-60.0)
    rnd.render([a, b])
    assert np.abs(a.spatial - b.spatial).max() < 2e-6 * np.abs(a.spatial).max()
    m = np.abs(a.spatial).mean()
    assert np.isclose(m, 10 ** ((-60.0 + 9.0) / 20.0), rtol=1e-4)   # mean |y| lands on ref_db + snr


# ---- awkward shapes -----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lx,lh,c,n", [(100, 7, 1, 1), (1023, 1025, 3, 1), (1024, 1024, 5, 1), (1025, 3000, 2, 4),
                                       (5000, 100, 7, 40), (2049, 2047, 6, 2), (300, 5000, 4, 3), (70000, 999, 9, 1)])
def test_awkward_shapes_vs_oracle(rnd, lx, lh, c, n):
    rng = np.random.default_rng(lx * 7 + lh + c + n)
    x = cases.make_audio(rng, lx)
    h = cases.make_irs(rng, c, n, lh)
    job = EventJob(audio=x, irs=h.astype(np.float32), n_channels=c, snr=11.0, ref_db=-65.0)
    if n > 1:
        job.ir_frames, job.n_frames = moving_frames(lx / 24000.0, 24000.0, n, lx)
    rnd.render([job])
    res = orc.render_event(x, h, 11.0, -65.0, is_moving=n > 1, duration=lx / 24000.0, sample_rate=24000.0, literal=False)
    err = np.abs(job.spatial - res.spatial).max()
    assert err <= TOL and err <= 5e-5 * max(np.abs(res.spatial).max(), 1e-30) + 1e-12


def test_strided_ir_views_device_and_host(rnd):
    """The reference hands per-event slices mic_ir[:, k:k+N, :] of one big array (synthesize.py:662)."""
    import torch
    rng = np.random.default_rng(8)
    big = cases.make_irs(rng, 4, 7, 3000).astype(np.float32)
    x = cases.make_audio(rng, 9000)
    views = [big[:, 0:1, :], big[:, 1:4, :], big[:, 4:7, :]]
    ref = []
    for v in views:
        j = EventJob(audio=x, irs=np.ascontiguousarray(v), n_channels=4, snr=10.0, ref_db=-65.0)
        if v.shape[1] > 1:
            j.ir_frames, j.n_frames = moving_frames(9000 / 24000.0, 24000.0, v.shape[1], 9000)
        rnd.render([j])
        ref.append(j.spatial.copy())
    jobs = []
    for v in views:
        j = EventJob(audio=x, irs=v, n_channels=4, snr=10.0, ref_db=-65.0)
        if v.shape[1] > 1:
            j.ir_frames, j.n_frames = moving_frames(9000 / 24000.0, 24000.0, v.shape[1], 9000)
        jobs.append(j)
    rnd.render(jobs)
    for j, r in zip(jobs, ref):
        assert np.array_equal(j.spatial, r)
    big_d, x_d = torch.from_numpy(big).cuda(), torch.from_numpy(x).cuda()
    jobs_d = []
    for lo, hi in [(0, 1), (1, 4), (4, 7)]:
        j = EventJob(audio=x_d, irs=big_d[:, lo:hi, :], n_channels=4, snr=10.0, ref_db=-65.0)
        if hi - lo > 1:
            j.ir_frames, j.n_frames = moving_frames(9000 / 24000.0, 24000.0, hi - lo, 9000)
        jobs_d.append(j)
    rnd.render(jobs_d)
    for j, r in zip(jobs_d, ref):
        assert np.array_equal(j.spatial.cpu().numpy(), r)


def test_misaligned_buffers(rnd):
    """Odd element offsets of audio / RIR / output buffers (no vector-alignment assumptions on caller memory)."""
    rng = np.random.default_rng(9)
    xa = np.zeros(7001, np.float32); xa[1:] = cases.make_audio(rng, 7000)
    ha = np.zeros(3 * 2 * 1501 + 1, np.float32)
    h = cases.make_irs(rng, 3, 2, 1501).astype(np.float32)
    ha[1:] = h.ravel()
    out = np.zeros(3 * 7000 + 1, np.float32)
    job = EventJob(audio=xa[1:], irs=ha[1:].reshape(3, 2, 1501), n_channels=3, snr=10.0, ref_db=-65.0,
                   spatial=out[1:].reshape(3, 7000))
    job.ir_frames, job.n_frames = moving_frames(7000 / 24000.0, 24000.0, 2, 7000)
    rnd.render([job])
    res = orc.render_event(xa[1:], h.astype(np.float64), 10.0, -65.0, is_moving=True, duration=7000 / 24000.0,
                           sample_rate=24000.0, literal=False)
    assert np.abs(job.spatial - res.spatial).max() <= TOL


def test_nonfinite_input_is_flagged(rnd):
    rng = np.random.default_rng(10)
    x = cases.make_audio(rng, 4000)
    h = cases.make_irs(rng, 2, 1, 900).astype(np.float32)
    h[1, 0, 10] = np.nan
    job = EventJob(audio=x, irs=h, n_channels=2, snr=10.0, ref_db=-65.0)
    rnd.render([job])
    assert job.stats["nonfinite"]


def test_many_events_one_call_matches_individual_calls(rnd):
    """Batching / chunking must not change results: 40 mixed events in one call == 40 single calls."""
    rng = np.random.default_rng(11)
    jobs, singles = [], []
    for i in range(40):
        lx = int(rng.integers(500, 9000)); lh = int(rng.integers(50, 4000)); n = int(rng.choice([1, 1, 2, 5]))
        x = cases.make_audio(rng, lx); h = cases.make_irs(rng, 4, n, lh).astype(np.float32)
        def mk():
            j = EventJob(audio=x, irs=h, n_channels=4, snr=10.0 + i % 7, ref_db=-65.0)
            if n > 1:
                j.ir_frames, j.n_frames = moving_frames(lx / 24000.0, 24000.0, n, lx)
            return j
        jobs.append(mk()); singles.append(mk())
    rnd.render(jobs)
    small = Renderer(0, workspace_limit=1 << 18)
    for s in singles:
        small.render([s])
    small.close()
    for a, b in zip(jobs, singles):
        assert np.array_equal(a.spatial, b.spatial)
