"""Round-2 parity cases (pytest -m gpu): the named configurations in full (all five em64 events + mix, two microphones,
the 128-scene batch bench.py times), run-to-run determinism, and the persistent producer/consumer launch
(k_mov_fused) against the default kernels and the oracle."""
import numpy as np
import pytest

import cases
from audiblelight_b200 import workload as wl
from audiblelight_b200.renderer import EventJob, Renderer, SceneJob, event_slice, moving_frames, scene_samples
from oracle import synth_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5  # BASELINE.json north_star: max-abs error <= 1e-5 of full scale (fp32)


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


def test_config4_em64_all_five_events_and_mix(rnd):
    """configs[3] in full: Eigenmike em64 (64 channels), 48 kHz, 2 s RIRs, five overlapping 10 s static events, 30 s mix."""
    spec = wl.c4_scene_spec(0)
    arrays, amb = wl.host_scene_arrays(spec, dtype=np.float32)
    jobs, sj = wl.scene_jobs(spec, arrays, amb, 0)
    rnd.render(jobs, [sj])
    assert sj.mix.shape == (64, 1440000) and len(jobs) == 5
    spatial, mix = _oracle_scene(spec, [(x, h.astype(np.float64)) for x, h in arrays], amb)
    for j, s in zip(jobs, spatial):
        err = np.abs(j.spatial - s).max()
        assert err <= TOL and err <= 3e-5 * np.abs(s).max()
    assert np.abs(sj.mix.astype(np.float64) - mix.scene).max() <= TOL


def test_config2_two_microphones_full_size(rnd):
    """configs[1] as written: FOA + MIC = two 4-channel arrays with independent RIR sets, 9 static events, ambience per
    microphone, 60 s at 24 kHz; both (scene, microphone) mixdowns in ONE call, as render_scenes issues them."""
    spec = wl.c2_scene_spec(7)
    T = scene_samples(spec.duration, spec.sr)
    all_jobs, scenes, oracle = [], [], []
    for mic in range(2):
        sp = wl.c2_scene_spec(7)
        sp.index = 700 + mic  # independent RIRs / ambience per microphone (same events, same dry audio lengths)
        arrays, amb = wl.host_scene_arrays(sp, dtype=np.float32)
        if mic == 1:  # the dry audio of an event is the same for both microphones
            arrays = [(x0, h) for (x0, _), (_, h) in zip(first_arrays, arrays)]
        else:
            first_arrays = arrays
        jobs, sj = wl.scene_jobs(sp, arrays, amb, mic)
        all_jobs += jobs
        scenes.append(sj)
        oracle.append(_oracle_scene(sp, [(x, h.astype(np.float64)) for x, h in arrays], amb))
    rnd.render(all_jobs, scenes)
    for mic in range(2):
        spatial, mix = oracle[mic]
        for j, s in zip(all_jobs[9 * mic:9 * mic + 9], spatial):
            assert np.abs(j.spatial - s).max() <= TOL
        assert scenes[mic].mix.shape == (4, T)
        assert np.abs(scenes[mic].mix.astype(np.float64) - mix.scene).max() <= TOL
    assert not np.array_equal(scenes[0].mix, scenes[1].mix)


def test_c5_batch_of_128_scenes_equals_per_scene_renders():
    """The batch bench.py times (128 C5 scenes per GPU, device-resident, one alr_render call) against the same scenes
    rendered one call per scene: bit-identical events and mixes; two of the scenes also against the oracle."""
    import torch
    dev = torch.device("cuda", 0)
    S = 128
    specs = [wl.c3_scene_spec(i, augment=True) for i in range(S)]
    r = Renderer(0)
    jobs, scenes, per_scene = [], [], []
    for si, sp in enumerate(specs):
        arrays, amb = wl.device_scene_arrays(sp, dev)
        j, sj = wl.scene_jobs(sp, arrays, amb, si)
        jobs += j
        scenes.append(sj)
        per_scene.append((arrays, amb))
    r.render(jobs, scenes)
    ev_of = [jobs[9 * si:9 * si + 9] for si in range(S)]
    check = list(range(0, S, 8)) + [S - 1]
    r1 = Renderer(0)
    for si in check:
        arrays, amb = per_scene[si]
        j1, s1 = wl.scene_jobs(specs[si], arrays, amb, 0)
        r1.render(j1, [s1])
        assert torch.equal(s1.mix, scenes[si].mix), f"scene {si}: mix differs between batch and single render"
        for a, b in zip(j1, ev_of[si]):
            assert torch.equal(a.spatial, b.spatial)
    for si in (0, 77):  # against the float64 oracle
        arrays, amb = per_scene[si]
        arrays64 = [(x.cpu().numpy(), h.cpu().numpy().astype(np.float64)) for x, h in arrays]
        spatial, mix = _oracle_scene(specs[si], arrays64, amb.cpu().numpy())
        for j, s_ in zip(ev_of[si], spatial):
            assert np.abs(j.spatial.cpu().numpy() - s_).max() <= TOL
        assert np.abs(scenes[si].mix.cpu().numpy().astype(np.float64) - mix.scene).max() <= TOL
    r.close()
    r1.close()


def _mixed_batch(seed=11, n=40):
    rng = np.random.default_rng(seed)
    jobs = []
    for i in range(n):
        lx = int(rng.integers(500, 9000)); lh = int(rng.integers(50, 4000)); k = int(rng.choice([1, 1, 2, 5, 12]))
        x = cases.make_audio(rng, lx); h = cases.make_irs(rng, 4, k, lh).astype(np.float32)
        j = EventJob(audio=x, irs=h, n_channels=4, snr=10.0 + i % 7, ref_db=-65.0)
        if k > 1:
            j.ir_frames, j.n_frames = moving_frames(lx / 24000.0, 24000.0, k, lx)
        jobs.append(j)
    return jobs


def test_fifty_repeats_are_bit_identical(rnd):
    """Run-to-run determinism: the 40-event batch that exposed round 1's energy-reduction race, 50 times over; every
    run must reproduce the first bit for bit (no atomics on sample data, no scheduling-dependent summation order)."""
    jobs = _mixed_batch()
    rnd.render(jobs)
    first = [j.spatial.copy() for j in jobs]
    gains = [j.stats["gain"] for j in jobs]
    for rep in range(50):
        for j in jobs:
            j.spatial[...] = 0
        rnd.render(jobs)
        for j, f, g in zip(jobs, first, gains):
            assert np.array_equal(j.spatial, f), f"repeat {rep}: output changed"
            assert j.stats["gain"] == g


# ---- the persistent launches for moving events (alr_set_option "fused": 1 = k_mov_fused, 2 = k_mov_sweep) -------------------
@pytest.fixture(scope="module", params=[1, 2], ids=["ring", "sweep"])
def rnd_fused(request):
    r = Renderer(0, fused=request.param)
    yield r
    r.close()


def test_fused_launch_matches_oracle_on_a_moving_scene(rnd_fused, rnd):
    spec = wl.c3_scene_spec(5)
    arrays, amb = wl.host_scene_arrays(spec, dtype=np.float32)
    jobs, sj = wl.scene_jobs(spec, arrays, amb, 0)
    rnd_fused.render(jobs, [sj])
    assert rnd_fused.profile()["ms_fused"] >= 0.0
    spatial, mix = _oracle_scene(spec, [(x, h.astype(np.float64)) for x, h in arrays], amb)
    for j, s in zip(jobs, spatial):
        err = np.abs(j.spatial - s).max()
        assert err <= TOL and err <= 3e-5 * np.abs(s).max()
    assert np.abs(sj.mix.astype(np.float64) - mix.scene).max() <= TOL
    # and against the default kernels: same contraction, a_l applied at a different place -> rounding-level differences
    jobs2, sj2 = wl.scene_jobs(spec, arrays, amb, 0)
    rnd.render(jobs2, [sj2])
    assert np.abs(sj.mix - sj2.mix).max() <= 1e-6


@pytest.mark.parametrize("mode", [1, 2], ids=["ring", "sweep"])
@pytest.mark.parametrize("ring_mb,lookahead", [(2, 0), (3, 1), (8, 4), (64, 2)])
def test_fused_launch_tight_rings_and_lookaheads(ring_mb, lookahead, mode):
    """Small rings force the planner to shrink the lookahead on the spot, make producers wait for consumers and push
    events that do not fit back to the unfused kernels; results must not depend on any of it."""
    rng = np.random.default_rng(21)
    jobs, ref_jobs = [], []
    for i in range(12):
        lx = int(rng.integers(20000, 90000)); lh = int(rng.integers(3000, 9000)); n = int(rng.choice([2, 7, 19, 38]))
        x = cases.make_audio(rng, lx); h = cases.make_irs(rng, 4, n, lh).astype(np.float32)
        for lst in (jobs, ref_jobs):
            j = EventJob(audio=x, irs=h, n_channels=4, snr=8.0 + i, ref_db=-65.0)
            j.ir_frames, j.n_frames = moving_frames(lx / 24000.0, 24000.0, n, lx)
            lst.append(j)
    r = Renderer(0, fused=mode, ring_bytes=ring_mb << 20, lookahead=lookahead)
    r.render(jobs)
    first = [j.spatial.copy() for j in jobs]
    r.render(jobs)
    for j, f in zip(jobs, first):
        assert np.array_equal(j.spatial, f)  # deterministic whatever the task interleaving was
    r.close()
    plain = Renderer(0)
    plain.render(ref_jobs)
    plain.close()
    for a, b in zip(jobs, ref_jobs):
        assert np.abs(a.spatial - b.spatial).max() <= 2e-6 * max(np.abs(b.spatial).max(), 1e-30) + 1e-9


def test_fused_launch_with_dry_audio_and_odd_channel_counts(rnd_fused):
    """a_0 of a fused moving event feeds its dry / direct-path render; C = 3 and C = 6 exercise partial capsule groups."""
    rng = np.random.default_rng(33)
    for c in (3, 6):
        x = cases.make_audio(rng, 30000)
        h = cases.make_irs(rng, c, 9, 6000)
        job = EventJob(audio=x, irs=h.astype(np.float32), n_channels=c, snr=15.0, ref_db=-65.0, dry=(1, 144, 1560))
        job.ir_frames, job.n_frames = moving_frames(30000 / 24000.0, 24000.0, 9, 30000)
        rnd_fused.render([job])
        res = orc.render_event(x, h, 15.0, -65.0, is_moving=True, duration=30000 / 24000.0, sample_rate=24000.0,
                               literal=False, ref_ir_channel=1, direct_path_time_ms=(6, 65))
        assert np.abs(job.spatial - res.spatial).max() <= TOL
        assert np.abs(job.dry_out - res.dry).max() <= TOL


# ---- k_small_rir: RIRs of at most one partition ----------------------------------------------------------------------------
@pytest.mark.parametrize("lx,lh,c", [(30000, 700, 2), (5000, 2048, 1), (100, 7, 1), (70001, 1999, 2), (2049, 2047, 2), (3000, 100, 4)])
def test_small_rir_kernel_vs_oracle_and_general_pipeline(rnd, lx, lh, c):
    """Short static RIRs take k_small_rir (spectra in registers); same result as the oracle and, to rounding, as the
    general partitioned pipeline (small_rir = 0)."""
    rng = np.random.default_rng(lx + lh + c)
    x = cases.make_audio(rng, lx)
    h = cases.make_irs(rng, c, 1, lh)
    job = EventJob(audio=x, irs=h.astype(np.float32), n_channels=c, snr=7.0, ref_db=-60.0)
    rnd.render([job])
    res = orc.render_event(x, h, 7.0, -60.0, is_moving=False)
    assert np.abs(job.spatial - res.spatial).max() <= TOL
    general = Renderer(0, small_rir=0)
    job2 = EventJob(audio=x, irs=h.astype(np.float32), n_channels=c, snr=7.0, ref_db=-60.0)
    general.render([job2])
    general.close()
    assert np.abs(job.spatial - job2.spatial).max() <= 2e-6 * np.abs(job2.spatial).max() + 1e-9
    assert abs(job.stats["gain"] - job2.stats["gain"]) <= 1e-5 * abs(job2.stats["gain"])


@pytest.mark.parametrize("n_irs", [1, 9])
def test_dry_window_through_small_rir_kernel(rnd, n_irs):
    """compute_dry_audio (synthesize.py:432-504): the 6 ms / 65 ms window around the direct-path peak is 1 704 taps at
    24 kHz — one partition — whatever the RIR length; static and moving parents, raw length Lx + Lh - 1."""
    rng = np.random.default_rng(90 + n_irs)
    x = cases.make_audio(rng, 40000)
    h = cases.make_irs(rng, 4, n_irs, 24000)
    job = EventJob(audio=x, irs=h.astype(np.float32), n_channels=4, snr=12.0, ref_db=-65.0, dry=(2, 144, 1560))
    if n_irs > 1:
        job.ir_frames, job.n_frames = moving_frames(40000 / 24000.0, 24000.0, n_irs, 40000)
    rnd.render([job])
    res = orc.render_event(x, h, 12.0, -65.0, is_moving=n_irs > 1, duration=40000 / 24000.0, sample_rate=24000.0,
                           literal=False, ref_ir_channel=2, direct_path_time_ms=(6, 65))
    assert job.dry_out.shape == (40000 + 24000 - 1,)
    assert np.abs(job.dry_out - res.dry).max() <= TOL
    assert np.abs(job.spatial - res.spatial).max() <= TOL
    general = Renderer(0, small_rir=0)
    job2 = EventJob(audio=x, irs=h.astype(np.float32), n_channels=4, snr=12.0, ref_db=-65.0, dry=(2, 144, 1560))
    if n_irs > 1:
        job2.ir_frames, job2.n_frames = moving_frames(40000 / 24000.0, 24000.0, n_irs, 40000)
    general.render([job2])
    general.close()
    assert job.stats["dry_peak"] == job2.stats["dry_peak"]
    assert np.abs(job.dry_out - job2.dry_out).max() <= 2e-6 * np.abs(job2.dry_out).max() + 1e-9
    outside = np.ones(job.dry_out.shape[0], bool)
    pk = job.stats["dry_peak"]
    outside[max(pk - 144, 0):pk + 1560 + 40000] = False
    assert np.all(job.dry_out[outside] == 0.0)  # exact zeros where no windowed tap can reach
