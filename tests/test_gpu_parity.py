"""GPU parity tests (run on the B200 box: pytest -m gpu). Everything goes through the C-ABI library
(audiblelight_b200/libalrender.so); results are compared with the reference's golden outputs
(tests/golden/*.npz, produced by the unmodified reference) and with the CPU oracle on seeded inputs."""
import os

import numpy as np
import pytest

import cases
import gpu_util
from gpu_util import TOL
from oracle import synth_oracle as orc

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def rnd():
    import torch
    assert torch.cuda.is_available(), "the gpu tests need a CUDA device"
    from audiblelight_b200.renderer import Renderer
    r = Renderer(0)
    yield r
    r.close()


@pytest.fixture(scope="module")
def gev():
    return np.load(os.path.join(G, "events.npz"))


@pytest.fixture(scope="module")
def gsc():
    return np.load(os.path.join(G, "scenes.npz"))


# ---- FFT core (negacyclic fold+twist transform, alr_fft.cuh) ----------------------------------------------------------
@pytest.mark.parametrize("n_valid", [4096, 1000, 1, 513])
def test_fft_core_forward_inverse(rnd, n_valid):
    import torch
    P = rnd._lib.alr_partition_size()
    n_valid = min(n_valid, P)
    rng = np.random.default_rng(n_valid)
    x = rng.standard_normal((37, n_valid)).astype(np.float32)
    spec = rnd.debug_rfft(torch.from_numpy(x).cuda())
    zeta = np.exp(1j * np.pi * np.arange(P) / (2 * P))
    ref = np.fft.fft(np.pad(x.astype(np.float64), ((0, 0), (0, P - n_valid))) * zeta, axis=-1)
    got = spec.cpu().numpy()
    got = got[..., 0].astype(np.float64) + 1j * got[..., 1]
    assert np.abs(got - ref).max() < 2e-6 * np.abs(ref).max()
    back = rnd.debug_irfft(spec).cpu().numpy()  # (n_blocks, 2P): first half = real parts, second half = overlap tail
    assert np.abs(back[:, :n_valid] - x).max() < 2e-6 * np.abs(x).max()
    assert np.abs(back[:, n_valid:]).max() < 2e-6 * np.abs(x).max()


def test_fft_core_block_convolution(rnd):
    """Pointwise product of two block spectra == linear convolution of the two P-sample blocks (2P-1 samples)."""
    import torch
    rng = np.random.default_rng(3)
    P = rnd._lib.alr_partition_size()
    a = rng.standard_normal((5, P)).astype(np.float32)
    b = rng.standard_normal((5, P)).astype(np.float32)
    A = rnd.debug_rfft(torch.from_numpy(a).cuda())
    B = rnd.debug_rfft(torch.from_numpy(b).cuda())
    Ac, Bc = torch.view_as_complex(A.contiguous()), torch.view_as_complex(B.contiguous())
    Y = torch.view_as_real(Ac * Bc).contiguous()
    y = rnd.debug_irfft(Y).cpu().numpy()
    for i in range(5):
        ref = np.convolve(a[i].astype(np.float64), b[i].astype(np.float64))
        assert np.abs(y[i, :2 * P - 1] - ref).max() < 3e-6 * np.abs(ref).max()


# ---- render_event_audio vs the reference's golden output --------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.EVENT_CASES))
def test_render_event_golden(rnd, gev, name):
    spec = cases.EVENT_CASES[name]
    audio, irs = cases.event_inputs(spec)
    job = gpu_util.event_job(spec, audio, irs)
    rnd.render([job])
    ref = gev[f"{name}__spatial"].astype(np.float64)
    got = job.spatial.astype(np.float64)
    assert got.shape == ref.shape
    err = np.abs(got - ref).max()
    assert err <= TOL, f"max-abs error {err:.3e} > {TOL}"
    # and much tighter relative to the signal itself (fp32 FFT noise only)
    assert err <= 2e-5 * max(np.abs(ref).max(), 1e-30) + 1e-12
    assert not job.stats["nonfinite"]
    if f"{name}__dry" in gev.files:
        dref = gev[f"{name}__dry"]
        assert job.dry_out.shape == dref.shape
        derr = np.abs(job.dry_out.astype(np.float64) - dref).max()
        assert derr <= TOL and derr <= 2e-5 * np.abs(dref).max()
        irs_n = orc.normalize_irs(irs.transpose(1, 0, 2)).transpose(1, 0, 2)
        assert job.stats["dry_peak"] == int(np.argmax(irs_n[spec["ref_ir_channel"], 0]))


def test_event_stats_match_oracle(rnd):
    spec = cases.EVENT_CASES["static_4ch"]
    audio, irs = cases.event_inputs(spec)
    job = gpu_util.event_job(spec, audio, irs)
    rnd.render([job])
    res = orc.render_event(audio, irs, spec["snr"], spec["ref_db"], is_moving=False)
    assert np.isclose(job.stats["event_scale"], res.event_scale, rtol=1e-5)


# ---- raw convolutions (time_invariant_convolution / time_variant_convolution) ---------------------------------------
def test_raw_static_convolution_full_length(rnd):
    from audiblelight_b200.renderer import ALR_GAIN_NONE
    rng = np.random.default_rng(7)
    audio = cases.make_audio(rng, 5000)
    irs = cases.make_irs(rng, 4, 1, 2100)
    job = gpu_util.event_job(dict(snr=1.0, ref_db=0.0, sr=24000), audio, irs, normalize_irs=False,
                             gain_mode=ALR_GAIN_NONE, n_out=5000 + 2100 - 1)
    rnd.render([job])
    want = orc.time_invariant_convolution(audio, irs[:, 0].T)
    assert job.spatial.shape == want.shape
    assert np.abs(job.spatial - want).max() < 2e-6 * np.abs(want).max()


def test_raw_moving_convolution(rnd):
    from audiblelight_b200.renderer import ALR_GAIN_NONE
    spec = cases.EVENT_CASES["moving_5ir"]
    audio, irs = cases.event_inputs(spec)
    g = np.load(os.path.join(G, "conv_primitives.npz"))["tvc_out"]
    job = gpu_util.event_job(spec, audio, irs, normalize_irs=False, gain_mode=ALR_GAIN_NONE, n_out=g.shape[1])
    rnd.render([job])
    assert np.abs(job.spatial - g).max() < 3e-6 * np.abs(g).max()


# ---- larger seeded cases vs the oracle's closed form --------------------------------------------------------------
@pytest.mark.parametrize("lx,lh,c,n,sr", [(48000, 24000, 4, 21, 24000), (120000, 24000, 4, 1, 24000),
                                          (30000, 9000, 2, 64, 24000), (96000, 48000, 8, 1, 48000),
                                          # dense trajectories with long RIRs: more than 64 RIRs per run of output blocks
                                          # (several list-building windows in k_cmac) and more than 448 (RIR, partition)
                                          # items per window (several passes)
                                          (40000, 40000, 2, 150, 24000), (60000, 30000, 1, 300, 24000)])
def test_render_event_vs_oracle_large(rnd, lx, lh, c, n, sr):
    rng = np.random.default_rng(lx + n)
    audio = cases.make_audio(rng, lx)
    irs = cases.make_irs(rng, c, n, lh)
    spec = dict(sr=sr, snr=17.0, ref_db=-65)
    job = gpu_util.event_job(spec, audio, irs)
    rnd.render([job])
    res = orc.render_event(audio, irs, 17.0, -65, is_moving=n > 1, duration=lx / float(sr), sample_rate=float(sr),
                           literal=False)
    err = np.abs(job.spatial - res.spatial).max()
    assert err <= TOL and err <= 2e-5 * np.abs(res.spatial).max()


# ---- scenes --------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.SCENE_CASES))
def test_scene_mix_golden(rnd, gsc, name):
    spec = cases.SCENE_CASES[name]
    jobs, scene = gpu_util.scene_jobs(spec)
    rnd.render(jobs, [scene])
    ref = gsc[f"{name}__scene"]
    assert scene.mix.shape == ref.shape and scene.mix.dtype == np.float32
    err = np.abs(scene.mix.astype(np.float64) - ref).max()
    assert err <= TOL and err <= 2e-5 * np.abs(ref).max()
    got_slices = np.array([(j.scene_start, j.scene_end) for j in jobs], dtype=np.int64)
    assert np.array_equal(got_slices, gsc[f"{name}__slices"])  # event timings bit-exact


def test_device_resident_matches_host(rnd):
    import torch
    spec = cases.SCENE_CASES["scene_moving_two_ambiences"]
    jobs, scene = gpu_util.scene_jobs(spec)
    rnd.render(jobs, [scene])
    jobs_d, scene_d = gpu_util.scene_jobs(spec)
    for j in jobs_d:
        j.audio = torch.from_numpy(j.audio).cuda()
        j.irs = torch.from_numpy(j.irs).cuda() if j.irs is not None else None
    scene_d.ambience = [torch.from_numpy(a).cuda() for a in scene_d.ambience]
    rnd.render(jobs_d, [scene_d])
    assert torch.equal(scene_d.mix.cpu(), torch.from_numpy(scene.mix))
    for a, b in zip(jobs, jobs_d):
        assert torch.equal(b.spatial.cpu(), torch.from_numpy(a.spatial))


def test_small_workspace_chunks_give_same_result(gsc):
    from audiblelight_b200.renderer import Renderer
    spec = cases.SCENE_CASES["scene_moving_two_ambiences"]
    r = Renderer(0, workspace_limit=1 << 17)
    jobs, scene = gpu_util.scene_jobs(spec)
    r.render(jobs, [scene])
    assert r.profile()["n_chunks"] >= 2
    ref = gsc["scene_moving_two_ambiences__scene"]
    assert np.abs(scene.mix.astype(np.float64) - ref).max() <= TOL
    r.close()
