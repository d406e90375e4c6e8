"""CPU tests of the C++ host planner (alr_debug_plan): a float64 numpy model of the device algorithm driven by the
planner's output must reproduce the oracle (closed form == the reference's STFT-domain result)."""
import os

import numpy as np
import pytest

import cases
import upols_model
from oracle import synth_oracle as orc

pytestmark = pytest.mark.skipif(
    not os.path.exists(os.path.join(os.path.dirname(os.path.dirname(__file__)), "audiblelight_b200", "libalrender.so")),
    reason="libalrender.so not built (run __graft_entry__.build())")


def _job(audio, irs, sr, n_out=None):
    from audiblelight_b200.renderer import EventJob, moving_frames
    n = irs.shape[1]
    job = EventJob(audio=audio, irs=np.ascontiguousarray(irs, dtype=np.float32), n_channels=irs.shape[0], n_out=n_out)
    if n > 1:
        job.ir_frames, job.n_frames = moving_frames(len(audio) / float(sr), float(sr), n, len(audio))
    return job


@pytest.mark.parametrize("name", [n for n, s in cases.EVENT_CASES.items() if s["n"] >= 1])
def test_plan_model_matches_oracle(name):
    from audiblelight_b200.renderer import debug_plan
    spec = cases.EVENT_CASES[name]
    audio, irs = cases.event_inputs(spec)
    moving = spec["n"] > 1
    plan = debug_plan(_job(audio, irs, spec["sr"]))
    a = orc.ir_scales(irs)
    scales = a * (512.0 if moving else 1.0)
    got = upols_model.model_convolve(audio.astype(np.float64), irs, plan, scales, moving, len(audio))
    irs_n = orc.normalize_irs(irs.transpose(1, 0, 2)).transpose(1, 0, 2)
    if moving:
        want = orc.time_variant_convolution_closed(irs_n, audio, len(audio) / float(spec["sr"]), float(spec["sr"]))
    else:
        want = orc.linear_convolve(audio.astype(np.float64)[None, :], irs_n[:, 0, :])
    want = orc.pad_or_truncate(want, len(audio))
    scale = max(1e-30, np.abs(want).max())
    # moving: the plan stores the cross-fade weights as float32 (6e-8 relative)
    assert np.abs(got - want).max() < (5e-7 if moving else 1e-9) * scale


@pytest.mark.parametrize("lx,lh,n,sr", [(30000, 5000, 13, 24000), (5000, 24000, 3, 24000), (20480, 1024, 9, 48000),
                                         (1023, 1025, 2, 16000), (48000, 3000, 21, 24000), (2049, 100, 40, 24000)])
def test_plan_model_random_shapes(lx, lh, n, sr):
    from audiblelight_b200.renderer import debug_plan
    rng = np.random.default_rng(lx + lh + n)
    audio = cases.make_audio(rng, lx)
    irs = cases.make_irs(rng, 2, n, lh)
    plan = debug_plan(_job(audio, irs, sr))
    scales = np.full(n, 512.0)
    got = upols_model.model_convolve(audio.astype(np.float64), irs, plan, scales, True, lx)
    want = orc.pad_or_truncate(orc.time_variant_convolution_closed(irs, audio, lx / float(sr), float(sr)), lx)
    assert np.abs(got - want).max() < 5e-7 * np.abs(want).max()
    # structural invariants of the plan
    ir = plan["irs"]
    act = ir[ir[:, 1] > 0]
    assert np.all(np.diff(act[:, 0]) >= 0) and np.all(np.diff(act[:, 0] + act[:, 1]) >= 0) and np.all(ir[:, 1] >= 0)
    assert np.array_equal(ir[:, 2], np.concatenate([[0], np.cumsum(ir[:, 1])[:-1]]))


@pytest.mark.parametrize("lx,lh,n,sr,c", [(30000, 5000, 13, 24000, 2), (60000, 30000, 300, 24000, 1),
                                           (40000, 40000, 150, 24000, 1), (9000, 2000, 1, 24000, 2)])
def test_cmac_item_lists_reproduce_the_block_sums(lx, lh, n, sr, c):
    """k_cmac's item lists (one record per (RIR, partition) with a validity mask over the run's 8 outputs, built per window
    of 64 RIRs and consumed in passes of 448) must cover exactly the (source block, partition) pairs of every output
    block: the item-wise accumulation equals the block-wise one. The dense cases need several windows and passes."""
    from audiblelight_b200.renderer import debug_plan
    rng = np.random.default_rng(lx + lh + n)
    audio = cases.make_audio(rng, lx)
    irs = cases.make_irs(rng, c, n, lh)
    plan = debug_plan(_job(audio, irs, sr))
    scales = np.full(n, 512.0 if n > 1 else 1.0)
    a = upols_model.model_convolve(audio.astype(np.float64), irs, plan, scales, n > 1, lx, cmac="blocks")
    b = upols_model.model_convolve(audio.astype(np.float64), irs, plan, scales, n > 1, lx, cmac="items")
    assert np.abs(a - b).max() <= 1e-12 * max(1e-30, np.abs(a).max())
    kc = upols_model.kernel_constants()
    runs = -(-plan["B_valid"] // kc["G"])
    if n >= 150:  # dense trajectory: more than one window / pass per run
        assert upols_model.model_convolve.last_lists > runs
    else:
        assert upols_model.model_convolve.last_lists <= runs


def test_plan_static_full_convolution():
    from audiblelight_b200.renderer import debug_plan
    rng = np.random.default_rng(5)
    audio = cases.make_audio(rng, 2500)
    irs = cases.make_irs(rng, 3, 1, 1300)
    n_out = 2500 + 1300 - 1
    plan = debug_plan(_job(audio, irs, 24000, n_out=n_out))
    assert plan["n_valid"] == n_out and plan["xlimit"] == 2500
    got = upols_model.model_convolve(audio.astype(np.float64), irs, plan, np.ones(1), False, n_out)
    want = orc.linear_convolve(audio.astype(np.float64)[None, :], irs[:, 0, :])
    assert np.abs(got - want).max() < 1e-10 * np.abs(want).max()


def test_moving_frames_half_to_even():
    from audiblelight_b200.renderer import moving_frames
    fr, n_frames = moving_frames(1.0, 24000.0, 11, 24000)
    assert fr[-1] == 188 and n_frames == 188  # 188.5 rounds to even, not 189
    w = orc.interpolation_matrix(np.linspace(0, 1.0, 11), 24000.0)
    assert w.shape[0] == 188
