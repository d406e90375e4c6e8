"""f4: the STFT / visibility front-end of acoustic imaging (imaging.py:455-719) — oracle against goldens produced by the
unmodified reference functions (CPU), and the GPU path through the C-ABI against both."""
import os

import numpy as np
import pytest

from oracle import imaging_oracle as io

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "imaging.npz")
NAMES = ["tetra_24k_default_bands", "em32_48k_100ms", "ragged_16k_low_band"]


def _case(g, name):
    p = g[f"{name}__params"]
    sr, c, bw, t_sti, per = float(p[0]), int(p[1]), float(p[2]), float(p[3]), int(p[4])
    return g[f"{name}__data"], sr, c, bw, t_sti, per, list(p[5:])


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_goldens(name):
    g = np.load(GOLD)
    data, sr, c, bw, t_sti, per, fcs = _case(g, name)
    for bi, fc in enumerate(fcs):
        want = g[f"{name}__vis{bi}"]
        got = io.form_visibility(data, sr, fc, bw, t_sti, per * t_sti)
        assert got.shape == want.shape and got.dtype == want.dtype
        assert np.abs(got - want).max() <= 1e-12 * max(np.abs(want).max(), 1e-300)
    want = g[f"{name}__ext0"]
    assert np.abs(io.extract_visibilities(data, sr, t_sti, fcs[0], bw, 0.3) - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_visibilities_match_reference_goldens(name):
    import torch
    from audiblelight_b200 import imaging
    from audiblelight_b200.renderer import Renderer
    g = np.load(GOLD)
    data, sr, c, bw, t_sti, per, fcs = _case(g, name)
    rnd = Renderer(0)
    mix = np.ascontiguousarray(data.T)
    got = imaging.visibility_bands(mix, sr, fcs, bw, t_sti, per, renderer=rnd)
    assert got.shape[0] == len(fcs) and got.dtype == np.complex128
    for bi in range(len(fcs)):
        want = g[f"{name}__vis{bi}"]
        assert got[bi].shape == want.shape
        # float32 samples, float64 accumulation: rounding-level agreement with the reference's complex128 FFT
        assert np.abs(got[bi] - want).max() <= 1e-9 * max(np.abs(want).max(), 1e-30)
    # drop-in signature of form_visibility
    v = imaging.form_visibility(data, sr, fcs[0], bw, t_sti, per * t_sti, renderer=rnd)
    assert np.abs(v - g[f"{name}__vis0"]).max() <= 1e-9 * max(np.abs(g[f"{name}__vis0"]).max(), 1e-30)
    # extract_visibilities (one frame per block) with a Tukey(0.3) window
    e = imaging.visibility_bands(mix, sr, [fcs[0]], bw, t_sti, 1, alpha=0.3, renderer=rnd)[0]
    want = g[f"{name}__ext0"]
    assert e.shape == want.shape and np.abs(e - want).max() <= 1e-9 * np.abs(want).max()
    # device-resident input gives the same numbers
    d = imaging.visibility_bands(torch.from_numpy(mix).cuda(), sr, fcs, bw, t_sti, per, renderer=rnd)
    assert np.array_equal(d.cpu().numpy(), got)
    rnd.close()


@pytest.mark.gpu
def test_gpu_visibilities_of_a_rendered_mix_and_errors():
    """The consumer side: visibilities of a mix that was rendered on the device, without a host round trip."""
    import torch
    from audiblelight_b200 import imaging, workload as wl
    from audiblelight_b200.renderer import Renderer
    rnd = Renderer(0)
    spec = wl.c3_scene_spec(3, duration=10.0, n_static=2, n_moving=1)
    arrays, amb = wl.device_scene_arrays(spec, torch.device("cuda", 0))
    jobs, sj = wl.scene_jobs(spec, arrays, amb, 0)
    rnd.render(jobs, [sj])
    freqs = imaging.band_frequencies()
    v = imaging.visibility_bands(sj.mix, spec.sr, freqs, renderer=rnd)
    assert tuple(v.shape) == (9, 100, 4, 4)
    want = np.stack([io.form_visibility(sj.mix.cpu().numpy().T, spec.sr, fc, 50.0, 10e-3, 100e-3) for fc in freqs])
    assert np.abs(v.cpu().numpy() - want).max() <= 1e-9 * np.abs(want).max()
    h = v.cpu().numpy()
    assert np.allclose(h, np.conj(np.swapaxes(h, -1, -2)))  # Hermitian by construction
    with pytest.raises(ValueError, match="Not enough samples per time frame"):
        imaging.visibility_bands(sj.mix, spec.sr, freqs, t_sti=1e-6, renderer=rnd)
    with pytest.raises(TypeError):
        imaging.visibility_bands(sj.mix.double(), spec.sr, freqs, renderer=rnd)
    rnd.close()
