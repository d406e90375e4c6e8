"""Golden vectors for the visibility front-end (f4) from the UNMODIFIED reference functions
`audiblelight.imaging.extract_visibilities` / `form_visibility` (imaging.py:455-492, 697-719).

    python tests/golden/make_golden_imaging.py      # needs /root/reference; writes tests/golden/imaging.npz

imaging.py imports packages that are not installed offline. None of them takes part in the arithmetic of these two
functions except scikit-image's `view_as_blocks` / `view_as_windows`, which are pure re-indexing: they are provided
here through numpy's `sliding_window_view` with the same semantics. The pyunlocbox / astropy stubs only have to let the
module import (its solver classes subclass pyunlocbox types).
"""
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_loader  # noqa: E402


def _view_as_windows(arr_in, window_shape, step=1):
    from numpy.lib.stride_tricks import sliding_window_view
    if isinstance(step, int):
        step = (step,) * arr_in.ndim
    v = sliding_window_view(arr_in, window_shape)
    return v[tuple(slice(None, None, s) for s in step)]


def _view_as_blocks(arr_in, block_shape):
    return _view_as_windows(arr_in, block_shape, block_shape)


def load_reference_imaging():
    ref_loader.load_reference_synthesize()  # installs the common stubs and the reference on sys.path
    sk = types.ModuleType("skimage")
    sku = types.ModuleType("skimage.util")
    sku.view_as_blocks, sku.view_as_windows = _view_as_blocks, _view_as_windows
    sk.util = sku
    sys.modules["skimage"], sys.modules["skimage.util"] = sk, sku
    opt = types.ModuleType("pyunlocbox")
    opt.functions = types.ModuleType("pyunlocbox.functions")
    opt.acceleration = types.ModuleType("pyunlocbox.acceleration")
    opt.functions.func = type("func", (), {})
    opt.functions.dummy = type("dummy", (), {})
    opt.acceleration.accel = type("accel", (), {})
    sys.modules["pyunlocbox"], sys.modules["pyunlocbox.functions"], sys.modules["pyunlocbox.acceleration"] = \
        opt, opt.functions, opt.acceleration
    for name in ("astropy", "astropy.coordinates", "astropy.units"):
        m = MagicMock(name=name)
        m.__path__ = []
        sys.modules[name] = m
    import audiblelight.imaging as img  # noqa
    return img


CASES = [  # (name, sr, channels, seconds, fc list, bw, t_sti, per_block)
    ("tetra_24k_default_bands", 24000, 4, 1.0, list(np.linspace(1500, 4500, 9)), 50.0, 10e-3, 10),
    ("em32_48k_100ms", 48000, 32, 0.45, [1500.0, 3000.0], 120.0, 100e-3, 2),
    ("ragged_16k_low_band", 16000, 3, 0.777, [10.0, 400.0, 7990.0], 50.0, 12.5e-3, 3),
]


def main():
    img = load_reference_imaging()
    out = {}
    for ci, (name, sr, c, secs, fcs, bw, t_sti, per) in enumerate(CASES):
        rng = np.random.default_rng(7100 + ci)
        n = int(secs * sr)
        data = (rng.standard_normal((n, c)) * np.exp(-np.arange(n)[:, None] / (0.6 * n))).astype(np.float32)
        out[f"{name}__data"] = data
        for bi, fc in enumerate(fcs):
            out[f"{name}__vis{bi}"] = img.form_visibility(data, sr, fc, bw, t_sti, per * t_sti)
        out[f"{name}__ext0"] = img.extract_visibilities(data, sr, t_sti, fcs[0], bw, alpha=0.3)
        out[f"{name}__params"] = np.array([sr, c, bw, t_sti, per] + list(fcs), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "imaging.npz"), **out)
    print("wrote imaging.npz:", {k: v.shape for k, v in out.items() if "__vis0" in k})


if __name__ == "__main__":
    main()
