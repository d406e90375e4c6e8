"""Import the UNMODIFIED reference `audiblelight.synthesize`.

Source tree, first that exists: $ALR_REFERENCE_ROOT, /root/reference (the build container), baseline/_ref (the copy
`baseline/install_reference.py` makes; git-ignored, shipped to the GPU box by gpurun).

Used by `tests/golden/make_golden.py` (golden-vector generation), the optional `tests/test_oracle_vs_reference.py`
and `bench.py --impl reference` / its `cpu_baseline` leg (baseline/reference_arm.py). Nothing in the product
(audiblelight_b200/), the `-m gpu` tests or `smoke()` imports this.

The reference needs packages that are not installed here (librosa, soundfile, trimesh, pedalboard, ...).
None of them is used by the arithmetic of the synthesis hot path (synthesize.py:40-677), so they are
replaced by MagicMock modules; `librosa.util.valid_audio` is re-implemented as the finite/float check it is.
"""
import importlib.metadata
import os
import sys
import types
from collections import OrderedDict
from unittest.mock import MagicMock

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    env = os.environ.get("ALR_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isdir(os.path.join(cand, "audiblelight")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()

_STUBS = [
    "librosa", "librosa.util", "librosa.effects", "deepdiff", "soundfile", "pedalboard", "matplotlib",
    "matplotlib.pyplot", "trimesh", "trimesh.visual", "pysofaconventions", "rlr_audio_propagation",
    "pedalboard.io", "pyvista", "vtk", "cv2", "pyroomacoustics", "h5py", "netCDF4", "gdown",
]


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "audiblelight"))


def _valid_audio(y, **_):
    y = np.asarray(y)
    if not np.issubdtype(y.dtype, np.floating):
        raise ValueError("Audio data must be floating-point")
    if not np.isfinite(y).all():
        raise ValueError("Audio buffer is not finite everywhere")
    return True


def load_reference_synthesize():
    """Returns the reference's `audiblelight.synthesize` module (cached in sys.modules)."""
    if "audiblelight.synthesize" in sys.modules and getattr(
        sys.modules["audiblelight.synthesize"], "_alr_is_reference", False
    ):
        return sys.modules["audiblelight.synthesize"]
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for name in _STUBS:
        try:
            __import__(name)
        except Exception:
            m = MagicMock(name=name)
            m.__path__ = []
            m.__name__ = name
            m.__spec__ = None
            sys.modules[name] = m
    real_version = importlib.metadata.version

    def _version(name):
        try:
            return real_version(name)
        except importlib.metadata.PackageNotFoundError:
            return "0.0.0"

    importlib.metadata.version = _version
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import audiblelight.synthesize as syn  # noqa

    syn.librosa.util.valid_audio = _valid_audio
    syn.Ambience = RefAmbience
    syn._alr_is_reference = True
    return syn


# ---- duck-typed stand-ins for the reference's Event / Ambience / Scene (event.py, ambience.py, core.py) ----
class RefEvent:
    def __init__(self, audio, sample_rate, n_emitters, snr, scene_start=0.0, alias="ev",
                 ref_ir_channel=None, direct_path_time_ms=None):
        self.audio = np.asarray(audio, dtype=np.float32)
        self.sample_rate = float(sample_rate)
        self.n_emitters = int(n_emitters)
        self.snr = snr
        self.alias = alias
        self.duration = len(self.audio) / self.sample_rate
        self.scene_start = float(scene_start)
        self.scene_end = self.scene_start + self.duration
        self.is_moving = self.n_emitters > 1
        self.ref_ir_channel = ref_ir_channel
        self.direct_path_time_ms = direct_path_time_ms
        self.spatial_audio = OrderedDict()
        self._spatial_audio_padded = OrderedDict()
        self._spatial_audio_dry = OrderedDict()
        self._spatial_audio_dry_padded = OrderedDict()

    def load_audio(self, ignore_cache=False, normalize=True):
        return self.audio

    def __len__(self):
        return self.n_emitters


class RefAmbience:
    def __init__(self, noise, ref_db):
        self.noise = noise
        self.ref_db = ref_db

    def load_ambience(self, normalize=True):
        return self.noise


class RefScene:
    def __init__(self, duration, sample_rate, ref_db, events, ambience=None, mics=("mic000",)):
        self.duration = duration
        self.sample_rate = sample_rate
        self.ref_db = ref_db
        self.events = OrderedDict((e.alias, e) for e in events)
        self.ambience = OrderedDict(ambience or {})
        self.audio = OrderedDict()
        self.state = types.SimpleNamespace(microphones=OrderedDict((m, None) for m in mics))
