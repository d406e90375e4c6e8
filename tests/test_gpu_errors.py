"""Error behaviour of the C-ABI on a real device: invalid descriptors are rejected with ALR_ERR_INVALID and a message,
before any work is queued, and the context stays usable (the reference raises Python exceptions at the same points;
its Python-level messages are reproduced by audiblelight_b200/synthesize.py and tested in test_dropin.py)."""
import numpy as np
import pytest

import cases
import gpu_util
from audiblelight_b200 import _lib
from audiblelight_b200 import augment as A
from audiblelight_b200.renderer import EventJob, Renderer, SceneJob

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rnd():
    r = Renderer(0)
    yield r
    r.close()


def _good(seed=0, c=2, n=1, lx=3000, lh=500):
    rng = np.random.default_rng(seed)
    return gpu_util.event_job(dict(sr=8000.0, snr=6.0, ref_db=-65), cases.make_audio(rng, lx), cases.make_irs(rng, c, n, lh))


def test_invalid_descriptors_are_rejected_and_context_survives(rnd):
    ref = _good()
    rnd.render([ref])
    want = ref.spatial.copy()

    bad = _good(n=3)
    bad.ir_frames = np.array([5, 3, 9], dtype=np.int32)  # decreasing
    with pytest.raises(_lib.AlrenderError, match="ir_frames must be non-decreasing"):
        rnd.render([bad])
    bad = _good(n=3)
    bad.ir_frames = np.array([0, 3, 9], dtype=np.int32)
    with pytest.raises(_lib.AlrenderError, match=r"ir_frames\[0\] must be >= 1"):
        rnd.render([bad])

    j = _good()
    j.scene = 3
    with pytest.raises(_lib.AlrenderError, match="scene index 3 out of range"):
        rnd.render([j], [SceneJob(n_channels=2, n_samples=4000)])
    j = _good(c=2)
    j.scene, j.scene_start, j.scene_end = 0, 0, 3000
    with pytest.raises(_lib.AlrenderError, match="2 channels but scene 0 has 3"):
        rnd.render([j], [SceneJob(n_channels=3, n_samples=4000)])
    j = _good()
    j.scene, j.scene_start, j.scene_end = 0, 100, 5000
    with pytest.raises(_lib.AlrenderError, match="outside the scene"):
        rnd.render([j], [SceneJob(n_channels=2, n_samples=4000)])

    j = _good()
    j.dry = (5, 10, 10)  # reference channel beyond the capsules (synthesize.py:470)
    with pytest.raises(_lib.AlrenderError, match="Reference channel index out of range"):
        rnd.render([j])
    j = _good(n=0)
    j.dry = (0, 10, 10)
    j.dry_out = np.zeros(10, np.float32)
    with pytest.raises((_lib.AlrenderError, TypeError, AttributeError)):
        rnd.render([j])

    j = _good()
    j.aug_ops = [A.gain_db(1.0)] * 9
    with pytest.raises(_lib.AlrenderError, match="more than 8 augmentations"):
        rnd.render([j])
    j = _good()
    j.aug_ops = [A.AugOp(99)]
    with pytest.raises(_lib.AlrenderError, match="bad augmentation type 99"):
        rnd.render([j])

    # a failing event anywhere in a batch fails the whole call before anything is written
    ok = _good(seed=1)
    ok.spatial = np.full((2, 3000), 7.0, np.float32)
    bad = _good(n=3)
    bad.ir_frames = np.array([5, 3, 9], dtype=np.int32)
    with pytest.raises(_lib.AlrenderError):
        rnd.render([ok, bad])
    assert (ok.spatial == 7.0).all()

    # the context is still good
    again = _good()
    rnd.render([again])
    assert np.array_equal(again.spatial, want)


def test_python_layer_shape_errors(rnd):
    j = _good()
    j.spatial = np.zeros((2, 10), np.float32)
    with pytest.raises(ValueError, match="spatial buffer has shape"):
        rnd.render([j])
    j = _good()
    j.n_channels = 3
    with pytest.raises(ValueError, match="irs.shape\\[0\\] != n_channels"):
        rnd.render([j])
    with pytest.raises(ValueError, match="Scene ambient noise does not match expected shape"):
        rnd.render([_good()], [SceneJob(n_channels=2, n_samples=100, ambience=[np.zeros((2, 99), np.float32)],
                                        ambience_ref_db=[-65.0])])
    rnd.render([], [])  # an empty call is a no-op


def test_ambience_only_scene_and_event_outside_any_scene(rnd):
    """C-ABI corner cases the reference never reaches (it refuses scenes without events): a scene that holds only
    ambience layers, next to an event that is rendered but not mixed anywhere."""
    from oracle import synth_oracle as orc
    rng = np.random.default_rng(5)
    amb = [np.ascontiguousarray(cases.make_ambience(rng, 3, 5000), dtype=np.float32) for _ in range(2)]
    sc = SceneJob(n_channels=3, n_samples=5000, ambience=amb, ambience_ref_db=[-40.0, -55.0])
    ev = _good(seed=3)
    rnd.render([ev], [sc])
    want = np.zeros((3, 5000), np.float32)
    for a, db in zip(amb, (-40.0, -55.0)):
        want += (orc.db_to_multiplier(db, np.mean(np.abs(a))) * a).astype(np.float32)
    assert np.abs(sc.mix - want).max() < 1e-6 * np.abs(want).max()
    alone = _good(seed=3)
    rnd.render([alone])
    assert np.array_equal(ev.spatial, alone.spatial)
    # no ambience, no events: the mix is silence
    empty = SceneJob(n_channels=2, n_samples=1000, mix=np.ones((2, 1000), np.float32))
    rnd.render([], [empty])
    assert not empty.mix.any()
