"""CPU: the C-ABI library loads and exports every symbol include/alrender.h declares; no compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audiblelight_b200", "libalrender.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="libalrender.so not built")


def _declared():
    text = open(os.path.join(ROOT, "include", "alrender.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(alr_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from audiblelight_b200 import _lib
    names = _declared()
    assert "alr_render" in names and "alr_create" in names
    lib = C.CDLL(LIB)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in alrender.h but not exported"
    assert sorted(s[0] for s in _lib.SYMBOLS) == names


def test_struct_sizes_match_header():
    from audiblelight_b200 import _lib
    lib = _lib.load()  # load() itself raises on a mismatch
    for which, mirror in enumerate((_lib.AlrEvent, _lib.AlrScene, _lib.AlrEventStats, _lib.AlrProfile, _lib.AlrAugOp)):
        assert lib.alr_struct_size(which) == C.sizeof(mirror)
    assert lib.alr_struct_size(99) == -1


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from audiblelight_b200 import _lib
    from audiblelight_b200.renderer import Renderer
    with pytest.raises(_lib.AlrenderError, match="no CUDA device|no CPU fallback"):
        Renderer(0)


def test_pack_rejects_non_float32_buffers():
    """The binding hands raw pointers to the library: float64 input must fail loudly, not be reinterpreted."""
    import pytest
    from audiblelight_b200.renderer import EventJob, Renderer, SceneJob
    r = Renderer.__new__(Renderer)  # pack() needs no device context
    ok = EventJob(audio=np.zeros(10, np.float32), irs=np.zeros((2, 1, 5), np.float32), n_channels=2)
    r.pack([ok], [])
    with pytest.raises(TypeError, match="audio must be float32"):
        r.pack([EventJob(audio=np.zeros(10), irs=np.zeros((2, 1, 5), np.float32), n_channels=2)], [])
    with pytest.raises(TypeError, match="irs must be float32"):
        r.pack([EventJob(audio=np.zeros(10, np.float32), irs=np.zeros((2, 1, 5)), n_channels=2)], [])
    with pytest.raises(TypeError, match="ambience must be float32"):
        r.pack([ok], [SceneJob(n_channels=2, n_samples=20, ambience=[np.zeros((2, 20))], ambience_ref_db=[-65.0])])
    with pytest.raises(ValueError, match="audio must be C-contiguous"):
        r.pack([EventJob(audio=np.zeros(20, np.float32)[::2], irs=None, n_channels=2)], [])


def test_pinned_pool_reuses_blocks_best_fit():
    """PinnedPool (host staging for the drop-in): blocks are recycled by best fit, so batches whose arrays differ a
    little in size do not allocate again (cudaHostAlloc per batch was slower than pageable copies)."""
    from audiblelight_b200.renderer import PinnedPool

    class FakeLib:  # stands in for libalrender's alr_pinned_alloc / alr_pinned_free
        def __init__(self):
            self.allocs, self.frees, self.keep = 0, 0, []

        def alr_pinned_alloc(self, h, nbytes, out):
            buf = (C.c_byte * nbytes)()
            self.keep.append(buf)
            out._obj.value = C.addressof(buf)
            self.allocs += 1
            return 0

        def alr_pinned_free(self, h, p):
            self.frees += 1

    lib = FakeLib()
    pool = PinnedPool(lib, None)
    rng = np.random.default_rng(0)
    seen = []
    for it in range(6):
        pool.recycle()
        arrs = [pool.take((4, int(rng.integers(21, 101)), 2400)) for _ in range(12)]
        for k, a in enumerate(arrs):
            assert a.dtype == np.float32 and a.flags.c_contiguous
            a[...] = k  # blocks handed out in one round must not overlap
        assert all((a == k).all() for k, a in enumerate(arrs))
        seen.append(lib.allocs)
    # 6 rounds x 12 arrays of random sizes: far fewer allocations than arrays, and (almost) none once warmed up
    assert seen[-1] <= 2 * 12 and seen[-1] - seen[-2] <= 1
    pcm = pool.take((1000, 4), np.int16)
    assert pcm.dtype == np.int16 and pcm.shape == (1000, 4)
    pool.close()
    assert lib.frees == lib.allocs


def test_pooled_float32_conversion_matches_numpy():
    import audiblelight_b200.synthesize as syn
    a = np.random.default_rng(1).standard_normal((4, 50, 12000))  # float64, above the threading threshold
    view = a[:, 7:40, :]                                             # non-contiguous slice, as the drop-in gets them
    out = syn._as_f32(view)
    assert out.dtype == np.float32 and out.flags.c_contiguous and np.array_equal(out, view.astype(np.float32))
    small = np.arange(10, dtype=np.float32)
    assert syn._as_f32(small) is small
