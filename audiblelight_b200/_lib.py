"""ctypes binding of include/alrender.h (the C-ABI drop-in boundary)."""
import ctypes as C
import os

from .build import library_path

ALR_OK = 0
ALR_ERR_NO_DEVICE = -4
ALR_MEM_HOST, ALR_MEM_DEVICE = 0, 1
ALR_GAIN_EVENT, ALR_GAIN_NONE = 0, 1


class AlrEvent(C.Structure):
    _fields_ = [
        ("audio", C.c_void_p), ("n_audio", C.c_int64),
        ("irs", C.c_void_p), ("ir_stride_c", C.c_int64), ("ir_stride_n", C.c_int64),
        ("n_channels", C.c_int32), ("n_irs", C.c_int32), ("n_ir_samples", C.c_int64),
        ("ir_frames", C.c_void_p), ("n_frames", C.c_int32), ("normalize_irs", C.c_int32),
        ("gain_mode", C.c_int32), ("snr", C.c_double), ("ref_db", C.c_double),
        ("dry_channel", C.c_int32), ("dry_low", C.c_int32), ("dry_high", C.c_int32), ("reserved0", C.c_int32),
        ("dry", C.c_void_p),
        ("spatial", C.c_void_p), ("n_out", C.c_int64),
        ("scene", C.c_int32), ("reserved1", C.c_int32), ("scene_start", C.c_int64), ("scene_end", C.c_int64),
        ("aug_ops", C.c_void_p), ("n_aug_ops", C.c_int32), ("normalize_audio", C.c_int32), ("audio_out", C.c_void_p),
    ]


class AlrAugOp(C.Structure):
    _fields_ = [("type", C.c_int32), ("fade_in_shape", C.c_int32), ("fade_out_shape", C.c_int32),
                ("fade_in_samples", C.c_int32), ("fade_out_samples", C.c_int32), ("reserved", C.c_int32),
                ("p", C.c_double * 6)]


class AlrScene(C.Structure):
    _fields_ = [
        ("n_channels", C.c_int32), ("n_ambience", C.c_int32), ("n_samples", C.c_int64),
        ("ambience", C.c_void_p), ("ambience_ref_db", C.c_void_p), ("mix", C.c_void_p), ("ambience_seed", C.c_void_p),
        ("pcm16", C.c_void_p),
    ]


class AlrEventStats(C.Structure):
    _fields_ = [("peak", C.c_double), ("mean_abs", C.c_double), ("gain", C.c_double), ("event_scale", C.c_double),
                ("nonfinite", C.c_int32), ("dry_peak", C.c_int32)]


class AlrProfile(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_ir_fft", C.c_double), ("ms_x_fft", C.c_double),
                ("ms_cmac", C.c_double), ("ms_cmac_static", C.c_double), ("ms_ifft", C.c_double), ("ms_mix", C.c_double), ("ms_other", C.c_double), ("ms_fused", C.c_double), ("ms_host_plan", C.c_double),
                ("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("workspace_bytes", C.c_int64), ("n_chunks", C.c_int64)]


# every symbol include/alrender.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("alr_version", C.c_int, []),
    ("alr_last_error", C.c_char_p, []),
    ("alr_struct_size", C.c_int, [C.c_int]),
    ("alr_create", C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    ("alr_destroy", None, [C.c_void_p]),
    ("alr_set_workspace_limit", C.c_int, [C.c_void_p, C.c_int64]),
    ("alr_set_option", C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    ("alr_set_profiling", C.c_int, [C.c_void_p, C.c_int]),
    ("alr_render", C.c_int, [C.c_void_p, C.POINTER(AlrEvent), C.c_int64, C.POINTER(AlrScene), C.c_int64, C.c_int,
                             C.POINTER(AlrEventStats), C.c_void_p]),
    ("alr_get_profile", C.c_int, [C.c_void_p, C.POINTER(AlrProfile)]),
    ("alr_visibilities", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_int32,
                                   C.c_double, C.c_int32, C.c_double, C.c_void_p, C.c_int, C.c_void_p]),
    ("alr_partition_size", C.c_int, []),
    ("alr_pinned_alloc", C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    ("alr_pinned_free", None, [C.c_void_p, C.c_void_p]),
    ("alr_debug_rfft", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    ("alr_debug_irfft", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    ("alr_debug_plan_movers", C.c_int, [C.POINTER(AlrEvent), C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    ("alr_debug_plan", C.c_int, [C.POINTER(AlrEvent), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_int64]),
]

_lib = None


class AlrenderError(RuntimeError):
    pass


def load():
    """Loads libalrender.so (built by `audiblelight_b200.build_library()` / `__graft_entry__.build()`).
    Fails loudly when it is missing — there is no Python/CPU fallback for the compute path."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise AlrenderError(
            f"{path} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(nvcc, sm_100a). audiblelight_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    for which, mirror in enumerate((AlrEvent, AlrScene, AlrEventStats, AlrProfile, AlrAugOp)):
        if lib.alr_struct_size(which) != C.sizeof(mirror):
            raise AlrenderError(f"ABI mismatch: {mirror.__name__} is {C.sizeof(mirror)} bytes in Python, "
                                f"{lib.alr_struct_size(which)} in {path}")
    _lib = lib
    return lib


def check(rc: int):
    if rc != ALR_OK:
        msg = load().alr_last_error().decode("utf-8", "replace")
        raise AlrenderError(f"alrender error {rc}: {msg}")
