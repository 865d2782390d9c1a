"""ctypes binding of `libiunet_b200.so` (C ABI declared in `include/iunet_b200.h`).

The library is the product: if it is missing, stale or cannot create an engine (no sm_100 GPU) the
callers raise -- there is no CPU or PyTorch fallback for the prediction path.
"""
import ctypes
import os
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# IU_LIB: development override (kernel-variant A/B runs built by tools/build_variant.py)
LIB_PATH = os.environ.get("IU_LIB") or os.path.join(_PKG_DIR, "libiunet_b200.so")

ABI_VERSION = 2
IU_OK, IU_ERR_INVALID, IU_ERR_CUDA, IU_ERR_OOM, IU_ERR_STATE = 0, 1, 2, 3, 4
FLAG_ASYNC = 1
DTYPE_U8, DTYPE_F32 = 0, 1
PRECISIONS = {"fp16": 0, "bf16": 1}

_c = ctypes
_engine_p = _c.c_void_p

# name -> (restype, argtypes); mirrors include/iunet_b200.h one to one
SIGNATURES = {
    "iu_abi_version": (_c.c_int, []),
    "iu_last_error": (_c.c_char_p, [_engine_p]),
    "iu_engine_create": (_c.c_int, [_c.c_int, _c.POINTER(_engine_p)]),
    "iu_engine_destroy": (None, [_engine_p]),
    "iu_engine_stream": (_c.c_void_p, [_engine_p]),
    "iu_engine_synchronize": (_c.c_int, [_engine_p]),
    "iu_engine_load_weights": (_c.c_int, [_engine_p, _c.c_int, _c.c_int, _c.POINTER(_c.c_char_p),
                                          _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_int64)]),
    "iu_engine_num_classes": (_c.c_int, [_engine_p]),
    "iu_engine_set_precision": (_c.c_int, [_engine_p, _c.c_int]),
    "iu_engine_precision": (_c.c_int, [_engine_p]),
    "iu_engine_set_max_batch": (_c.c_int, [_engine_p, _c.c_int]),
    "iu_engine_auto_batch": (_c.c_int, [_engine_p, _c.c_int, _c.c_int, _c.c_int]),
    "iu_engine_workspace_bytes": (_c.c_int64, [_engine_p, _c.c_int, _c.c_int, _c.c_int]),
    "iu_engine_forward": (_c.c_int, [_engine_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_uint]),
    "iu_engine_predict_axis": (_c.c_int, [_engine_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                          _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_uint]),
    "iu_engine_predict_slices": (_c.c_int, [_engine_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int64,
                                            _c.c_int64, _c.c_int64, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_uint]),
    "iu_engine_gather_slices": (_c.c_int, [_engine_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                           _c.c_void_p, _c.c_uint]),
    "iu_engine_reduce": (_c.c_int, [_engine_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_int), _c.c_int,
                                    _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_float, _c.c_float,
                                    _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_uint]),
    "iu_engine_reduce_planes": (_c.c_int, [_engine_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_int), _c.c_int,
                                           _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p,
                                           _c.c_float, _c.c_float, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_uint]),
    "iu_engine_predict_volume": (_c.c_int, [_engine_p, _c.c_void_p, _c.c_int, _c.c_int, _c.POINTER(_c.c_int),
                                            _c.c_int, _c.c_void_p, _c.c_float, _c.c_float, _c.c_void_p, _c.c_void_p,
                                            _c.c_void_p, _c.c_uint]),
    "iu_engine_predict_tiled": (_c.c_int, [_engine_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                           _c.POINTER(_c.c_int), _c.POINTER(_c.c_int), _c.c_int, _c.c_void_p, _c.c_float,
                                           _c.c_float, _c.c_void_p, _c.c_void_p, _c.c_uint]),
    "iu_engine_extract_block": (_c.c_int, [_engine_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                           _c.c_int, _c.c_int, _c.c_void_p, _c.c_uint]),
    "iu_engine_blend_block": (_c.c_int, [_engine_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_int), _c.c_int,
                                         _c.c_int, _c.c_int, _c.c_void_p, _c.c_float, _c.c_float, _c.c_void_p,
                                         _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.POINTER(_c.c_int), _c.c_uint]),
    "iu_engine_finalise": (_c.c_int, [_engine_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_void_p,
                                      _c.c_void_p, _c.c_uint]),
    "iu_engine_to_chunks": (_c.c_int, [_engine_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                       _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_uint]),
    "iu_engine_from_chunks": (_c.c_int, [_engine_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                         _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_uint]),
    "iu_engine_zoom_nearest": (_c.c_int, [_engine_p, _c.c_void_p, _c.POINTER(_c.c_int), _c.c_void_p,
                                          _c.POINTER(_c.c_int), _c.POINTER(_c.c_int), _c.POINTER(_c.c_int),
                                          _c.POINTER(_c.c_int), _c.POINTER(_c.c_int), _c.c_int, _c.c_uint]),
    "iu_engine_conv_test": (_c.c_int, [_engine_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int,
                                       _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p,
                                       _c.c_int, _c.c_int, _c.c_void_p]),
    "iu_engine_launch_count": (_c.c_int64, [_engine_p]),
    "iu_engine_release_workspace": (_c.c_int, [_engine_p]),
    "iu_engine_held_bytes": (_c.c_int64, [_engine_p]),
    "iu_engine_debug_counters": (_c.c_int, [_engine_p, _c.POINTER(_c.c_uint64), _c.c_int, _c.c_int]),
    "iu_engine_profile": (_c.c_int, [_engine_p, _c.c_int]),
    "iu_engine_profile_read": (_c.c_int, [_engine_p, _c.POINTER(_c.c_double), _c.POINTER(_c.c_int64), _c.c_int]),
}
PROF_CLASSES = ("gather", "stem", "pool", "conv", "reduce")

_lib = None
_lock = threading.Lock()


class EngineError(RuntimeError):
    """Failure reported by the native library (code in `.code`)."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


def load():
    """Load the shared library (once) and declare every prototype.  Raises if it is not built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is not built; run `python __graft_entry__.py build` (or interactive-unet_b200/build.py). "
                "The prediction path has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        if lib.iu_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libiunet_b200.so ABI {lib.iu_abi_version()} != expected {ABI_VERSION}; rebuild")
        _lib = lib
        return lib


def check(lib, handle, rc):
    """Map a non-zero return code to the exception the reference's callers expect."""
    if rc == IU_OK:
        return
    msg = lib.iu_last_error(handle)
    msg = msg.decode("utf-8", "replace") if msg else f"iunet_b200 error {rc}"
    if rc == IU_ERR_OOM and "out of memory" not in msg:
        msg = "CUDA out of memory. " + msg          # `find_max_batch_size` matches this text (predict.py:67-72)
    raise EngineError(rc, msg)
