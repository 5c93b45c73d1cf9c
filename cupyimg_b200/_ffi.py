"""ctypes binding of libsepfilt_b200.so (the C ABI in include/sepfilt.h).

There is no CPU fallback and no JIT: if the shared library has not been built
(``python -m cupyimg_b200._build``) every filter call raises ``RuntimeError``.
"""
import ctypes
import os
import threading

import numpy as np

from . import _build

MAX_NDIM = 8
PARAM_TAPS = 129
MAX_TAPS = 4096
FAST_MAX_RADIUS = 16

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_SCRATCH, ERR_CUDA, ERR_VALUE = 0, -1, -2, -3, -4, -5
ACC_F64_EXACT, ACC_F32 = 0, 1

MODE_CODES = {
    "reflect": 0, "grid-mirror": 0,
    "constant": 1, "grid-constant": 1,
    "nearest": 2,
    "mirror": 3,
    "wrap": 4, "grid-wrap": 4,
}

DTYPE_CODES = {np.dtype(k): v for k, v in {
    "int8": 0, "uint8": 1, "int16": 2, "uint16": 3, "int32": 4, "uint32": 5,
    "int64": 6, "uint64": 7, "float32": 8, "float64": 9, "bool": 10}.items()}


class Tensor(ctypes.Structure):
    _fields_ = [
        ("ptr", ctypes.c_void_p),
        ("dtype", ctypes.c_int32),
        ("ndim", ctypes.c_int32),
        ("shape", ctypes.c_int64 * MAX_NDIM),
        ("stride_bytes", ctypes.c_int64 * MAX_NDIM),
        ("device", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class Pass(ctypes.Structure):
    _fields_ = [
        ("axis", ctypes.c_int32),
        ("ntaps", ctypes.c_int32),
        ("taps", ctypes.POINTER(ctypes.c_double)),
        ("origin", ctypes.c_int32),
        ("mode", ctypes.c_int32),
        ("uniform", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class Halo(ctypes.Structure):
    _fields_ = [
        ("lo", ctypes.c_void_p),
        ("hi", ctypes.c_void_p),
        ("planes_lo", ctypes.c_int32),
        ("planes_hi", ctypes.c_int32),
        ("ready_lo", ctypes.c_void_p),
        ("ready_hi", ctypes.c_void_p),
        ("done_lo", ctypes.c_void_p),
        ("done_hi", ctypes.c_void_p),
        ("cta_counter", ctypes.c_void_p),
        ("epoch", ctypes.c_uint32),
        ("reserved", ctypes.c_uint32),
    ]


EXPORTS = (
    "sepfilt_version", "sepfilt_last_error", "sepfilt_correlate1d", "sepfilt_separable_f32",
    "sepfilt_separable_f32_supported", "sepfilt_gradmag_step", "sepfilt_copy_cast", "sepfilt_correlate_nd",
    "sepfilt_last_launch_count", "sepfilt_separable_f32_halo", "sepfilt_stream_write32", "sepfilt_stream_write32x2", "sepfilt_stream_wait32_geq",
    "sepfilt_multiply", "sepfilt_ssim_map",
)

_lib = None
_lock = threading.Lock()
LAUNCHES = 0          # kernels launched through this binding (bench.py's gpu_launches)


def count_launch(n=1):
    global LAUNCHES
    LAUNCHES += n


def lib():
    """Load (once) and return the ctypes handle; raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("SEPFILT_LIB", _build.LIB_PATH)
        if not os.path.exists(path):
            raise RuntimeError(
                "libsepfilt_b200.so is not built (%s missing); run `python -m cupyimg_b200._build`. "
                "cupyimg_b200 has no CPU or eager fallback." % path)
        L = ctypes.CDLL(path)
        vp, i64, ci, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
        TP, PP = ctypes.POINTER(Tensor), ctypes.POINTER(Pass)
        L.sepfilt_version.restype = ci
        L.sepfilt_last_error.restype = ctypes.c_char_p
        L.sepfilt_correlate1d.argtypes = [TP, TP, PP, i64, dbl, ci, vp, ctypes.c_size_t, vp]
        L.sepfilt_correlate1d.restype = ci
        L.sepfilt_separable_f32.argtypes = [TP, TP, PP, ci, PP, ci, i64, dbl, vp]
        L.sepfilt_separable_f32.restype = ci
        L.sepfilt_separable_f32_supported.argtypes = [TP, TP, PP, ci, ci, dbl]
        L.sepfilt_separable_f32_supported.restype = ci
        L.sepfilt_last_launch_count.restype = ci
        L.sepfilt_separable_f32_halo.argtypes = [TP, TP, PP, ci, PP, ci, ctypes.POINTER(Halo), i64, dbl, vp]
        L.sepfilt_separable_f32_halo.restype = ci
        L.sepfilt_stream_write32.argtypes = [vp, vp, ctypes.c_uint32]
        L.sepfilt_stream_write32.restype = ci
        L.sepfilt_stream_write32x2.argtypes = [vp, vp, vp, ctypes.c_uint32]
        L.sepfilt_stream_write32x2.restype = ci
        L.sepfilt_multiply.argtypes = [vp, vp, vp, i64, ci, vp]
        L.sepfilt_multiply.restype = ci
        L.sepfilt_ssim_map.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ctypes.POINTER(i64), ci, dbl, dbl, dbl, ci, vp]
        L.sepfilt_ssim_map.restype = ci
        L.sepfilt_stream_wait32_geq.argtypes = [vp, vp, ctypes.c_uint32]
        L.sepfilt_stream_wait32_geq.restype = ci
        L.sepfilt_gradmag_step.argtypes = [vp, vp, i64, ci, ci, vp]
        L.sepfilt_gradmag_step.restype = ci
        i32p = ctypes.POINTER(ctypes.c_int32)
        L.sepfilt_correlate_nd.argtypes = [TP, TP, ctypes.POINTER(dbl), i32p, i32p, ci, dbl, vp, ctypes.c_size_t, vp]
        L.sepfilt_correlate_nd.restype = ci
        L.sepfilt_copy_cast.argtypes = [TP, TP, vp]
        L.sepfilt_copy_cast.restype = ci
        _lib = L
    return _lib


def last_error():
    msg = lib().sepfilt_last_error()
    return msg.decode() if msg else ""


def check(rc):
    """Map a sepfilt_status to the Python exception classes the reference uses."""
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_VALUE:                  # axis / origin out of range: ValueError in the reference
        raise ValueError(msg)
    if rc == ERR_INVALID:
        raise RuntimeError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def make_pass(axis, taps, origin, mode_code, uniform=False, size=0):
    """Build a ``Pass``; returns (struct, keepalive)."""
    p = Pass()
    p.axis = int(axis)
    p.origin = int(origin)
    p.mode = int(mode_code)
    p.uniform = int(uniform)               # 0 taps, 1 mean, 2 minimum, 3 maximum
    keep = None
    if uniform:
        p.ntaps = int(size)
        p.taps = None
    else:
        keep = np.ascontiguousarray(taps, dtype=np.float64)
        p.ntaps = int(keep.size)
        p.taps = keep.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    return p, keep
