"""Shared helpers of the skimage-level consumers: float conversion and the elementwise C-ABI calls."""
import numpy as np
import torch

from .. import _array, _ffi
from ..scipy.ndimage import filters as _filters


def elementwise_multiply(a, b, out=None):
    """out = a * b on contiguous float32 / float64 device arrays (``sepfilt_multiply``)."""
    if a.dtype != b.dtype or a.shape != b.shape:
        raise ValueError("multiply needs arrays of equal shape and dtype")
    if out is None:
        out = _array.empty(a.shape, a.dtype, a.device)
    if not (a.c_contiguous() and b.c_contiguous() and out.c_contiguous()):
        raise ValueError("multiply needs C-contiguous arrays")
    _ffi.check(_ffi.lib().sepfilt_multiply(a.ptr, b.ptr, out.ptr, a.size, _ffi.DTYPE_CODES[a.dtype],
                                           _array.current_stream(a.device)))
    _ffi.count_launch()
    return out


def convert_to_float(image, preserve_range):
    """``skimage._shared.utils.convert_to_float`` (reference skimage/_shared/utils.py) for the dtypes this path
    serves: float arrays pass through; with ``preserve_range`` integers become float64 unscaled; otherwise
    unsigned integers are multiplied by 1 / max like ``img_as_float`` (``np.multiply(image, 1. / imax)``)."""
    inp = _array.ingest(image)
    if inp.dtype.kind == "f":
        if inp.dtype.itemsize < 4:
            raise NotImplementedError("float16 images are not supported (like scipy.ndimage)")
        return inp
    out = _array.empty(inp.shape, np.float64, inp.device)
    if preserve_range:
        _filters._copy_cast(inp, out)
        return out
    if inp.dtype.kind == "u":
        scale = 1.0 / float(np.iinfo(inp.dtype).max)
    elif inp.dtype.kind == "b":
        scale = 1.0
    else:
        raise NotImplementedError("img_as_float of signed integer images is outside this path; pass "
                                  "preserve_range=True or convert first")
    if inp.ndim == 0 or inp.size == 0:
        _filters._copy_cast(inp, out)
        return out
    # one exact single-tap pass: float64(x) * scale, separately rounded — the bits of np.multiply(image, scale)
    spec = _filters._PassSpec(inp.ndim - 1, np.array([scale]), 0, _filters._check_mode("nearest"))
    _filters._launch_pass(inp, out, spec, 0.0, True)
    return out
