"""``convolve_separable`` (reference cupyimg/_misc.py:39-77): n-d convolution by one
``convolve1d`` per axis, with one shared or one per-axis 1-D kernel."""
from . import _array
from .scipy.ndimage.filters import convolve1d

__all__ = ["convolve_separable"]


def convolve_separable(x, w, axes=None, **kwargs):
    """Filter ``x`` along ``axes`` (default: all) with the 1-D kernel(s) ``w``."""
    ndim = _array.ingest(x, "x").ndim
    if axes is None:
        axes = range(ndim)
    axes = tuple(axes)
    if any(ax < -ndim or ax > ndim - 1 for ax in axes):
        raise ValueError("axis out of range")
    if _array.is_device_array(w) or (hasattr(w, "ndim") and getattr(w, "ndim") == 1):
        w = [w] * len(axes)
    elif len(w) != len(axes):
        raise ValueError("user should supply one filter per axis")
    for ax, w0 in zip(axes, w):
        w0 = _array.host_weights(w0)
        if w0.ndim != 1:
            raise ValueError("w must be a 1d array (or sequence of 1d arrays)")
        x = convolve1d(x, w0, axis=ax, **kwargs)
    return x
