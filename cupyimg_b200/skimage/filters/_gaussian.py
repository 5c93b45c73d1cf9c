"""``skimage.filters.gaussian`` / ``difference_of_gaussians`` (reference cupyimg/skimage/filters/_gaussian.py:13-145,
:178-290): thin callers of ``gaussian_filter`` — the float conversion, the channel rule and the subtraction run
through the C ABI, the filtering through the fused kernels."""
import warnings
from collections.abc import Iterable

import numpy as np

from ... import _array
from ...scipy.ndimage import filters as _filters
from .._util import convert_to_float

__all__ = ["gaussian", "difference_of_gaussians"]


def _guess_spatial_dimensions(image):
    """_gaussian.py:148-175."""
    if image.ndim == 2:
        return 2
    if image.ndim == 3 and image.shape[-1] != 3:
        return 3
    if image.ndim == 3 and image.shape[-1] == 3:
        return None
    if image.ndim == 4 and image.shape[-1] == 3:
        return 3
    raise ValueError("Expected 2D, 3D, or 4D array, got %iD." % image.ndim)


def gaussian(image, sigma=1, output=None, mode="nearest", cval=0, multichannel=None, preserve_range=False,
             truncate=4.0):
    """Multi-dimensional Gaussian filter with skimage's conventions (default mode 'nearest', integer images
    converted to float, the channel axis of an (M, N[, P], 3) image left unfiltered)."""
    inp = _array.ingest(image)
    try:
        spatial_dims = _guess_spatial_dimensions(inp)
    except ValueError:
        spatial_dims = inp.ndim
    if spatial_dims is None and multichannel is None:
        warnings.warn(RuntimeWarning("Images with dimensions (M, N, 3) are interpreted as 2D+RGB by default. Use "
                                     "`multichannel=False` to interpret as 3D image with last dimension of length 3."))
        multichannel = True
    sigma_msg = "Sigma values less than zero are not valid"
    if not isinstance(sigma, Iterable):
        if sigma < 0:
            raise ValueError(sigma_msg)
    elif any(s < 0 for s in sigma):
        raise ValueError(sigma_msg)
    if multichannel:
        if not isinstance(sigma, Iterable):
            sigma = [sigma] * (inp.ndim - 1)
        if len(sigma) != inp.ndim:
            sigma = tuple(sigma) + (0,)             # zero on the channel axis: do not filter across channels
        sigma = tuple(sigma)
    img = convert_to_float(inp, preserve_range)
    if output is not None and _array.is_device_array(output) and _array.ingest(output).dtype.kind != "f":
        raise ValueError("Provided output data type is not float")
    res = _filters.gaussian_filter(img, sigma, output=output, mode=mode, cval=cval, truncate=truncate)
    if output is not None and _array.is_device_array(output):
        return res                                  # the caller's array, like the reference
    return _array.export(_array.ingest(res), inp)


def difference_of_gaussians(image, low_sigma, high_sigma=None, *, mode="nearest", cval=0, multichannel=False,
                            truncate=4.0):
    """Band-pass: gaussian(low_sigma) - gaussian(high_sigma) (_gaussian.py:178-290; high defaults to 1.6 x low)."""
    inp = _array.ingest(image)
    img = convert_to_float(inp, preserve_range=False)
    low = np.array(low_sigma, dtype="float", ndmin=1)
    high = low * 1.6 if high_sigma is None else np.array(high_sigma, dtype="float", ndmin=1)
    spatial_dims = inp.ndim - 1 if multichannel else inp.ndim
    if len(low) != 1 and len(low) != spatial_dims:
        raise ValueError("low_sigma must have length equal to number of spatial dimensions of input")
    if len(high) != 1 and len(high) != spatial_dims:
        raise ValueError("high_sigma must have length equal to number of spatial dimensions of input")
    low = low * np.ones(spatial_dims)
    high = high * np.ones(spatial_dims)
    if any(high < low):
        raise ValueError("high_sigma must be equal to or larger than low_sigma for all axes")
    kw = dict(mode=mode, cval=cval, multichannel=multichannel, truncate=truncate, preserve_range=True)
    im1 = _array.ingest(gaussian(img, tuple(low), **kw))
    im2 = _array.ingest(gaussian(img, tuple(high), **kw))
    _filters._accumulate(im1, im2, 4)               # im1 -= im2, separately rounded in the array dtype
    return _array.export(im1, inp)
