"""``skimage.metrics.structural_similarity`` (reference cupyimg/skimage/metrics/_structural_similarity.py:17-260).
The five filters (of im1, im2, im1², im2², im1·im2, :197-207) are this library's uniform / Gaussian filters; the
products are one ``sepfilt_multiply`` each and everything after the filters (:208-230: variances, A1 A2 B1 B2, the
SSIM map and its cropped mean) is ONE ``sepfilt_ssim_map`` launch instead of ~20 elementwise kernels."""
import ctypes

import numpy as np
import torch

from ... import _array, _ffi
from ...scipy.ndimage import filters as _filters
from .._util import elementwise_multiply

__all__ = ["structural_similarity"]

_DTYPE_RANGE = {np.dtype(k): v for k, v in {
    "bool": (False, True), "uint8": (0, 255), "uint16": (0, 65535), "uint32": (0, 2 ** 32 - 1),
    "uint64": (0, 2 ** 64 - 1), "int8": (-128, 127), "int16": (-32768, 32767), "int32": (-2 ** 31, 2 ** 31 - 1),
    "int64": (-2 ** 63, 2 ** 63 - 1), "float32": (-1, 1), "float64": (-1, 1)}.items()}


def structural_similarity(im1, im2, *, win_size=None, gradient=False, data_range=None, multichannel=False,
                          gaussian_weights=False, full=False, data_dtype=np.float64, **kwargs):
    """Mean structural similarity index of two images (same arguments and defaults as the reference).
    Returns ``mssim`` (a 0-d float64 device tensor) or ``(mssim, S)`` with ``full=True``."""
    a, b = _array.ingest(im1), _array.ingest(im2)
    if a.shape != b.shape:
        raise ValueError("Input images must have the same dimensions.")
    if gradient:
        raise NotImplementedError("gradient=True (Avanaki 2009, _structural_similarity.py:232-244) is not built")
    if multichannel:
        nch = a.shape[-1]
        args = dict(win_size=win_size, data_range=data_range, multichannel=False, gaussian_weights=gaussian_weights,
                    full=full, data_dtype=data_dtype, **kwargs)
        ta, tb = a.obj if isinstance(a.obj, torch.Tensor) else None, b.obj if isinstance(b.obj, torch.Tensor) else None
        if ta is None or tb is None:
            raise NotImplementedError("multichannel SSIM takes torch tensors")
        vals, maps = [], []
        for ch in range(nch):
            r = structural_similarity(ta[..., ch].contiguous(), tb[..., ch].contiguous(), **args)
            if full:
                vals.append(r[0]); maps.append(r[1])
            else:
                vals.append(r)
        m = torch.stack(vals).mean()
        return (m, torch.stack(maps, dim=-1)) if full else m

    K1 = kwargs.pop("K1", 0.01)
    K2 = kwargs.pop("K2", 0.03)
    sigma = kwargs.pop("sigma", 1.5)
    if K1 < 0:
        raise ValueError("K1 must be positive")
    if K2 < 0:
        raise ValueError("K2 must be positive")
    if sigma < 0:
        raise ValueError("sigma must be positive")
    use_sample_covariance = kwargs.pop("use_sample_covariance", True)
    truncate = 3.5 if gaussian_weights else None       # an 11-tap filter at sigma 1.5, as Wang et al. 2004
    if win_size is None:
        win_size = 2 * int(truncate * sigma + 0.5) + 1 if gaussian_weights else 7
    if any(s < win_size for s in a.shape):
        raise ValueError("win_size exceeds image extent.  If the input is a multichannel (color) image, set "
                         "multichannel=True.")
    if not (win_size % 2 == 1):
        raise ValueError("Window size must be odd.")
    if data_range is None:
        dmin, dmax = _DTYPE_RANGE[a.dtype]
        data_range = dmax - dmin
    ndim = a.ndim
    if ndim < 1 or ndim > 3:
        raise NotImplementedError("SSIM takes 1-D to 3-D images (plus a channel axis with multichannel=True)")
    dt = np.dtype(data_dtype)
    if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise ValueError("data_dtype must be float32 or float64")

    def as_float(x):                                    # im.astype(data_dtype, copy=False), :193-194
        if x.dtype == dt and x.c_contiguous():
            return x
        y = _array.empty(x.shape, dt, x.device)
        _filters._copy_cast(x, y)
        return y

    x, y = as_float(a), as_float(b)
    if gaussian_weights:
        filt = lambda t: _array.ingest(_filters.gaussian_filter(t, sigma, mode="reflect", truncate=truncate))
    else:
        filt = lambda t: _array.ingest(_filters.uniform_filter(t, size=win_size, mode="reflect"))
    NP = win_size ** ndim
    cov_norm = NP / (NP - 1) if use_sample_covariance else 1.0
    ux, uy = filt(x), filt(y)
    tmp = _array.empty(x.shape, dt, x.device)
    uxx = filt(elementwise_multiply(x, x, tmp))
    uyy = filt(elementwise_multiply(y, y, tmp))
    uxy = filt(elementwise_multiply(x, y, tmp))
    R = data_range
    C1, C2 = (K1 * R) ** 2, (K2 * R) ** 2
    pad = (win_size - 1) // 2
    S = _array.empty(x.shape, dt, x.device) if full else None
    dev = torch.device("cuda", x.device)
    total = torch.zeros(1, dtype=torch.float64, device=dev)
    shape = (ctypes.c_int64 * 3)(*(list(x.shape) + [1] * (3 - ndim)))
    _ffi.check(_ffi.lib().sepfilt_ssim_map(ux.ptr, uy.ptr, uxx.ptr, uyy.ptr, uxy.ptr, S.ptr if full else None,
                                           total.data_ptr(), ndim, shape, pad, float(cov_norm), float(C1), float(C2),
                                           _ffi.DTYPE_CODES[dt], _array.current_stream(x.device)))
    _ffi.count_launch()
    count = 1
    for s in x.shape:
        count *= s - 2 * pad
    mssim = (total / float(count)).reshape(())
    if full:
        return mssim, _array.export(S, a)
    return mssim
