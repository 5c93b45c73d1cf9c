"""CPU restatement (numpy, on top of oracle.oracle's filters) of the skimage-level consumers — TEST INFRASTRUCTURE
ONLY, like the rest of oracle/: imported by tests/ and nothing under cupyimg_b200/.

Each function follows the cited reference lines.  Pinned by tests/test_consumers.py against the literal examples
of the reference's docstrings (gaussian: _gaussian.py:86-101; structure_tensor: corner.py:89-98) and against the
same formulas evaluated with scipy.ndimage's filters."""
from itertools import combinations_with_replacement

import numpy as np

from . import oracle


def convert_to_float(image, preserve_range):
    """skimage/_shared/utils.py convert_to_float + util.dtype.img_as_float for float and unsigned inputs."""
    image = np.asarray(image)
    if image.dtype.kind == "f":
        return image
    if preserve_range:
        return image.astype(np.float64)
    if image.dtype.kind == "u":
        return np.multiply(image, 1.0 / np.iinfo(image.dtype).max, dtype=np.float64)
    if image.dtype.kind == "b":
        return image.astype(np.float64)
    raise NotImplementedError


def gaussian(image, sigma=1, mode="nearest", cval=0, multichannel=False, preserve_range=False, truncate=4.0):
    """skimage/filters/_gaussian.py:13-145."""
    image = np.asarray(image)
    if multichannel:
        if np.isscalar(sigma):
            sigma = [sigma] * (image.ndim - 1)
        if len(sigma) != image.ndim:
            sigma = tuple(sigma) + (0,)
    image = convert_to_float(image, preserve_range)
    return oracle.gaussian_filter(image, sigma, mode=mode, cval=cval, truncate=truncate)


def difference_of_gaussians(image, low_sigma, high_sigma=None, mode="nearest", cval=0, truncate=4.0):
    """_gaussian.py:178-290 (single-channel)."""
    image = convert_to_float(np.asarray(image), False)
    low = np.array(low_sigma, dtype="float", ndmin=1) * np.ones(image.ndim)
    high = low * 1.6 if high_sigma is None else np.array(high_sigma, dtype="float", ndmin=1) * np.ones(image.ndim)
    im1 = oracle.gaussian_filter(image, tuple(low), mode=mode, cval=cval, truncate=truncate)
    im2 = oracle.gaussian_filter(image, tuple(high), mode=mode, cval=cval, truncate=truncate)
    return im1 - im2


def structure_tensor(image, sigma=1, mode="constant", cval=0, order="rc"):
    """skimage/feature/corner.py:17-41, :44-136."""
    image = convert_to_float(np.asarray(image), False)
    derivatives = [oracle.sobel(image, axis=i, mode=mode, cval=cval) for i in range(image.ndim)]
    if order == "xy":
        derivatives = list(reversed(derivatives))
    return [oracle.gaussian_filter(d0 * d1, sigma, mode=mode, cval=cval)
            for d0, d1 in combinations_with_replacement(derivatives, 2)]


def structural_similarity(im1, im2, win_size=None, data_range=None, gaussian_weights=False, full=False,
                          data_dtype=np.float64, K1=0.01, K2=0.03, sigma=1.5, use_sample_covariance=True,
                          filters=None):
    """skimage/metrics/_structural_similarity.py:143-260 (single channel, no gradient).  ``filters``: a namespace
    with uniform_filter / gaussian_filter (default: the oracle's; tests also pass scipy.ndimage)."""
    f = filters or oracle
    im1, im2 = np.asarray(im1), np.asarray(im2)
    truncate = 3.5
    if win_size is None:
        win_size = 2 * int(truncate * sigma + 0.5) + 1 if gaussian_weights else 7
    if data_range is None:
        rng = {"u": lambda d: np.iinfo(d).max - np.iinfo(d).min, "i": lambda d: np.iinfo(d).max - np.iinfo(d).min,
               "f": lambda d: 2}[im1.dtype.kind]
        data_range = rng(im1.dtype)
    ndim = im1.ndim
    if gaussian_weights:
        filt = lambda a: f.gaussian_filter(a, sigma, mode="reflect", truncate=truncate)
    else:
        filt = lambda a: f.uniform_filter(a, size=win_size, mode="reflect")
    im1 = im1.astype(data_dtype, copy=False)
    im2 = im2.astype(data_dtype, copy=False)
    NP = win_size ** ndim
    cov_norm = NP / (NP - 1) if use_sample_covariance else 1.0
    ux, uy = filt(im1), filt(im2)
    uxx, uyy, uxy = filt(im1 * im1), filt(im2 * im2), filt(im1 * im2)
    T = np.dtype(data_dtype).type
    cov, C1, C2, two = T(cov_norm), T((K1 * data_range) ** 2), T((K2 * data_range) ** 2), T(2)
    vx = cov * (uxx - ux * ux)
    vy = cov * (uyy - uy * uy)
    vxy = cov * (uxy - ux * uy)
    A1, A2, B1, B2 = two * ux * uy + C1, two * vxy + C2, ux * ux + uy * uy + C1, vx + vy + C2
    D = B1 * B2
    S = (A1 * A2) / D
    pad = (win_size - 1) // 2
    crop = S[tuple(slice(pad, s - pad) for s in S.shape)]
    mssim = crop.astype(np.float64).mean()
    return (mssim, S) if full else mssim
