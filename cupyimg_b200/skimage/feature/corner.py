"""``skimage.feature.structure_tensor`` (reference cupyimg/skimage/feature/corner.py:17-41, :44-136): Sobel
derivatives, their pairwise products, and a Gaussian filter of every product — all separable filters of this
library plus one elementwise multiply per tensor element."""
import warnings
from itertools import combinations_with_replacement

from ... import _array
from ...scipy.ndimage import filters as _filters
from .._util import convert_to_float, elementwise_multiply

__all__ = ["structure_tensor"]


def _compute_derivatives(img, mode="constant", cval=0):
    """corner.py:17-41: one Sobel derivative per axis."""
    return [_array.ingest(_filters.sobel(img, axis=i, mode=mode, cval=cval)) for i in range(img.ndim)]


def structure_tensor(image, sigma=1, mode="constant", cval=0, order=None):
    """Upper-diagonal elements of the structure tensor (a list of arrays, like the reference)."""
    inp = _array.ingest(image)
    if order == "xy" and inp.ndim > 2:
        raise ValueError('Only "rc" order is supported for dim > 2.')
    if order is None:
        if inp.ndim == 2:
            warnings.warn('deprecation warning: the default order of the structure tensor values will be '
                          '"row-column" instead of "xy" starting in skimage version 0.20. Use order="rc" or '
                          'order="xy" to set this explicitly.  (Specify order="xy" to maintain the old behavior.)',
                          category=FutureWarning, stacklevel=2)
            order = "xy"
        else:
            order = "rc"
    img = convert_to_float(inp, preserve_range=False)          # _prepare_grayscale_input_nD: img_as_float
    derivatives = _compute_derivatives(img, mode=mode, cval=cval)
    if order == "xy":
        derivatives = list(reversed(derivatives))
    elems = []
    for d0, d1 in combinations_with_replacement(derivatives, 2):
        prod = elementwise_multiply(d0, d1)
        res = _filters.gaussian_filter(prod, sigma, mode=mode, cval=cval)
        elems.append(_array.export(_array.ingest(res), inp))
    return elems
