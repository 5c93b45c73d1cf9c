"""cupyimg_b200 — B200-native separable n-d correlation behind the cupyimg /
scipy.ndimage API.

One hot path of mritools/cupyimg, rebuilt from scratch for sm_100a:
``cupyimg.scipy.ndimage.filters.correlate1d`` / ``convolve1d`` and the filters built on
them.  Python (this package) validates arguments and hands plain pointers to the C ABI of
``libsepfilt_b200.so`` (include/sepfilt.h); PyTorch only allocates memory and names streams.

    from cupyimg_b200.scipy import ndimage as ndi
    y = ndi.gaussian_filter(x_cuda, sigma=2)          # torch.Tensor or cupy.ndarray in / out
"""
from . import scipy  # noqa: F401
from ._misc import convolve_separable  # noqa: F401

__version__ = "0.1.0"
__all__ = ["convolve_separable", "scipy"]
