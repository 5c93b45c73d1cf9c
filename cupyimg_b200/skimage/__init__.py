"""skimage-level callers of the separable filters: the parts of ``cupyimg.skimage`` whose cost is
separable filtering, with their elementwise stages fused into hand-written kernels (csrc/consumers.cu)."""
from . import feature, filters, metrics  # noqa: F401
