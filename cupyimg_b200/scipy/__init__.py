"""``cupyimg_b200.scipy`` — mirrors the ``cupyimg.scipy`` namespace for the hot path."""
from . import ndimage  # noqa: F401
