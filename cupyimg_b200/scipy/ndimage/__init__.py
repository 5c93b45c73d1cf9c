"""``cupyimg_b200.scipy.ndimage`` — the separable-correlation subset of
``cupyimg.scipy.ndimage`` (reference scipy/ndimage/__init__.py:1-15)."""
from .filters import *  # noqa: F401,F403
from .filters import __all__  # noqa: F401
