from .corner import structure_tensor  # noqa: F401

__all__ = ["structure_tensor"]
