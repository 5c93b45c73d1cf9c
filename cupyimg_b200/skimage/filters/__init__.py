from ._gaussian import difference_of_gaussians, gaussian  # noqa: F401

__all__ = ["gaussian", "difference_of_gaussians"]
