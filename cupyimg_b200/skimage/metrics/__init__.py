from ._structural_similarity import structural_similarity  # noqa: F401

__all__ = ["structural_similarity"]
