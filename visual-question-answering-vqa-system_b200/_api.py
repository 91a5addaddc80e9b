"""Public names of the package (kept import-light: the CUDA library loads lazily)."""
__all__ = []
