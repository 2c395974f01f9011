from .matrix import LocalMatrix, UpdateMatrix, VirtualRows, as_matrix  # noqa: F401
