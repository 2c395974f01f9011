"""Coordinate-wise trimmed mean (reference aggregators/trimmedmean.py:23-42):
drop the ``nb`` largest and ``nb`` smallest values per coordinate, average the rest.
If ``N - 2*nb <= 0`` the trim silently shrinks until valid (quirk Q4)."""
from .base import _BaseAggregator

__all__ = ["Trimmedmean"]


class Trimmedmean(_BaseAggregator):
    def __init__(self, nb: int = 5):
        super().__init__()
        self.b = nb

    def effective_trim(self, n: int) -> int:
        b = self.b
        while n - 2 * b <= 0:
            b -= 1
        if b < 0:
            raise RuntimeError("trimmed mean needs at least one row")
        return b

    def aggregate(self, matrix):
        return matrix.trimmed_mean(self.effective_trim(matrix.n_rows))

    def __str__(self):
        return "Trimmed Mean (b={})".format(self.b)
