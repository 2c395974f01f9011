"""Coordinate-wise median (reference aggregators/median.py:21-25).

The reference computes ``(median(U) - median(-U)) / 2`` = mean of the lower and
upper medians.  Here that is one pass of the coordinate-select kernel
(csrc/cuda/coord_select.cu) with k = floor((N-1)/2), ceil((N-1)/2)."""
from .base import _BaseAggregator

__all__ = ["Median"]


class Median(_BaseAggregator):
    def aggregate(self, matrix):
        return matrix.median()

    def __str__(self):
        return "Median"
