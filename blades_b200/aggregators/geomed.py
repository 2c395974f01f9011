"""Geometric median by Weiszfeld iterations (reference aggregators/geomed.py:35-84),
solved in the Gram domain: one tcgen05 Gram pass + one weighted row-combine,
independent of ``maxiter``.  ``compat=True`` keeps the compounding-weights quirk Q5."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _gramops as gops
from .base import _BaseAggregator

__all__ = ["Geomed", "smoothed_weiszfeld"]


def smoothed_weiszfeld(weights, alphas, z, eps=1e-6, T=5):
    """Textbook smoothed Weiszfeld on explicit vectors (reference geomed.py:14-32; unused there)."""
    if len(alphas) != len(weights):
        raise ValueError("alphas and weights must have equal length")
    if eps < 0:
        raise ValueError("eps must be non-negative")
    pts = torch.stack(list(weights))
    a = torch.as_tensor(alphas, dtype=pts.dtype)
    for _ in range(T):
        dist = (pts - z).norm(dim=1).clamp_min(eps)
        beta = (a / dist).clamp_min(eps)
        z = (beta[:, None] * pts).sum(0) / beta.sum()
    return z


class Geomed(_BaseAggregator):
    def __init__(self, maxiter: Optional[int] = 100, eps: Optional[float] = 1e-6,
                 ftol: Optional[float] = 1e-10, compat: bool = True):
        super().__init__()
        self.maxiter = maxiter
        self.eps = eps
        self.ftol = ftol
        self.compat = compat
        self._iters = 0            # int, or a 1-element device tensor written by the on-device solver

    def _geometric_median_objective(self, median, points, alphas):
        """``sum_i alpha_i * ||median - p_i||`` (reference geomed.py:61-62; the solver itself works on the Gram matrix)."""
        return sum(float(a) * torch.linalg.norm(median - p) for a, p in zip(alphas, points))

    @property
    def last_iterations(self) -> int:
        """Weiszfeld iterations of the last call (reading it after a device solve synchronises on that kernel)."""
        if torch.is_tensor(self._iters):
            self._iters = int(self._iters.item())
        return self._iters

    @last_iterations.setter
    def last_iterations(self, v) -> None:
        self._iters = v

    def weights_from_gram(self, G: np.ndarray, alphas=None) -> np.ndarray:
        w, self._iters = gops.weiszfeld_weights(
            G, alphas, self.maxiter, self.eps, self.ftol, compounding=self.compat)
        return w

    def aggregate(self, matrix, weights=None):
        dg = matrix.gram_device()
        if dg is not None:
            # Weiszfeld iterations on the device (csrc/cuda/gram_solve.cu): no D2H copy of G, no host sync
            from ..ops import gram_solve
            w, self._iters = gram_solve.weiszfeld_weights(dg, weights, self.maxiter, self.eps, self.ftol, self.compat)
            return matrix.combine(w)
        return matrix.combine(self.weights_from_gram(matrix.gram(), weights))

    def __call__(self, inputs, weights=None):
        return self.aggregate(self._matrix(inputs), weights)

    def __str__(self):
        return "GeoMed"
