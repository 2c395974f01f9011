"""AutoGM (reference aggregators/autogm.py:15-65): alternate a weighted geometric
median with a water-filling update of the weights.  Entirely in the Gram domain."""
from __future__ import annotations

from typing import Optional

from . import _gramops as gops
from .base import _BaseAggregator
from .geomed import Geomed

__all__ = ["Autogm"]


class Autogm(_BaseAggregator):
    def __init__(self, lamb: Optional[float] = None, maxiter: Optional[int] = 100,
                 eps: Optional[float] = 1e-6, ftol: Optional[float] = 1e-10, compat: bool = True):
        super().__init__()
        self.lamb = lamb
        self.maxiter = maxiter
        self.eps = eps
        self.ftol = ftol
        self.compat = compat
        self.gm_agg = Geomed(maxiter=maxiter, eps=eps, ftol=ftol, compat=compat)

    def geometric_median_objective(self, median, points, alphas):
        """``sum_i alpha_i * ||median - p_i||`` (reference autogm.py:33-34)."""
        return self.gm_agg._geometric_median_objective(median, points, alphas)

    def aggregate(self, matrix, weights=None):
        dg = matrix.gram_device()
        if dg is not None:               # water filling + Weiszfeld on the device (csrc/cuda/gram_solve.cu): no host sync
            from ..ops import gram_solve
            return matrix.combine(gram_solve.autogm_weights(dg, self.lamb, self.maxiter, self.eps, self.ftol,
                                                            self.compat, self.compat))
        w = gops.autogm_weights(matrix.gram(), self.lamb, self.maxiter, self.eps, self.ftol,
                                sort_by_index=self.compat, compounding=self.compat)
        return matrix.combine(w)

    def __call__(self, inputs, weights=None):
        return self.aggregate(self._matrix(inputs))

    def __str__(self):
        return "AutoGM"
