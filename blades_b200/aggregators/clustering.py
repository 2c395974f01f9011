"""Cosine-similarity clustering (reference aggregators/clustering.py:13-44).

The N x N cosine matrix is derived from the Gram pass; two-cluster complete
linkage runs on the host; the majority cluster is averaged with one row-combine.
``compat=True`` keeps quirk Q7 (the *similarity* matrix is handed to the
clustering as if it were a distance); ``compat=False`` clusters on 1 - cos."""
from __future__ import annotations

import numpy as np

from . import _gramops as gops
from .base import _BaseAggregator

__all__ = ["Clustering"]


class Clustering(_BaseAggregator):
    def __init__(self, compat: bool = True):
        super().__init__()
        self.compat = compat
        self.last_labels = None

    def aggregate(self, matrix):
        n = matrix.n_rows
        sim = gops.cosine_matrix(matrix.gram())
        np.fill_diagonal(sim, 1.0)
        sim[sim == -np.inf] = -1
        sim[sim == np.inf] = 1
        sim[np.isnan(sim)] = -1
        labels = gops.complete_linkage_2(sim if self.compat else 1.0 - sim)
        self.last_labels = labels
        keep = gops.majority_cluster(labels)
        w = keep.astype(np.float64) / max(int(keep.sum()), 1)
        return matrix.combine(w)

    def __str__(self):
        return "Clustering"
