"""Clipped clustering (reference aggregators/clippedclustering.py:20-66).

Stateful: keeps every update norm ever seen; the clip threshold is ``tau`` or the
median of that history.  Cosine distance is scale-invariant, so clipping only
changes the final average: rows of the majority cluster are combined with weights
``s_i / |majority|`` where ``s_i = min(1, thr / (||u_i|| + 1e-6))`` for rows whose
norm exceeds the threshold (torch_utils.clip_tensor_norm_ semantics)."""
from __future__ import annotations

import numpy as np

from . import _gramops as gops
from .base import _BaseAggregator

__all__ = ["Clippedclustering"]


class Clippedclustering(_BaseAggregator):
    def __init__(self, tau=None) -> None:
        super().__init__()
        self.tau = tau
        self.l2norm_his = []
        self.last_labels = None

    def aggregate(self, matrix):
        G = matrix.gram()
        norms = np.sqrt(np.maximum(np.diag(G), 0.0))
        self.l2norm_his.extend(float(x) for x in norms)
        thr = self.tau if self.tau else float(np.median(self.l2norm_his))
        scale = np.where(norms > thr, np.minimum(1.0, thr / (norms + 1e-6)), 1.0)
        dist = 1.0 - gops.cosine_matrix(G)
        np.fill_diagonal(dist, 0.0)
        dist[dist == -np.inf] = 0
        dist[dist == np.inf] = 2
        dist[np.isnan(dist)] = 2
        labels = gops.complete_linkage_2(dist)
        self.last_labels = labels
        keep = gops.majority_cluster(labels)
        w = np.where(keep, scale, 0.0) / max(int(keep.sum()), 1)
        return matrix.combine(w)

    def state_dict(self):
        return {"l2norm_his": list(self.l2norm_his)}

    def load_state_dict(self, state):
        self.l2norm_his = list(state.get("l2norm_his", []))

    def __str__(self):
        return "Clippedclustering (tau={})".format(self.tau)
