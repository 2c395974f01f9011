"""Sample mean and the reference's private async/decentralised variants
(/root/reference/src/blades/aggregators/mean.py:62-116)."""
from __future__ import annotations

import logging

import torch

from .base import _BaseAggregator, _BaseAsyncAggregator

__all__ = ["Mean", "_BaseAggregator", "_BaseAsyncAggregator", "_AsyncMean", "_DecentralizedAggregator"]


class Mean(_BaseAggregator):
    r"""Column mean of the update matrix."""

    def aggregate(self, matrix):
        return matrix.mean()

    def __str__(self):
        return "Mean"


class _AsyncMean(_BaseAsyncAggregator):
    """Mean over the inputs that arrived (``None`` = straggler), divided by ALL slots."""

    def __call__(self, inputs):
        present = [x for x in inputs if x is not None]
        return torch.stack(present, dim=0).sum(dim=0) / len(inputs)

    def __str__(self):
        return "_AsyncMean"


class _DecentralizedAggregator(_BaseAggregator):
    """Gossip-style mixing: ``inputs[0]`` is the node's own vector, the rest its neighbours'."""

    def __init__(self, node, weights):
        super().__init__()
        assert weights.dim() == 1
        self.node = node
        self.weights = weights
        logging.getLogger("debug").info(f"Aggregator: node={node.index} weights={weights}")

    def __call__(self, inputs):
        assert len(inputs) == 1 + len(self.node.edges)
        acc = self.weights[self.node.index] * inputs[0]
        for edge, vec in zip(self.node.edges, inputs[1:]):
            acc = acc + self.weights[edge.theother(self.node).index] * vec
        return acc

    def __str__(self):
        return "_DecentralizedAggregator"
