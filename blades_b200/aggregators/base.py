"""Aggregator base class.

Call convention is the reference's (aggregators/mean.py:21-28): an aggregator is a
callable ``agg(inputs) -> Tensor[d]`` where ``inputs`` is a list of clients, a list
of tensors, or a dense ``Tensor[N, d]``.  New here: ``inputs`` may also be an
``UpdateMatrix`` (``parallel.matrix``), which is what the engine passes so the
aggregation runs as fused sm_100a kernels over (possibly sharded) device memory.

Subclasses implement ``aggregate(matrix)`` against the matrix primitives; they do
not touch dense ``[N, d]`` tensors unless they call ``matrix.rows()``.
"""
from __future__ import annotations


import torch

from ..parallel.matrix import UpdateMatrix, as_matrix

__all__ = ["_BaseAggregator", "_BaseAsyncAggregator"]


class _BaseAggregator:
    #: set by Simulator: Gram-capable aggregators may fold a VirtualRows attack in
    supports_virtual_rows = True
    #: reference-quirk compatibility switch; default = reference behaviour
    compat = True
    #: the returned vector is exactly the output of the last matrix primitive, so the server step
    #: ``theta += lr * agg`` may be fused into that kernel's epilogue
    fusable_final = True

    def __init__(self, *args, **kwargs):
        pass

    def _get_updates(self, inputs) -> torch.Tensor:
        """Dense ``[N, d]`` view of the inputs (reference helper, kept for subclasses)."""
        return as_matrix(inputs).rows()

    def _matrix(self, inputs) -> UpdateMatrix:
        return as_matrix(inputs)

    def aggregate(self, matrix: UpdateMatrix) -> torch.Tensor:
        raise NotImplementedError

    def __call__(self, inputs):
        return self.aggregate(self._matrix(inputs))

    # checkpointing of stateful aggregators
    def state_dict(self) -> dict:
        return {}

    def load_state_dict(self, state: dict) -> None:
        return None

    def __str__(self) -> str:
        return type(self).__name__


class _BaseAsyncAggregator:
    """Async aggregator base (reference mean.py:42-59; never wired in the reference)."""

    def __init__(self):
        pass

    def __call__(self, inputs):
        raise NotImplementedError
