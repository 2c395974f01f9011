"""FLTrust (reference aggregators/fltrust.py:8-38): exactly one trusted client;
trust score = relu(cos(trusted, u_i)); every untrusted update is rescaled to the
trusted norm; weighted average by trust score.  Gram pass + one row-combine.

The reference only accepts client lists (needs ``is_trusted()``).  Here a matrix /
tensor input works too when ``trusted_index`` is given."""
from __future__ import annotations

from typing import Optional

from . import _gramops as gops
from .base import _BaseAggregator

__all__ = ["Fltrust"]


class Fltrust(_BaseAggregator):
    def __init__(self, trusted_index: Optional[int] = None):
        super().__init__()
        self.trusted_index = trusted_index

    def __call__(self, inputs):
        from ..client import BladesClient
        seq = list(inputs) if not hasattr(inputs, "n_rows") and not hasattr(inputs, "dim") else None
        if seq is not None and len(seq) and all(isinstance(c, BladesClient) or callable(getattr(c, "is_trusted", None)) for c in seq):
            trusted = [i for i, c in enumerate(seq) if c.is_trusted()]
            assert len(trusted) == 1, "FLTrust needs exactly one trusted client"
            return self.aggregate(self._matrix(seq), trusted[0])
        assert self.trusted_index is not None, "pass clients or set trusted_index"
        return self.aggregate(self._matrix(inputs), self.trusted_index)

    def aggregate(self, matrix, trusted: Optional[int] = None):
        t = self.trusted_index if trusted is None else trusted
        dg = matrix.gram_device()
        if dg is not None:               # trust scores on the device (csrc/cuda/gram_solve.cu): no host sync
            from ..ops import gram_solve
            return matrix.combine(gram_solve.fltrust_weights(dg, t))
        return matrix.combine(gops.fltrust_weights(matrix.gram(), t))

    def __str__(self):
        return "FLTrust"
