"""``FLDataset``: per-client train batch streams + per-client test sets
(reference datasets/dataset.py:80-115).  Client ids are re-keyed to 0..n-1 in the
order given (reference :102-105).

B200 addition: ``get_train_batches`` returns the batches of MANY clients stacked
into one pinned host tensor ``[n_clients, k, B, ...]`` so a trainer shard uploads a
whole round's input with a single async H2D copy (the reference does one small
``.to(device)`` per batch per client, client.py:186)."""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple
from warnings import warn

import torch

__all__ = ["FLDataset"]


class FLDataset:
    def __init__(self, train_dataloaders: list, test_dataloaders: list = None) -> None:
        if not test_dataloaders:
            warn("No test data is given. Model evaluation will be based on train data. ")
            test_dataloaders = train_dataloaders
        if len(train_dataloaders) != len(test_dataloaders):
            raise Exception("Invalid Input: Numbers of train dataloaders and test dataloaders should be equal. ")
        self._train_dls = dict(enumerate(train_dataloaders))
        self._test_dls = dict(enumerate(test_dataloaders))
        self._clients = list(range(len(self._train_dls)))
        self._pinned = {}

    def get_clients(self) -> List[int]:
        return self._clients

    def get_train_data(self, u_id, num_batches: int):
        stream = self._train_dls[u_id]
        return [next(stream) for _ in range(num_batches)]

    def get_all_test_data(self, u_id):
        return self._test_dls[u_id]

    # ------------------------------------------------------------------ batched access
    def get_train_batches(self, client_ids: Sequence[int], num_batches: int, pin: bool = True
                          ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Stack ``num_batches`` batches of each listed client:
        ``X[len(ids), k, B, ...]`` float32 and ``y[len(ids), k, B]`` int64, in reusable pinned memory.
        All clients must yield equal batch shapes (true for the built-in generators
        unless a client's shard is smaller than one batch)."""
        rows = [self.get_train_data(c, num_batches) for c in client_ids]
        x0, y0 = rows[0][0]
        shape_x = (len(client_ids), num_batches) + tuple(x0.shape)
        shape_y = (len(client_ids), num_batches) + tuple(y0.shape)
        key = (shape_x, shape_y)
        buf = self._pinned.get(key)
        if buf is None:
            can_pin = pin and torch.cuda.is_available()
            bx = torch.empty(shape_x, dtype=torch.float32, pin_memory=can_pin)
            by = torch.empty(shape_y, dtype=torch.int64, pin_memory=can_pin)
            buf = self._pinned[key] = (bx, by)
        bx, by = buf
        for i, batches in enumerate(rows):
            for j, (x, y) in enumerate(batches):
                if tuple(x.shape) != tuple(x0.shape):
                    raise ValueError("ragged batch shapes; use the time-sliced engine")
                bx[i, j].copy_(x)
                by[i, j].copy_(y)
        return bx, by

    def state_dict(self) -> dict:
        """Data cursors for checkpoint/resume (generators that expose ``state()``)."""
        out = {}
        for u, s in self._train_dls.items():
            if hasattr(s, "state"):
                out[u] = s.state()
        return out

    def load_state_dict(self, state: dict) -> None:
        for u, st in state.items():
            if hasattr(self._train_dls.get(u), "load_state"):
                self._train_dls[u].load_state(st)
