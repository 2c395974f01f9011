"""``FLDataset``: per-client train batch streams + per-client test sets
(reference datasets/dataset.py:80-115).  Client ids are re-keyed to 0..n-1 in the
order given (reference :102-105).

B200 addition: ``get_train_batches`` returns the batches of MANY clients stacked
into one pinned host tensor ``[n_clients, k, B, ...]`` so a trainer shard uploads a
whole round's input with a single async H2D copy (the reference does one small
``.to(device)`` per batch per client, client.py:186)."""
from __future__ import annotations

from typing import List, Sequence, Tuple
from warnings import warn

import numpy as np
import torch

__all__ = ["FLDataset", "RaggedBatches"]


class RaggedBatches(Exception):
    """Raised by ``FLDataset.get_train_batches`` when the requested batches cannot be stacked into one tensor
    (a client reached the short tail batch of its shard, or shards of different sizes are out of phase).  The
    batches have ALREADY been drawn from the streams; they travel with the exception so that nothing is lost:
    ``batches[client_id] = [(X[b, ...] float32, y[b] int64), ...]`` (host tensors, transforms applied)."""

    def __init__(self, batches):
        super().__init__("batches of different sizes cannot be stacked; train them per size group")
        self.batches = batches


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
    def get_train_batches(self, client_ids: Sequence[int], num_batches: int, pin: bool = True, slot: int = 0
                          ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Stack ``num_batches`` batches of each listed client:
        ``X[len(ids), k, B, ...]`` float32 and ``y[len(ids), k, B]`` int64, in reusable pinned memory.
        All batches must have one shape; otherwise (the reference simply trains on the short tail batch of a
        shard, basedataset.py:76-86) ``RaggedBatches`` is raised carrying the batches that were drawn."""
        fast = self._native_gather(client_ids, num_batches, pin, slot)
        if fast is not None:
            return fast
        rows = [self.get_train_data(c, num_batches) for c in client_ids]
        x0, y0 = rows[0][0]
        if any(tuple(x.shape) != tuple(x0.shape) for batches in rows for x, _ in batches):
            raise RaggedBatches({c: batches for c, batches in zip(client_ids, rows)})
        shape_x = (len(client_ids), num_batches) + tuple(x0.shape)
        shape_y = (len(client_ids), num_batches) + tuple(y0.shape)
        key = (shape_x, shape_y, slot)          # ``slot``: independent staging buffers for double buffering
        buf = self._pinned.get(key)
        if buf is None:
            can_pin = pin and torch.cuda.is_available()
            bx = torch.empty(shape_x, dtype=torch.float32, pin_memory=can_pin)
            by = torch.empty(shape_y, dtype=torch.int64, pin_memory=can_pin)
            buf = self._pinned[key] = (bx, by)
        bx, by = buf
        for i, batches in enumerate(rows):
            for j, (x, y) in enumerate(batches):
                bx[i, j].copy_(x)
                by[i, j].copy_(y)
        return bx, by

    def _native_gather(self, client_ids, num_batches, pin, slot=0):
        """Multi-threaded C++ batch assembly (csrc/host: bl_gather_batches) straight into the pinned
        staging buffer -- used when every stream is an untransformed float32 ``BatchStream``."""
        from .basedataset import BatchStream
        try:
            from ..ops import host
        except Exception:
            return None
        if not host.available():
            return None
        streams = [self._train_dls[c] for c in client_ids]
        if not all(isinstance(s, BatchStream) and s.transform is None and s.data.dtype == np.float32
                   and s.data.flags.c_contiguous and s.labels.dtype == np.int64 and s.labels.flags.c_contiguous
                   and len(s.labels) >= s.batch_size for s in streams):
            return None
        bs = streams[0].batch_size
        shp = streams[0].data.shape[1:]
        if not all(s.batch_size == bs and s.data.shape[1:] == shp for s in streams):
            return None
        n = len(streams)
        drawn = [[s.next_indices() for _ in range(num_batches)] for s in streams]
        lens = {len(sl) for per_client in drawn for sl in per_client}
        if len(lens) != 1:
            # short tail batches somewhere: hand the drawn batches to the caller (nothing is re-drawn or lost)
            raise RaggedBatches({c: [(torch.from_numpy(np.ascontiguousarray(s.data[sl])).float(),
                                      torch.from_numpy(np.ascontiguousarray(s.labels[sl])).long()) for sl in per_client]
                                 for c, s, per_client in zip(client_ids, streams, drawn)})
        bs = lens.pop()                       # every batch has this size (the shard's tail batch may be shorter)
        per = num_batches * bs
        idx = np.empty((n, per), dtype=np.int64)
        for i, per_client in enumerate(drawn):
            for j, sl in enumerate(per_client):
                idx[i, j * bs:(j + 1) * bs] = sl
        shape_x = (n, num_batches, bs) + tuple(shp)
        shape_y = (n, num_batches, bs)
        key = (shape_x, shape_y, slot)
        buf = self._pinned.get(key)
        if buf is None:
            can_pin = pin and torch.cuda.is_available()
            buf = self._pinned[key] = (torch.empty(shape_x, dtype=torch.float32, pin_memory=can_pin),
                                       torch.empty(shape_y, dtype=torch.int64, pin_memory=can_pin))
        bx, by = buf
        import ctypes as C
        src_x = (C.c_void_p * n)(*[s.data.ctypes.data for s in streams])
        src_y = (C.c_void_p * n)(*[s.labels.ctypes.data for s in streams])
        sample_bytes = int(np.prod(shp)) * 4
        host._lib().bl_gather_batches(C.cast(src_x, C.c_void_p), C.cast(src_y, C.c_void_p), idx.reshape(-1), n, per,
                                      sample_bytes, bx.data_ptr(), by.data_ptr())
        return bx, by

    # ------------------------------------------------------------------ zero-copy device gather
    def device_gather_plan(self, client_ids: Sequence[int]):
        """If every listed client streams untransformed float32 samples, pin the per-client arrays once and
        return ``(streams, sample_shape, batch_size)`` so a GPU kernel can gather mini-batches straight from
        pinned host memory (``ops.gather``); otherwise ``None``."""
        from .basedataset import BatchStream
        streams = [self._train_dls[c] for c in client_ids]
        if not streams or not all(isinstance(s, BatchStream) and s.transform is None and s.data.dtype == np.float32
                                  and s.labels.dtype == np.int64 and len(s.labels) >= s.batch_size
                                  and len(s.labels) % s.batch_size == 0 for s in streams):
            return None                                  # (ragged tail batches go through the host path)
        bs, shp = streams[0].batch_size, streams[0].data.shape[1:]
        if not all(s.batch_size == bs and s.data.shape[1:] == shp for s in streams):
            return None
        for s in streams:
            if not getattr(s, "_pinned", False):
                px = torch.from_numpy(np.ascontiguousarray(s.data)).pin_memory()
                py = torch.from_numpy(np.ascontiguousarray(s.labels)).pin_memory()
                s._pin_keep = (px, py)                 # keep the pinned storage alive
                s.data, s.labels = px.numpy(), py.numpy()
                s._pinned = True
        return streams, tuple(shp), bs

    def state_dict(self) -> dict:
        """Data cursors for checkpoint/resume (generators that expose ``state()``)."""
        out = {}
        for u, s in self._train_dls.items():
            if hasattr(s, "state"):
                out[u] = s.state()
        return out

    def stream_state(self, u_id):
        """Cursor of one client's stream (``None`` when the stream is not resumable)."""
        s = self._train_dls.get(u_id)
        return s.state() if hasattr(s, "state") else None

    def load_state_dict(self, state: dict) -> None:
        for u, st in state.items():
            if st is not None and hasattr(self._train_dls.get(u), "load_state"):
                self._train_dls[u].load_state(st)
