"""``ShardedMatrix``: the update matrix spread over G trainer shards, rows in NVLink-addressable
symmetric memory, aggregation coordinate-sharded.

Every primitive is ONE pull-mode kernel per rank: rank g owns coordinates ``[c_g, c_{g+1})``, its
CTAs read that coordinate range of EVERY row -- local rows from HBM, peer rows with plain global
loads / TMA on NVLink-mapped pointers -- reduce in registers / tensor cores, and store the result
range into every replica (``agg`` and, when the server step is fused, ``theta``).  The reference's
gather (Ray pickles -> torch.stack, simulator.py:235 + mean.py:23) and broadcast (model pickled to
every actor, simulator.py:222-233) therefore never exist as separate steps.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from ..ops import attack as k_attack
from ..ops import combine as k_combine
from ..ops import select as k_select
from .matrix import UpdateMatrix, VirtualRows

__all__ = ["ShardedMatrix"]


class ShardedMatrix(UpdateMatrix):
    def __init__(self, symm, virtual: Optional[VirtualRows] = None):
        self.symm = symm
        self.n_rows = symm.n_total
        self.n_cols = symm.d
        self.device = symm.world.device
        self.virtual = virtual
        self.server_step = None          # (lr,) -> theta += lr*agg fused into the final primitive
        self.step_applied = False
        self._synced = False

    # ------------------------------------------------------------------ helpers
    def _pre(self):
        if not self._synced:             # all ranks finished writing their rows
            self.symm.barrier()
            self._synced = True

    def _epilogue(self, final: bool):
        s = self.symm
        if final and self.server_step is not None:
            self.step_applied = True
            return k_select.make_epilogue(s.agg_ptrs(), s.theta_ptrs(), s.theta.data_ptr(), float(self.server_step[0]))
        return k_select.make_epilogue(s.agg_ptrs())

    def _finish(self) -> torch.Tensor:
        self.symm.barrier()              # every shard of agg/theta has landed everywhere
        return self.symm.agg

    def materialize_virtual(self):
        v = self.virtual
        if v is not None and v.count:
            self._pre()
            s = self.symm
            byz = set(v.byzantine)
            honest = [i for i in range(self.n_rows) if i not in byz]
            c0, c1 = s.my_cols
            k_attack.attack_rows(s.row_ptrs(honest), s.row_ptrs(list(v.replaced)), v.kind, v.param, c0, c1,
                                 self.device)
            self.virtual = None
            s.barrier()

    # ------------------------------------------------------------------ primitives
    def _select(self, mode: int, b: int) -> torch.Tensor:
        self._pre()
        s = self.symm
        c0, c1 = s.my_cols
        v = self.virtual
        if v is not None and v.count:
            byz, rep = set(v.byzantine), set(v.replaced)
            stat = s.row_ptrs([i for i in range(self.n_rows) if i not in byz])
            other = s.row_ptrs([i for i in range(self.n_rows) if i in byz and i not in rep])
            k_select.launch_select(stat, other, v.count, v.kind, v.param, mode, b, c0, c1,
                                   self._epilogue(True), self.device)
        else:
            k_select.launch_select(s.row_ptrs(range(self.n_rows)), [], 0, None, 0.0, mode, b, c0, c1,
                                   self._epilogue(True), self.device)
        return self._finish()

    def trimmed_mean(self, b: int) -> torch.Tensor:
        if self.n_rows - 2 * b <= 0:
            raise ValueError(f"trim {b} too large for {self.n_rows} rows")
        return self._select(0, b)

    def median(self) -> torch.Tensor:
        return self._select(1, 0)

    def combine(self, weights, extra: Optional[torch.Tensor] = None, extra_weight: float = 0.0) -> torch.Tensor:
        self.materialize_virtual()
        self._pre()
        s = self.symm
        w = [float(x) for x in (weights.tolist() if hasattr(weights, "tolist") else weights)]
        rows = s.row_ptrs(range(self.n_rows))
        if extra is not None and extra_weight != 0.0:
            rows = rows + [extra.contiguous().data_ptr()]
            w = w + [float(extra_weight)]
        c0, c1 = s.my_cols
        k_combine.launch_combine(rows, w, c0, c1, self._epilogue(True), self.device)
        return self._finish()

    def gram(self, extra: Optional[torch.Tensor] = None) -> np.ndarray:
        from ..ops import _gram_impl, gram as k_gram
        self.materialize_virtual()
        self._pre()
        s = self.symm
        blocks = []
        for (ptr, ld, rows) in s.block_descs():
            blocks += _gram_impl.split_blocks(ptr, ld, rows)
        e = 0
        if extra is not None:
            extra = extra.reshape(-1, self.n_cols)
            e = extra.shape[0]
            buf = torch.zeros(e, s.ld, device=self.device, dtype=torch.float32)
            buf[:, : self.n_cols] = extra
            blocks += _gram_impl.split_blocks(buf.data_ptr(), s.ld, e)
        total_pad = sum((b[2] + 7) // 8 * 8 for b in blocks)
        assert total_pad <= 512 and len(blocks) <= 9, "too many rows/blocks for one Gram pass"
        tile_rows = (total_pad + 127) // 128 * 128
        ldg = (total_pad + 31) // 32 * 32
        out = torch.zeros(tile_rows, ldg, device=self.device, dtype=torch.float32)
        c0, c1 = s.my_cols
        starts = _gram_impl.launch_gram(blocks, self.n_cols, c0, c1, out, k_gram.PRECISION == "tf32x3", self.device)
        dist.all_reduce(out)             # N x N partials (<= 1 MB): NCCL is plumbing here
        idx = []
        for (_, _, rows), st in zip(blocks, starts):
            idx += list(range(st, st + rows))
        idx_t = torch.tensor(idx, device=self.device)
        G = out[idx_t][:, idx_t]
        return (0.5 * (G + G.T)).double().cpu().numpy()

    def rows(self) -> torch.Tensor:
        self.materialize_virtual()
        self._pre()
        s = self.symm
        pad = s.local_full.contiguous()
        outs = [torch.empty_like(pad) for _ in range(s.world.size)]
        dist.all_gather(outs, pad)
        return torch.cat([o[:k, : self.n_cols] for o, k in zip(outs, s.shard_sizes)], 0)
