"""NVLink-addressable symmetric memory for the update rows, the aggregate and the parameter
replica of every trainer shard.

One symmetric allocation per rank (``torch.distributed._symmetric_memory``: CUDA VMM + IPC handle
exchange; NVLS multicast object when the fabric supports it), laid out identically everywhere:

    [ U_g : nmax rows x ld floats ][ agg : ld ][ theta : ld ][ scratch : 512 x 512 ][ recv : n_total x recv_ld ]

``recv`` is the landing zone of the push-mode aggregation: row i holds client i's values of the coordinates THIS rank
aggregates (its shard of every window), written by the owner of row i with copy-engine DMAs over NVLink.

(``scratch`` holds the N x N Gram partials that are summed inside the NVSwitch, ``reduce_scratch``)

so every kernel can address any peer's rows/agg/theta as ``peer_base[r] + same offset``.  ``ld`` is
``d`` rounded up to 64 floats: rows are 256 B aligned for 16 B vector loads and TMA.

Synchronisation: ``barrier()`` is the symmetric-memory device barrier (signal-pad flags over NVLink,
enqueued on the current stream -- no host round trip).  Protocol per round:
    train (writes own rows) -> barrier -> fused pull-aggregate kernels (read all rows, write
    agg/theta shards to all peers) -> barrier -> next round.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist

__all__ = ["SymmetricUpdates", "round_up", "coordinate_shards"]

SCRATCH_FLOATS = 512 * 512          # one padded N x N Gram (N <= 512)


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def coordinate_shards(d: int, world_size: int, align: int = 128):
    """Coordinate ranges ``[(c_0, c_1), ...]`` owned by each rank for coordinate-sharded aggregation: contiguous,
    disjoint, covering ``[0, d)``, interior boundaries aligned to ``align`` floats (512 B: whole 128 B row segments
    per warp and TMA-box friendly).  Ranks at the end may own an empty range when ``d`` is tiny."""
    cuts = [min(d, round_up(d * r // world_size, align)) for r in range(world_size)] + [d]
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


class SymmetricUpdates:
    def __init__(self, world, shard_sizes: Sequence[int], d: int):
        import torch.distributed._symmetric_memory as symm_mem
        self.world = world
        self.shard_sizes = list(shard_sizes)
        self.d = d
        self.ld = round_up(d, 64)
        self.nmax = max(self.shard_sizes)
        self.n_local = self.shard_sizes[world.rank]
        self.n_total = sum(self.shard_sizes)
        self.scratch_floats = SCRATCH_FLOATS
        # push-mode landing zone: every rank's shard of every aggregation window (<= 16 windows, 128-float alignment)
        self.recv_ld = round_up((d + world.size - 1) // world.size + 128 * 17, 64) if world.size > 1 else 0
        total = self.nmax * self.ld + 2 * self.ld + self.scratch_floats + self.n_total * self.recv_ld
        self.buf = symm_mem.empty((total,), dtype=torch.float32, device=world.device)
        from ..ops import nvls
        nvls.zero_(self.buf)
        self.handle = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        self.barrier_timeout_ms = max(0, int(os.environ.get("BLADES_BARRIER_TIMEOUT_MS", "120000")))
        self.base_ptrs: List[int] = [int(p) for p in self.handle.buffer_ptrs]
        self.off_agg = self.nmax * self.ld
        self.off_theta = self.off_agg + self.ld
        self.local_full = self.buf[: self.nmax * self.ld].view(self.nmax, self.ld)
        self.local = self.local_full[: self.n_local, :d]
        self.agg = self.buf[self.off_agg: self.off_agg + d]
        self.theta = self.buf[self.off_theta: self.off_theta + d]
        self.off_scratch = self.off_theta + self.ld
        self.scratch = self.buf[self.off_scratch: self.off_scratch + self.scratch_floats]
        self.off_recv = self.off_scratch + self.scratch_floats
        self.row0 = [sum(self.shard_sizes[:r]) for r in range(world.size)]       # first global row of every rank
        self.multicast_ptr = int(self.handle.multicast_ptr) if getattr(self.handle, "multicast_ptr", 0) else 0
        if os.environ.get("BLADES_MULTIMEM", "1") == "0":       # A/B switch: per-peer stores / P2P reduction
            self.multicast_ptr = 0
        # global row -> (rank, local row)
        self.row_owner = []
        for r, k in enumerate(self.shard_sizes):
            self.row_owner += [(r, i) for i in range(k)]
        self.col_ranges = coordinate_shards(d, world.size)

    # ------------------------------------------------------------------ addressing
    def row_ptr(self, global_row: int) -> int:
        r, i = self.row_owner[global_row]
        return self.base_ptrs[r] + i * self.ld * 4

    def row_ptrs(self, rows: Sequence[int]) -> List[int]:
        return [self.row_ptr(g) for g in rows]

    def agg_ptrs(self) -> List[int]:
        return [b + self.off_agg * 4 for b in self.base_ptrs]

    def theta_ptrs(self) -> List[int]:
        return [b + self.off_theta * 4 for b in self.base_ptrs]

    def mc_agg_ptr(self) -> int:
        """NVLS multicast address of ``agg`` (0 when the fabric has no multicast object)."""
        return self.multicast_ptr + self.off_agg * 4 if self.multicast_ptr else 0

    def mc_theta_ptr(self) -> int:
        return self.multicast_ptr + self.off_theta * 4 if self.multicast_ptr else 0

    def reduce_scratch(self, count: int) -> None:
        """Sum the first ``count`` floats of ``scratch`` over all ranks, result in every replica: one two-shot
        kernel per rank (``multimem.ld_reduce`` + ``multimem.st`` through the switch, or peer loads/stores), bracketed
        by the device barrier.  Replaces the NCCL all-reduce of the Gram partials."""
        from ..ops import nvls
        count = round_up(count, 4)
        assert count <= self.scratch_floats
        self.barrier()
        nvls.allreduce([b + self.off_scratch * 4 for b in self.base_ptrs],
                       self.multicast_ptr + self.off_scratch * 4 if self.multicast_ptr else 0,
                       count, self.world.rank, self.world.size, self.world.device)
        self.barrier()

    def push_streams(self):
        """Side streams for the push-mode DMAs (one per destination rank, ``BLADES_PUSH_STREAMS`` caps the number)."""
        if getattr(self, "_push_streams", None) is None:
            k = max(1, min(self.world.size, int(os.environ.get("BLADES_PUSH_STREAMS", "8"))))
            self._push_streams = [torch.cuda.Stream(device=self.world.device) for _ in range(k)]
        return self._push_streams

    def window_span(self, lo: int, hi: int) -> int:
        """Columns a window occupies in every rank's ``recv`` rows (identical on all ranks): the longest shard."""
        return max(c1 - c0 for c0, c1 in coordinate_shards(hi - lo, self.world.size))

    def recv_row_ptr(self, rank: int, global_row: int, col: int = 0) -> int:
        """Address (in this process: local or peer mapping) of ``recv[global_row][col]`` on ``rank``."""
        return self.base_ptrs[rank] + (self.off_recv + global_row * self.recv_ld + col) * 4

    def block_descs(self):
        """(base_ptr, ld, rows) of every rank's row block, in global row order."""
        return [(self.base_ptrs[r], self.ld, k) for r, k in enumerate(self.shard_sizes) if k > 0]

    @property
    def my_cols(self):
        return self.col_ranges[self.world.rank]

    def barrier(self, channel: int = 0) -> None:
        """Device barrier over the signal pads (stream ordered, capturable).  Failure detection: a rank that died or
        never arrives would leave the others spinning forever inside the barrier kernel; with a timeout
        (``BLADES_BARRIER_TIMEOUT_MS``, default 120 000, 0 = wait forever) the waiting kernels trap instead, the CUDA
        error surfaces at the next synchronisation and the job aborts instead of hanging the node."""
        self.handle.barrier(channel=channel, timeout_ms=self.barrier_timeout_ms)
