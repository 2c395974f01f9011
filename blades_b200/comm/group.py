"""Process-group plumbing: one process per GPU, ``torch.distributed`` (NCCL on GPUs,
gloo on CPU) for bootstrap, barriers, small host-side exchanges and the *baseline*
collectives.  The product path does its bulk data movement inside our own kernels
over NVLink peer pointers (``comm.symm``); NCCL never carries update matrices there.

Replaces the reference's Ray runtime (actor pool + plasma object store + gRPC,
SURVEY 2.9 X1-X6): there is no driver process; every rank runs the same
``Simulator`` program (SPMD) and owns the clients ``np.array_split`` assigns to it.
"""
from __future__ import annotations

import datetime
import os
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["World", "init_world", "get_world", "split_clients", "shutdown"]


@dataclass
class World:
    rank: int = 0
    size: int = 1
    local_rank: int = 0
    device: torch.device = torch.device("cpu")
    backend: Optional[str] = None
    group: Optional[object] = None

    @property
    def distributed(self) -> bool:
        return self.size > 1

    def barrier(self) -> None:
        if self.distributed:
            if self.backend == "nccl":
                dist.barrier(device_ids=[self.device.index])
            else:
                dist.barrier()

    def all_reduce_max(self, value: float) -> float:
        if not self.distributed:
            return value
        t = torch.tensor([value], dtype=torch.float64, device=self.device if self.backend == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_reduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if self.distributed:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def broadcast_object(self, obj, src: int = 0):
        if not self.distributed:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=src)
        return box[0]

    def all_gather_object(self, obj) -> list:
        if not self.distributed:
            return [obj]
        out = [None] * self.size
        dist.all_gather_object(out, obj)
        return out


_WORLD: Optional[World] = None


def init_world(use_cuda: Optional[bool] = None, timeout_s: int = 600) -> World:
    """Initialise from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).
    Without those variables this is a single-process world."""
    global _WORLD
    if _WORLD is not None:
        return _WORLD
    size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if use_cuda is None:
        use_cuda = torch.cuda.is_available()
    device = torch.device("cpu")
    if use_cuda:
        device = torch.device("cuda", local % max(torch.cuda.device_count(), 1))
        torch.cuda.set_device(device)
    backend = None
    if size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        backend = "nccl" if use_cuda else "gloo"
        if not dist.is_initialized():
            kw = {}
            if backend == "nccl":
                kw["device_id"] = device
            dist.init_process_group(backend=backend, rank=rank, world_size=size,
                                    timeout=datetime.timedelta(seconds=timeout_s), **kw)
    _WORLD = World(rank, size, local, device, backend, dist.group.WORLD if size > 1 else None)
    return _WORLD


def get_world() -> World:
    return _WORLD if _WORLD is not None else init_world()


def shutdown() -> None:
    global _WORLD
    if dist.is_available() and dist.is_initialized():
        try:
            dist.destroy_process_group()
        except Exception:
            pass
    _WORLD = None


def split_clients(n_clients: int, n_shards: int) -> List[np.ndarray]:
    """Contiguous client->shard assignment, identical to the reference's
    ``np.array_split(clients, n_actors)`` (simulator.py:223):
    100 clients / 8 shards -> [13,13,13,13,12,12,12,12]."""
    return np.array_split(np.arange(n_clients), n_shards)
