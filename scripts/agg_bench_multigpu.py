"""Sharded aggregation alone, under torchrun: device barrier, pull-mode select (peer loads over NVLink), push-mode
(copy-engine DMA + local select), with the achieved fraction of the NVLink roofline (measured peer-copy bandwidth
770 GB/s per direction per GPU, B200_PROFILING.md) -- the headline shape: 100 clients (20 ALIE virtual rows), ResNet-18.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/agg_bench_multigpu.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from blades_b200.comm.group import init_world, shutdown
from blades_b200.comm.symm import SymmetricUpdates
from blades_b200.parallel.matrix import VirtualRows
from blades_b200.parallel.sharded import ShardedMatrix

world = init_world(use_cuda=True)
n, f, d = 100, 20, 11181642
sizes = [len(a) for a in np.array_split(np.arange(n), world.size)]
symm = SymmetricUpdates(world, sizes, d)
symm.local_full.normal_(0.0, 0.01)
torch.cuda.synchronize()
world.barrier()
virt = VirtualRows("alie", 0.2858, list(range(f)))


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    world.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return world.all_reduce_max(a.elapsed_time(b) / iters)


def select(push):
    os.environ["BLADES_AGG_PUSH"] = "1" if push else "0"
    m = ShardedMatrix(symm, virtual=virt)
    m.trimmed_mean(f)


def push_only():
    m = ShardedMatrix(symm, virtual=virt)
    m.push = True
    m._push_rows()


res = {"gpus": world.size, "clients": n, "virtual_rows": f, "d": d}
res["barrier_us"] = timed(symm.barrier, iters=50) * 1e3
res["select_pull_ms"] = timed(lambda: select(False))
res["select_push_ms"] = timed(lambda: select(True))
res["push_dma_only_ms"] = timed(push_only)
G = world.size
peer_bytes = (n - f) * (G - 1) / G * (d / G) * 4              # honest rows owned elsewhere x this rank's coordinates
res["nvlink_bytes_per_gpu"] = peer_bytes
res["nvlink_floor_ms_at_770GBps"] = peer_bytes / 770e9 * 1e3
for k in ("select_pull_ms", "select_push_ms", "push_dma_only_ms"):
    res[k.replace("_ms", "_frac_of_nvlink_roofline")] = res["nvlink_floor_ms_at_770GBps"] / max(res[k] - (2e-3 * res["barrier_us"] if "select" in k else 0.0), 1e-9)
if world.rank == 0:
    print(json.dumps(res))
shutdown()
