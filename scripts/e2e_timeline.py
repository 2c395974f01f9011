"""Host-side timeline of the end-to-end round loop (per rank): where does the wall time of one round go?
Run: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/e2e_timeline.py"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blades_b200 import Simulator
from blades_b200.comm.group import init_world
from blades_b200.datasets import synthetic_fldataset
from blades_b200.models import resnet18

world = init_world(use_cuda=True)
ds = synthetic_fldataset(100, shape=(3, 32, 32), num_classes=10, train_bs=32, train_per_client=64, test_per_client=32, seed=1)
sim = Simulator(ds, num_byzantine=20, attack="alie", attack_kws={"num_clients": 100, "num_byzantine": 20},
                aggregator="trimmedmean", aggregator_kws={"nb": 20}, use_cuda=True, seed=1,
                log_path=tempfile.mkdtemp(), progress=False, wipe_logs=True)
sim.prepare(resnet18(10), "SGD", "SGD", "crossentropy", server_lr=1.0, client_lr=0.1)
eng = sim.engine
clients = sim.get_clients()
T = {}


def wrap(obj, name):
    fn = getattr(obj, name)

    def w(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        T.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
        return r
    setattr(obj, name, w)


wrap(eng, "stage_batches")
wrap(eng, "flush_prefetch")
wrap(eng, "static_round")
for r in range(6):
    sim.train_actor(r, 1, clients, 0.1)
    eng.last_client_losses.cpu()
torch.cuda.synchronize()
world.barrier()
T.clear()
rounds, cpu_ms, tot = 20, [], []
for r in range(rounds):
    t0 = time.perf_counter()
    sim.train_actor(r, 1, clients, 0.1)
    t1 = time.perf_counter()
    eng.last_client_losses.cpu()
    t2 = time.perf_counter()
    cpu_ms.append((t2 - t1) * 1e3)
    tot.append((t2 - t0) * 1e3)
torch.cuda.synchronize()
fmt = lambda v: " ".join(f"{x:.2f}" for x in v[:12])
for rk in range(world.size):
    world.barrier()
    if rk == world.rank:
        print(f"[rank {rk}] total/round ms: {fmt(tot)}")
        print(f"[rank {rk}]   .cpu() wait   : {fmt(cpu_ms)}")
        for k, v in T.items():
            print(f"[rank {rk}]   {k:14s}: {fmt(v)}")
        sys.stdout.flush()
