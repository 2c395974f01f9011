"""Per-phase CUDA-event times of the headline round under torchrun (train / aggregate incl. barriers / apply), max over
ranks -- the missing measurement behind the NVLink roofline fraction of the sharded aggregation.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/phase_times_multigpu.py
"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blades_b200 import Simulator
from blades_b200.comm.group import init_world, shutdown
from blades_b200.datasets import synthetic_fldataset
from blades_b200.models import resnet18

world = init_world(use_cuda=True)
n, f = 100, 20
ds = synthetic_fldataset(n, shape=(3, 32, 32), num_classes=10, train_bs=32, train_per_client=64, test_per_client=32)
sim = Simulator(ds, num_byzantine=f, attack="alie", attack_kws={"num_clients": n, "num_byzantine": f},
                aggregator="trimmedmean", aggregator_kws={"nb": f}, use_cuda=True, seed=1,
                log_path=tempfile.mkdtemp(), progress=False, profile=True)          # profile=True: eager phases
sim.prepare(resnet18(10), "SGD", "SGD", "crossentropy", 1.0, 0.1)
eng = sim.engine
eng.prestaged = eng.stage_batches(None, 1)
for r in range(8):
    sim.train_actor(r, 1, sim.get_clients(), 0.1)
torch.cuda.synchronize()
recs = eng.timer.records[-4:]
keys = sorted({k for r in recs for k in r})
mine = {k: sum(r.get(k, 0.0) for r in recs) / len(recs) for k in keys}
allr = world.all_gather_object(mine)
if world.rank == 0:
    worst = {k: max(a.get(k, 0.0) for a in allr) for k in keys}
    d = eng.d
    ingress = (world.size - 1) / world.size * (n - f) * (d / world.size) * 4      # bytes of peer rows read per GPU
    print(json.dumps({"gpus": world.size, "phase_ms_max_over_ranks": worst,
                      "peer_bytes_per_gpu": ingress, "nvlink_roofline_ms_at_900GBps": ingress / 900e9 * 1e3}))
shutdown()
