"""GPU time of one headline round by kernel (torch profiler / CUPTI, eager launches so every kernel is attributed)."""
import os, sys, tempfile
os.environ.setdefault("BLADES_GRAPH", "0")
os.environ.setdefault("BLADES_ROUND_GRAPH", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from blades_b200 import Simulator
from blades_b200.datasets import synthetic_fldataset
from blades_b200.models import resnet18

n = 100
ds = synthetic_fldataset(n, shape=(3, 32, 32), num_classes=10, train_bs=32, train_per_client=64, test_per_client=32, seed=1)
sim = Simulator(ds, num_byzantine=20, attack="alie", attack_kws={"num_clients": n, "num_byzantine": 20},
                aggregator="trimmedmean", aggregator_kws={"nb": 20}, use_cuda=True, seed=1,
                log_path=tempfile.mkdtemp(), progress=False)
sim.prepare(resnet18(10), "SGD", "SGD", "crossentropy", 1.0, 0.1)
clients = sim.get_clients()
for r in range(4):
    sim.train_actor(r, 1, clients, 0.1)
torch.cuda.synchronize()
R = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for r in range(R):
        sim.train_actor(r, 1, clients, 0.1)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
ev.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in ev)
print(f"GPU kernel time per round: {tot / R / 1e3:.3f} ms over {sum(e.count for e in ev) // R} launches")
for e in ev[:40]:
    print(f"{e.device_time_total / R / 1e3:8.3f} ms  {e.count // R:5d}x  {e.key[:130]}")
# per-launch durations of our kernels in launch order (last profiled round)
mine = [e for e in prof.events() if e.device_time_total > 0 and any(k in e.name for k in ("tcgen05", "client_", "im2col", "coord_select", "pool", "pad_rows"))]
mine.sort(key=lambda e: e.time_range.start)
per = len(mine) // R
print("--- per-launch (one round, launch order = backward order of the layers) ---")
for e in mine[-per:]:
    print(f"{e.device_time_total:9.1f} us  {e.name[:60]}")
