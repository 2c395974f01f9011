"""Where do the layout/copy kernels of a round come from?  (aten::copy_ / clone / contiguous with shapes + python stack)"""
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
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
    sim.train_actor(5, 1, clients, 0.1)
    torch.cuda.synchronize()
rows = []
for e in prof.events():
    if e.name in ("aten::copy_", "aten::clone", "aten::contiguous", "aten::add_", "aten::add", "aten::clamp_min",
                  "aten::clamp_min_", "aten::relu", "aten::relu_", "aten::threshold_backward", "aten::mul", "aten::fill_",
                  "aten::zero_") and e.device_time_total > 5:
        stack = [s for s in (e.stack or []) if "blades_b200" in s or "scripts/" in s][:3]
        rows.append((e.device_time_total, e.name, str(e.input_shapes)[:90], " <- ".join(s.split("/")[-1][:60] for s in stack)))
rows.sort(reverse=True)
tot = {}
for t, nme, shp, st in rows:
    tot[(nme, st)] = tot.get((nme, st), 0) + t
print("--- by (op, stack) ---")
for (nme, st), t in sorted(tot.items(), key=lambda kv: -kv[1])[:25]:
    print(f"{t:9.1f} us  {nme:28s} {st}")
print("--- top individual ---")
for t, nme, shp, st in rows[:25]:
    print(f"{t:9.1f} us  {nme:24s} {shp:90s} {st}")
