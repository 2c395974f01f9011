"""Per-phase CUDA-event timing of the headline round (train / aggregate / apply) + kernel micro-benchmarks."""
import os, sys, tempfile, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blades_b200 import Simulator
from blades_b200.datasets import synthetic_fldataset
from blades_b200.models import resnet18
from blades_b200.ops import select, combine

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts)//2]

n, f = 100, 20
ds = synthetic_fldataset(n, shape=(3, 32, 32), num_classes=10, train_bs=32, train_per_client=64, test_per_client=32)
sim = Simulator(ds, num_byzantine=f, attack="alie", attack_kws={"num_clients": n, "num_byzantine": f},
                aggregator="trimmedmean", aggregator_kws={"nb": f}, use_cuda=True, seed=1,
                log_path=tempfile.mkdtemp(), progress=False, profile=True)
model = resnet18(10)
sim.prepare(model, "SGD", "SGD", "crossentropy", 1.0, 0.1)
eng = sim.engine
eng.prestaged = eng.stage_batches(None, 1)
for r in range(6):
    sim.train_actor(r, 1, sim.get_clients(), 0.1)
print("phase ms (last 3 rounds):", json.dumps(eng.timer.records[-3:]))
U = eng.U
d = U.shape[1]
print("U", tuple(U.shape), "GB", U.numel() * 4 / 1e9)
from blades_b200.parallel.matrix import VirtualRows
v = VirtualRows("alie", 0.2858, list(range(f)))
for name, fn, nbytes in [
    ("trimmed_mean N=100 b=20", lambda: select.trimmed_mean(U, 20), n * d * 4),
    ("trimmed_mean+ALIE virtual f=20", lambda: select.trimmed_mean(U, 20, virtual=v), (n - f) * d * 4),
    ("median N=100", lambda: select.median(U), n * d * 4),
    ("row_combine N=100", lambda: combine.row_combine(U, [0.01] * n), n * d * 4),
    ("torch topk-trimmed-mean (reference formulation)", lambda: torch.cat([U, -torch.topk(U, 20, 0).values, torch.topk(-U, 20, 0).values]).sum(0), n * d * 4),
    ("torch mean(0)", lambda: U.mean(0), n * d * 4),
]:
    best, med = timeit(fn, iters=3 if "torch topk" in name else 6)
    print(f"{name:50s} best {best:8.3f} ms  median {med:8.3f} ms  -> {nbytes / best / 1e6:8.1f} GB/s")
# torch profile of one training phase
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    eng.train_local(1, 0.1)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
