import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blades_b200 import Simulator
from blades_b200.datasets import synthetic_fldataset
from blades_b200.models import resnet18
n = 16
ds = synthetic_fldataset(n, shape=(3, 32, 32), num_classes=10, train_bs=32, train_per_client=64, test_per_client=32)
sim = Simulator(ds, num_byzantine=3, attack="ipm", aggregator="median", use_cuda=True, seed=1, log_path=tempfile.mkdtemp(), progress=False)
model = resnet18(10)
sim.prepare(model, "SGD", "SGD", "crossentropy", 1.0, 0.1)
eng = sim.engine
for r in range(3):
    sim.train_actor(r, 5, sim.get_clients(), 0.1)
torch.cuda.synchronize()
t0 = time.perf_counter()
for r in range(3):
    sim.train_actor(r, 5, sim.get_clients(), 0.1)
torch.cuda.synchronize()
print("ms per client visit (5 steps):", (time.perf_counter() - t0) / 3 / n * 1e3)
t0 = time.perf_counter()
eng.train_local(5, 0.1)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("train_local: host %.1f ms, +sync %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    eng.train_local(5, 0.1)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
