"""Kernel names launched by ``__graft_entry__.smoke()`` (torch profiler / CUPTI), own vs foreign."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import __graft_entry__ as g
g.smoke()                                   # warm-up: lazy inits, graph capture policy
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.smoke()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
own = ("tcgen05", "client_", "coord_select", "row_combine", "im2col", "maxpool", "avgpool", "pad_rows", "gather_samples",
       "fill_normal", "attack_rows", "select", "gram")
for e in sorted(ev, key=lambda e: -e.device_time_total):
    tag = "own    " if any(k in e.key for k in own) else "FOREIGN"
    print(f"{tag} {e.count:4d}x {e.device_time_total:9.1f} us  {e.key[:110]}")
