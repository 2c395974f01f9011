import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blades_b200.ops import select
from blades_b200.parallel.matrix import VirtualRows
U = torch.randn(100, 11181642 // 4, device="cuda")     # quarter-size: ncu replays ~40x
v = VirtualRows("alie", 0.2858, list(range(20)))
for _ in range(2):
    select.trimmed_mean(U, 20, virtual=v)
torch.cuda.synchronize()
