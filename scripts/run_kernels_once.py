"""One launch of each hot kernel at (reduced) headline shapes, for `ncu --set full` captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blades_b200.ops import select, gram, wgrad, combine
from blades_b200.parallel.matrix import VirtualRows
d = 11181642 // 4
ld = (d + 63) // 64 * 64
U = (torch.randn(100, ld, device="cuda") * 0.01)[:, :d]
v = VirtualRows("alie", 0.2858, list(range(20)))
select.trimmed_mean(U, 20, virtual=v)
combine.row_combine(U, [0.01] * 100)
gram.gram(U, precision="tf32")
gram.gram(U, precision="tf32x3")
a_t = torch.randn(100, 512, 128, device="cuda"); b = torch.randn(100, 512, 1152, device="cuda"); out = torch.empty(100, 128, 1152, device="cuda")
wgrad.grouped_wgrad(a_t.transpose(1, 2), b, out, -0.1)
a_t = torch.randn(100, 32, 512, device="cuda"); b = torch.randn(100, 32, 4608, device="cuda"); out = torch.empty(100, 512, 4608, device="cuda")
wgrad.grouped_wgrad(a_t.transpose(1, 2), b, out, -0.1)
torch.cuda.synchronize()
