"""Timing bisect of the conv kernel on the full-size layer1 / layer2 shapes (BLADES_CONV_DBG set by the caller)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blades_b200.ops import conv as kc
dev = torch.device("cuda")
torch.manual_seed(0)
out = []
for (NB, H, Cin, Cout, k, s, p) in [(3200, 8, 64, 64, 3, 1, 1), (3200, 4, 128, 128, 3, 1, 1), (3200, 2, 256, 256, 3, 1, 1)]:
    x = torch.randn(NB, Cin, H, H, device=dev).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5).contiguous(memory_format=torch.channels_last)
    w2d = w.permute(0, 2, 3, 1).reshape(Cout, -1)
    for _ in range(3):
        kc.conv_fprop(x, w2d, (k, k), s, p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        kc.conv_fprop(x, w2d, (k, k), s, p)
    e1.record()
    torch.cuda.synchronize()
    out.append(f"{e0.elapsed_time(e1) / 20 * 1e3:7.1f}")
print("dbg", os.environ.get("BLADES_CONV_DBG", "0"), " l1/l2/l3 us:", " ".join(out))
