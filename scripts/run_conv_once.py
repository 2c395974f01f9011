"""Launch the full-size ResNet-18 layer1 / layer2 convolutions once each (for an ncu capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blades_b200.ops import conv as kc
dev = torch.device("cuda")
torch.manual_seed(0)
for (NB, H, Cin, Cout, k, s, p) in [(3200, 8, 64, 64, 3, 1, 1), (3200, 4, 128, 128, 3, 1, 1)]:
    x = torch.randn(NB, Cin, H, H, device=dev).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5).contiguous(memory_format=torch.channels_last)
    w2d = w.permute(0, 2, 3, 1).reshape(Cout, -1)
    for _ in range(3):
        y = kc.conv_fprop(x, w2d, (k, k), s, p)
    torch.cuda.synchronize()
print("done")
