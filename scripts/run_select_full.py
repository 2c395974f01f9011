"""Headline-size launches of the trimmed-mean kernels (80 real rows + 20 ALIE virtual rows, d = 11.18 M) for a full-size
ncu capture: LDG partition form, then the bulk-copy staged form (BLADES_SELECT_STAGED is read once per process, so the
second form runs in a child process)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blades_b200.ops import select, gram
from blades_b200.parallel.matrix import VirtualRows
d = 11181642
ld = (d + 63) // 64 * 64
U = (torch.randn(100, ld, device="cuda") * 0.01)[:, :d]
v = VirtualRows("alie", 0.2858, list(range(20)))
for _ in range(2):
    select.trimmed_mean(U, 20, virtual=v)
if os.environ.get("BLADES_SELECT_STAGED") != "1":
    gram.gram(U, precision="tf32")
    gram.gram(U, precision="tf32x3")
torch.cuda.synchronize()
if os.environ.get("BLADES_SELECT_STAGED") != "1" and "--both" in sys.argv:
    del U
    torch.cuda.empty_cache()
    subprocess.run([sys.executable, __file__], env=dict(os.environ, BLADES_SELECT_STAGED="1"), check=True)
