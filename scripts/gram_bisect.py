import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blades_b200.ops import gram
d = 11181642; ld = (d + 63) // 64 * 64
U = (torch.randn(100, ld, device="cuda") * 0.01)[:, :d]
def t(fn, it=4):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
print(os.environ.get("BLADES_GRAM_DBG"), os.environ.get("BLADES_GRAM_SLABS"), "tf32 %.3f ms  tf32x3 %.3f ms" % (t(lambda: gram.gram(U, precision="tf32")), t(lambda: gram.gram(U, precision="tf32x3"))))
