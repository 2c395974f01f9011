"""Log every aten op that touches a layer1-sized activation ([*,64,8,8]) with its input/output strides, in order:
finds where NHWC tensors silently become NCHW (and pay a layout copy)."""
import os, sys, tempfile
os.environ["BLADES_GRAPH"] = "0"
os.environ["BLADES_ROUND_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.utils._python_dispatch import TorchDispatchMode
from blades_b200 import Simulator
from blades_b200.datasets import synthetic_fldataset
from blades_b200.models import resnet18

n = int(os.environ.get("N", "4"))
ds = synthetic_fldataset(n, shape=(3, 32, 32), num_classes=10, train_bs=32, train_per_client=64, test_per_client=32, seed=1)
sim = Simulator(ds, num_byzantine=0, aggregator="mean", use_cuda=torch.cuda.is_available(), seed=1,
                log_path=tempfile.mkdtemp(), progress=False)
sim.prepare(resnet18(10), "SGD", "SGD", "crossentropy", 1.0, 0.1)
clients = sim.get_clients()
sim.train_actor(0, 1, clients, 0.1)


def fmt(t):
    if isinstance(t, torch.Tensor) and t.dim() == 4:
        lay = "NHWC" if (t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous()) else \
              ("NCHW" if t.is_contiguous() else "strided")
        return f"{tuple(t.shape)}:{lay}"
    return None


class Log(TorchDispatchMode):
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        flat = [a for a in args if isinstance(a, torch.Tensor)] + \
               [b for a in args if isinstance(a, (list, tuple)) for b in a if isinstance(b, torch.Tensor)]
        outs = [o for o in (out if isinstance(out, (list, tuple)) else [out]) if isinstance(o, torch.Tensor)]
        want = [t for t in flat + outs if t.dim() == 4 and t.shape[1] == 64 and t.shape[2] == 8]
        if want:
            print(f"{str(func):45s} in: {[fmt(a) for a in flat if fmt(a)]}  out: {[fmt(o) for o in outs if fmt(o)]}")
        return out


with Log():
    sim.train_actor(1, 1, clients, 0.1)
