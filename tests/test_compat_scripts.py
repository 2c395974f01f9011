"""The reference's own example scripts run UNCHANGED on this package through ``blades_b200.compat`` (module alias
``blades`` -> ``blades_b200`` + a no-op ``ray``): only the hard-coded round counts are shortened, and torchvision's
MNIST download is replaced by a deterministic fake.  Skipped when the reference's files are not available."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = next((d for d in (os.path.join(ROOT, "baseline", "_ref", "blades", "examples"),
                             "/root/reference/src/blades/examples") if os.path.isdir(d)), None)

pytestmark = pytest.mark.skipif(EXAMPLES is None, reason="reference examples not available")

_RUNNER = r'''
import re, sys, torch, torchvision
sys.path.insert(0, {root!r})

class FakeMNIST:
    def __init__(self, train=True, download=True, root=None, **kw):
        g = torch.Generator().manual_seed(11 if train else 12)
        n = 1200 if train else 200
        self.data = torch.randint(0, 256, (n, 28, 28), generator=g, dtype=torch.uint8)
        self.targets = torch.randint(0, 10, (n,), generator=g)
torchvision.datasets.MNIST = FakeMNIST

import blades_b200.compat as compat
def shorten(src):
    src = re.sub(r'"global_rounds":\s*\d+', '"global_rounds": 3', src)
    return re.sub(r'"local_steps":\s*\d+', '"local_steps": 2', src)
g = compat.run_script({script!r}, patch=shorten)
sim = g["simulator"]
import blades, blades.simulator, blades_b200.simulator
assert blades.simulator.Simulator is blades_b200.simulator.Simulator and type(sim) is blades_b200.simulator.Simulator
print("RAN", type(sim).__module__, len(sim.get_clients()), sum(c.is_byzantine() for c in sim.get_clients()))
'''


@pytest.mark.parametrize("script,clients,byz", [("mini_example.py", 10, 4), ("customize_attack.py", 10, 5)])
def test_reference_example_runs_unchanged(tmp_path, script, clients, byz):
    code = _RUNNER.format(root=ROOT, script=os.path.join(EXAMPLES, script))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=tmp_path)
    assert out.returncode == 0, out.stderr[-3000:]
    m = re.search(r"RAN (\S+) (\d+) (\d+)", out.stdout)
    assert m and m.group(1) == "blades_b200.simulator" and int(m.group(2)) == clients and int(m.group(3)) == byz


def test_alias_modules_are_the_same_objects():
    code = ("import sys; sys.path.insert(0, %r); import blades_b200.compat as c; c.install();"
            "import blades.aggregators.krum as a, blades_b200.aggregators.krum as b; assert a is b;"
            "from blades.attackers.alieclient import AlieClient; from blades.models.cifar10 import CCTNet;"
            "from blades.datasets import MNIST, CIFAR10; from blades.client import BladesClient, ByzantineClient;"
            "from blades.server import BladesServer; import ray; ray.init(num_gpus=0); print('ok')" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


_COMPARE = r'''
import re, sys, torch, torchvision
sys.path.insert(0, {root!r})

class FakeMNIST:
    def __init__(self, train=True, download=True, root=None, **kw):
        g = torch.Generator().manual_seed(11 if train else 12)
        n = 1200 if train else 200
        self.data = torch.randint(0, 256, (n, 28, 28), generator=g, dtype=torch.uint8)
        self.targets = torch.randint(0, 10, (n,), generator=g)
torchvision.datasets.MNIST = FakeMNIST

def patch(src):
    src = re.sub(r'"global_rounds":\s*\d+', '"global_rounds": 4', src)
    src = re.sub(r'"local_steps":\s*\d+', '"local_steps": 3', src)
    return re.sub(r'"num_actors":\s*\d+', '"num_actors": 1', src)      # one actor: deterministic call order

if {which!r} == "reference":
    import sklearn.cluster as skc                 # the reference passes sklearn's removed ``affinity=`` kwarg (Q7)
    _orig = skc.AgglomerativeClustering
    def _compat(*a, affinity=None, **k):
        if affinity is not None:
            k["metric"] = affinity
        return _orig(*a, **k)
    skc.AgglomerativeClustering = _compat
    from baseline import ref_arm
    ref_arm.import_reference(0)
    import os
    os.makedirs("data", exist_ok=True)                                # torchvision's download would create it
    glob = {{"__name__": "__main__", "__file__": {script!r}}}
    exec(compile(patch(open({script!r}).read()), {script!r}, "exec"), glob)
else:
    import blades_b200.compat as compat
    # same batch order as the reference for the comparison (the default streams use independent per-client RNGs)
    import blades_b200.datasets.basedataset as bd
    bd.BaseDataset.compat = True
    glob = compat.run_script({script!r}, patch=patch)
model = glob["run_params"]["model"]
torch.save(torch.cat([p.detach().reshape(-1) for p in model.parameters()]), {out!r})
'''


@pytest.mark.parametrize("script", ["mini_example.py", "customize_attack.py"])
def test_reference_example_gives_the_same_model_on_both_packages(tmp_path, script):
    """The same script text, executed once against the reference package and once against blades_b200 through
    ``blades_b200.compat``: identical global model at the end."""
    import torch
    outs = {}
    for which in ("reference", "ours"):
        wd = tmp_path / which
        wd.mkdir()
        out = str(wd / "theta.pt")
        code = _COMPARE.format(root=ROOT, script=os.path.join(EXAMPLES, script), which=which, out=out)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=wd)
        assert r.returncode == 0, r.stderr[-3000:]
        outs[which] = torch.load(out)
    err = (outs["ours"] - outs["reference"]).abs().max().item()
    assert err <= 1e-5 * max(1.0, outs["reference"].abs().max().item()), err
