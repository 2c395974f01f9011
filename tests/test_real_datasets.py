"""The real dataset classes (MNIST / CIFAR-10 / CIFAR-100, torchvision download replaced by deterministic fakes)
through whole simulations on the CPU: iid and Dirichlet partitions, shards that are not a multiple of the batch size
(short tail batches, several epochs), per-batch augmentation, fedsgd and fedavg."""
import numpy as np
import pytest
import torch

from blades_b200 import Simulator
from blades_b200.models.mnist import MLP


class _FakeMNIST:
    def __init__(self, train=True, download=True, root=None, **kw):
        g = torch.Generator().manual_seed(11 if train else 12)
        n = 1200 if train else 200
        self.data = torch.randint(0, 256, (n, 28, 28), generator=g, dtype=torch.uint8)
        self.targets = torch.randint(0, 10, (n,), generator=g)


def _fake_cifar(classes):
    class _FakeCIFAR:
        def __init__(self, train=True, download=True, root=None, **kw):
            rng = np.random.default_rng(1 if train else 2)
            n = 600 if train else 120
            self.data = rng.integers(0, 256, (n, 32, 32, 3), dtype=np.uint8)
            self.targets = rng.integers(0, classes, n).tolist()
    return _FakeCIFAR


@pytest.mark.parametrize("iid,local_steps", [(True, 1), (False, 1), (False, 3)])
def test_mnist_end_to_end_with_tail_batches(tmp_path, monkeypatch, iid, local_steps):
    import torchvision
    from blades_b200.datasets import MNIST
    monkeypatch.setattr(torchvision.datasets, "MNIST", _FakeMNIST)
    ds = MNIST(data_root=str(tmp_path / "d"), train_bs=32, num_clients=10, iid=iid, alpha=0.5, seed=1)   # 120 = 3*32 + 24
    sim = Simulator(ds, num_byzantine=2, attack="alie", attack_kws={"num_clients": 10, "num_byzantine": 2},
                    aggregator="trimmedmean", aggregator_kws={"nb": 2}, log_path=str(tmp_path / "l"), seed=1, progress=False)
    m = MLP()
    times = sim.run(m, global_rounds=9, local_steps=local_steps, server_lr=1.0, client_lr=0.1, validate_interval=3)
    assert len(times) == 9 and all(torch.isfinite(p).all() for p in m.parameters())


@pytest.mark.parametrize("name,classes", [("CIFAR10", 10), ("CIFAR100", 100)])
def test_cifar_end_to_end_with_augmentation(tmp_path, monkeypatch, name, classes):
    import torchvision
    import blades_b200.datasets as D
    from blades_b200.models.cifar10.cctnets import cct_2_3x2_32
    monkeypatch.setattr(torchvision.datasets, name, _fake_cifar(classes))
    ds = getattr(D, name)(data_root=str(tmp_path / "d"), train_bs=16, num_clients=6, seed=1)     # 100 = 6*16 + 4
    sim = Simulator(ds, num_byzantine=1, attack="ipm", aggregator="median", log_path=str(tmp_path / "l"), seed=1,
                    progress=False)
    m = cct_2_3x2_32(num_classes=classes)
    times = sim.run(m, global_rounds=8, local_steps=1, server_lr=1.0, client_lr=0.05, validate_interval=4)
    assert len(times) == 8 and all(torch.isfinite(p).all() for p in m.parameters())
