"""Federated CIFAR-100 (new; BASELINE config #5 names CIFAR-100/ResNet-50 -- the
reference ships no CIFAR-100).  Same pipeline as CIFAR10 with 100 classes;
uneven client splits are allowed so 512 clients work."""

import numpy as np

from .basedataset import partition
from .cifar10 import CIFAR10, _cifar_transforms

__all__ = ["CIFAR100"]


class CIFAR100(CIFAR10):
    stats = {"mean": (0.5071, 0.4865, 0.4409), "std": (0.2673, 0.2564, 0.2762)}
    num_classes = 100
    train_transform, test_transform = _cifar_transforms(stats["mean"], stats["std"])

    def _load_raw(self, path):
        import torchvision
        tr = torchvision.datasets.CIFAR100(train=True, download=True, root=path)
        te = torchvision.datasets.CIFAR100(train=False, download=True, root=path)
        return tr.data, np.array(tr.targets), te.data, np.array(te.targets)

    def generate_datasets(self, path='./data', iid=True, alpha=0.1, num_clients=20, seed=1):
        x_tr, y_tr, x_te, y_te = self._load_raw(path)
        x_tr = np.transpose(x_tr.astype('float32') / 255.0, (0, 3, 1, 2))
        x_te = np.transpose(x_te.astype('float32') / 255.0, (0, 3, 1, 2))
        return partition(x_tr, y_tr, x_te, y_te, num_clients, iid, alpha, seed, self.num_classes,
                         strict=False)
