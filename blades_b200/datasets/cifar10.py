"""Federated CIFAR-10 (reference datasets/cifar10.py:11-108): NCHW float images in
[0,1]; per-batch augmentation RandomResizedCrop(32,(0.75,1),(1,1)) + HFlip +
Normalize + RandomErasing(0.25).  The reference's generation typos (Q8:
``train_set.dat``, 1-D ``y[idx, :]``, dropped ``seed``) are fixed."""
from typing import Optional

import numpy as np

from .basedataset import BaseDataset, partition

__all__ = ["CIFAR10"]


def _cifar_transforms(mean, std):
    import torchvision.transforms as T
    test = T.Compose([T.Normalize(mean=mean, std=std)])
    test.deterministic = True          # evaluation may cache the transformed test shards (Simulator.test_actor)
    train = T.Compose([
        T.RandomResizedCrop(32, scale=(0.75, 1.0), ratio=(1.0, 1.0)),
        T.RandomHorizontalFlip(p=0.5),
        T.Normalize(mean=mean, std=std),
        T.RandomErasing(p=0.25),
    ])
    return train, test


class CIFAR10(BaseDataset):
    stats = {"mean": (0.4914, 0.4822, 0.4465), "std": (0.2023, 0.1994, 0.2010)}
    img_size = 32
    num_classes = 10
    train_transform, test_transform = _cifar_transforms(stats["mean"], stats["std"])

    def __init__(self, data_root: str = './data', train_bs: Optional[int] = 32, iid: Optional[bool] = True,
                 alpha: Optional[float] = 0.1, num_clients: Optional[int] = 20, seed: Optional[int] = 1):
        super().__init__(data_root, train_bs, iid, alpha, num_clients, seed)

    def _load_raw(self, path):
        import torchvision
        tr = torchvision.datasets.CIFAR10(train=True, download=True, root=path)
        te = torchvision.datasets.CIFAR10(train=False, download=True, root=path)
        return tr.data, np.array(tr.targets), te.data, np.array(te.targets)

    def generate_datasets(self, path='./data', iid=True, alpha=0.1, num_clients=20, seed=1):
        x_tr, y_tr, x_te, y_te = self._load_raw(path)
        x_tr = np.transpose(x_tr.astype('float32') / 255.0, (0, 3, 1, 2))
        x_te = np.transpose(x_te.astype('float32') / 255.0, (0, 3, 1, 2))
        return partition(x_tr, y_tr, x_te, y_te, num_clients, iid, alpha, seed, self.num_classes)
