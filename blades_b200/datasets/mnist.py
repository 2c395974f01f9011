"""Federated MNIST (reference datasets/mnist.py:9-81): pixels / 255, seeded shuffle,
equal split (iid) or Dirichlet(alpha) label split."""
from typing import Optional


from .basedataset import BaseDataset, partition

__all__ = ["MNIST"]


class MNIST(BaseDataset):
    num_classes = 10

    def __init__(self, data_root: str = './data', train_bs: Optional[int] = 32, iid: Optional[bool] = True,
                 alpha: Optional[float] = 0.1, num_clients: Optional[int] = 20, seed: Optional[int] = 1):
        super().__init__(data_root, train_bs, iid, alpha, num_clients, seed)

    def _load_raw(self, path):
        import torchvision
        tr = torchvision.datasets.MNIST(train=True, download=True, root=path)
        te = torchvision.datasets.MNIST(train=False, download=True, root=path)
        return tr.data.numpy(), tr.targets.numpy(), te.data.numpy(), te.targets.numpy()

    def generate_datasets(self, path='./data', iid=True, alpha=0.1, num_clients=20, seed=1):
        x_tr, y_tr, x_te, y_te = self._load_raw(path)
        x_tr = x_tr.astype('float32') / 255.0
        x_te = x_te.astype('float32') / 255.0
        return partition(x_tr, y_tr, x_te, y_te, num_clients, iid, alpha, seed, self.num_classes)
