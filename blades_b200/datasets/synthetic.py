"""Synthetic federated datasets of a named shape (no network on the GPU boxes;
benchmarks and tests use these -- SURVEY 7.4 "Data on a no-network GPU box").

``Synthetic`` follows the same cache/partition/stream pipeline as the real datasets
(so the whole data path is exercised); the samples are drawn from a class-conditional
Gaussian so that a model can actually learn (loss decreases in the e2e tests).

``SyntheticMNIST / SyntheticCIFAR10 / SyntheticCIFAR100`` fix the shapes.
``synthetic_fldataset`` builds an ``FLDataset`` directly in memory (no cache file).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from .basedataset import BaseDataset, BatchStream, partition
from .customdataset import CustomTensorDataset
from .dataset import FLDataset

__all__ = ["Synthetic", "SyntheticMNIST", "SyntheticCIFAR10", "SyntheticCIFAR100", "synthetic_fldataset"]


def _make_samples(rng: np.random.Generator, n: int, shape: Tuple[int, ...], num_classes: int,
                  separation: float = 1.0):
    y = rng.integers(0, num_classes, size=n)
    dim = int(np.prod(shape))
    # low-rank class means keep generation cheap for image-sized inputs
    basis = rng.standard_normal((num_classes, min(dim, 64))).astype(np.float32)
    proj = rng.standard_normal((min(dim, 64), dim)).astype(np.float32) / np.sqrt(min(dim, 64))
    x = rng.standard_normal((n, dim), dtype=np.float32)
    x += separation * (basis[y] @ proj)
    return x.reshape((n,) + tuple(shape)), y.astype(np.int64)


class Synthetic(BaseDataset):
    shape: Tuple[int, ...] = (1, 28, 28)
    num_classes = 10
    train_per_client = 128
    test_per_client = 32

    def __init__(self, data_root: str = './data', train_bs: Optional[int] = 32, iid: Optional[bool] = True,
                 alpha: Optional[float] = 0.1, num_clients: Optional[int] = 20, seed: Optional[int] = 1,
                 train_per_client: Optional[int] = None, test_per_client: Optional[int] = None,
                 shape: Optional[Sequence[int]] = None, num_classes: Optional[int] = None):
        if train_per_client is not None:
            self.train_per_client = train_per_client
        if test_per_client is not None:
            self.test_per_client = test_per_client
        if shape is not None:
            self.shape = tuple(shape)
        if num_classes is not None:
            self.num_classes = num_classes
        super().__init__(data_root, train_bs, iid, alpha, num_clients, seed)

    def _extra_meta(self):
        return {"shape": tuple(self.shape), "classes": self.num_classes,
                "tpc": self.train_per_client, "tepc": self.test_per_client}

    def generate_datasets(self, path='./data', iid=True, alpha=0.1, num_clients=20, seed=1):
        rng = np.random.default_rng(seed)
        x_tr, y_tr = _make_samples(rng, self.train_per_client * num_clients, self.shape, self.num_classes)
        x_te, y_te = _make_samples(rng, self.test_per_client * num_clients, self.shape, self.num_classes)
        return partition(x_tr, y_tr, x_te, y_te, num_clients, iid, alpha, seed, self.num_classes,
                         strict=False)


class SyntheticMNIST(Synthetic):
    shape = (28, 28)


class SyntheticCIFAR10(Synthetic):
    shape = (3, 32, 32)


class SyntheticCIFAR100(Synthetic):
    shape = (3, 32, 32)
    num_classes = 100


def synthetic_fldataset(num_clients: int, shape: Sequence[int] = (3, 32, 32), num_classes: int = 10,
                        train_bs: int = 32, train_per_client: int = 64, test_per_client: int = 32,
                        seed: int = 1, separation: float = 1.0) -> FLDataset:
    """In-memory FLDataset of random class-conditional Gaussians."""
    rng = np.random.default_rng(seed)
    trains, tests = [], []
    import torch
    n_tr, n_te = train_per_client * num_clients, test_per_client * num_clients
    # one draw for everybody: all clients share the same class-conditional distribution
    X, Y = _make_samples(rng, n_tr + n_te, tuple(shape), num_classes, separation)
    for c in range(num_clients):
        sl = slice(c * train_per_client, (c + 1) * train_per_client)
        x, y = X[sl], Y[sl]
        trains.append(BatchStream(x, y, train_bs, seed=c))
        st = slice(n_tr + c * test_per_client, n_tr + (c + 1) * test_per_client)
        xt, yt = X[st], Y[st]
        tests.append(CustomTensorDataset(torch.from_numpy(xt), torch.from_numpy(yt)))
    return FLDataset(trains, tests)
