"""Federated dataset base: partition cache + per-client batch streams.

Behavioural contract: /root/reference/src/blades/datasets/basedataset.py:13-115
  * ctor compares a ``meta_info`` dict with the one stored in
    ``<data_root>/<ClassName>.obj`` and regenerates the partition iff they differ;
  * cache file = 5 consecutive pickles ``(meta, train_ids, train_data{id:{x,y}},
    test_ids, test_data)`` -- format kept byte-compatible so caches interchange;
  * training data is an infinite stream of ``(X[bs,...] float32, y int64)`` batches,
    reshuffled every epoch; ``train_transform`` is applied per batch.

Differences: the stream is a resumable iterator object (``BatchStream``) with its
own ``numpy.random.Generator`` instead of a generator function that reseeds the
process-global RNG (basedataset.py:67) -- so data cursors can be checkpointed and
clients do not perturb each other's randomness.  ``legacy_rng=True`` reproduces the
reference's exact batch order (global ``np.random`` seeded with 0 per stream).
Test data uses ``test_transform`` (fixes Q9) unless ``compat=True``.
"""
from __future__ import annotations

import os
import pickle
from abc import ABC, abstractmethod
from typing import Optional

import numpy as np
import torch

from .customdataset import CustomTensorDataset

__all__ = ["BaseDataset", "BatchStream"]


class BatchStream:
    """Infinite, resumable, reshuffling mini-batch iterator over one client's shard."""

    def __init__(self, data: np.ndarray, labels: np.ndarray, batch_size: int, seed: int = 0,
                 transform=None, legacy_rng: bool = False):
        self.data = np.asarray(data)
        self.labels = np.asarray(labels)
        self.batch_size = batch_size
        self.transform = transform
        self.legacy = legacy_rng
        self.seed = seed
        self._rng = np.random.default_rng(seed)
        self._started = False
        self._epoch = 0
        self._pos = 0
        self._perm: Optional[np.ndarray] = None

    def _reshuffle(self):
        n = len(self.labels)
        if self.legacy:
            if not self._started:
                from ..utils import set_random_seed
                set_random_seed(self.seed)
            # the reference permutes the ALREADY permuted arrays at every epoch (basedataset.py:69-75): compose
            idx = np.random.permutation(n)
            self._perm = idx if self._perm is None else self._perm[idx]
        else:
            self._perm = self._rng.permutation(n)
        self._started = True
        self._pos = 0

    def __iter__(self):
        return self

    def next_indices(self) -> np.ndarray:
        """Sample indices of the next mini-batch (advances the cursor)."""
        if self._perm is None:
            self._reshuffle()
        if self._pos * self.batch_size >= len(self.labels):
            self._epoch += 1
            self._reshuffle()
        sl = self._perm[self._pos * self.batch_size:(self._pos + 1) * self.batch_size]
        self._pos += 1
        return sl

    def __next__(self):
        sl = self.next_indices()
        X = torch.from_numpy(np.ascontiguousarray(self.data[sl])).float()
        if self.transform:
            X = self.transform(X)
        return X, torch.from_numpy(np.ascontiguousarray(self.labels[sl])).long()

    # -- checkpointing -------------------------------------------------------------
    def state(self) -> dict:
        return {"rng": self._rng.bit_generator.state, "epoch": self._epoch, "pos": self._pos,
                # ``_perm`` is replaced, never mutated in place: a reference is a consistent O(1) snapshot
                "perm": self._perm, "started": self._started}

    def load_state(self, st: dict) -> None:
        self._rng.bit_generator.state = st["rng"]
        self._epoch, self._pos, self._started = st["epoch"], st["pos"], st["started"]
        self._perm = None if st["perm"] is None else np.array(st["perm"])


class BaseDataset(ABC):
    train_transform = None
    test_transform = None
    #: replicate reference quirks (test data through train_transform, legacy RNG)
    compat = False

    def __init__(self, data_root: str = './data', train_bs: Optional[int] = 32, iid: Optional[bool] = True,
                 alpha: Optional[float] = 0.1, num_clients: Optional[int] = 20, seed=1):
        self.train_bs = train_bs
        self.num_clients = num_clients
        os.makedirs(data_root, exist_ok=True)
        self._data_path = os.path.join(data_root, type(self).__name__ + '.obj')
        meta_info = {"num_clients": num_clients, "data_root": data_root, "train_bs": train_bs,
                     "iid": iid, "alpha": alpha, "seed": seed}
        meta_info.update(self._extra_meta())
        if not self._cache_matches(meta_info):
            parts = self.generate_datasets(data_root, iid, alpha, num_clients, seed)
            with open(self._data_path, 'wb') as f:
                pickle.dump(meta_info, f)
                for obj in parts:
                    pickle.dump(obj, f)

    def _extra_meta(self) -> dict:
        return {}

    def _cache_matches(self, meta_info: dict) -> bool:
        if not os.path.exists(self._data_path):
            return False
        try:
            with open(self._data_path, 'rb') as f:
                return pickle.load(f) == meta_info
        except Exception:
            return False

    @abstractmethod
    def generate_datasets(self, path='./data', iid=True, alpha=0.1, num_clients=20, seed=1):
        """Return ``(train_ids, train_data, test_ids, test_data)``."""

    def _preprocess_train_data(self, data, labels, batch_size, seed=0) -> BatchStream:
        return BatchStream(data, labels, batch_size, seed=seed, transform=self.train_transform,
                           legacy_rng=self.compat)

    def _preprocess_test_data(self, data, labels) -> CustomTensorDataset:
        tf = self.train_transform if self.compat else self.test_transform
        return CustomTensorDataset(torch.as_tensor(np.asarray(data)).float(),
                                   torch.as_tensor(np.asarray(labels)).long(), transform_list=tf)

    def get_dls(self):
        assert os.path.isfile(self._data_path)
        with open(self._data_path, 'rb') as f:
            (_, train_clients, train_data, test_clients, test_data) = [pickle.load(f) for _ in range(5)]
        assert sorted(train_clients) == sorted(test_clients)
        train_dls, test_dls = [], []
        for idx, u_id in enumerate(train_clients):
            train_dls.append(self._preprocess_train_data(
                np.array(train_data[u_id]['x']), np.array(train_data[u_id]['y']), self.train_bs,
                seed=0 if self.compat else idx))
            test_dls.append(self._preprocess_test_data(
                np.array(test_data[u_id]['x']), np.array(test_data[u_id]['y'])))
        return train_dls, test_dls


# --------------------------------------------------------------------------- partitioning helpers
def seeded_shuffle(seed, *arrays):
    """``np.random.seed(seed)`` then one shared permutation of all arrays
    (what ``sklearn.utils.shuffle`` does in the reference, mnist.py:34-36)."""
    np.random.seed(seed)
    out = []
    for a in arrays_pairs(arrays):
        perm = np.random.permutation(len(a[0]))
        out.extend(x[perm] for x in a)
    return out


def arrays_pairs(arrays):
    # (x_train, y_train, x_test, y_test) -> [(x_train, y_train), (x_test, y_test)]
    return [arrays[i:i + 2] for i in range(0, len(arrays), 2)]


def split_even(x, y, num_clients, strict=True):
    """IID split.  ``strict`` = reference behaviour (``np.split`` raises unless the size divides);
    non-strict uses ``np.array_split`` (needed e.g. for 512 clients)."""
    fn = np.split if strict else np.array_split
    return fn(x, num_clients), fn(y, num_clients)


def split_dirichlet(y, num_clients, alpha, num_classes, min_size=10):
    """Dirichlet(alpha) label-proportion partition with the reference's balancing mask
    and ``min_size`` retry loop (mnist.py:47-66).  Returns a list of index lists."""
    N = y.shape[0]
    while True:
        idx_batch = [[] for _ in range(num_clients)]
        for k in range(num_classes):
            idx_k = np.where(y == k)[0]
            np.random.shuffle(idx_k)
            prop = np.random.dirichlet(np.repeat(alpha, num_clients))
            prop = np.array([p * (len(b) < N / num_clients) for p, b in zip(prop, idx_batch)])
            prop = prop / prop.sum()
            cuts = (np.cumsum(prop) * len(idx_k)).astype(int)[:-1]
            idx_batch = [b + part.tolist() for b, part in zip(idx_batch, np.split(idx_k, cuts))]
        if min(len(b) for b in idx_batch) >= min_size:
            break
    for b in idx_batch:
        np.random.shuffle(b)
    return idx_batch


def partition(x_train, y_train, x_test, y_test, num_clients, iid, alpha, seed, num_classes,
              strict=True):
    """Shared body of every ``generate_datasets``: shuffle, split, build the id->{x,y} dicts."""
    x_train, y_train, x_test, y_test = seeded_shuffle(seed, x_train, y_train, x_test, y_test)
    ids = [str(i) for i in range(num_clients)]
    xs_te, ys_te = split_even(x_test, y_test, num_clients, strict)
    if iid:
        xs_tr, ys_tr = split_even(x_train, y_train, num_clients, strict)
    else:
        parts = split_dirichlet(y_train, num_clients, alpha, num_classes)
        xs_tr = [x_train[p] for p in parts]
        ys_tr = [y_train[p] for p in parts]
    train = {u: {'x': xs_tr[i], 'y': np.asarray(ys_tr[i]).flatten()} for i, u in enumerate(ids)}
    test = {u: {'x': xs_te[i], 'y': np.asarray(ys_te[i]).flatten()} for i, u in enumerate(ids)}
    return ids, train, ids, test
