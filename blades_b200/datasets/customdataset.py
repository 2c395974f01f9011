"""Per-client test shard: a pair of tensors ``(X[N, ...], y[N])`` with an optional per-item transform.

Same constructor and attributes as the reference's class of this name (datasets/customdataset.py:4-21:
``CustomTensorDataset(data_X, data_y, transform_list=None)``, ``.tensors``, ``.transforms``), plus two things the
engine uses: ``batches(batch_size)`` yields whole mini-batches without going through a ``DataLoader`` (the
evaluation fast path caches deterministic shards on the device), and ``deterministic`` tells whether the transform
may be applied once and cached.
"""
from typing import Iterator, Tuple

import torch
from torch.utils.data import Dataset

__all__ = ["CustomTensorDataset"]


class CustomTensorDataset(Dataset):
    def __init__(self, data_X: torch.Tensor, data_y: torch.Tensor, transform_list=None):
        if len(data_X) != len(data_y):
            raise ValueError(f"{len(data_X)} samples but {len(data_y)} labels")
        self.tensors = (data_X, data_y)
        self.transforms = transform_list

    @property
    def deterministic(self) -> bool:
        """No transform, or one that is flagged as free of randomness (e.g. the CIFAR test normalisation)."""
        return self.transforms is None or bool(getattr(self.transforms, "deterministic", False))

    def __len__(self) -> int:
        return int(self.tensors[1].shape[0])

    def __getitem__(self, index):
        X, y = self.tensors
        sample = X[index] if self.transforms is None else self.transforms(X[index])
        return sample, y[index]

    def batches(self, batch_size: int) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        X, y = self.tensors
        for lo in range(0, len(self), batch_size):
            xb = X[lo: lo + batch_size]
            if self.transforms is not None:
                xb = torch.stack([self.transforms(v) for v in xb])
            yield xb, y[lo: lo + batch_size]
