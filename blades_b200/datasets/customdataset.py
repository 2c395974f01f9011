"""(x, y) tensor dataset with an optional per-item transform
(reference datasets/customdataset.py:4-21)."""
from torch.utils.data import Dataset

__all__ = ["CustomTensorDataset"]


class CustomTensorDataset(Dataset):
    def __init__(self, data_X, data_y, transform_list=None):
        self.tensors = (data_X, data_y)
        self.transforms = transform_list

    def __getitem__(self, index):
        x = self.tensors[0][index]
        if self.transforms:
            x = self.transforms(x)
        return x, self.tensors[1][index]

    def __len__(self):
        return self.tensors[1].size(0)
