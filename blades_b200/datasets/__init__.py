from .basedataset import BaseDataset, BatchStream
from .cifar10 import CIFAR10
from .cifar100 import CIFAR100
from .customdataset import CustomTensorDataset
from .dataset import FLDataset
from .mnist import MNIST
from .synthetic import (Synthetic, SyntheticCIFAR10, SyntheticCIFAR100, SyntheticMNIST,
                        synthetic_fldataset)

__all__ = ["BaseDataset", "BatchStream", "MNIST", "CIFAR10", "CIFAR100", "FLDataset", "CustomTensorDataset", "Synthetic",
           "SyntheticMNIST", "SyntheticCIFAR10", "SyntheticCIFAR100", "synthetic_fldataset"]
