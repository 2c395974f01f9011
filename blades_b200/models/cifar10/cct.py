"""``CCTNet`` = ``cct_2_3x2_32`` wrapper used by the CIFAR-10 experiments
(reference models/cifar10/cct.py:6-12).  The attribute is spelled ``mdoel`` in the
reference, so state-dict keys are ``mdoel.*``; kept for checkpoint compatibility."""
import torch.nn as nn

from .cctnets import cct_2_3x2_32

__all__ = ["CCTNet"]


class CCTNet(nn.Module):
    def __init__(self, num_classes: int = 10):
        super().__init__()
        self.mdoel = cct_2_3x2_32(num_classes=num_classes)

    def forward(self, x):
        return self.mdoel(x)


def create_model():
    return CCTNet(), nn.CrossEntropyLoss()
