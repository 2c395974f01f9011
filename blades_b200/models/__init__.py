"""Model zoo.  ``MLP`` / ``CCTNet`` mirror the reference (models/mnist/dnn.py,
models/cifar10/cct.py); ``resnet18`` / ``resnet50`` are new (BASELINE configs)."""
from .cifar10 import CCTNet
from .mnist import MLP
from .resnet import resnet18, resnet50, ResNet

__all__ = ["MLP", "CCTNet", "resnet18", "resnet50", "ResNet"]
