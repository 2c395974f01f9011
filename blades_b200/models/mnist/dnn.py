"""MNIST MLP 784-64-128-10 (reference models/mnist/dnn.py:5-22).  Parameter names
(``layer1..3``) match the reference so state-dicts interchange.  The forward ends in
``log_softmax`` although training uses CrossEntropyLoss (quirk Q18: a harmless double
log-softmax); kept so trajectories match."""
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["MLP", "create_model"]


class MLP(nn.Module):
    def __init__(self, in_features: int = 28 * 28, hidden=(64, 128), num_classes: int = 10):
        super().__init__()
        self.flatten = nn.Flatten()
        self.layer1 = nn.Linear(in_features, hidden[0])
        self.layer2 = nn.Linear(hidden[0], hidden[1])
        self.layer3 = nn.Linear(hidden[1], num_classes)

    def forward(self, x):
        x = self.flatten(x)
        x = F.relu(self.layer1(x))
        x = F.relu(self.layer2(x))
        return F.log_softmax(self.layer3(x), dim=1)


def create_model():
    return MLP(), nn.CrossEntropyLoss()
