"""``CCTNet``: the model of the reference's CIFAR-10 experiments -- ``cct_2_3x2_32`` (2 conv layers of 3x3 in the
tokenizer, 2 encoder layers, 32x32 inputs; 283 723 parameters) behind a wrapper module.

The reference stores the network under the misspelt attribute ``mdoel`` (models/cifar10/cct.py:6-12), which makes
every ``state_dict`` key start with ``mdoel.``; the spelling is kept so checkpoints interchange (parity test:
``tests/test_reference_parity.py::test_cctnet_and_mlp_match_reference``).
"""
from torch import nn

from .cctnets import cct_2_3x2_32

__all__ = ["CCTNet", "create_model"]


class CCTNet(nn.Module):
    def __init__(self, num_classes: int = 10, **cct_kwargs):
        super().__init__()
        # ``cct_kwargs`` reach the factory, e.g. attention_dropout=0.0, stochastic_depth=0.0 for deterministic tests
        self.add_module("mdoel", cct_2_3x2_32(num_classes=num_classes, **cct_kwargs))

    def forward(self, images):
        return self.mdoel(images)


def create_model():
    """(model, loss) pair in the style of the LEAF model files."""
    return CCTNet(), nn.CrossEntropyLoss()
