"""Label-flipping attacker (reference attackers/labelflippingclient.py:13-23):
``target <- num_classes - 1 - target`` on every training batch.

The batched engine applies the same transform to the label rows of Byzantine
clients before the shared forward pass (``batch_label_transform``)."""
from ..client import ByzantineClient

__all__ = ["LabelflippingClient"]


class LabelflippingClient(ByzantineClient):
    def __init__(self, num_classes=10, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.num_classes = num_classes

    def on_train_batch_begin(self, data, target, logs=None):
        return data, self.num_classes - 1 - target

    def __str__(self) -> str:
        return "LableFlippingWorker"
