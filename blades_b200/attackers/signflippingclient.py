"""Sign-flipping attacker (reference attackers/signflippingclient.py:10-21): every
local step ascends the loss (gradients negated before ``optimizer.step``); the loss
clamp is 1e5 instead of 1e6.

Batched engine: with plain SGD, negating every gradient is equivalent to stepping
with ``-lr``; the engine uses ``grad_sign = -1`` for these clients' rows."""
from ..client import ByzantineClient

__all__ = ["SignflippingClient"]


class SignflippingClient(ByzantineClient):
    loss_clamp = 1e5
    #: consumed by engine.batched: per-client gradient sign
    grad_sign = -1.0

    def _post_backward(self) -> None:
        for _, p in self.model.named_parameters():
            if p.grad is not None:
                p.grad.neg_()
