"""Inner-product-manipulation attacker (reference attackers/ipmclient.py:4-16):
uploads ``-epsilon * mean(honest updates)``.  Fusable as virtual rows like ALIE."""
import torch

from ..client import ByzantineClient

__all__ = ["IpmClient"]


class IpmClient(ByzantineClient):
    def __init__(self, epsilon: float = 0.5, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.epsilon = epsilon

    def fused_spec(self):
        return {"kind": "ipm", "param": float(self.epsilon)}

    def omniscient_callback(self, simulator):
        honest = [c.get_update() for c in simulator.get_clients() if not c.is_byzantine()]
        self.save_update(-self.epsilon * torch.stack(honest, 0).mean(0))
