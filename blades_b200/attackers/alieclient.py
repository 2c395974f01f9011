""""A Little Is Enough" attacker (reference attackers/alieclient.py:8-37).

z_max = Phi^-1((n - f - s)/(n - f)),  s = floor(n/2 + 1) - f; the malicious update is
``mean - z_max * std`` (unbiased std) over the honest clients' updates, identical
for every ALIE client.

B200 path: instead of f callbacks each re-deriving the same d-vector on the host,
the attack is described by ``fused_spec()`` and evaluated inside the aggregation
kernel's prologue as *virtual rows* (csrc/cuda/coord_select.cu), or once on device
by ops.attack.alie_rows for Gram-based aggregators."""
import math

import torch

from ..client import ByzantineClient

__all__ = ["AlieClient", "alie_z_max"]


def _norm_ppf(p: float) -> float:
    try:
        from scipy.stats import norm
        return float(norm.ppf(p))
    except Exception:  # pragma: no cover - scipy is present in the image
        return math.sqrt(2.0) * float(torch.erfinv(torch.tensor(2.0 * p - 1.0, dtype=torch.float64)))


def alie_z_max(num_clients: int, num_byzantine: int) -> float:
    s = math.floor(num_clients / 2 + 1) - num_byzantine
    good = num_clients - num_byzantine
    return _norm_ppf((good - s) / good)


class AlieClient(ByzantineClient):
    def __init__(self, num_clients: int, num_byzantine: int, z=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.z_max = z if z is not None else alie_z_max(num_clients, num_byzantine)
        self.n_good = num_clients - num_byzantine

    def fused_spec(self):
        return {"kind": "alie", "param": float(self.z_max)}

    def omniscient_callback(self, simulator):
        honest = [c.get_update() for c in simulator._clients.values() if not c.is_byzantine()]
        stacked = torch.stack(honest, 0)
        self._gradient = stacked.mean(0) - self.z_max * stacked.std(0)
        self.save_update(self._gradient)
