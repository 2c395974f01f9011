"""Gaussian-noise attacker (reference attackers/noiseclient.py:8-25): after local
training the saved update is replaced by i.i.d. N(mean, std) noise.

On the engine path the noise is generated directly into the client's row of the
device update matrix (in-kernel Philox, ops.attack.fill_normal) -- nothing crosses
the host."""
from typing import Optional

import torch

from ..client import ByzantineClient

__all__ = ["NoiseClient"]


class NoiseClient(ByzantineClient):
    #: the callback only rewrites this client's own row: on several GPUs the owning rank runs it in place and the
    #: update matrix is never gathered (Simulator.train_actor)
    row_local_attack = True

    def __init__(self, mean: Optional[float] = 0.1, std: Optional[float] = 0.1, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._noise_mean = mean
        self._noise_std = std

    def omniscient_callback(self, simulator):
        cur = self._get_saved_update()
        if cur.is_cuda:
            from ..ops import attack as _k
            # Philox stream position from (round, client): the same noise whichever rank owns the client, and a
            # resumed run continues the stream instead of replaying it
            clients = simulator.get_clients() if hasattr(simulator, "get_clients") else []
            idx = next((i for i, c in enumerate(clients) if c is self), 0)
            rnd = int(getattr(simulator, "_round_index", 0))
            per = (cur.numel() + 3) // 4
            _k.fill_normal_(cur, self._noise_mean, self._noise_std, offset=(rnd * max(len(clients), 1) + idx) * per)
            self._state["saved_update"] = cur
        else:
            noise = torch.normal(self._noise_mean, self._noise_std, size=cur.shape)
            self.save_update(noise.to(cur.dtype))
