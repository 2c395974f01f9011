"""ByzantineSGD filter of Alistarh et al. (reference aggregators/byzantinesgd.py:16-80).

Stateful per worker: ``A_i`` accumulates <g_i, theta - theta_0>, ``B_i`` accumulates
the gradients themselves (an N x d state -- kept as one dense matrix here instead
of a Python list of vectors).  Not reachable by name in the reference either
(class name != module.capitalize()); exported for completeness."""
from __future__ import annotations

import torch

from .base import _BaseAggregator

__all__ = ["ByzantineSGD"]


def _get_vectorized_parameters(optimizer) -> torch.Tensor:
    return torch.cat([p.data.reshape(-1) for g in optimizer.param_groups for p in g["params"]])


class ByzantineSGD(_BaseAggregator):
    fusable_final = False
    def __init__(self, m, th_A, th_B, th_V, optimizer):
        super().__init__()
        self.m = m
        self.th_A = th_A
        self.th_B = th_B
        self.th_V = th_V
        self.optimizer = optimizer
        self.init_model = _get_vectorized_parameters(optimizer).clone()
        self.A = torch.zeros(m, dtype=torch.float64)
        self.B = None                      # [m, d]
        self.good = list(range(m))
        self.debug_message = ""

    def vector_median(self, vs: torch.Tensor, threshold: float):
        """First row with more than m/2 rows within ``threshold`` of it (rows scanned in order)."""
        dist = torch.cdist(vs[None].double(), vs[None].double())[0]
        for i in range(self.m):
            count = 0
            for j in range(self.m):
                count += int(dist[i, j] <= threshold)
                if count > self.m / 2:
                    return i, vs[i]
        raise NotImplementedError("No median found")

    def __call__(self, inputs):
        grads = self._get_updates(inputs)
        diff = _get_vectorized_parameters(self.optimizer) - self.init_model
        self.A += (grads.double() @ diff.double().to(grads.device)).cpu()
        self.B = grads.clone() if self.B is None else self.B + grads
        A_med = self.A.median().item() if self.m % 2 else \
            0.5 * (self.A.sort().values[self.m // 2 - 1] + self.A.sort().values[self.m // 2]).item()
        _, B_med = self.vector_median(self.B, self.th_B)
        _, g_med = self.vector_median(grads, 2 * self.th_V)
        keep = []
        for i in self.good:
            a_ok = abs(self.A[i].item() - A_med) <= self.th_A
            b_ok = (self.B[i] - B_med).norm().item() <= self.th_B
            g_ok = (grads[i] - g_med).norm().item() <= 4 * self.th_V
            if a_ok and b_ok and g_ok:
                keep.append(i)
        self.good = keep
        return grads[self.good].sum(0) / len(self.good)

    def state_dict(self):
        return {"A": self.A, "B": self.B, "good": list(self.good), "init_model": self.init_model}

    def load_state_dict(self, state):
        self.A, self.B, self.good = state["A"], state["B"], list(state["good"])
        self.init_model = state["init_model"]
