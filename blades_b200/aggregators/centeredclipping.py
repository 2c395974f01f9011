"""Centered clipping (reference aggregators/centeredclipping.py:13-49) and the
private anchor/async variants (:52-137).

Stateful: the momentum vector survives across rounds.  All ``n_iter`` clipping
iterations run on the (N+1)x(N+1) Gram matrix of ``[U; m_prev]``; the update
matrix itself is read twice in total (Gram + combine) instead of ``2*n_iter`` times.
Unlike the reference (Q16) any input convention is accepted, not only client lists.
"""
from __future__ import annotations

import logging
import types
from typing import Optional

import numpy as np
import torch

from . import _gramops as gops
from .base import _BaseAggregator, _BaseAsyncAggregator

__all__ = ["Centeredclipping", "_AnchorClipping", "_AsyncCenteredClipping"]

debug_logger = logging.getLogger("debug")


class Centeredclipping(_BaseAggregator):
    def __init__(self, tau: Optional[float] = 10.0, n_iter: Optional[int] = 5):
        super().__init__()
        self.tau = tau
        self.n_iter = n_iter
        self.momentum: Optional[torch.Tensor] = None

    def clip(self, v: torch.Tensor) -> torch.Tensor:
        nrm = torch.norm(v)
        return v * min(1.0, (self.tau / nrm).item() if nrm > 0 else 1.0)

    def aggregate(self, matrix):
        n, d = matrix.n_rows, matrix.n_cols
        if n + 1 <= 512:
            m = (self.momentum.to(matrix.device) if self.momentum is not None
                 else torch.zeros(d, device=matrix.device, dtype=torch.float32))        # m_0 = 0
            dg = matrix.gram_device(extra=m)
            if dg is not None:
                # clipping iterations on the device (csrc/cuda/gram_solve.cu); the last coefficient weighs m itself
                from ..ops import gram_solve
                c = gram_solve.centered_clip_coeffs(dg, self.tau, self.n_iter)
                new_m = matrix.combine(c, extra=m)
                if self.momentum is not None and self.momentum.is_cuda and self.momentum.shape == new_m.shape \
                        and self.momentum.dtype == new_m.dtype and self.momentum.device == new_m.device:
                    self.momentum.copy_(new_m)       # in place: a captured round keeps reading this very buffer
                else:
                    self.momentum = new_m.detach().clone()
                return new_m
        if self.momentum is None:
            # m = 0: Gram row/col of zeros, no need to touch the device for it
            G = matrix.gram()
            G_aug = np.zeros((n + 1, n + 1))
            G_aug[:n, :n] = G
            c = gops.centered_clip_coeffs(G_aug, self.tau, self.n_iter)
            new_m = matrix.combine(c[:n])
        else:
            m = self.momentum.to(matrix.device)
            G_aug = matrix.gram(extra=m)
            c = gops.centered_clip_coeffs(G_aug, self.tau, self.n_iter)
            new_m = matrix.combine(c[:n], extra=m.to(torch.float32) if m.is_cuda else m, extra_weight=float(c[n]))
        self.momentum = new_m.detach().clone()      # the returned vector may be a reused device buffer
        return new_m

    def state_dict(self):
        return {"momentum": None if self.momentum is None else self.momentum.detach().cpu()}

    def load_state_dict(self, state):
        self.momentum = state.get("momentum")

    def __str__(self):
        return "Clipping (tau={}, n_iter={})".format(self.tau, self.n_iter)


def _flat_state(model: torch.nn.Module) -> torch.Tensor:
    return torch.cat([v.detach().reshape(-1) for v in model.state_dict().values()])


class _AnchorClipping(Centeredclipping):
    """Decentralised variant: clip every neighbour around an anchor that tracks the
    node's own optimizer steps (reference centeredclipping.py:52-103, unwired there)."""

    def __init__(self, node, weights, opt, model, tau, n_iter=1):
        super().__init__(tau, n_iter)
        assert n_iter == 1 and weights.dim() == 1
        self._anchor_buffer = _flat_state(model).clone()
        self.node = node
        self.weights = weights
        self.opt = self._wrap_step(opt, model)

    def _wrap_step(self, opt, model):
        if hasattr(opt, "_core_step") or hasattr(opt, "anchorclipping"):
            raise NotImplementedError("optimizer already wrapped")
        debug_logger.info("Wrap the step function of opt")
        opt._core_step = types.MethodType(type(opt).step, opt)
        opt.anchorclipping = self

        def step(this, closure=None):
            before = _flat_state(model).clone()
            this._core_step(closure=closure)
            this.anchorclipping._anchor_buffer.add_(_flat_state(model) - before)

        opt.step = types.MethodType(step, opt)
        return opt

    def __call__(self, inputs):
        assert len(inputs) == 1 + len(self.node.edges)
        a = self._anchor_buffer
        acc = self.weights[self.node.index] * (a + self.clip(inputs[0] - a))
        for edge, vec in zip(self.node.edges, inputs[1:]):
            acc = acc + self.weights[edge.theother(self.node).index] * (a + self.clip(vec - a))
        return acc

    def __str__(self):
        return "_AnchorClipping(tau={}, n_iter={})".format(self.tau, self.n_iter)


class _AsyncCenteredClipping(_BaseAsyncAggregator):
    """Like ``Centeredclipping`` but divides by the number of slots, counting stragglers."""

    def __init__(self, tau, n_iter=1):
        super().__init__()
        self.tau = tau
        self.n_iter = n_iter
        self.momentum = 0

    def clip(self, v):
        nrm = torch.norm(v)
        return v * min(1.0, (self.tau / nrm).item() if nrm > 0 else 1.0)

    def __call__(self, inputs):
        n = len(inputs)
        present = [x for x in inputs if x is not None]
        for _ in range(self.n_iter):
            self.momentum = sum(self.clip(v - self.momentum) for v in present) / n + self.momentum
        return torch.clone(self.momentum).detach()

    def __str__(self):
        return "_AsyncCenteredClipping (tau={}, n_iter={})".format(self.tau, self.n_iter)
