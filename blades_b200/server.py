"""The FL server: global model + server optimizer + aggregator.

Contract: /root/reference/src/blades/server.py:22-75 -- ``BladesServer(optimizer,
model, aggregator)`` with ``get_opt / zero_grad / get_model / apply_update``.
``apply_update(agg)`` treats ``-agg`` as the pseudo-gradient and takes one
optimizer step, so plain SGD gives ``theta <- theta + lr * agg``.

B200 design: when the model's parameters are views into one flat buffer
(``engine.flat.FlatParams``) and the optimizer is momentum-free SGD, the step is
one fused axpy on the flat vector (or already happened inside the aggregation
kernel's epilogue -- see ``ops.fused_round``); the per-parameter Python walk of the
reference only remains as the general fallback (any torch optimizer).
"""
from __future__ import annotations

from typing import Callable

import torch

__all__ = ["BladesServer"]


def _is_plain_sgd(opt: torch.optim.Optimizer) -> bool:
    if type(opt) is not torch.optim.SGD:
        return False
    for g in opt.param_groups:
        if g.get("momentum", 0) != 0 or g.get("weight_decay", 0) != 0 or g.get("nesterov", False) \
                or g.get("dampening", 0) != 0 or g.get("maximize", False):
            return False
    return True


class BladesServer:
    def __init__(self, optimizer: torch.optim.Optimizer, model: torch.nn.Module,
                 aggregator: Callable[[list], torch.Tensor], *args, flat=None, **kwargs):
        self.optimizer = optimizer
        self.model = model
        self.aggregator = aggregator
        self.flat = flat  # Optional[FlatParams]

    def get_opt(self) -> torch.optim.Optimizer:
        return self.optimizer

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.optimizer.zero_grad(set_to_none=set_to_none)

    def get_model(self) -> torch.nn.Module:
        return self.model

    # ------------------------------------------------------------------
    def _flat_fast_path_ok(self) -> bool:
        if self.flat is None or not _is_plain_sgd(self.optimizer):
            return False
        # optimizer must walk the parameters in flat order (true when it was built
        # from model.parameters(), which is what Simulator.run does)
        it = iter(self.flat.parameters())
        for g in self.optimizer.param_groups:
            for p in g["params"]:
                if not p.requires_grad:
                    continue
                if next(it, None) is not p:
                    return False
        return next(it, None) is None and len({g["lr"] for g in self.optimizer.param_groups}) == 1

    def current_lr(self) -> float:
        return float(self.optimizer.param_groups[0]["lr"])

    @torch.no_grad()
    def apply_update(self, update: torch.Tensor) -> None:
        """One global optimisation step with pseudo-gradient ``-update``."""
        if self._flat_fast_path_ok():
            theta = self.flat.theta
            theta.add_(update.to(theta.device, theta.dtype), alpha=self.current_lr())
            return
        self.zero_grad()
        if self.flat is not None and getattr(self.flat, "channels_last", False):
            # the flat vector is in physical (channels_last) order: map through the parameter specs
            views = {id(p): s.view(update.to(self.flat.theta.device, self.flat.theta.dtype))
                     for p, s in zip(self.flat.parameters(), self.flat.specs)}
            for group in self.optimizer.param_groups:
                for p in group["params"]:
                    if p.requires_grad and id(p) in views:
                        if p.grad is None:
                            p.grad = views[id(p)].neg()
                        else:
                            torch.neg(views[id(p)], out=p.grad)
            self.optimizer.step()
            return
        beg = 0
        for group in self.optimizer.param_groups:
            for p in group["params"]:
                if not p.requires_grad:
                    continue
                end = beg + p.numel()
                piece = update[beg:end].to(p.device, p.dtype).reshape(p.shape)
                if p.grad is None:
                    p.grad = piece.neg()
                else:
                    torch.neg(piece, out=p.grad)
                beg = end
        self.optimizer.step()
