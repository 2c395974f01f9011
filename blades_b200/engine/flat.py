"""Flat parameter storage: every trainable parameter of a module becomes a view
into one contiguous fp32 vector ``theta[d]`` (and optionally ``grad[d]``).

Why: the reference walks ``optimizer.param_groups`` slicing the aggregated vector
per parameter (server.py:66-75) and concatenates all parameters twice per client
per round (client.py:216-228).  With views into one buffer both become a single
axpy / no-op, and the fused aggregation kernels can write ``theta += lr*agg``
directly.

Ordering = ``model.named_parameters()`` restricted to ``requires_grad`` -- the
same order the reference uses for client updates (client.py:219-226), so update
vectors are coordinate-compatible with the reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

__all__ = ["ParamSpec", "FlatParams", "flatten_module", "param_layout"]


@dataclass(frozen=True)
class ParamSpec:
    name: str
    shape: Tuple[int, ...]
    offset: int
    numel: int


def param_layout(model: nn.Module, align: int = 1) -> Tuple[List[ParamSpec], int]:
    """Offsets of every trainable parameter in the flat vector.

    ``align`` > 1 would pad offsets; the public update vector is always dense
    (align=1) so it stays coordinate-compatible with the reference.
    """
    specs: List[ParamSpec] = []
    off = 0
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if align > 1:
            off = (off + align - 1) // align * align
        specs.append(ParamSpec(name, tuple(p.shape), off, p.numel()))
        off += p.numel()
    return specs, off


class FlatParams:
    """Owns ``theta`` (and lazily ``grad``) and re-points module parameters at views."""

    def __init__(self, model: nn.Module, device=None, dtype=torch.float32,
                 storage: Optional[torch.Tensor] = None):
        self.model = model
        self.specs, self.numel = param_layout(model)
        first = next(model.parameters())
        self.device = torch.device(device) if device is not None else first.device
        self.dtype = dtype
        if storage is None:
            storage = torch.empty(self.numel, device=self.device, dtype=dtype)
        else:
            assert storage.numel() >= self.numel and storage.dtype == dtype
            storage = storage.view(-1)[: self.numel]
        self.theta = storage
        self.grad: Optional[torch.Tensor] = None
        self._by_name: Dict[str, nn.Parameter] = dict(model.named_parameters())
        with torch.no_grad():
            # move buffers / frozen params to the device, then alias trainables
            for name, p in self._by_name.items():
                if not p.requires_grad:
                    p.data = p.data.to(self.device)
            for buf_name, buf in model.named_buffers():
                buf.data = buf.data.to(self.device)
            for s in self.specs:
                p = self._by_name[s.name]
                view = self.theta[s.offset: s.offset + s.numel].view(s.shape)
                view.copy_(p.data.to(self.device, dtype))
                p.data = view

    # -- views -----------------------------------------------------------------
    def view_of(self, vec: torch.Tensor, spec: ParamSpec) -> torch.Tensor:
        return vec[..., spec.offset: spec.offset + spec.numel].view(*vec.shape[:-1], *spec.shape)

    def named_views(self, vec: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {s.name: self.view_of(vec, s) for s in self.specs}

    def attach_grad(self) -> torch.Tensor:
        """Allocate ``grad[d]`` and make every ``p.grad`` a view of it."""
        if self.grad is None:
            self.grad = torch.zeros_like(self.theta)
        for s in self.specs:
            self._by_name[s.name].grad = self.grad[s.offset: s.offset + s.numel].view(s.shape)
        return self.grad

    def realias(self) -> None:
        """Re-point parameters at ``theta`` (after someone replaced ``p.data``)."""
        for s in self.specs:
            p = self._by_name[s.name]
            want = self.theta[s.offset: s.offset + s.numel].view(s.shape)
            if p.data.data_ptr() != want.data_ptr():
                want.copy_(p.data)
                p.data = want

    def is_aliased(self) -> bool:
        return all(
            self._by_name[s.name].data.data_ptr() == self.theta[s.offset:].data_ptr()
            for s in self.specs)

    def parameters(self) -> List[nn.Parameter]:
        return [self._by_name[s.name] for s in self.specs]


def flatten_module(model: nn.Module, device=None) -> FlatParams:
    return FlatParams(model, device=device)
