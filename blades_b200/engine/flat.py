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
    #: True -> a 4-D (conv) weight stored physically as [Cout, kh, kw, Cin] (channels_last): the layout the
    #: NHWC tensor-core convolutions and the per-client wgrad GEMM ([Cout] x [kh*kw*Cin]) produce natively.
    channels_last: bool = False

    def view(self, flat: torch.Tensor) -> torch.Tensor:
        """Logical-shape view of this parameter inside a flat vector (last dim = d)."""
        lead = flat.shape[:-1]
        seg = flat[..., self.offset: self.offset + self.numel]
        if self.channels_last:
            co, ci, kh, kw = self.shape
            nl = len(lead)
            return seg.view(*lead, co, kh, kw, ci).permute(*range(nl), nl, nl + 3, nl + 1, nl + 2)
        return seg.view(*lead, *self.shape)


def param_layout(model: nn.Module, align: int = 1, channels_last: bool = False) -> Tuple[List[ParamSpec], int]:
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
        cl = channels_last and p.dim() == 4 and (p.shape[2] > 1 or p.shape[3] > 1)
        specs.append(ParamSpec(name, tuple(p.shape), off, p.numel(), cl))
        off += p.numel()
    return specs, off


class FlatParams:
    """Owns ``theta`` (and lazily ``grad``) and re-points module parameters at views."""

    def __init__(self, model: nn.Module, device=None, dtype=torch.float32,
                 storage: Optional[torch.Tensor] = None, channels_last: Optional[bool] = None):
        self.model = model
        first = next(model.parameters())
        self.device = torch.device(device) if device is not None else first.device
        # conv weights are stored channels_last on the GPU (flat coordinate order = physical order; use
        # to_reference_order() for the reference's named_parameters() flattening)
        self.channels_last = (self.device.type == "cuda") if channels_last is None else channels_last
        self.specs, self.numel = param_layout(model, channels_last=self.channels_last)
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
            if self.device.type == "cuda":
                # lay the vector out on the HOST (the channels_last permutation of the conv weights is a strided CPU
                # copy) and upload it with ONE memcpy: no per-parameter strided-copy kernels on the device
                host = torch.empty(self.numel, dtype=dtype)
                for s in self.specs:
                    s.view(host).copy_(self._by_name[s.name].data.detach().to("cpu", dtype))
                self.theta.copy_(host)
                for s in self.specs:
                    self._by_name[s.name].data = s.view(self.theta)
            else:
                for s in self.specs:
                    p = self._by_name[s.name]
                    view = s.view(self.theta)
                    view.copy_(p.data.to(self.device, dtype))
                    p.data = view

    # -- views -----------------------------------------------------------------
    def view_of(self, vec: torch.Tensor, spec: ParamSpec) -> torch.Tensor:
        return spec.view(vec)

    def to_reference_order(self, vec: torch.Tensor) -> torch.Tensor:
        """Flat vector(s) in this layout -> the reference's coordinate order (each parameter flattened
        contiguously in ``named_parameters()`` order, client.py:216-228).  Identity unless channels_last."""
        if not self.channels_last:
            return vec
        out = torch.empty_like(vec)
        for s in self.specs:
            out[..., s.offset: s.offset + s.numel] = s.view(vec).reshape(*vec.shape[:-1], s.numel)
        return out

    def from_reference_order(self, vec: torch.Tensor) -> torch.Tensor:
        if not self.channels_last:
            return vec
        out = torch.empty_like(vec)
        for s in self.specs:
            s.view(out).copy_(vec[..., s.offset: s.offset + s.numel].view(*vec.shape[:-1], *s.shape))
        return out

    def named_views(self, vec: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {s.name: self.view_of(vec, s) for s in self.specs}

    def attach_grad(self) -> torch.Tensor:
        """Allocate ``grad[d]`` and make every ``p.grad`` a view of it."""
        if self.grad is None:
            self.grad = torch.zeros_like(self.theta)
        for s in self.specs:
            self._by_name[s.name].grad = s.view(self.grad)
        return self.grad

    def realias(self) -> None:
        """Re-point parameters at ``theta`` (after someone replaced ``p.data``)."""
        for s in self.specs:
            p = self._by_name[s.name]
            want = s.view(self.theta)
            if p.data.data_ptr() != want.data_ptr() or p.data.stride() != want.stride():
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
