"""State-dict / tensor norm helpers (reference aggregators/torch_utils.py:12-98).

``torch._six`` (removed from torch) is not used.  ``clip_tensor_norm_`` keeps the
reference behaviour of acting on -- and returning -- the first float tensor only."""
from __future__ import annotations

import math
from typing import Iterable, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

inf = math.inf
_tensor_or_tensors = Union[torch.Tensor, Iterable[torch.Tensor]]

__all__ = ["HLoss", "l2dist", "l2norm", "cos_sim", "clip_para_norm_", "clip_tensor_norm_"]


class HLoss(nn.Module):
    """Entropy of the softmax (summed over the batch)."""

    def forward(self, x):
        return -(F.softmax(x, dim=1) * F.log_softmax(x, dim=1)).sum()


def _float_items(sd):
    return [(k, v) for k, v in sd.items() if v.dtype != torch.int64]


def l2norm(model: dict) -> torch.Tensor:
    return torch.sqrt(sum(v.double().pow(2).sum() for _, v in _float_items(model))).float()


def l2dist(model1: dict, model2: dict) -> torch.Tensor:
    return torch.sqrt(sum((v.double() - model2[k].double()).pow(2).sum()
                          for k, v in _float_items(model1))).float()


def cos_sim(model1: dict, model2: dict) -> torch.Tensor:
    dot = sum((model1[k] * model2[k]).sum() for k in model1)
    return dot / torch.clamp(l2norm(model1) * l2norm(model2), min=1e-5)


def _total_norm(tensors, norm_type: float) -> torch.Tensor:
    tensors = [t.detach() for t in tensors]
    if norm_type == inf:
        return torch.stack([t.abs().max() for t in tensors]).max()
    per = torch.stack([torch.norm(t, norm_type) for t in tensors if t.dtype != torch.int64])
    return torch.norm(per, norm_type)


def _clip_coef(total_norm, max_norm, norm_type, error_if_nonfinite):
    if error_if_nonfinite and not torch.isfinite(total_norm):
        raise RuntimeError(f"The total norm of order {norm_type} is non-finite, so it cannot be clipped.")
    return torch.clamp(float(max_norm) / (total_norm + 1e-6), max=1.0)


def clip_para_norm_(parameters: dict, max_norm: float, norm_type: float = 2.0,
                    error_if_nonfinite: bool = False) -> torch.Tensor:
    """Scale every float tensor of a state-dict so the global norm is <= max_norm."""
    tensors = list(parameters.values())
    if not tensors:
        return torch.tensor(0.0)
    total = _total_norm(tensors, float(norm_type))
    coef = _clip_coef(total, max_norm, norm_type, error_if_nonfinite)
    for t in tensors:
        if t.dtype != torch.int64:
            t.detach().mul_(coef.to(t.device))
    return total


def clip_tensor_norm_(parameters: _tensor_or_tensors, max_norm: float, norm_type: float = 2.0,
                      error_if_nonfinite: bool = False) -> torch.Tensor:
    """Clip in place and return the first float tensor (see module docstring)."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    parameters = list(parameters)
    if not parameters:
        return torch.tensor(0.0)
    total = _total_norm(parameters, float(norm_type))
    coef = _clip_coef(total, max_norm, norm_type, error_if_nonfinite)
    for t in parameters:
        if t.dtype != torch.int64:
            return t.detach().mul_(coef.to(t.device))
