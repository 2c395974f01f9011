"""Launcher for the weighted row-combine kernel (csrc/cuda/row_combine.cu)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _loader, _structs
from .select import make_epilogue, row_pointers

__all__ = ["row_combine", "launch_combine"]

_SMS = {}


def _num_sms(device) -> int:
    idx = torch.device(device).index or 0
    if idx not in _SMS:
        _SMS[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SMS[idx]


def launch_combine(rows: Sequence[int], weights, c0: int, c1: int,
                   ep: _structs.Epilogue, device=None) -> None:
    """``weights``: a sequence of floats, or a CUDA fp32 tensor with one weight per row (written by the on-device Gram
    solvers): then the kernel reads the weights from device memory and skips zero-weight rows itself."""
    lib = _loader.cuda_lib()
    if torch.is_tensor(weights):
        assert weights.is_cuda and weights.dtype == torch.float32 and weights.is_contiguous()
        assert weights.numel() == len(rows) <= _structs.MAX_ROWS + 1
        p = _structs.CombineParams()
        for i, r in enumerate(rows):
            p.rows[i] = r
        p.n_rows, p.c0, p.c1, p.ep, p.w_dev = len(rows), c0, c1, ep, weights.data_ptr()
        _loader.check(lib.bl_row_combine(C.byref(p), _num_sms(device), _loader.stream_ptr(device)), "row_combine")
        _loader.count_launch()
        return
    keep = [(r, float(w)) for r, w in zip(rows, weights) if float(w) != 0.0]
    if not keep:                      # all-zero weights: still must produce zeros
        keep = [(rows[0], 0.0)]
    assert len(keep) <= _structs.MAX_ROWS + 1
    p = _structs.CombineParams()
    for i, (r, w) in enumerate(keep):
        p.rows[i] = r
        p.w[i] = w
    p.n_rows, p.c0, p.c1, p.ep = len(keep), c0, c1, ep
    _loader.check(lib.bl_row_combine(C.byref(p), _num_sms(device), _loader.stream_ptr(device)), "row_combine")
    _loader.count_launch()


def row_combine(data: torch.Tensor, weights, extra: Optional[torch.Tensor] = None,
                extra_weight: float = 0.0) -> torch.Tensor:
    """``sum_i w_i data[i]`` (+ ``extra_weight * extra``) on the device of ``data``."""
    assert data.is_cuda and data.dtype == torch.float32 and data.stride(1) == 1
    n, d = data.shape
    w = [float(x) for x in (weights.tolist() if hasattr(weights, "tolist") else weights)]
    assert len(w) == n
    rows = row_pointers(data)
    if extra is not None and extra_weight != 0.0:
        extra = extra.contiguous()
        rows = rows + [extra.data_ptr()]
        w = w + [float(extra_weight)]
    out = torch.empty(d, device=data.device, dtype=torch.float32)
    launch_combine(rows, w, 0, d, make_epilogue([out.data_ptr()]), data.device)
    return out
