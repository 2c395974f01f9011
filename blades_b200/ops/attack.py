"""Launchers for the on-device attacker kernels (csrc/cuda/attack.cu)."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch

from . import _loader, _structs

__all__ = ["fill_normal_", "attack_rows"]

_counter = [0]


def fill_normal_(row: torch.Tensor, mean: float, std: float, seed: int = None, offset: int = None) -> torch.Tensor:
    """In-place N(mean, std) fill with the in-kernel Philox generator.  ``offset``: position in the Philox stream
    (in 128-bit blocks); default = a process-wide running counter."""
    assert row.is_cuda and row.dtype == torch.float32 and row.is_contiguous()
    lib = _loader.cuda_lib()
    if seed is None:
        seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
    if offset is None:
        offset = _counter[0]
        _counter[0] += (row.numel() + 3) // 4
    lib.bl_fill_normal.argtypes = [C.c_void_p, C.c_longlong, C.c_float, C.c_float, C.c_ulonglong,
                                   C.c_ulonglong, C.c_void_p]
    _loader.check(lib.bl_fill_normal(row.data_ptr(), row.numel(), float(mean), float(std), seed, offset,
                                     _loader.stream_ptr(row.device)), "fill_normal")
    _loader.count_launch()
    return row


def attack_rows(honest_rows: Sequence[int], out_rows: Sequence[int], kind: str, param: float,
                c0: int, c1: int, device=None) -> None:
    """Write the ALIE / IPM malicious value into every row of ``out_rows`` for coords [c0,c1)."""
    lib = _loader.cuda_lib()
    p = _structs.AttackRowParams()
    assert len(honest_rows) <= _structs.MAX_ROWS and len(out_rows) <= _structs.MAX_ROWS
    for i, r in enumerate(honest_rows):
        p.rows[i] = r
    for i, r in enumerate(out_rows):
        p.out[i] = r
    p.n_stat, p.kind, p.param = len(honest_rows), {"alie": 1, "ipm": 2}[kind], float(param)
    p.c0, p.c1, p.n_out = c0, c1, len(out_rows)
    _loader.check(lib.bl_attack_rows(C.byref(p), _loader.stream_ptr(device)), "attack_rows")
    _loader.count_launch()
