"""Launcher of the zero-copy sample gather (csrc/cuda/gather.cu): GPU reads pinned host datasets over PCIe."""
from __future__ import annotations

import ctypes as C

import torch

from . import _loader


class GatherParams(C.Structure):
    _fields_ = [("src_x", C.c_void_p), ("src_y", C.c_void_p), ("idx", C.c_void_p), ("dst_x", C.c_void_p),
                ("dst_y", C.c_void_p), ("per_client", C.c_int), ("sample_floats", C.c_int), ("total", C.c_longlong)]


def gather_samples(src_x_table: torch.Tensor, src_y_table: torch.Tensor, idx: torch.Tensor, dst_x: torch.Tensor,
                   dst_y: torch.Tensor, per_client: int, sample_floats: int) -> None:
    """``src_*_table``: device uint64 tensors of pinned-host base pointers; ``idx``: device int64 indices."""
    lib = _loader.cuda_lib()
    assert lib.bl_sizeof_gather_params() == C.sizeof(GatherParams)
    p = GatherParams(src_x_table.data_ptr(), src_y_table.data_ptr(), idx.data_ptr(), dst_x.data_ptr(),
                     dst_y.data_ptr(), per_client, sample_floats, idx.numel())
    _loader.check(lib.bl_gather_samples(C.byref(p), _loader.stream_ptr(dst_x.device)), "gather_samples")
    _loader.count_launch()
