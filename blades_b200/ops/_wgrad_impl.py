"""ctypes launcher of the tcgen05 grouped weight-gradient kernel (csrc/cuda/wgrad_tcgen05.cu)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _loader

_SMS = {}
MIN_OUTPUT_ELEMS = 4096          # tiny layers (biases, fc heads) stay on cuBLAS


def launch(lib, a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, alpha: float) -> bool:
    """a: [n, M, T] given as a transposed view of a contiguous [n, T, M]; b: [n, T, N] contiguous;
    out: [n, M, N] with out.stride() == (batch_stride, N, 1).  Returns False if the shape is unsupported."""
    n, M, T = a.shape
    N = b.shape[2]
    if a.dtype != torch.float32 or b.dtype != torch.float32 or out.dtype != torch.float32:
        return False
    a_t = a.transpose(1, 2)                 # [n, T, M] view; rows may be padded (row stride lda >= M)
    lda, ldb = a_t.stride(1), b.stride(1)
    if a_t.stride(2) != 1 or b.stride(2) != 1 or a_t.stride(0) != T * lda or b.stride(0) != T * ldb:
        return False
    if out.stride(2) != 1 or out.stride(1) != N or M * N < MIN_OUTPUT_ELEMS:
        return False
    if lda % 4 or ldb % 4 or T % 8 or a_t.data_ptr() % 16 or b.data_ptr() % 16:
        return False
    idx = out.device.index or 0
    if idx not in _SMS:
        _SMS[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    lib.bl_grouped_wgrad.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_longlong, C.c_longlong, C.c_longlong, C.c_float, C.c_int, C.c_void_p]
    code = lib.bl_grouped_wgrad(a_t.data_ptr(), b.data_ptr(), out.data_ptr(), n, T, M, N, lda, ldb, out.stride(0),
                                float(alpha), _SMS[idx], _loader.stream_ptr(out.device))
    if code == -1 or code == -2:
        return False
    _loader.check(code, "grouped_wgrad")
    _loader.count_launch()
    return True
