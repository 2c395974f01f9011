"""Launcher of the in-fabric all-reduce (csrc/cuda/nvls.cu): two-shot sum of a symmetric buffer over all ranks,
``multimem.ld_reduce`` + ``multimem.st`` through the NVSwitch when a multicast address is given, peer loads/stores
otherwise."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

from . import _loader, _structs

__all__ = ["allreduce", "copy2d", "zero_"]


class NvlsReduceParams(C.Structure):
    _fields_ = [("peers", C.c_void_p * _structs.MAX_PEERS), ("mc", C.c_void_p), ("count", C.c_longlong),
                ("rank", C.c_int), ("world", C.c_int)]


def allreduce(peer_ptrs: Sequence[int], mc_ptr: int, count: int, rank: int, world: int, device=None) -> None:
    lib = _loader.cuda_lib()
    assert lib.bl_sizeof_nvls_reduce_params() == C.sizeof(NvlsReduceParams)
    assert len(peer_ptrs) == world <= _structs.MAX_PEERS and count % 4 == 0
    p = NvlsReduceParams()
    for i, b in enumerate(peer_ptrs):
        p.peers[i] = b
    p.mc = mc_ptr or None
    p.count, p.rank, p.world = count, rank, world
    _loader.check(lib.bl_nvls_allreduce(C.byref(p), _loader.stream_ptr(device)), "nvls_allreduce")
    _loader.count_launch()


def copy2d(dst_ptr: int, dst_pitch: int, src_ptr: int, src_pitch: int, width_bytes: int, height: int, device=None) -> None:
    """``height`` rows of ``width_bytes`` from ``src`` to ``dst`` (byte pitches) on the copy engines, stream ordered;
    either side may be NVLink peer memory (symmetric-memory mapping)."""
    lib = _loader.cuda_lib()
    lib.bl_copy2d_async.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong,
                                    C.c_void_p]
    _loader.check(lib.bl_copy2d_async(dst_ptr, dst_pitch, src_ptr, src_pitch, width_bytes, height,
                                      _loader.stream_ptr(device)), "copy2d_async")


def zero_(t: "torch.Tensor") -> "torch.Tensor":
    """Zero a contiguous CUDA tensor with ``cudaMemsetAsync`` (stream ordered; no fill kernel)."""
    assert t.is_cuda and t.is_contiguous()
    lib = _loader.cuda_lib()
    lib.bl_memset_zero_async.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p]
    _loader.check(lib.bl_memset_zero_async(t.data_ptr(), t.numel() * t.element_size(), _loader.stream_ptr(t.device)),
                  "memset_zero")
    return t
