"""ctypes launcher of the tcgen05 split-K Gram kernel (csrc/cuda/gram_tcgen05.cu)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _loader


class GramBlockDesc(C.Structure):
    _fields_ = [("base", C.c_void_p), ("ld", C.c_longlong), ("rows", C.c_int)]


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _num_sms(device) -> int:
    return torch.cuda.get_device_properties(device).multi_processor_count


def launch_gram(blocks: Sequence[Tuple[int, int, int]], d: int, col0: int, col1: int, out: torch.Tensor,
                split3: bool, device) -> List[int]:
    """blocks: (base_ptr, ld_floats, rows).  Accumulates into ``out`` ([tile_rows, ld] fp32, zeroed by the
    caller) over coordinates [col0, col1).  Returns, per block, the first padded row index."""
    lib = _loader.cuda_lib()
    arr = (GramBlockDesc * len(blocks))()
    starts, row = [], 0
    for i, (ptr, ld, rows) in enumerate(blocks):
        arr[i].base, arr[i].ld, arr[i].rows = ptr, ld, rows
        starts.append(row)
        row += _pad8(rows)
    lib.bl_gram_tcgen05.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_longlong, C.c_void_p,
                                    C.c_int, C.c_int, C.c_int, C.c_void_p]
    code = lib.bl_gram_tcgen05(C.cast(arr, C.c_void_p), len(blocks), d, col0, col1, out.data_ptr(),
                               out.stride(0), 1 if split3 else 0, _num_sms(device), _loader.stream_ptr(device))
    if code == -7:                       # 3xTF32 staging does not fit in shared memory for this row count
        return None
    _loader.check(code, "gram_tcgen05")
    _loader.count_launch()
    return starts


def split_blocks(ptr: int, ld: int, rows: int, max_rows: int = 256) -> List[Tuple[int, int, int]]:
    out, r = [], 0
    while r < rows:
        k = min(max_rows, rows - r)
        out.append((ptr + r * ld * 4, ld, k))
        r += k
    return out


def gram_tcgen05(lib, data: torch.Tensor, extra: Optional[torch.Tensor], precision: str) -> Optional[torch.Tensor]:
    r = gram_padded(lib, data, extra, precision)
    if r is None:
        return None
    out, idx = r
    idx_t = torch.tensor(idx, device=data.device)
    G = out[idx_t][:, idx_t]
    return 0.5 * (G + G.T)


def gram_padded(lib, data: torch.Tensor, extra: Optional[torch.Tensor], precision: str):
    """(padded accumulators ``[tile_rows, ld]``, logical -> padded row list) or None when the kernel does not apply."""
    n, d = data.shape
    if data.stride(1) != 1 or data.stride(0) % 4 != 0 or data.data_ptr() % 16 != 0:
        return None
    blocks = split_blocks(data.data_ptr(), data.stride(0), n)
    e = 0
    keep = [data]
    if extra is not None:
        extra = extra.reshape(-1, d)
        e = extra.shape[0]
        ld = (d + 3) // 4 * 4
        buf = torch.zeros(e, ld, device=data.device, dtype=torch.float32)
        buf[:, :d] = extra
        keep.append(buf)
        blocks += split_blocks(buf.data_ptr(), ld, e)
    total_pad = sum(_pad8(b[2]) for b in blocks)
    if total_pad > 512 or len(blocks) > 9:
        return None
    tile_rows = (total_pad + 127) // 128 * 128
    ldg = (total_pad + 31) // 32 * 32
    out = torch.zeros(tile_rows, ldg, device=data.device, dtype=torch.float32)
    starts = launch_gram(blocks, d, 0, d, out, precision == "tf32x3", data.device)
    if starts is None:
        return None
    idx = []
    for (ptr, ld, rows), s in zip(blocks, starts):
        idx += list(range(s, s + rows))
    return out, idx
