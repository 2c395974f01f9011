"""Row-major im2col launcher (csrc/cuda/im2col.cu); CPU fallback via F.unfold for the oracle."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from . import _loader

__all__ = ["im2col_rows", "im2col_nhwc"]


def im2col_rows(x: torch.Tensor, kernel, stride, padding, dilation, out_hw) -> torch.Tensor:
    """x [NB, Cin, H, W] -> [NB*Ho*Wo, Cin*kh*kw] (row = (b, ho, wo), col = (cin, r, s))."""
    NB, Cin, H, W = x.shape
    kh, kw = kernel
    Ho, Wo = out_hw
    if not x.is_cuda or x.dtype != torch.float32:
        cols = F.unfold(x, kernel, dilation=dilation, padding=padding, stride=stride)      # [NB, K, L]
        return cols.transpose(1, 2).reshape(NB * Ho * Wo, Cin * kh * kw)
    x = x.contiguous()
    out = torch.empty(NB * Ho * Wo, Cin * kh * kw, device=x.device, dtype=torch.float32)
    lib = _loader.cuda_lib()
    lib.bl_im2col_rows.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 14 + [C.c_void_p]
    _loader.check(lib.bl_im2col_rows(x.data_ptr(), out.data_ptr(), NB, Cin, H, W, kh, kw, stride[0], stride[1],
                                     padding[0], padding[1], dilation[0], dilation[1], Ho, Wo,
                                     _loader.stream_ptr(x.device)), "im2col_rows")
    _loader.count_launch()
    return out


def im2col_nhwc(x: torch.Tensor, kernel, stride, padding, dilation, out_hw) -> torch.Tensor:
    """x: channels_last ``[NB, Cin, H, W]`` -> ``[NB*Ho*Wo, K]`` view of a ``[rows, ldk]`` buffer with
    ``K = kh*kw*Cin`` ordered (r, s, cin) (= a channels_last conv weight) and ``ldk = K`` rounded up to 4."""
    NB, Cin, H, W = x.shape
    kh, kw = kernel
    Ho, Wo = out_hw
    K = Cin * kh * kw
    if not x.is_cuda or x.dtype != torch.float32:
        cols = F.unfold(x, kernel, dilation=dilation, padding=padding, stride=stride)      # [NB, Cin*kh*kw, L]
        return cols.view(NB, Cin, kh * kw, Ho * Wo).permute(0, 3, 2, 1).reshape(NB * Ho * Wo, K)
    ldk = (K + 3) // 4 * 4
    out = torch.empty(NB * Ho * Wo, ldk, device=x.device, dtype=torch.float32)
    lib = _loader.cuda_lib()
    lib.bl_im2col_nhwc.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 15 + [C.c_longlong] * 4 + [C.c_void_p]
    # any strided [NB, Cin, H, W] input: a channels_last tensor moves float4 channel runs, an NCHW one (the network
    # input) is read element-wise -- no layout conversion pass either way
    _loader.check(lib.bl_im2col_nhwc(x.data_ptr(), out.data_ptr(), NB, Cin, H, W, kh, kw, stride[0], stride[1],
                                     padding[0], padding[1], dilation[0], dilation[1], Ho, Wo, ldk,
                                     x.stride(0), x.stride(2), x.stride(3), x.stride(1),
                                     _loader.stream_ptr(x.device)), "im2col_nhwc")
    _loader.count_launch()
    return out[:, :K]
