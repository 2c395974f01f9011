"""Grouped (per-client) weight-gradient GEMM with a fused update-row epilogue:

    out[c] = alpha * a[c] @ b[c]        a: [n, M, K]   b: [n, K, N]   out: [n, M, N]

``out`` is a strided window into the shard's update matrix ``U_g[n, d]`` (batch stride
= d), so the GEMM epilogue IS the client's update write (SGD step + update
diff + save_update fused).  CUDA: tcgen05 kernel (csrc/cuda/wgrad_tcgen05.cu) when
shapes allow, else cuBLAS strided-batched through ``torch.baddbmm``.  CPU: baddbmm.
"""
from __future__ import annotations

import torch

__all__ = ["grouped_wgrad"]

import os

#: BLADES_WGRAD=cublas forces the library fallback (A/B comparisons, debugging)
_USE_KERNEL = os.environ.get("BLADES_WGRAD", "tcgen05") != "cublas"


def grouped_wgrad(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    if a.is_cuda and _USE_KERNEL:
        from . import _loader
        lib = _loader.cuda_lib(optional=True)
        if lib is not None and hasattr(lib, "bl_grouped_wgrad"):
            from ._wgrad_impl import launch
            if launch(lib, a, b, out, alpha):
                return out
    torch.baddbmm(out, a, b, beta=0.0, alpha=alpha, out=out)
    return out


#: BLADES_IMPLICIT_WGRAD=0 forces the explicit im2col + grouped GEMM path for convolutions
_USE_IMPLICIT = os.environ.get("BLADES_IMPLICIT_WGRAD", "1") != "0"


_IMPLICIT_MAX_T = int(os.environ.get("BLADES_IMPLICIT_MAX_T", "4096"))


def conv_wgrad_implicit(gy: torch.Tensor, x: torch.Tensor, out: torch.Tensor, n_clients: int, kernel, stride,
                        padding, dilation, alpha: float, force: bool = False) -> bool:
    """Per-client conv weight gradient WITHOUT materialising im2col: the tcgen05 kernel gathers its B operand
    straight from the NHWC activation with strided 4-D TMA boxes (padding = TMA out-of-bounds zero fill).

    gy: channels_last ``[NB, Cout, Ho, Wo]``; x: channels_last ``[NB, Cin, H, W]``;
    out: ``[n, Cout, kh*kw*Cin]`` window of the update matrix (physical channels_last weight order).
    Returns False when the shape is not supported (caller falls back to im2col + grouped GEMM), or -- unless
    ``force`` -- when the per-client reduction length T = B*Ho*Wo exceeds ``BLADES_IMPLICIT_MAX_T`` (default 4096).
    History of that threshold: in round 1 the 4-D gather ran the long-K / small-output layers (ResNet layer1: T = 2048,
    64 x 576 outputs) at 329 us vs 162 us + ~100 us im2col for the explicit pair and the limit was 1024; with the
    warp-uniform issue loops and running stage counters of round 2 the implicit form wins there too -- headline round
    6.67 -> 6.03 ms (150 -> 166 rounds/s, profiles/README.md) -- and saves the four im2col launches and matrices."""
    if not (_USE_KERNEL and _USE_IMPLICIT and gy.is_cuda):
        return False
    import ctypes as C
    from . import _loader
    lib = _loader.cuda_lib()
    NB, Cin, H, W = x.shape
    _, Cout, Ho, Wo = gy.shape
    kh, kw = kernel
    if tuple(dilation) != (1, 1) or stride[0] != stride[1] or padding[0] != padding[1]:
        return False
    if Cin % 32 or Cout % 4 or Wo > 32 or out.stride(2) != 1 or out.stride(1) != kh * kw * Cin:
        return False
    if not force and (NB // n_clients) * Ho * Wo > _IMPLICIT_MAX_T:
        return False
    xp, gp = x.permute(0, 2, 3, 1), gy.permute(0, 2, 3, 1)
    if not (xp.is_contiguous() and gp.is_contiguous()) or x.dtype != torch.float32 or gy.dtype != torch.float32:
        return False
    from ._wgrad_impl import _SMS
    idx = out.device.index or 0
    if idx not in _SMS:
        _SMS[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    lib.bl_conv_wgrad_implicit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 12 + \
        [C.c_longlong, C.c_float, C.c_int, C.c_void_p]
    code = lib.bl_conv_wgrad_implicit(gp.data_ptr(), xp.data_ptr(), out.data_ptr(), n_clients, NB // n_clients, H, W,
                                      Cin, Ho, Wo, Cout, kh, kw, stride[0], padding[0], out.stride(0), float(alpha),
                                      _SMS[idx], _loader.stream_ptr(out.device))
    if code in (-1, -2):
        return False
    _loader.check(code, "conv_wgrad_implicit")
    _loader.count_launch()
    return True
