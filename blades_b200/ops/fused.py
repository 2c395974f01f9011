"""Launchers for the small fused pieces of the training pass (csrc/cuda/fused_ops.cu): NHWC max / average pooling,
the per-client cross-entropy (loss + logits gradient in one launch), per-client column sums (bias gradients) and
the row-padding copy.  All of them replace ATen / cuDNN launches of the reference's autograd graph
(/root/reference/src/blades/client.py:178-193)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _loader

__all__ = ["maxpool_fwd", "maxpool_bwd", "avgpool_fwd", "avgpool_bwd", "client_ce", "group_eval", "client_colsum",
           "pad_rows"]

_typed = False


def _lib():
    global _typed
    lib = _loader.cuda_lib()
    if not _typed:
        vp, i, ll, f = C.c_void_p, C.c_int, C.c_longlong, C.c_float
        lib.bl_maxpool_nhwc_fwd.argtypes = [vp, vp, vp] + [i] * 9 + [vp]
        lib.bl_maxpool_nhwc_bwd.argtypes = [vp, vp, vp] + [i] * 9 + [vp]
        lib.bl_avgpool_nhwc_fwd.argtypes = [vp, vp, ll, i, i, vp]
        lib.bl_avgpool_nhwc_bwd.argtypes = [vp, vp, ll, i, i, vp]
        lib.bl_client_ce.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, i, vp]
        lib.bl_client_colsum.argtypes = [vp, vp, i, i, i, ll, ll, f, vp]
        lib.bl_pad_rows.argtypes = [vp, vp, ll, i, ll, i, vp]
        lib.bl_diff_rows.argtypes = [vp, vp, vp, ll, vp]
        _typed = True
    return lib


def _cl_empty(shape, device, dtype=torch.float32):
    return torch.empty(shape, device=device, dtype=dtype, memory_format=torch.channels_last)


def maxpool_fwd(x: torch.Tensor, k: int, s: int, p: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """x: channels_last ``[NB, C, H, W]`` -> (y channels_last ``[NB, C, Ho, Wo]``, argmax positions uint8, same layout)."""
    NB, Cc, H, W = x.shape
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    y = _cl_empty((NB, Cc, Ho, Wo), x.device)
    idx = _cl_empty((NB, Cc, Ho, Wo), x.device, torch.uint8)
    _loader.check(_lib().bl_maxpool_nhwc_fwd(x.data_ptr(), y.data_ptr(), idx.data_ptr(), NB, H, W, Cc, Ho, Wo, k, s, p,
                                             _loader.stream_ptr(x.device)), "maxpool_fwd")
    _loader.count_launch()
    return y, idx


def maxpool_bwd(gy: torch.Tensor, idx: torch.Tensor, in_hw: Tuple[int, int], k: int, s: int, p: int) -> torch.Tensor:
    NB, Cc, Ho, Wo = gy.shape
    H, W = in_hw
    gx = _cl_empty((NB, Cc, H, W), gy.device)
    _loader.check(_lib().bl_maxpool_nhwc_bwd(gy.data_ptr(), gx.data_ptr(), idx.data_ptr(), NB, H, W, Cc, Ho, Wo, k, s, p,
                                             _loader.stream_ptr(gy.device)), "maxpool_bwd")
    _loader.count_launch()
    return gx


def avgpool_fwd(x: torch.Tensor) -> torch.Tensor:
    """Global average pool: channels_last ``[NB, C, H, W]`` -> ``[NB, C]``."""
    NB, Cc, H, W = x.shape
    y = torch.empty(NB, Cc, device=x.device, dtype=torch.float32)
    _loader.check(_lib().bl_avgpool_nhwc_fwd(x.data_ptr(), y.data_ptr(), NB, H * W, Cc, _loader.stream_ptr(x.device)),
                  "avgpool_fwd")
    _loader.count_launch()
    return y


def avgpool_bwd(gy: torch.Tensor, hw: Tuple[int, int]) -> torch.Tensor:
    NB, Cc = gy.shape
    gx = _cl_empty((NB, Cc, hw[0], hw[1]), gy.device)
    _loader.check(_lib().bl_avgpool_nhwc_bwd(gy.data_ptr(), gx.data_ptr(), NB, hw[0] * hw[1], Cc,
                                             _loader.stream_ptr(gy.device)), "avgpool_bwd")
    _loader.count_launch()
    return gx


def client_ce(logits: torch.Tensor, target: torch.Tensor, n: int, clamp: torch.Tensor
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """logits ``[n*B, C]`` (row stride free), int64 targets ``[n*B]``, per-client clamps ``[n]`` ->
    (per-client mean CE ``[n]``, d(sum_c clamp(loss_c))/d(logits) as a ``[n*B, C]`` view of a buffer whose rows are
    padded with zeros to a multiple of 4 floats -- directly usable as a TMA source by ``ops.conv.linear_dgrad``)."""
    M, Cc = logits.shape
    assert logits.stride(1) == 1 and target.dtype == torch.int64 and target.is_contiguous() and M % n == 0
    ldg = (Cc + 3) // 4 * 4
    loss = torch.empty(n, device=logits.device, dtype=torch.float32)
    g = torch.empty(M, ldg, device=logits.device, dtype=torch.float32)
    _loader.check(_lib().bl_client_ce(logits.data_ptr(), target.data_ptr(), clamp.data_ptr(), loss.data_ptr(),
                                      g.data_ptr(), None, n, M // n, Cc, logits.stride(0), ldg,
                                      _loader.stream_ptr(logits.device)), "client_ce")
    _loader.count_launch()
    return loss, g[:, :Cc]


def group_eval(logits: torch.Tensor, target: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Evaluation twin of ``client_ce``: per-group mean cross-entropy ``[n]`` and top-1 hit counts ``[n]`` of ``n``
    equally sized groups of samples (one launch, no gradient)."""
    M, Cc = logits.shape
    assert logits.stride(1) == 1 and target.dtype == torch.int64 and target.is_contiguous() and M % n == 0
    loss = torch.empty(n, device=logits.device, dtype=torch.float32)
    hits = torch.empty(n, device=logits.device, dtype=torch.float32)
    _loader.check(_lib().bl_client_ce(logits.data_ptr(), target.data_ptr(), None, loss.data_ptr(), None,
                                      hits.data_ptr(), n, M // n, Cc, logits.stride(0), 0,
                                      _loader.stream_ptr(logits.device)), "group_eval")
    _loader.count_launch()
    return loss, hits


def client_colsum(g: torch.Tensor, n: int, out_view: torch.Tensor, alpha: float) -> None:
    """``out_view[c, j] = alpha * sum_t g[c*T + t, j]``; g ``[n*T, C]`` (row stride free), out_view ``[n, C]`` window of
    the update matrix."""
    M, Cc = g.shape
    assert g.stride(1) == 1 and out_view.stride(1) == 1
    _loader.check(_lib().bl_client_colsum(g.data_ptr(), out_view.data_ptr(), n, M // n, Cc, g.stride(0),
                                          out_view.stride(0), float(alpha), _loader.stream_ptr(g.device)),
                  "client_colsum")
    _loader.count_launch()


def pad_rows(src: torch.Tensor, ld: Optional[int] = None) -> torch.Tensor:
    """``[rows, cols]`` -> contiguous ``[rows, ld]`` (ld = cols rounded up to 4) with zero pad columns."""
    rows, cols = src.shape
    ld = (cols + 3) // 4 * 4 if ld is None else ld
    assert src.stride(1) == 1
    dst = torch.empty(rows, ld, device=src.device, dtype=torch.float32)
    _loader.check(_lib().bl_pad_rows(src.data_ptr(), dst.data_ptr(), rows, cols, src.stride(0), ld,
                                     _loader.stream_ptr(src.device)), "pad_rows")
    _loader.count_launch()
    return dst


def diff_rows(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """``out = nan_to_num(a - b)`` for flat fp32 CUDA vectors (the time-sliced client update), one own launch."""
    assert a.is_cuda and a.dtype == b.dtype == out.dtype == torch.float32
    assert a.is_contiguous() and b.is_contiguous() and out.is_contiguous() and a.numel() == b.numel() == out.numel()
    _loader.check(_lib().bl_diff_rows(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(),
                                      _loader.stream_ptr(a.device)), "diff_rows")
    _loader.count_launch()
    return out
