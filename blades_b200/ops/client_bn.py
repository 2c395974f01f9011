"""Launchers for the fused per-client BatchNorm kernels (csrc/cuda/client_bn.cu)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from . import _loader


class ClientBNParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("gy", C.c_void_p), ("y", C.c_void_p), ("gamma", C.c_void_p),
                ("beta", C.c_void_p), ("mean", C.c_void_p), ("rstd", C.c_void_p), ("dgamma", C.c_void_p),
                ("dbeta", C.c_void_p), ("ld", C.c_longlong), ("n", C.c_int), ("B", C.c_int), ("C", C.c_int),
                ("HW", C.c_int), ("eps", C.c_float), ("alpha", C.c_float),
                ("res", C.c_void_p), ("act", C.c_void_p), ("gmask", C.c_void_p), ("relu", C.c_int), ("pad_", C.c_int)]


_checked = False


def supported(x: torch.Tensor) -> bool:
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous()):
        return False
    hw = x.shape[2] * x.shape[3]
    return 1 <= hw <= 256 and (hw & (hw - 1)) == 0


def is_nhwc(x: torch.Tensor) -> bool:
    """channels_last-contiguous 4-D fp32 CUDA tensor (C > 1 or trivially both layouts)."""
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] % 4 == 0
            and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous())


def _same_layout(a: torch.Tensor, x: torch.Tensor) -> bool:
    """Same shape and both NHWC-dense (strides of size-1 dimensions are arbitrary in torch, so compare contiguity)."""
    cl = torch.channels_last
    return a.shape == x.shape and a.dtype == x.dtype and a.is_contiguous(memory_format=cl) \
        and x.is_contiguous(memory_format=cl)


#: BLADES_BN_REMASK=0: always read the forward output for the ReLU mask
_REMASK = os.environ.get("BLADES_BN_REMASK", "1") != "0"


def _lib():
    global _checked
    lib = _loader.cuda_lib()
    if not _checked:
        assert lib.bl_sizeof_bn_params() == C.sizeof(ClientBNParams)
        _checked = True
    return lib


def forward(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, n: int, eps: float,
            res: Optional[torch.Tensor] = None, relu: bool = False, nhwc: Optional[bool] = None
            ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``y = relu?(bn_per_client(x) + res)``; the fused residual / ReLU exist in the NHWC kernels only."""
    NB, Cc, H, W = x.shape
    nhwc = is_nhwc(x) if nhwc is None else nhwc
    assert nhwc or (res is None and not relu)
    if nhwc:                                # float4 parameter loads need 16 B alignment
        gamma = gamma if gamma.data_ptr() % 16 == 0 else gamma.clone()
        beta = beta if beta.data_ptr() % 16 == 0 else beta.clone()
    y = torch.empty_like(x)                 # preserves the memory format
    mean = torch.empty(n, Cc, device=x.device, dtype=torch.float32)
    rstd = torch.empty(n, Cc, device=x.device, dtype=torch.float32)
    p = ClientBNParams(x.data_ptr(), None, y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(),
                       rstd.data_ptr(), None, None, 0, n, NB // n, Cc, H * W, float(eps), 1.0)
    if res is not None:
        assert _same_layout(res, x)
        p.res = res.data_ptr()
    p.relu = 1 if relu else 0
    fn = _lib().bl_client_bn_nhwc_fwd if nhwc else _lib().bl_client_bn_fwd
    _loader.check(fn(C.byref(p), _loader.stream_ptr(x.device)), "client_bn_fwd")
    _loader.count_launch()
    return y, mean, rstd


def backward(gy: torch.Tensor, x: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, gamma: torch.Tensor,
             n: int, dgamma_view: torch.Tensor, dbeta_view: torch.Tensor, alpha: float, need_dx: bool,
             act: Optional[torch.Tensor] = None, gmask: Optional[torch.Tensor] = None, nhwc: Optional[bool] = None,
             beta: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """``dgamma_view`` / ``dbeta_view``: strided ``[n, C]`` windows of the update matrix (row stride ld).
    ``act`` (NHWC only): the forward OUTPUT of a fused ReLU -- the incoming gradient is masked with ``act > 0``
    first; ``gmask`` (may be ``gy`` itself) receives that masked gradient (what a residual branch needs).
    ``beta`` (NHWC, with ``act``): the unit had NO residual input, so the kernel recomputes the pre-activation from
    ``x`` (bit-identical to the forward pass) for the mask instead of reading ``act`` -- one tensor less to stream."""
    NB, Cc, H, W = x.shape
    nhwc = is_nhwc(x) if nhwc is None else nhwc
    assert nhwc or (act is None and gmask is None)
    if nhwc and gamma.data_ptr() % 16:
        gamma = gamma.clone()
    gy = gy.contiguous(memory_format=torch.channels_last) if nhwc else gy.contiguous()
    dx = torch.empty_like(x) if need_dx else None
    assert dgamma_view.stride(1) == 1 and dgamma_view.stride(0) == dbeta_view.stride(0)
    p = ClientBNParams(x.data_ptr(), gy.data_ptr(), dx.data_ptr() if need_dx else None, gamma.data_ptr(), None,
                       mean.data_ptr(), rstd.data_ptr(), dgamma_view.data_ptr(), dbeta_view.data_ptr(),
                       dgamma_view.stride(0), n, NB // n, Cc, H * W, 0.0, float(alpha))
    if act is not None:
        assert _same_layout(act, x)
        p.act, p.relu = act.data_ptr(), 1
        if beta is not None and nhwc and _REMASK:
            beta = beta if beta.data_ptr() % 16 == 0 else beta.clone()
            p.beta, p.act = beta.data_ptr(), None
    if gmask is not None:
        assert _same_layout(gmask, x)
        p.gmask = gmask.data_ptr()
    fn = _lib().bl_client_bn_nhwc_bwd if nhwc else _lib().bl_client_bn_bwd
    _loader.check(fn(C.byref(p), _loader.stream_ptr(x.device)), "client_bn_bwd")
    _loader.count_launch()
    return dx
