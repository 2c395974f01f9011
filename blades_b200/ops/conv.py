"""Implicit-GEMM convolution / linear forward and input-gradient on the tcgen05 tensor cores
(csrc/cuda/conv_tcgen05.cu) -- the shared-weight half of the client-batched training pass (the per-client half is
``ops.wgrad``).  Replaces cuDNN / cuBLAS at the reference's ``model(data)`` / ``loss.backward()`` call sites
(/root/reference/src/blades/client.py:178-193).

Everything here works on NHWC ("channels_last") activations ``[NB, H, W, C]`` and the channels_last weight matrix
``[Cout, kh*kw*Cin]`` (exactly the physical layout of the flat parameter vector on the GPU, ``engine/flat.py``).

The host side is a *planner*: a convolution (or its input gradient) becomes a list of phases, each with a tap list
(source offset + weight tap index); the kernel only ever sees that description (``ConvDesc``).  ``emulate`` executes
the same description with plain PyTorch ops -- the CPU tests check every plan against ``F.conv2d`` / autograd, the
GPU tests check the kernel against both.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _loader

__all__ = ["plan_fprop", "plan_dgrad", "emulate", "conv_fprop", "conv_dgrad", "linear_fprop", "linear_dgrad",
           "supported_conv", "ENABLED"]

MAX_TAPS = 25
MAX_PHASES = 4

#: BLADES_CONV=cudnn routes convolutions / linears back to the library kernels (A/B comparisons, debugging)
ENABLED = os.environ.get("BLADES_CONV", "tcgen05") != "cudnn"


class ConvPhase(C.Structure):
    _fields_ = [("ntaps", C.c_int), ("oh_off", C.c_int), ("ow_off", C.c_int), ("pad_", C.c_int),
                ("dy", C.c_short * MAX_TAPS), ("dx", C.c_short * MAX_TAPS), ("widx", C.c_short * MAX_TAPS),
                ("pad2_", C.c_short)]


class ConvDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("w", C.c_void_p), ("out", C.c_void_p), ("add", C.c_void_p),
                ("bias", C.c_void_p),
                ("NB", C.c_int), ("Hs", C.c_int), ("Ws", C.c_int), ("Cs", C.c_int), ("lds", C.c_int),
                ("w_rows", C.c_int), ("w_cols", C.c_int), ("ldw", C.c_int),
                ("mode", C.c_int), ("N", C.c_int), ("wtap_stride", C.c_int),
                ("Hout", C.c_int), ("Wout", C.c_int), ("ldc", C.c_int),
                ("Ht", C.c_int), ("Wt", C.c_int), ("cs", C.c_int), ("ostep", C.c_int), ("n_phases", C.c_int),
                ("accumulate_only", C.c_int), ("num_sms", C.c_int),
                ("ph", ConvPhase * MAX_PHASES)]


class Plan:
    """Geometry of one launch: what the kernel needs besides the pointers."""

    def __init__(self, mode, Ht, Wt, cs, ostep, Hout, Wout, phases):
        self.mode, self.Ht, self.Wt, self.cs, self.ostep = mode, Ht, Wt, cs, ostep
        self.Hout, self.Wout = Hout, Wout
        self.phases: List[Tuple[int, int, List[Tuple[int, int, int]]]] = phases     # (oh_off, ow_off, [(dy, dx, widx)])

    def max_taps(self) -> int:
        return max((len(t) for _, _, t in self.phases), default=0)


def _live(off: int, n_grid: int, step: int, n_src: int) -> bool:
    """Does any grid index g in [0, n_grid) read an in-range source coordinate ``g*step + off``?"""
    return any(0 <= g * step + off < n_src for g in range(n_grid))


def plan_fprop(H: int, W: int, kh: int, kw: int, stride: int, pad: int) -> Plan:
    Ho = (H + 2 * pad - kh) // stride + 1
    Wo = (W + 2 * pad - kw) // stride + 1
    taps = []
    for r in range(kh):
        for s in range(kw):
            dy, dx = r - pad, s - pad
            if _live(dy, Ho, stride, H) and _live(dx, Wo, stride, W):       # drop taps that only ever see padding
                taps.append((dy, dx, r * kw + s))
    return Plan(0, Ho, Wo, stride, 1, Ho, Wo, [(0, 0, taps)])


def plan_dgrad(H: int, W: int, kh: int, kw: int, stride: int, pad: int) -> Plan:
    """Input gradient of a conv with input ``H x W``: a gather over the output gradient ``Ho x Wo`` per input-pixel
    parity class (ph, pw) = (h % stride, w % stride)."""
    Ho = (H + 2 * pad - kh) // stride + 1
    Wo = (W + 2 * pad - kw) // stride + 1
    Ht, Wt = -(-H // stride), -(-W // stride)
    phases = []
    for ph in range(stride):
        for pw in range(stride):
            taps = []
            for r in range(kh):
                if (ph + pad - r) % stride:
                    continue
                dy = (ph + pad - r) // stride
                if not any(stride * g + ph < H and 0 <= g + dy < Ho for g in range(Ht)):
                    continue
                for s in range(kw):
                    if (pw + pad - s) % stride:
                        continue
                    dx = (pw + pad - s) // stride
                    if not any(stride * g + pw < W and 0 <= g + dx < Wo for g in range(Wt)):
                        continue
                    taps.append((dy, dx, r * kw + s))
            phases.append((ph, pw, taps))
    return Plan(1, Ht, Wt, 1, stride, H, W, phases)


def emulate(plan: Plan, src: torch.Tensor, w2d: torch.Tensor, N: int, wtap: int, add: Optional[torch.Tensor] = None,
            bias: Optional[torch.Tensor] = None, accumulate_only: bool = False,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Execute ``plan`` with plain torch ops.  src: [NB, Hs, Ws, Cs]; w2d: [rows, cols]; returns [NB, Hout, Wout, N]."""
    NB, Hs, Ws, Cs = src.shape
    res = torch.zeros(NB, plan.Hout, plan.Wout, N, dtype=src.dtype, device=src.device) if out is None else out
    gh = torch.arange(plan.Ht, device=src.device)
    gw = torch.arange(plan.Wt, device=src.device)
    for oh_off, ow_off, taps in plan.phases:
        if not taps and accumulate_only:
            continue
        acc = torch.zeros(NB, plan.Ht, plan.Wt, N, dtype=src.dtype, device=src.device)
        for dy, dx, widx in taps:
            hs, ws = gh * plan.cs + dy, gw * plan.cs + dx
            okh, okw = (hs >= 0) & (hs < Hs), (ws >= 0) & (ws < Ws)
            g = src[:, hs.clamp(0, Hs - 1)][:, :, ws.clamp(0, Ws - 1)]
            g = g * (okh[:, None] & okw[None, :]).to(src.dtype)[None, :, :, None]
            if plan.mode == 0:
                wt = w2d[:N, widx * wtap: widx * wtap + Cs]                 # [N, Cs]
                acc += g @ wt.t()
            else:
                wt = w2d[:Cs, widx * wtap: widx * wtap + N]                 # [Cs (K), N]
                acc += g @ wt
        oh, ow = gh * plan.ostep + oh_off, gw * plan.ostep + ow_off
        okh, okw = oh < plan.Hout, ow < plan.Wout
        tgt_h, tgt_w = oh[okh], ow[okw]
        val = acc[:, okh][:, :, okw]
        if bias is not None:
            val = val + bias
        if add is not None:
            val = val + add[:, tgt_h][:, :, tgt_w]
        res[:, tgt_h[:, None], tgt_w[None, :]] = val
    return res


# ------------------------------------------------------------------------------------------------ launch
_SMS = {}
_checked = False


def _lib():
    global _checked
    lib = _loader.cuda_lib()
    if not _checked:
        assert lib.bl_sizeof_conv_desc() == C.sizeof(ConvDesc), (lib.bl_sizeof_conv_desc(), C.sizeof(ConvDesc))
        lib.bl_conv_tc.argtypes = [C.c_void_p, C.c_void_p]
        _checked = True
    return lib


def _num_sms(device) -> int:
    idx = device.index or 0
    if idx not in _SMS:
        _SMS[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SMS[idx]


def tile_rows(Wt: int, Ht: int, NB: int, cs: int) -> int:
    """Pixels per tile the kernel would use (0 = unsupported grid).  Pure arithmetic twin of ``bl_conv_tile_box``."""
    if Wt < 1 or Ht < 1 or NB < 1 or Wt > 128 or Wt * cs > 256:
        return 0
    h = min(128 // Wt, Ht)
    while h * cs > 256:
        h -= 1
    b = 1
    if h == Ht:
        b = min(max(128 // (Wt * Ht), 1), 256)
    return Wt * h * b


def _launch(plan: Plan, src: torch.Tensor, w2d: torch.Tensor, out: torch.Tensor, N: int, wtap: int,
            add: Optional[torch.Tensor], bias: Optional[torch.Tensor], accumulate_only: bool) -> bool:
    """src / out / add: NHWC-contiguous 4-D tensors ``[NB, H, W, C]`` (C may be the padded row length).  The pixel
    pitch is the last dimension's size (strides of size-1 dimensions are arbitrary in torch)."""
    NB, Hs, Ws, Cs = src.shape
    if plan.max_taps() > MAX_TAPS or len(plan.phases) > MAX_PHASES:
        return False
    if tile_rows(plan.Wt, plan.Ht, NB, plan.cs) == 0:
        return False
    lds, ldw, ldc = src.shape[3], w2d.stride(0), out.shape[3]
    if lds % 4 or ldw % 4 or src.data_ptr() % 16 or w2d.data_ptr() % 16 or src.stride(3) != 1 or w2d.stride(1) != 1:
        return False
    if plan.mode == 0 and Cs % 32 and plan.max_taps() > 1:
        return False
    d = ConvDesc()
    d.src, d.w, d.out = src.data_ptr(), w2d.data_ptr(), out.data_ptr()
    d.add = add.data_ptr() if add is not None else None
    d.bias = bias.data_ptr() if bias is not None else None
    d.NB, d.Hs, d.Ws, d.Cs, d.lds = NB, Hs, Ws, Cs, lds
    d.w_rows, d.w_cols, d.ldw = w2d.shape[0], w2d.shape[1], ldw
    d.mode, d.N, d.wtap_stride = plan.mode, N, wtap
    d.Hout, d.Wout, d.ldc = plan.Hout, plan.Wout, ldc
    d.Ht, d.Wt, d.cs, d.ostep = plan.Ht, plan.Wt, plan.cs, plan.ostep
    d.n_phases = len(plan.phases)
    d.accumulate_only = 1 if accumulate_only else 0
    d.num_sms = _num_sms(src.device)
    for i, (oh_off, ow_off, taps) in enumerate(plan.phases):
        ph = d.ph[i]
        ph.ntaps, ph.oh_off, ph.ow_off = len(taps), oh_off, ow_off
        for j, (dy, dx, widx) in enumerate(taps):
            ph.dy[j], ph.dx[j], ph.widx[j] = dy, dx, widx
    code = _lib().bl_conv_tc(C.byref(d), _loader.stream_ptr(src.device))
    if code in (-1, -2):
        return False
    _loader.check(code, "conv_tc")
    _loader.count_launch()
    return True


def _nhwc(x: torch.Tensor) -> Optional[torch.Tensor]:
    """[NB, H, W, C] view of a channels_last 4-D tensor (None if the memory is not NHWC)."""
    v = x.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else None


def supported_conv(x: torch.Tensor, weight: torch.Tensor, stride, padding, dilation) -> bool:
    """Can ``conv_fprop`` / ``conv_dgrad`` take this layer?  (x: channels_last [NB, Cin, H, W])"""
    if not (ENABLED and x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32):
        return False
    if tuple(dilation) != (1, 1) or stride[0] != stride[1] or padding[0] != padding[1]:
        return False
    Cout, Cin, kh, kw = weight.shape
    NB, _, H, W = x.shape
    if Cin % 32 or Cout % 4 or kh * kw > MAX_TAPS or stride[0] * stride[0] > MAX_PHASES:
        return False
    pf = plan_fprop(H, W, kh, kw, stride[0], padding[0])
    pd = plan_dgrad(H, W, kh, kw, stride[0], padding[0])
    return tile_rows(pf.Wt, pf.Ht, NB, pf.cs) > 0 and tile_rows(pd.Wt, pd.Ht, NB, 1) > 0


def conv_fprop(x: torch.Tensor, w2d: torch.Tensor, kernel: Sequence[int], stride: int, pad: int,
               bias: Optional[torch.Tensor] = None, add: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """x: channels_last ``[NB, Cin, H, W]``; w2d: ``[Cout, kh*kw*Cin]`` (K ordered (r, s, cin)).
    Returns channels_last ``[NB, Cout, Ho, Wo]`` or None when the shape is not supported."""
    xs = _nhwc(x)
    if xs is None:
        return None
    NB, H, W, Cin = xs.shape
    kh, kw = kernel
    plan = plan_fprop(H, W, kh, kw, stride, pad)
    Cout = w2d.shape[0]
    y = torch.empty((NB, Cout, plan.Hout, plan.Wout), device=x.device, dtype=torch.float32,
                    memory_format=torch.channels_last)
    a = _nhwc(add) if add is not None else None
    if not _launch(plan, xs, w2d, y.permute(0, 2, 3, 1), Cout, Cin, a, bias, False):
        return None
    return y


def conv_dgrad(gy: torch.Tensor, w2d: torch.Tensor, kernel: Sequence[int], stride: int, pad: int,
               in_hw: Tuple[int, int], Cin: int, add: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """gy: channels_last ``[NB, Cout, Ho, Wo]`` -> gx channels_last ``[NB, Cin, H, W]`` (``+ add``).
    With ``out`` given and ``add is out`` the result is accumulated in place and pixel classes that no tap reaches
    (1x1 stride-2 convolutions) are left untouched."""
    gs = _nhwc(gy)
    if gs is None:
        return None
    NB = gs.shape[0]
    H, W = in_hw
    kh, kw = kernel
    plan = plan_dgrad(H, W, kh, kw, stride, pad)
    if out is None:
        out = torch.empty((NB, Cin, H, W), device=gy.device, dtype=torch.float32, memory_format=torch.channels_last)
    a = _nhwc(add) if add is not None else None
    acc_only = add is not None and add.data_ptr() == out.data_ptr()
    if not _launch(plan, gs, w2d, out.permute(0, 2, 3, 1), Cin, Cin, a, None, acc_only):
        return None
    return out


def linear_fprop(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """x: ``[M, K]`` row-major, weight ``[N, K]`` -> ``[M, N]`` (a 1x1 convolution over 1x1 images).  The result is a
    view of a buffer whose rows are padded to a multiple of 4 floats."""
    if not (ENABLED and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
            and weight.is_contiguous() and x.shape[1] % 4 == 0):
        return None
    M, K = x.shape
    N = weight.shape[0]
    ldc = (N + 3) // 4 * 4
    y = torch.empty(M, ldc, device=x.device, dtype=torch.float32)
    plan = plan_fprop(1, 1, 1, 1, 1, 0)
    if not _launch(plan, x.view(M, 1, 1, K), weight, y.view(M, 1, 1, ldc), N, K, None, bias, False):
        return None
    return y[:, :N]


def linear_dgrad(gy: torch.Tensor, weight: torch.Tensor) -> Optional[torch.Tensor]:
    """gy: ``[M, N]`` (row stride a multiple of 4 floats; the pad columns of a padded buffer must be zero),
    weight ``[N, K]`` -> ``[M, K]``."""
    if not (ENABLED and gy.is_cuda and gy.dtype == torch.float32 and gy.dim() == 2 and gy.stride(1) == 1
            and gy.stride(0) % 4 == 0 and weight.is_contiguous() and weight.shape[1] % 4 == 0):
        return None
    M = gy.shape[0]
    N, K = weight.shape
    ld = gy.stride(0)
    src = torch.as_strided(gy, (M, 1, 1, ld), (ld, ld, ld, 1))         # the pixel pitch is the last dimension
    gx = torch.empty(M, K, device=gy.device, dtype=torch.float32)
    plan = plan_dgrad(1, 1, 1, 1, 1, 0)
    if not _launch(plan, src, weight, gx.view(M, 1, 1, K), K, K, None, None, False):
        return None
    return gx
