"""Explicit forward/backward schedule of the client-batched fedsgd step for the ResNet family on B200 -- every launch
is one of this repo's kernels, nothing goes through autograd, cuDNN, cuBLAS or ATen.

The generic engine (``engine/batched.py``) swaps module forwards and lets autograd drive the backward pass; that keeps
every model working but leaves the shared-weight GEMMs to the libraries and the glue (ReLU, residual adds, pooling,
layout copies, gradient accumulation at the residual forks) to ~70 ATen launches per ResNet-18 step.  Here the step of
``models.resnet.ResNet`` (BasicBlock / Bottleneck, per-client "ghost" BatchNorm) is written out as a fixed schedule:

    forward  per conv unit:   implicit-GEMM conv (tcgen05, ops.conv)  ->  per-client BN (+ residual, + ReLU fused)
    backward per conv unit:   per-client BN backward (ReLU mask + dgamma/dbeta into the update rows fused)
                              -> per-client wgrad GEMM with the update-row epilogue (tcgen05, ops.wgrad)
                              -> implicit-GEMM dgrad (tcgen05) with the residual-branch gradient accumulated in its epilogue

Same math as the reference client's ``loss.backward(); optimizer.step(); update = after - before``
(/root/reference/src/blades/client.py:127-131,178-198) for all clients of the shard at once.
Activations are NHWC fp32; GEMMs run tf32 with fp32 accumulation (the reference's convolutions run TF32 under torch
defaults as well)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from ..models.resnet import BasicBlock, Bottleneck, ResNet
from ..ops import client_bn as kbn
from ..ops import conv as kc
from ..ops import fused as kf
from ..ops import wgrad as kw
from ..ops.im2col import im2col_nhwc

__all__ = ["supports", "supports_eval", "step", "forward_logits"]


def _model_ok(model: nn.Module, x: torch.Tensor) -> bool:
    if not (kc.ENABLED and isinstance(model, ResNet) and model.norm_kind == "batch" and x.is_cuda
            and x.dtype == torch.float32):
        return False
    if type(model.maxpool) is not nn.MaxPool2d or model.fc.bias is None:
        return False
    for m in model.modules():
        if isinstance(m, nn.Conv2d) and (m.bias is not None or m.groups != 1 or m.dilation != (1, 1)):
            return False
        if isinstance(m, nn.BatchNorm2d) and (m.track_running_stats or not m.affine):
            return False
    for name, p in model.named_parameters():
        if p.dim() > 1 and (p.data_ptr() % 16 or p.dtype != torch.float32):
            return False
        if p.dim() == 4 and not p.data.permute(0, 2, 3, 1).is_contiguous():      # physical layout = channels_last
            return False
    return True


def supports(model: nn.Module, sink, x: torch.Tensor) -> bool:
    return bool(model.training and sink.channels_last and sink.out.dtype == torch.float32 and _model_ok(model, x))


def supports_eval(model: nn.Module, x: torch.Tensor) -> bool:
    """Forward-only use (evaluation): the model's BatchNorms have no running statistics, so ``model.eval()`` still
    normalises with the statistics of each forward batch -- exactly the per-group statistics of the fused pass."""
    return _model_ok(model, x)


class _Unit:
    """conv -> per-client BN (+ residual) (+ ReLU): the tensors the backward pass needs."""
    __slots__ = ("conv", "bn", "wname", "gname", "bname", "x", "c", "mean", "rstd", "y", "relu", "cols", "w2d", "has_res")


def _w2d(conv: nn.Conv2d) -> torch.Tensor:
    w = conv.weight.data
    v = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)
    assert v.data_ptr() == w.data_ptr(), "conv weight is not stored channels_last"
    return v


class _Pass:
    def __init__(self, model: ResNet, sink, n: int, keep: bool = True):
        self.model, self.sink, self.n = model, sink, n
        self.keep = keep            # False: forward only -- drop what the backward pass would need
        self.names = {id(m): name for name, m in model.named_modules()}

    # ---------------------------------------------------------------- forward pieces
    def conv_fwd(self, u: _Unit, x: torch.Tensor) -> torch.Tensor:
        conv = u.conv
        kh, kw_ = conv.kernel_size
        s, p = conv.stride[0], conv.padding[0]
        Cout, Cin = conv.out_channels, conv.in_channels
        u.x, u.cols = x, None
        if Cin % 32 == 0:
            u.w2d = _w2d(conv)
            c = kc.conv_fprop(x, u.w2d, (kh, kw_), s, p)
            if c is None:
                raise RuntimeError(f"conv_fprop declined {tuple(x.shape)} k{kh} s{s}")
            return c
        # few input channels (the stem): explicit im2col (kept for the wgrad GEMM) + the same kernel as a 1x1 conv
        NB, _, H, W = x.shape
        Ho, Wo = (H + 2 * p - kh) // s + 1, (W + 2 * p - kw_) // s + 1
        u.cols = im2col_nhwc(x, (kh, kw_), (s, s), (p, p), (1, 1), (Ho, Wo))      # [NB*Ho*Wo, K] view, ld = K rounded to 4
        K = Cin * kh * kw_
        w = conv.weight.data.permute(0, 2, 3, 1).reshape(Cout, K)
        u.w2d = w
        wp = w if K % 4 == 0 else kf.pad_rows(w)
        ld = u.cols.stride(0)
        src = torch.as_strided(u.cols, (NB * Ho * Wo, 1, 1, ld), (ld, ld, ld, 1))
        c = torch.empty((NB, Cout, Ho, Wo), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
        out = c.permute(0, 2, 3, 1).reshape(NB * Ho * Wo, 1, 1, Cout)
        assert out.data_ptr() == c.data_ptr()
        if not kc._launch(kc.plan_fprop(1, 1, 1, 1, 1, 0), src, wp, out, Cout, ld, None, None, False):
            raise RuntimeError("stem GEMM declined")
        return c

    def unit_fwd(self, conv: nn.Conv2d, bn: nn.BatchNorm2d, x: torch.Tensor, relu: bool,
                 res: Optional[torch.Tensor] = None) -> _Unit:
        u = _Unit()
        u.conv, u.bn, u.relu = conv, bn, relu
        u.has_res = res is not None
        u.wname = self.names[id(conv)] + ".weight"
        u.gname, u.bname = self.names[id(bn)] + ".weight", self.names[id(bn)] + ".bias"
        u.c = self.conv_fwd(u, x)
        u.y, u.mean, u.rstd = kbn.forward(u.c, bn.weight.data, bn.bias.data, self.n, bn.eps, res=res, relu=relu,
                                          nhwc=True)
        if not self.keep:
            u.x = u.c = u.cols = None
        return u

    # ---------------------------------------------------------------- backward pieces
    def bn_bwd(self, u: _Unit, gy: torch.Tensor, want_masked: bool) -> torch.Tensor:
        """BN backward of the unit (dgamma / dbeta go to the update rows); returns d/d(conv output).  With
        ``want_masked`` the ReLU-masked incoming gradient is written back into ``gy`` (the residual branch's share)."""
        s = self.sink
        assert gy.is_contiguous(memory_format=torch.channels_last)      # the in-place mask must hit the caller's tensor
        gc = kbn.backward(gy, u.c, u.mean, u.rstd, u.bn.weight.data, self.n, s.view(u.gname), s.view(u.bname), s.alpha,
                          True, act=u.y if u.relu else None, gmask=gy if (want_masked and u.relu) else None, nhwc=True,
                          beta=None if u.has_res else u.bn.bias.data)
        s.written.update((u.gname, u.bname))
        return gc

    def wgrad(self, u: _Unit, gc: torch.Tensor) -> None:
        s = self.sink
        conv = u.conv
        kh, kw_ = conv.kernel_size
        Cout = conv.out_channels
        out3 = s.view(u.wname).view(self.n, Cout, -1)
        NB, _, Ho, Wo = gc.shape
        T = (NB // self.n) * Ho * Wo
        if u.cols is None and kw.conv_wgrad_implicit(gc, u.x, out3, self.n, (kh, kw_), conv.stride, conv.padding, (1, 1),
                                                     s.alpha):
            s.written.add(u.wname)
            return
        cols = u.cols
        if cols is None:
            cols = im2col_nhwc(u.x, (kh, kw_), conv.stride, conv.padding, (1, 1), (Ho, Wo))
        b = cols.as_strided((self.n, T, cols.shape[1]), (T * cols.stride(0), cols.stride(0), 1))
        a_t = gc.permute(0, 2, 3, 1).reshape(self.n, T, Cout)
        s.put_bmm(u.wname, a_t.transpose(1, 2), b)

    def dgrad(self, u: _Unit, gc: torch.Tensor, add: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
        conv = u.conv
        gx = kc.conv_dgrad(gc, u.w2d, conv.kernel_size, conv.stride[0], conv.padding[0], tuple(u.x.shape[2:]),
                           conv.in_channels, add=add, out=out)
        if gx is None:
            raise RuntimeError(f"conv_dgrad declined {tuple(gc.shape)}")
        return gx


def _block_fwd(ps: _Pass, blk, x: torch.Tensor):
    if isinstance(blk, BasicBlock):
        convs = [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2)]
    else:
        convs = [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2), (blk.conv3, blk.bn3)]
    ds = None
    idt = x
    if blk.downsample is not None:
        ds = ps.unit_fwd(blk.downsample[0], blk.downsample[1], x, relu=False)
        idt = ds.y
    units: List[_Unit] = []
    h = x
    for i, (cv, bn) in enumerate(convs):
        last = i == len(convs) - 1
        u = ps.unit_fwd(cv, bn, h, relu=True, res=idt if last else None)
        units.append(u)
        h = u.y
    return h, (units, ds)


def _block_bwd(ps: _Pass, saved, g: torch.Tensor) -> torch.Tensor:
    """g: gradient w.r.t. the block output (post-ReLU); returns the gradient w.r.t. the block input."""
    units, ds = saved
    # last unit: out = relu(bn(c) + idt): its BN backward masks g in place -> g is also the identity branch's gradient
    gc = ps.bn_bwd(units[-1], g, want_masked=True)
    for i in range(len(units) - 1, 0, -1):
        ps.wgrad(units[i], gc)
        ga = ps.dgrad(units[i], gc)
        gc = ps.bn_bwd(units[i - 1], ga, want_masked=False)
    ps.wgrad(units[0], gc)
    if ds is None:
        return ps.dgrad(units[0], gc, add=g)                       # + identity branch, fused in the dgrad epilogue
    gx = ps.dgrad(units[0], gc)
    gcd = ps.bn_bwd(ds, g, want_masked=False)
    ps.wgrad(ds, gcd)
    return ps.dgrad(ds, gcd, add=gx, out=gx)                        # accumulate the shortcut's share in place


def _forward(ps: _Pass, x: torch.Tensor):
    model = ps.model
    stem = ps.unit_fwd(model.conv1, model.bn1, x, relu=True)
    mp = model.maxpool
    as_int = lambda v: v if isinstance(v, int) else v[0]
    k, st, pd = as_int(mp.kernel_size), as_int(mp.stride), as_int(mp.padding)
    h, pool_idx = kf.maxpool_fwd(stem.y, k, st, pd)
    saved = []
    for li, layer in enumerate((model.layer1, model.layer2, model.layer3, model.layer4)):
        for blk in layer:
            h, sv = _block_fwd(ps, blk, h)
            saved.append((li + 1, sv) if ps.keep else None)
    NB, Cf, Hf, Wf = h.shape
    feat = h.reshape(NB, Cf) if Hf * Wf == 1 else kf.avgpool_fwd(h)
    if Hf * Wf == 1:
        assert feat.data_ptr() == h.data_ptr()
    logits = kc.linear_fprop(feat, model.fc.weight.data, model.fc.bias.data)
    if logits is None:
        raise RuntimeError("linear_fprop declined the classifier")
    return logits, feat, (Hf, Wf), saved, stem, pool_idx, (k, st, pd)


def forward_logits(model: ResNet, x: torch.Tensor, n_groups: int) -> torch.Tensor:
    """Forward only: logits of ``x`` (``[n_groups*B, Cin, H, W]``) with BatchNorm statistics taken per group of B
    samples -- what ``model(x_group)`` computes group by group (evaluation, reference client.py:144-176)."""
    with torch.no_grad():
        return _forward(_Pass(model, None, n_groups, keep=False), x)[0]


def _final_from(sink, prefixes) -> int:
    """Smallest flat offset of the parameters under the given module-name prefixes."""
    return min(sp.offset for name, sp in sink.by_name.items() if name.startswith(prefixes))


def step(model: ResNet, sink, x: torch.Tensor, y: torch.Tensor, n: int, clamp: torch.Tensor,
         progress=None) -> torch.Tensor:
    """One fedsgd step of ``n`` clients (x: ``[n*B, Cin, H, W]`` in any layout, y: ``[n*B]`` int64): fills the rows of
    ``sink.out`` with ``-lr * grad_c`` and returns the per-client mean losses.

    ``progress(lo)``: called during the backward pass each time a stage of the network is done, with the flat offset
    ``lo`` from which on every update coordinate is final (the backward pass visits fc, layer4, ..., layer1, the stem,
    and the flat vector stores them in the opposite order, so the finished part is always a suffix).  The round engine
    uses it to start aggregating those coordinates on a side stream while the rest of the backward pass still runs."""
    ps = _Pass(model, sink, n)
    s = sink
    with torch.no_grad():
        logits, feat, (Hf, Wf), saved, stem, pool_idx, (k, st, pd) = _forward(ps, x)
        NB, Cf = feat.shape
        fc = model.fc
        loss, glogits = kf.client_ce(logits, y, n, clamp)
        # ------------------------------------------------------------------ backward
        T = NB // n
        s.put_bmm("fc.weight", glogits.view(n, T, -1).transpose(1, 2), feat.view(n, T, Cf))
        kf.client_colsum(glogits, n, s.view("fc.bias"), s.alpha)
        s.written.add("fc.bias")
        gfeat = kc.linear_dgrad(glogits, fc.weight.data)
        if gfeat is None:
            raise RuntimeError("linear_dgrad declined the classifier")
        g = gfeat.view(NB, Cf, 1, 1) if Hf * Wf == 1 else kf.avgpool_bwd(gfeat, (Hf, Wf))
        if Hf * Wf == 1:
            g = g.contiguous(memory_format=torch.channels_last)      # no-op for 1x1 maps (both layouts coincide)
        ordered = progress is not None and _final_from(s, ("fc.",)) > _final_from(s, ("layer4.",)) > \
            _final_from(s, ("layer3.",)) > _final_from(s, ("layer2.",)) > _final_from(s, ("layer1.",)) > 0
        for i in range(len(saved) - 1, -1, -1):
            li, sv = saved[i]
            g = _block_bwd(ps, sv, g)
            if ordered and (i == 0 or saved[i - 1][0] != li):        # first block of stage li done: the stage is final
                progress(_final_from(s, (f"layer{li}.",)))
        gy = kf.maxpool_bwd(g, pool_idx, tuple(stem.y.shape[2:]), k, st, pd)
        gc = ps.bn_bwd(stem, gy, want_masked=False)
        ps.wgrad(stem, gc)
    return loss
