"""Client-batched local training (fedsgd fast path).

In a fedsgd round every client on a shard starts from the SAME parameters and takes
ONE SGD step on its own mini-batch, so ``delta_c = -lr * grad_c(theta)``.  Instead of
time-slicing ``n`` tiny forward/backward passes (reference actor.py:23-33: one Python
loop iteration, ~200 kernel launches and a deep-copied model per client) we run ONE
forward/backward over the concatenated batch ``[n*B, ...]``:

* activations / input-gradients: shared weights -> ordinary big-batch kernels;
* parameter gradients: **per client** -- every parametrised layer's backward computes
  ``dW_c`` for each client as a grouped GEMM over that client's ``B`` samples and
  writes ``alpha_c * dW_c`` straight into row ``c`` of the shard's update matrix
  ``U_g[n, d]`` at the parameter's flat offset (``alpha_c = -lr`` -- the SGD step, the
  update diff ``theta_after - theta_before`` and ``save_update`` of the reference
  (client.py:127-131,178-198) all collapse into the GEMM epilogue);
* BatchNorm uses per-client batch statistics (each client normalises over its own
  ``B`` samples exactly as it would in isolation).

Mechanics: supported leaf modules get their ``forward`` swapped for the duration of
the pass with a version built on custom ``autograd.Function`` s; stray parameters used
by broadcasting (e.g. CCT's ``positional_emb``) are re-parametrised the same way.
Models containing anything else fall back to the time-sliced engine.

The grouped weight-gradient GEMM is ``ops.wgrad`` (tcgen05, TMA-fed, fused scale +
nan-to-num epilogue into ``U_g``) on B200 and ``torch.baddbmm`` elsewhere.
"""
from __future__ import annotations

import contextlib
import types
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .flat import ParamSpec

__all__ = ["GradSink", "client_batched", "is_batchable", "batched_loss", "batched_step", "BatchedUnsupported"]


class BatchedUnsupported(RuntimeError):
    pass


class GradSink:
    """Where per-client parameter gradients land: ``out[c, off:off+numel] = alpha[c] * grad_c``.

    ``out`` is the shard's update matrix (fedsgd: alpha = -lr) or a scratch gradient
    matrix.  ``alpha`` is a python float (uniform) -- per-client sign flips are applied
    afterwards on the few affected rows.
    """

    def __init__(self, out: torch.Tensor, specs: Sequence[ParamSpec], n_clients: int, alpha: float = 1.0):
        assert out.dim() == 2 and out.shape[0] >= n_clients
        self.out = out
        self.channels_last = any(s.channels_last for s in specs)
        self.n = n_clients
        self.alpha = float(alpha)
        self.by_name: Dict[str, ParamSpec] = {s.name: s for s in specs}
        self.written = set()

    def view(self, name: str) -> torch.Tensor:
        s = self.by_name[name]
        return self.out[: self.n, s.offset: s.offset + s.numel]

    def put(self, name: str, grad: torch.Tensor) -> None:
        """grad: [n, *shape] (logical shape, any strides)."""
        s = self.by_name[name]
        if s.channels_last:            # physical order [Cout, kh, kw, Cin]
            grad = grad.reshape((self.n,) + s.shape).permute(0, 1, 3, 4, 2)
        v = self.view(name)
        torch.mul(grad.reshape(self.n, -1), self.alpha, out=v)
        self.written.add(name)

    def put_bmm(self, name: str, a: torch.Tensor, b: torch.Tensor) -> None:
        """out_view[n, M, N] = alpha * a[n, M, K] @ b[n, K, N] (the grouped wgrad GEMM)."""
        s = self.by_name[name]
        v = self.view(name)
        M, N = a.shape[1], b.shape[2]
        v3 = v.view(self.n, M, N)
        from ..ops import wgrad as _w
        _w.grouped_wgrad(a, b, v3, self.alpha)
        self.written.add(name)


# ================================================================================ layer functions
class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, sink, wname, bname):
        ctx.save_for_backward(x, weight)
        ctx.sink, ctx.wname, ctx.bname = sink, wname, bname
        if x.is_cuda and x.dtype == torch.float32:
            # own tcgen05 GEMM (a 1x1 convolution over 1x1 images); declines odd strides / alignments
            from ..ops import conv as kc
            x2 = x.reshape(-1, x.shape[-1])
            y = kc.linear_fprop(x2, weight, bias) if x2.is_contiguous() else None
            if y is not None:
                return y.reshape(tuple(x.shape[:-1]) + (weight.shape[0],))
        return F.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        sink: GradSink = ctx.sink
        n = sink.n
        gy2 = gy.reshape(-1, gy.shape[-1])
        x2 = x.reshape(-1, x.shape[-1])
        T = gy2.shape[0] // n
        # dW_c[out, in] = gy_c^T @ x_c
        sink.put_bmm(ctx.wname, gy2.view(n, T, -1).transpose(1, 2), x2.view(n, T, -1))
        if ctx.bname is not None:
            sink.put(ctx.bname, gy2.view(n, T, -1).sum(1))
        gx = None
        if ctx.needs_input_grad[0]:
            if gy2.is_cuda and gy2.dtype == torch.float32:
                from ..ops import conv as kc
                g2 = kc.linear_dgrad(gy2, weight)
                if g2 is not None:
                    gx = g2.reshape(x.shape)
            if gx is None:
                gx = gy @ weight
        return gx, None, None, None, None, None


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _conv_dgrad(gy, x, weight, stride, padding, dilation, w2d=None):
    """Input gradient of conv2d.  ``torch.nn.grad.conv2d_input`` hands cuDNN a stride-0 dummy of the input's
    shape, which the backend then materialises with ``.contiguous()`` -- a full activation-sized copy per layer
    (measured: 20 copies, 0.31 ms of an 11 ms ResNet-18 round).  Passing the real saved input costs nothing and
    keeps the result channels_last."""
    if w2d is not None:                 # the forward ran on the own tcgen05 kernel: so does the input gradient
        from ..ops import conv as kc
        gx = kc.conv_dgrad(gy.contiguous(memory_format=torch.channels_last), w2d, tuple(weight.shape[2:]),
                           _pair(stride)[0], _pair(padding)[0], tuple(x.shape[2:]), x.shape[1])
        if gx is not None:
            return gx

    def pair(v):
        return [v, v] if isinstance(v, int) else list(v)
    return torch.ops.aten.convolution_backward(gy, x, weight, None, pair(stride), pair(padding), pair(dilation),
                                               False, [0, 0], 1, [True, False, False])[0]


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, sink, wname, bname, stride, padding, dilation):
        # NHWC end to end when the flat parameter layout is channels_last (GPU): cuDNN runs its native
        # tensor-core kernels without layout conversions and the wgrad GEMM's [Cout] x [kh*kw*Cin] output
        # IS the physical weight layout.
        if x.is_cuda and sink.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        ctx.save_for_backward(x, weight)
        ctx.sink, ctx.wname, ctx.bname = sink, wname, bname
        ctx.conf = (stride, padding, dilation)
        ctx.w2d = None
        if x.is_cuda and sink.channels_last:
            from ..ops import conv as kc
            if kc.supported_conv(x, weight, _pair(stride), _pair(padding), _pair(dilation)):
                wp = weight.permute(0, 2, 3, 1)
                if wp.is_contiguous():                         # physical layout = channels_last: [Cout, kh*kw*Cin] view
                    w2d = wp.reshape(weight.shape[0], -1)
                    y = kc.conv_fprop(x, w2d, tuple(weight.shape[2:]), _pair(stride)[0], _pair(padding)[0], bias=bias)
                    if y is not None:
                        ctx.w2d = w2d
                        return y
        return F.conv2d(x, weight, bias, stride, padding, dilation, 1)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        sink: GradSink = ctx.sink
        stride, padding, dilation = ctx.conf
        n = sink.n
        NB, Cout, Ho, Wo = gy.shape
        B = NB // n
        kh, kw = weight.shape[2], weight.shape[3]
        from ..ops.im2col import im2col_nhwc, im2col_rows
        L = Ho * Wo
        # both operands row-major over the client's T = B*L rows ("MN-major" for the tensor cores)
        if sink.by_name[ctx.wname].channels_last or (sink.channels_last and kh == 1 and kw == 1):
            # K ordered (r, s, cin) = physical order of the channels_last weight
            if gy.is_cuda:
                gy = gy.contiguous(memory_format=torch.channels_last)
                from ..ops import wgrad as _w
                out3 = sink.view(ctx.wname).view(n, Cout, -1)
                if _w.conv_wgrad_implicit(gy, x, out3, n, (kh, kw), stride, padding, dilation, sink.alpha):
                    sink.written.add(ctx.wname)           # implicit GEMM: no im2col matrix at all
                    if ctx.bname is not None:
                        sink.put(ctx.bname, gy.reshape(n, B, Cout, L).sum((1, 3)))
                    gx = None
                    if ctx.needs_input_grad[0]:
                        gx = _conv_dgrad(gy, x, weight, stride, padding, dilation, ctx.w2d)
                    return gx, None, None, None, None, None, None, None, None
            cols = im2col_nhwc(x, (kh, kw), stride, padding, dilation, (Ho, Wo))      # [NB*L, K] (padded rows)
            b = cols.as_strided((n, B * L, cols.shape[1]), (B * L * cols.stride(0), cols.stride(0), 1))
            a_t = gy.permute(0, 2, 3, 1).reshape(n, B * L, Cout)                      # view: no copy
        else:
            b = im2col_rows(x, (kh, kw), stride, padding, dilation, (Ho, Wo)).view(n, B * L, -1)   # [n, T, K]
            a_t = gy.reshape(n, B, Cout, L).permute(0, 1, 3, 2).reshape(n, B * L, Cout)           # [n, T, Cout]
        sink.put_bmm(ctx.wname, a_t.transpose(1, 2), b)
        if ctx.bname is not None:
            sink.put(ctx.bname, gy.reshape(n, B, Cout, L).sum((1, 3)))
        gx = None
        if ctx.needs_input_grad[0]:
            gx = _conv_dgrad(gy, x, weight, stride, padding, dilation, ctx.w2d)
        return gx, None, None, None, None, None, None, None, None


class _ClientBNFn(torch.autograd.Function):
    """BatchNorm2d with statistics per (client, channel) over that client's B*H*W values."""

    @staticmethod
    def forward(ctx, x, weight, bias, sink, wname, bname, eps):
        n = sink.n
        NB, C, H, W = x.shape
        ctx.sink, ctx.wname, ctx.bname = sink, wname, bname
        ctx.fused = False
        if x.is_cuda:
            from ..ops import client_bn as kbn
            xc = x if kbn.is_nhwc(x) else x.contiguous()
            if (kbn.is_nhwc(xc) or kbn.supported(xc)) and sink.out.dtype == torch.float32:
                y, mean, rstd = kbn.forward(xc, weight, bias, n, eps)
                ctx.save_for_backward(xc, mean, rstd, weight)
                ctx.fused = True
                return y
        x5 = x.view(n, NB // n, C, H * W)
        var, mean = torch.var_mean(x5, dim=(1, 3), unbiased=False, keepdim=True)
        rstd = torch.rsqrt(var + eps)
        xhat = (x5 - mean) * rstd
        y = torch.empty_like(x)
        torch.addcmul(bias.view(1, 1, C, 1), xhat, weight.view(1, 1, C, 1), out=y.view(x5.shape))
        ctx.save_for_backward(xhat, rstd, weight)
        ctx.sink, ctx.wname, ctx.bname = sink, wname, bname
        return y

    @staticmethod
    def backward(ctx, gy):
        sink: GradSink = ctx.sink
        n = sink.n
        if ctx.fused:
            from ..ops import client_bn as kbn
            x, mean, rstd, weight = ctx.saved_tensors
            gx = kbn.backward(gy, x, mean, rstd, weight, n, sink.view(ctx.wname), sink.view(ctx.bname),
                              sink.alpha, ctx.needs_input_grad[0])
            sink.written.update((ctx.wname, ctx.bname))
            return gx, None, None, None, None, None, None
        xhat, rstd, weight = ctx.saved_tensors
        NB, C, H, W = gy.shape
        g5 = gy.reshape(n, NB // n, C, H * W)
        dbeta = g5.sum((1, 3))                      # [n, C]
        dgamma = (g5 * xhat).sum((1, 3))            # [n, C]
        sink.put(ctx.wname, dgamma)
        sink.put(ctx.bname, dbeta)
        gx = None
        if ctx.needs_input_grad[0]:
            m = g5.shape[1] * g5.shape[3]
            w = weight.view(1, 1, C, 1)
            gx = torch.empty_like(gy)
            torch.mul(w * rstd, g5 - (dbeta.view(n, 1, C, 1) + xhat * dgamma.view(n, 1, C, 1)) / m,
                      out=gx.view(g5.shape))
        return gx, None, None, None, None, None, None


class _AffineFn(torch.autograd.Function):
    """y = xhat * w + b where xhat is already normalised per sample (LayerNorm / GroupNorm):
    only the affine parameters need per-client gradients.  ``keep`` = dims of ``xhat`` that
    the affine parameters span (LayerNorm: trailing dims; GroupNorm: dim 1)."""

    @staticmethod
    def forward(ctx, xhat, weight, bias, sink, wname, bname, keep):
        shape = [xhat.shape[i] if i in keep else 1 for i in range(xhat.dim())]
        ctx.save_for_backward(xhat, weight)
        ctx.sink, ctx.wname, ctx.bname, ctx.shape, ctx.keep = sink, wname, bname, shape, keep
        y = xhat * weight.view(shape)
        return y + bias.view(shape) if bias is not None else y

    @staticmethod
    def backward(ctx, gy):
        xhat, weight = ctx.saved_tensors
        sink: GradSink = ctx.sink
        n = sink.n
        per = gy.reshape((n, gy.shape[0] // n) + tuple(gy.shape[1:]))
        xh = xhat.reshape(per.shape)
        # reduce the per-client sample dim (1) and every non-kept dim (shifted by the client dim)
        red = [1] + [i + 1 for i in range(1, gy.dim()) if i not in ctx.keep]
        sink.put(ctx.wname, (per * xh).sum(red))
        if ctx.bname is not None:
            sink.put(ctx.bname, per.sum(red))
        return gy * weight.view(ctx.shape), None, None, None, None, None, None


class _BroadcastParamFn(torch.autograd.Function):
    """A parameter with leading dim 1 that the model broadcasts against the batch."""

    @staticmethod
    def forward(ctx, p, sink, name, total):
        ctx.sink, ctx.name = sink, name
        return p.expand((total,) + tuple(p.shape[1:]))

    @staticmethod
    def backward(ctx, g):
        sink: GradSink = ctx.sink
        n = sink.n
        sink.put(ctx.name, g.reshape((n, g.shape[0] // n) + tuple(g.shape[1:])).sum(1))
        return None, None, None, None


# ================================================================================ forward swaps
def _linear_forward(self, x):
    s = self._cb_sink
    return _LinearFn.apply(x, self.weight, self.bias, s, self._cb_names[0], self._cb_names[1])


def _conv_forward(self, x):
    s = self._cb_sink
    return _ConvFn.apply(x, self.weight, self.bias, s, self._cb_names[0], self._cb_names[1],
                         self.stride, self.padding, self.dilation)


def _bn_forward(self, x):
    s = self._cb_sink
    if not self.training and self.track_running_stats:
        return F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, False, 0.0, self.eps)
    return _ClientBNFn.apply(x, self.weight, self.bias, s, self._cb_names[0], self._cb_names[1], self.eps)


def _ln_forward(self, x):
    s = self._cb_sink
    xhat = F.layer_norm(x, self.normalized_shape, None, None, self.eps)
    keep = tuple(range(x.dim() - len(self.normalized_shape), x.dim()))
    return _AffineFn.apply(xhat, self.weight, self.bias, s, self._cb_names[0], self._cb_names[1], keep)


def _gn_forward(self, x):
    s = self._cb_sink
    xhat = F.group_norm(x, self.num_groups, None, None, self.eps)
    return _AffineFn.apply(xhat, self.weight, self.bias, s, self._cb_names[0], self._cb_names[1], (1,))


def _supported(m: nn.Module) -> Optional[Callable]:
    if type(m) is nn.Linear:
        return _linear_forward
    if type(m) is nn.Conv2d and m.groups == 1 and m.padding_mode == "zeros" and not isinstance(m.padding, str):
        return _conv_forward
    if type(m) is nn.BatchNorm2d and m.affine:
        return _bn_forward
    if type(m) is nn.LayerNorm and m.elementwise_affine:
        return _ln_forward
    if type(m) is nn.GroupNorm and m.affine:
        return _gn_forward
    return None


def _plan(model: nn.Module):
    """(leaf swaps, stray broadcast params) or raise BatchedUnsupported."""
    swaps: List[Tuple[nn.Module, Callable, Tuple[str, Optional[str]]]] = []
    covered = set()
    for mname, m in model.named_modules():
        own = [(k, p) for k, p in m._parameters.items() if p is not None and p.requires_grad]
        if not own:
            continue
        fwd = _supported(m)
        prefix = mname + "." if mname else ""
        if fwd is not None:
            names = (prefix + "weight", prefix + "bias" if m._parameters.get("bias") is not None else None)
            swaps.append((m, fwd, names))
            covered.update(x for x in names if x)
        else:
            for k, p in own:
                if p.dim() >= 2 and p.shape[0] == 1:
                    covered.add(prefix + k)
                    swaps.append((m, None, (k, prefix + k)))
                else:
                    raise BatchedUnsupported(f"parameter {prefix + k} of {type(m).__name__} is not batchable")
    return swaps


def is_batchable(model: nn.Module) -> bool:
    try:
        _plan(model)
        return True
    except BatchedUnsupported:
        return False


@contextlib.contextmanager
def client_batched(model: nn.Module, sink: GradSink, total_batch: int):
    """Swap forwards so that one backward pass fills ``sink`` with per-client gradients."""
    plan = _plan(model)
    undo = []
    try:
        for m, fwd, names in plan:
            if fwd is not None:
                m._cb_sink, m._cb_names = sink, names
                m.forward = types.MethodType(fwd, m)
                undo.append(("fwd", m, None, None))
            else:
                attr, full = names
                p = m._parameters.pop(attr)
                setattr(m, attr, _BroadcastParamFn.apply(p, sink, full, total_batch))
                undo.append(("param", m, attr, p))
        yield
    finally:
        for kind, m, attr, p in undo:
            if kind == "fwd":
                del m.forward
                del m._cb_sink, m._cb_names
            else:
                delattr(m, attr)
                m._parameters[attr] = p


def batched_loss(logits: torch.Tensor, target: torch.Tensor, n: int, clamp: torch.Tensor) -> torch.Tensor:
    """Sum over clients of clamp(mean CE over the client's batch, 0, clamp_c) -> scalar,
    plus the per-client losses (detached) for logging.  ``clamp``: [n] tensor."""
    per_sample = F.cross_entropy(logits, target, reduction="none")
    per_client = per_sample.view(n, -1).mean(1)
    clamped = torch.minimum(per_client.clamp_min(0), clamp)
    return clamped.sum(), per_client.detach()


def batched_step(model: nn.Module, sink: GradSink, x: torch.Tensor, y: torch.Tensor, n: int,
                 clamp: torch.Tensor, progress=None) -> torch.Tensor:
    """One client-batched fedsgd step: fills ``sink`` with ``alpha * grad_c`` for the ``n`` clients whose samples are
    concatenated in ``x`` / ``y`` and returns the per-client mean losses.  The ResNet family runs the explicit
    all-own-kernels schedule (``engine/resnet_fused.py``) on B200; everything else the swapped-forward autograd pass."""
    from . import resnet_fused as rf
    if rf.supports(model, sink, x):
        return rf.step(model, sink, x, y, n, clamp, progress)      # ``progress``: see resnet_fused.step
    with client_batched(model, sink, x.shape[0]):
        logits = model(x)
        if logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1 \
                and y.dtype == torch.int64 and sink.out.dtype == torch.float32:
            # per-client clamped mean cross-entropy and its logits gradient in ONE own launch; the gradient rows are
            # padded to 16 B so the classifier's input-gradient GEMM can read them by TMA
            from ..ops import fused as kf
            per_client, g = kf.client_ce(logits.detach(), y.contiguous(), n, clamp)
            logits.backward(g)
        else:
            loss, per_client = batched_loss(logits, y, n, clamp)
            loss.backward()
    return per_client
