"""The implicit-GEMM convolution planner (ops/conv.py) executed with plain torch ops must reproduce F.conv2d and its
input gradient for every geometry the engine meets -- the tcgen05 kernel consumes the very same description."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from blades_b200.ops import conv as kc

GEOMS = [  # (H, W, Cin, Cout, k, stride, pad)
    (8, 8, 8, 12, 3, 1, 1), (8, 8, 8, 12, 3, 2, 1), (4, 4, 8, 12, 3, 2, 1), (2, 2, 8, 12, 3, 1, 1),
    (1, 1, 8, 12, 3, 1, 1), (2, 2, 8, 12, 3, 2, 1), (8, 8, 8, 12, 1, 2, 0), (8, 8, 8, 12, 1, 1, 0),
    (7, 9, 4, 8, 3, 2, 1), (9, 7, 4, 8, 5, 1, 2), (6, 6, 4, 8, 3, 1, 0), (5, 5, 4, 8, 2, 2, 0), (16, 16, 4, 4, 3, 2, 1),
]


def _w2d(w):
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()          # [Cout, kh*kw*Cin], K = (r, s, cin)


@pytest.mark.parametrize("H,W,Cin,Cout,k,stride,pad", GEOMS)
def test_fprop_plan_equals_conv2d(H, W, Cin, Cout, k, stride, pad):
    torch.manual_seed(0)
    x = torch.randn(3, Cin, H, W, dtype=torch.float64)
    w = torch.randn(Cout, Cin, k, k, dtype=torch.float64)
    b = torch.randn(Cout, dtype=torch.float64)
    ref = F.conv2d(x, w, b, stride, pad)
    plan = kc.plan_fprop(H, W, k, k, stride, pad)
    got = kc.emulate(plan, x.permute(0, 2, 3, 1).contiguous(), _w2d(w), Cout, Cin, bias=b)
    assert torch.allclose(got.permute(0, 3, 1, 2), ref, atol=1e-10)
    assert plan.max_taps() <= k * k


@pytest.mark.parametrize("H,W,Cin,Cout,k,stride,pad", GEOMS)
def test_dgrad_plan_equals_autograd(H, W, Cin, Cout, k, stride, pad):
    torch.manual_seed(1)
    x = torch.randn(2, Cin, H, W, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cout, Cin, k, k, dtype=torch.float64)
    y = F.conv2d(x, w, None, stride, pad)
    gy = torch.randn_like(y)
    (ref,) = torch.autograd.grad(y, x, gy)
    plan = kc.plan_dgrad(H, W, k, k, stride, pad)
    assert len(plan.phases) == stride * stride
    add = torch.randn(2, H, W, Cin, dtype=torch.float64)
    got = kc.emulate(plan, gy.permute(0, 2, 3, 1).contiguous(), _w2d(w), Cin, Cin, add=add)
    assert torch.allclose(got.permute(0, 3, 1, 2), ref + add.permute(0, 3, 1, 2), atol=1e-10)
    # in-place accumulation: pixel classes without taps keep their old value
    out = add.clone()
    kc.emulate(plan, gy.permute(0, 2, 3, 1).contiguous(), _w2d(w), Cin, Cin, add=out, accumulate_only=True, out=out)
    assert torch.allclose(out.permute(0, 3, 1, 2), ref + add.permute(0, 3, 1, 2), atol=1e-10)


def test_dead_taps_are_dropped():
    assert kc.plan_fprop(1, 1, 3, 3, 1, 1).max_taps() == 1            # ResNet layer4 on 32x32 inputs: centre tap only
    assert kc.plan_fprop(2, 2, 3, 3, 1, 1).max_taps() == 9
    assert kc.plan_dgrad(1, 1, 3, 3, 1, 1).max_taps() == 1
    p = kc.plan_dgrad(8, 8, 1, 1, 2, 0)                                # 1x1 stride-2 shortcut: one class has the tap
    assert sorted(len(t) for _, _, t in p.phases) == [0, 0, 0, 1]


def test_linear_is_a_1x1_convolution():
    torch.manual_seed(2)
    x, w, b = torch.randn(7, 20, dtype=torch.float64), torch.randn(5, 20, dtype=torch.float64), torch.randn(5, dtype=torch.float64)
    got = kc.emulate(kc.plan_fprop(1, 1, 1, 1, 1, 0), x.view(7, 1, 1, 20), w, 5, 20, bias=b)
    assert torch.allclose(got.view(7, 5), F.linear(x, w, b), atol=1e-12)
    gy = torch.randn(7, 5, dtype=torch.float64)
    gx = kc.emulate(kc.plan_dgrad(1, 1, 1, 1, 1, 0), gy.view(7, 1, 1, 5), w, 20, 20)
    assert torch.allclose(gx.view(7, 20), gy @ w, atol=1e-12)


def test_tile_box_twin_and_struct_size():
    from blades_b200.ops import _loader
    lib = _loader.cuda_lib(optional=True)
    if lib is None:
        pytest.skip("native library not built")
    assert lib.bl_sizeof_conv_desc() == ctypes.sizeof(kc.ConvDesc)
    bw, bh, bb = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    for Wt, Ht, NB, cs in [(8, 8, 3200, 1), (4, 4, 3200, 2), (2, 2, 100, 1), (1, 1, 3200, 1), (16, 16, 64, 2),
                           (28, 28, 10, 1), (7, 9, 5, 1), (130, 4, 2, 1), (128, 128, 1, 2), (3, 5, 1000, 1)]:
        r = lib.bl_conv_tile_box(Wt, Ht, NB, cs, ctypes.byref(bw), ctypes.byref(bh), ctypes.byref(bb))
        assert r == kc.tile_rows(Wt, Ht, NB, cs), (Wt, Ht, NB, cs)
        if r:
            assert r == bw.value * bh.value * bb.value <= 128 and bw.value == Wt
