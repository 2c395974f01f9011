"""GPU numerics of the own-kernel training pass: tcgen05 implicit-GEMM conv (fprop / dgrad), fused BN activation,
pooling, per-client cross-entropy, and the explicit ResNet schedule built from them -- each against an fp64 reference
of the same op."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


CONV_CASES = [  # NB, H, W, Cin, Cout, k, stride, pad
    (64, 8, 8, 64, 64, 3, 1, 1), (64, 8, 8, 64, 128, 3, 2, 1), (64, 8, 8, 64, 128, 1, 2, 0), (96, 4, 4, 128, 256, 3, 2, 1),
    (96, 2, 2, 256, 256, 3, 1, 1), (160, 2, 2, 256, 512, 3, 2, 1), (160, 1, 1, 512, 512, 3, 1, 1), (32, 8, 8, 64, 256, 1, 1, 0),
    (5, 7, 9, 32, 48, 3, 1, 1), (5, 9, 7, 32, 40, 3, 2, 1), (9, 16, 16, 32, 32, 3, 1, 1), (7, 8, 8, 32, 64, 5, 1, 2),
    (3, 32, 32, 32, 32, 3, 1, 1), (300, 1, 1, 64, 20, 1, 1, 0),
]


@pytest.mark.parametrize("NB,H,W,Cin,Cout,k,s,p", CONV_CASES)
def test_conv_tcgen05_fprop_and_dgrad(NB, H, W, Cin, Cout, k, s, p):
    from blades_b200.ops import conv as kc
    torch.manual_seed(NB + H + Cin + k)
    x = _cl(torch.randn(NB, Cin, H, W, device=_dev()))
    w = _cl(torch.randn(Cout, Cin, k, k, device=_dev()) / (Cin * k * k) ** 0.5)
    w2d = w.permute(0, 2, 3, 1).reshape(Cout, -1)
    assert w2d.data_ptr() == w.data_ptr()
    ref = F.conv2d(x.double(), w.double(), None, s, p)
    y = kc.conv_fprop(x, w2d, (k, k), s, p)
    assert y is not None and y.is_contiguous(memory_format=torch.channels_last)
    assert _rel(y, ref) < 2e-3                                          # tf32 operands, fp32 accumulation
    gy = _cl(torch.randn(NB, Cout, ref.shape[2], ref.shape[3], device=_dev()))
    x64 = x.double().requires_grad_(True)
    (gref,) = torch.autograd.grad(F.conv2d(x64, w.double(), None, s, p), x64, gy.double())
    gx = kc.conv_dgrad(gy, w2d, (k, k), s, p, (H, W), Cin)
    assert gx is not None and _rel(gx, gref) < 2e-3
    acc = _cl(torch.randn(NB, Cin, H, W, device=_dev()))
    want = gref + acc.double()
    kc.conv_dgrad(gy, w2d, (k, k), s, p, (H, W), Cin, add=acc, out=acc)
    assert _rel(acc, want) < 2e-3
    other = _cl(torch.randn(NB, Cin, H, W, device=_dev()))
    got = kc.conv_dgrad(gy, w2d, (k, k), s, p, (H, W), Cin, add=other)
    assert _rel(got, gref + other.double()) < 2e-3


@pytest.mark.parametrize("M,K,N", [(3200, 512, 10), (256, 784, 512), (100, 64, 24), (77, 128, 300)])
def test_linear_tcgen05(M, K, N):
    from blades_b200.ops import conv as kc
    torch.manual_seed(M)
    x = torch.randn(M, K, device=_dev())
    w = torch.randn(N, K, device=_dev()) / K ** 0.5
    b = torch.randn(N, device=_dev())
    y = kc.linear_fprop(x, w, b)
    assert y is not None and _rel(y, F.linear(x.double(), w.double(), b.double())) < 2e-3
    ld = (N + 3) // 4 * 4
    g = torch.zeros(M, ld, device=_dev())
    g[:, :N] = torch.randn(M, N, device=_dev())
    gx = kc.linear_dgrad(g[:, :N], w)
    assert gx is not None and _rel(gx, g[:, :N].double() @ w.double()) < 2e-3


@pytest.mark.parametrize("n,B,C,H", [(5, 32, 64, 16), (3, 32, 64, 8), (4, 16, 128, 4), (6, 32, 256, 2), (7, 32, 512, 1), (2, 8, 24, 7)])
@pytest.mark.parametrize("with_res", [False, True])
def test_client_bn_fused_relu_residual(n, B, C, H, with_res):
    from blades_b200.ops import client_bn as kbn
    torch.manual_seed(n * C + H)
    x = _cl(torch.randn(n * B, C, H, H, device=_dev()) * 2 + 0.5)
    res = _cl(torch.randn(n * B, C, H, H, device=_dev())) if with_res else None
    gamma, beta = torch.rand(C, device=_dev()) + 0.5, torch.randn(C, device=_dev()) * 0.3
    y, mean, rstd = kbn.forward(x, gamma, beta, n, 1e-5, res=res, relu=True, nhwc=True)
    x5 = x.double().view(n, B, C, H * H).requires_grad_(True)
    var, mu = torch.var_mean(x5, dim=(1, 3), unbiased=False, keepdim=True)
    pre = (x5 - mu) / torch.sqrt(var + 1e-5) * gamma.double().view(1, 1, C, 1) + beta.double().view(1, 1, C, 1)
    r5 = res.double().view(n, B, C, H * H).requires_grad_(True) if with_res else None
    yref = torch.relu(pre + r5) if with_res else torch.relu(pre)
    assert torch.allclose(y.double().view_as(yref), yref, atol=1e-4, rtol=1e-4)
    gy = _cl(torch.randn(n * B, C, H, H, device=_dev()))
    # mask from OUR output (ties at exactly 0 aside, identical to the reference's)
    mask = (y > 0).double().view_as(yref)
    gmasked_ref = gy.double().view_as(yref) * mask
    xhat = ((x5 - mu) / torch.sqrt(var + 1e-5)).detach()
    (gx_ref,) = torch.autograd.grad(pre, x5, gmasked_ref)
    U = torch.zeros(n, 2 * C + 64, device=_dev())
    g_inplace = gy.clone(memory_format=torch.channels_last)
    # without a residual the kernel is told beta and recomputes the ReLU mask from x (no read of the forward output):
    # the masked gradient must be EXACTLY the one the act-based mask gives
    dx = kbn.backward(g_inplace, x, mean, rstd, gamma, n, U[:, 8:8 + C], U[:, 8 + C:8 + 2 * C], -0.1, True, act=y,
                      gmask=g_inplace, nhwc=True, beta=None if with_res else beta)
    assert torch.equal(g_inplace, gy * (y > 0))                                              # bit-identical mask
    assert torch.allclose(g_inplace.double().view_as(yref), gmasked_ref, atol=1e-6)         # residual branch's share
    assert torch.allclose(dx.double().view_as(gx_ref), gx_ref, atol=2e-4, rtol=1e-3)
    assert torch.allclose(U[:, 8:8 + C].double(), -0.1 * (gmasked_ref * xhat).sum((1, 3)), atol=2e-3, rtol=1e-3)
    assert torch.allclose(U[:, 8 + C:8 + 2 * C].double(), -0.1 * gmasked_ref.sum((1, 3)), atol=2e-3, rtol=1e-3)


@pytest.mark.parametrize("NB,C,H,W,k,s,p", [(6, 64, 16, 16, 3, 2, 1), (3, 8, 9, 7, 3, 2, 1), (4, 16, 8, 8, 2, 2, 0), (2, 4, 5, 5, 3, 1, 1)])
def test_maxpool_nhwc(NB, C, H, W, k, s, p):
    from blades_b200.ops import fused as kf
    torch.manual_seed(H)
    x = _cl(torch.randn(NB, C, H, W, device=_dev()))
    y, idx = kf.maxpool_fwd(x, k, s, p)
    xr = x.double().requires_grad_(True)
    yr = F.max_pool2d(xr, k, s, p)
    assert torch.equal(y.double(), yr.detach())
    gy = _cl(torch.randn_like(y))
    (gr,) = torch.autograd.grad(yr, xr, gy.double())
    gx = kf.maxpool_bwd(gy, idx, (H, W), k, s, p)
    assert torch.allclose(gx.double(), gr, atol=1e-6)


def test_avgpool_ce_colsum_padrows():
    from blades_b200.ops import fused as kf
    torch.manual_seed(0)
    x = _cl(torch.randn(6, 32, 3, 5, device=_dev()))
    assert torch.allclose(kf.avgpool_fwd(x), x.mean((2, 3)), atol=1e-6)
    g = torch.randn(6, 32, device=_dev())
    assert torch.allclose(kf.avgpool_bwd(g, (3, 5)), (g / 15)[:, :, None, None].expand(6, 32, 3, 5), atol=1e-7)
    for n, B, C in [(5, 32, 10), (3, 7, 100), (4, 130, 3)]:
        logits = torch.randn(n * B, C, device=_dev()) * 3
        tgt = torch.randint(0, C, (n * B,), device=_dev())
        clamp = torch.full((n,), 1e6, device=_dev())
        clamp[0] = 0.01                                            # this client's loss is clamped: zero gradient
        loss, gl = kf.client_ce(logits, tgt, n, clamp)
        lr_ = logits.double().requires_grad_(True)
        per = F.cross_entropy(lr_, tgt, reduction="none").view(n, B).mean(1)
        obj = torch.minimum(per.clamp_min(0), clamp.double()).sum()
        (gref,) = torch.autograd.grad(obj, lr_)
        assert torch.allclose(loss.double(), per.detach(), atol=1e-5, rtol=1e-5)
        assert torch.allclose(gl.double(), gref, atol=1e-6)
        assert gl.stride(0) % 4 == 0 and float(gl[0].abs().max()) == 0.0
        U = torch.zeros(n, 2 * C + 8, device=_dev())
        kf.client_colsum(gl, n, U[:, 4:4 + C], -0.5)
        assert torch.allclose(U[:, 4:4 + C].double(), -0.5 * gref.view(n, B, C).sum(1), atol=1e-6)
    w = torch.randn(64, 147, device=_dev())
    wp = kf.pad_rows(w)
    assert wp.shape == (64, 148) and torch.equal(wp[:, :147], w) and float(wp[:, 147].abs().max()) == 0.0


def _per_client_rows(model, X, y, lr):
    rows = []
    for c in range(X.shape[0]):
        m = copy.deepcopy(model)
        m.train()
        loss = F.cross_entropy(m(X[c]), y[c])
        g = torch.autograd.grad(loss, [p for p in m.parameters()])
        rows.append(torch.cat([-lr * t.reshape(-1) for t in g]))
    return torch.stack(rows)


def _step_rows(model, X, y, lr):
    from blades_b200.engine import batched as cb
    from blades_b200.engine.flat import FlatParams
    n, B = X.shape[:2]
    flat = FlatParams(model)
    ld = (flat.numel + 63) // 64 * 64
    U = torch.zeros(n, ld, device=X.device)[:, :flat.numel]
    sink = cb.GradSink(U, flat.specs, n, alpha=-lr)
    model.train()
    losses = cb.batched_step(model, sink, X.reshape((n * B,) + tuple(X.shape[2:])), y.reshape(-1), n,
                             torch.full((n,), 1e6, device=X.device))
    missing = [s.name for s in flat.specs if s.name not in sink.written]
    return flat.to_reference_order(U).double().cpu(), losses, missing, flat


@pytest.mark.parametrize("arch", ["resnet18", "bottleneck"])
def test_fused_resnet_schedule_matches_fp64_and_generic_path(arch, monkeypatch):
    """The explicit all-own-kernels ResNet step vs (a) the fp64 per-client truth and (b) the swapped-forward autograd
    pass of the same engine; and: no library kernel is launched by it."""
    from blades_b200.engine import resnet_fused as rf
    from blades_b200.models.resnet import Bottleneck, ResNet, resnet18
    from blades_b200.ops import conv as kc
    torch.manual_seed(0)
    model = resnet18(10) if arch == "resnet18" else ResNet(Bottleneck, [1, 1, 1, 1], num_classes=20)
    n, B, lr = 4, 32, 0.1
    Xc = torch.randn(n, B, 3, 32, 32)
    yc = torch.randint(0, 10, (n, B))
    truth = _per_client_rows(copy.deepcopy(model).double(), Xc.double(), yc, lr)
    X, y = Xc.to(_dev()), yc.to(_dev())
    torch_rows = _per_client_rows(copy.deepcopy(model).to(_dev()), X, y, lr).double().cpu()
    calls = {"n": 0}
    orig = rf.step

    def counted(*a, **k):
        calls["n"] += 1
        return orig(*a, **k)
    monkeypatch.setattr(rf, "step", counted)
    ours, losses, missing, flat = _step_rows(copy.deepcopy(model).to(_dev()), X, y, lr)
    assert calls["n"] == 1 and not missing
    monkeypatch.setattr(kc, "ENABLED", False)
    generic, losses_g, _, _ = _step_rows(copy.deepcopy(model).to(_dev()), X, y, lr)
    monkeypatch.setattr(kc, "ENABLED", True)
    assert calls["n"] == 1
    e_torch = ((torch_rows - truth).norm() / truth.norm()).item()
    e_ours = ((ours - truth).norm() / truth.norm()).item()
    e_gen = ((generic - truth).norm() / truth.norm()).item()
    worst = []
    for sp in flat.specs:
        sl = slice(sp.offset, sp.offset + sp.numel)
        worst.append((((ours[:, sl] - truth[:, sl]).norm() / (truth[:, sl].norm() + 1e-30)).item(), sp.name))
    worst.sort(reverse=True)
    assert e_ours < 3 * e_torch + 2e-3, (e_ours, e_torch, e_gen, worst[:5])
    ref_loss = torch.stack([F.cross_entropy(copy.deepcopy(model).double().train()(Xc[c].double()), yc[c]) for c in range(n)])
    assert torch.allclose(losses.double().cpu(), ref_loss.detach(), atol=5e-3, rtol=5e-3)
    assert torch.allclose(losses.cpu(), losses_g.cpu(), atol=5e-3, rtol=5e-3)


def test_fused_resnet_step_launches_only_own_kernels():
    from torch.profiler import ProfilerActivity, profile
    from blades_b200.models import resnet18
    torch.manual_seed(1)
    model = resnet18(10).to(_dev())
    X = torch.randn(4, 16, 3, 32, 32, device=_dev())
    y = torch.randint(0, 10, (4, 16), device=_dev())
    _step_rows(model, X, y, 0.1)                       # warm-up (attribute setting, lazy inits)
    torch.cuda.synchronize()
    model2 = resnet18(10).to(_dev())
    from blades_b200.engine import batched as cb
    from blades_b200.engine.flat import FlatParams
    flat = FlatParams(model2)
    U = torch.zeros(4, (flat.numel + 63) // 64 * 64, device=_dev())[:, :flat.numel]
    sink = cb.GradSink(U, flat.specs, 4, alpha=-0.1)
    clamp = torch.full((4,), 1e6, device=_dev())
    model2.train()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        cb.batched_step(model2, sink, X.reshape(64, 3, 32, 32), y.reshape(-1), 4, clamp)
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
    foreign = [k for k in names if any(t in k for t in ("cutlass", "cudnn", "xmma", "at::native", "gemm", "Memcpy"))
               and "tcgen05" not in k]
    assert names and not foreign, foreign
    total = sum(e.count for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA)
    assert total <= 130, total


def test_fused_evaluation_matches_the_per_client_loop(tmp_path):
    """K10: evaluation of a batch-statistics ResNet through the own-kernel forward (one statistics group per reference
    evaluation batch, tail batches included) vs the reference-style per-client ``evaluate`` loop."""
    from blades_b200 import Simulator
    from blades_b200.datasets import synthetic_fldataset
    from blades_b200.models import resnet18
    ds = synthetic_fldataset(5, shape=(3, 32, 32), num_classes=10, train_bs=8, train_per_client=16, test_per_client=20,
                             seed=2, separation=1.5)
    sim = Simulator(ds, aggregator="mean", use_cuda=True, seed=1, log_path=str(tmp_path), progress=False)
    model = resnet18(10)
    sim.prepare(model, "SGD", "SGD", "crossentropy", server_lr=1.0, client_lr=0.05)
    sim.train_actor(0, 1, sim.get_clients(), 0.05)
    fused = sim._test_fused(0, 8)
    assert fused is not None and len(fused) == 5
    m = sim.server.get_model()
    for rec, c in zip(fused, sim.get_clients()):
        ref = c.evaluate(round_number=0, test_set=ds.get_all_test_data(c.id()), batch_size=8, metrics=sim.metrics,
                         use_actor=True, model=m)
        assert rec["Length"] == ref["Length"] == 20
        assert abs(rec["Loss"] - ref["Loss"]) < 2e-2 * max(1.0, abs(ref["Loss"]))
        assert abs(rec["top1"] - ref["top1"]) <= 10.0 + 1e-6           # at most 2 of 20 borderline samples flip
    m.train()


@pytest.mark.parametrize("agg,attack", [("trimmedmean", "alie"), ("median", "ipm"), ("mean", "alie"), ("mean", None)])
def test_pipelined_aggregation_equals_unpipelined(agg, attack, tmp_path, monkeypatch):
    """Whole-round graph with the aggregation of finished layers overlapping the rest of the backward pass (side
    stream, windows reported by the fused ResNet step) == the same graph with one aggregation after training."""
    from blades_b200 import Simulator
    from blades_b200.datasets import synthetic_fldataset
    from blades_b200.models import resnet18
    res, chunks = [], []
    for flag in ("1", "0"):
        monkeypatch.setenv("BLADES_AGG_PIPELINE", flag)
        ds = synthetic_fldataset(8, shape=(3, 32, 32), num_classes=10, train_bs=8, train_per_client=16, test_per_client=8,
                                 seed=2)
        akw = {"num_clients": 8, "num_byzantine": 2} if attack == "alie" else None
        sim = Simulator(ds, num_byzantine=2 if attack else 0, attack=attack, attack_kws=akw, aggregator=agg,
                        aggregator_kws={"nb": 2} if agg == "trimmedmean" else None, use_cuda=True, seed=1,
                        log_path=str(tmp_path / f"p{flag}"), progress=False)
        torch.manual_seed(3)
        m = resnet18(num_classes=10)
        sim.run(m, global_rounds=6, local_steps=1, server_lr=1.0, client_lr=0.05, validate_interval=6)
        sts = [st for st in sim.engine._round_graphs.values() if "graph" in st]
        assert sts, "round graph was not captured"
        chunks.append(sts[0].get("chunks"))
        res.append(sim.engine.gflat.theta.detach().cpu().clone())
    assert chunks[0] is not None and len(chunks[0]) >= 3, chunks[0]          # layer4+fc | layer3 | rest
    assert chunks[0][0][1] == res[0].numel() and chunks[0][-1][0] == 0
    assert all(a[0] == b[1] for a, b in zip(chunks[0][:-1], chunks[0][1:]))  # contiguous, descending
    assert chunks[1] is None
    assert torch.equal(res[0], res[1]), (res[0] - res[1]).abs().max()
