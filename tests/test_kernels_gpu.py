"""Numerics of the hand-written sm_100a kernels vs. plain PyTorch fp32/fp64 references."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


@pytest.fixture(scope="module", autouse=True)
def _lib():
    from blades_b200.ops import _loader, _structs
    lib = _loader.cuda_lib()
    _structs.verify(lib)
    return lib


def _trim_ref(U, b):
    s = U.double().sort(0).values
    return s[b: U.shape[0] - b].mean(0).float()


def _med_ref(U):
    n = U.shape[0]
    s = U.double().sort(0).values
    return ((s[(n - 1) // 2] + s[n // 2]) / 2).float()


@pytest.mark.parametrize("n,d,b", [(3, 1000, 1), (8, 4097, 2), (20, 10007, 5), (100, 50001, 20),
                                   (128, 3001, 30), (77, 12345, 0), (129, 2000, 10), (200, 3003, 40),
                                   (512, 1500, 100)])
def test_trimmed_mean_and_median(n, d, b):
    from blades_b200.ops import select
    g = torch.Generator(device="cuda").manual_seed(n * 7 + d)
    U = torch.randn(n, d, device=_dev(), generator=g)
    U[0, :5] = float("nan")
    U[1, 5:9] = float("inf")
    U[2, 9:12] = -float("inf")
    ref_in = torch.nan_to_num(U)
    out = select.trimmed_mean(U, b)
    assert torch.allclose(out, _trim_ref(ref_in, b), atol=1e-5, rtol=1e-5), (out - _trim_ref(ref_in, b)).abs().max()
    out = select.median(U)
    assert torch.allclose(out, _med_ref(ref_in), atol=1e-6, rtol=1e-6)


@pytest.mark.parametrize("kind,param", [("alie", 0.2858), ("ipm", 0.5)])
@pytest.mark.parametrize("n,f,d,b", [(10, 3, 5003, 3), (100, 20, 20011, 20), (100, 10, 7001, 5), (200, 40, 3000, 40)])
def test_fused_virtual_rows(kind, param, n, f, d, b):
    from blades_b200.ops import select
    from blades_b200.parallel.matrix import VirtualRows
    U = torch.randn(n, d, device=_dev())
    v = VirtualRows(kind, param, list(range(f)))
    honest = U[f:].double()
    val = honest.mean(0) - param * honest.std(0) if kind == "alie" else -param * honest.mean(0)
    Um = U.clone()
    Um[:f] = val.float()
    assert torch.allclose(select.trimmed_mean(U, b, virtual=v), _trim_ref(Um, b), atol=2e-5, rtol=1e-4)
    assert torch.allclose(select.median(U, virtual=v), _med_ref(Um), atol=2e-5, rtol=1e-4)


def test_strided_rows_and_partial_byzantine():
    from blades_b200.ops import select
    from blades_b200.parallel.matrix import VirtualRows
    store = torch.randn(12, 4096 + 64, device=_dev())
    U = store[:, :4001]
    assert torch.allclose(select.trimmed_mean(U, 2), _trim_ref(U, 2), atol=1e-5)
    # 4 byzantine clients, only 2 of them replaced by virtual rows
    v = VirtualRows("alie", 0.3, [0, 1], byzantine=[0, 1, 2, 3])
    honest = U[4:].double()
    Um = U.clone()
    Um[:2] = (honest.mean(0) - 0.3 * honest.std(0)).float()
    assert torch.allclose(select.trimmed_mean(U, 2, virtual=v), _trim_ref(Um, 2), atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("n,d", [(1, 100), (7, 4099), (100, 100003), (513 - 1, 2048)])
def test_row_combine(n, d):
    from blades_b200.ops import combine
    U = torch.randn(n, d, device=_dev())
    w = torch.rand(n, dtype=torch.float64)
    w[::3] = 0
    ref = (w.to(U.device)[:, None] * U.double()).sum(0).float()
    assert torch.allclose(combine.row_combine(U, w), ref, atol=1e-4, rtol=1e-4)
    extra = torch.randn(d, device=_dev())
    got = combine.row_combine(U, w, extra, 0.25)
    assert torch.allclose(got, ref + 0.25 * extra, atol=1e-4, rtol=1e-4)
    # unaligned row stride -> scalar path
    store = torch.randn(n, d + 3, device=_dev())
    V = store[:, 1:d + 1]
    ref = (w.to(U.device)[:, None] * V.double()).sum(0).float()
    assert torch.allclose(combine.row_combine(V, w), ref, atol=1e-4, rtol=1e-4)


def test_fill_normal_statistics():
    from blades_b200.ops import attack
    row = torch.empty(1 << 20, device=_dev())
    attack.fill_normal_(row, 0.1, 0.5, seed=123)
    assert abs(row.mean().item() - 0.1) < 5e-3 and abs(row.std().item() - 0.5) < 5e-3
    k = ((row - 0.1) / 0.5)
    assert abs((k ** 4).mean().item() - 3.0) < 0.1          # Gaussian kurtosis
    row2 = torch.empty(1 << 20, device=_dev())
    attack.fill_normal_(row2, 0.1, 0.5, seed=123)
    assert not torch.equal(row, row2)                        # counter advances


def test_attack_rows_kernel():
    from blades_b200.ops import attack, select
    U = torch.randn(30, 9001, device=_dev())
    honest = U[6:].double()
    attack.attack_rows(select.row_pointers(U, range(6, 30)), select.row_pointers(U, range(6)), "alie", 0.4,
                       0, 9001, U.device)
    ref = (honest.mean(0) - 0.4 * honest.std(0)).float()
    for r in range(6):
        assert torch.allclose(U[r], ref, atol=2e-5, rtol=1e-4)


def test_fused_epilogue_server_step():
    from blades_b200.ops import select
    U = torch.randn(16, 7777, device=_dev())
    theta = torch.randn(7777, device=_dev())
    theta0 = theta.clone()
    out = torch.empty(7777, device=_dev())
    ep = select.make_epilogue([out.data_ptr()], [theta.data_ptr()], theta.data_ptr(), 0.5)
    select.launch_select(select.row_pointers(U), [], 0, None, 0.0, 0, 3, 0, 7777, ep, U.device)
    ref = _trim_ref(U, 3)
    assert torch.allclose(out, ref, atol=1e-5)
    assert torch.allclose(theta, theta0 + 0.5 * ref, atol=1e-5)


@pytest.mark.parametrize("n,d", [(10, 4096), (100, 65536 + 40), (200, 30000), (512, 8192), (33, 1000)])
def test_gram(n, d):
    from blades_b200.ops import gram
    U = torch.randn(n, d, device=_dev()) * 0.01 + 0.003
    ref = (U.double() @ U.double().T).cpu().numpy()
    for prec in ("tf32x3", "tf32"):
        G = gram.gram(U, precision=prec)
        tol = 2e-6 if prec == "tf32x3" else 2e-3
        scale = np.abs(ref).max()
        assert np.abs(G - ref).max() <= tol * scale, (prec, np.abs(G - ref).max() / scale)
    extra = torch.randn(d, device=_dev()) * 0.01
    G = gram.gram(U, extra)
    full = torch.cat([U, extra[None]]).double()
    ref = (full @ full.T).cpu().numpy()
    assert G.shape == (n + 1, n + 1)
    assert np.abs(G - ref).max() <= 2e-6 * np.abs(ref).max()


def test_grouped_wgrad():
    from blades_b200.ops import wgrad
    n, M, K, N, d = 5, 64, 32, 576, 64 * 576 + 128
    a = torch.randn(n, M, K, device=_dev())
    b = torch.randn(n, K, N, device=_dev())
    U = torch.zeros(n, d, device=_dev())
    out = U[:, 64: 64 + M * N].view(n, M, N)
    wgrad.grouped_wgrad(a, b, out, -0.1)
    ref = -0.1 * (a.double() @ b.double())
    assert torch.allclose(out.double(), ref, atol=2e-2, rtol=2e-3)
    assert U[:, :64].abs().sum() == 0 and U[:, 64 + M * N:].abs().sum() == 0


def test_simulator_gpu_matches_cpu_oracle(tmp_path):
    from blades_b200 import Simulator
    from blades_b200.datasets import synthetic_fldataset
    from blades_b200.models import MLP
    res = []
    for use_cuda in (False, True):
        ds = synthetic_fldataset(10, shape=(28, 28), train_bs=8, seed=3, separation=2.0)
        sim = Simulator(ds, num_byzantine=3, attack="alie", attack_kws={"num_clients": 10, "num_byzantine": 3},
                        aggregator="trimmedmean", aggregator_kws={"nb": 3}, use_cuda=use_cuda, seed=1,
                        log_path=str(tmp_path / f"l{use_cuda}"), progress=False)
        torch.manual_seed(5)
        m = MLP()
        sim.run(m, global_rounds=3, local_steps=1, server_lr=1.0, client_lr=0.1, validate_interval=3)
        res.append(torch.cat([p.detach().cpu().reshape(-1) for p in m.parameters()]))
    assert torch.allclose(res[0], res[1], atol=2e-4, rtol=1e-3), (res[0] - res[1]).abs().max()


@pytest.mark.parametrize("agg,kws", [("median", None), ("krum", {"num_clients": 10, "num_byzantine": 2}),
                                     ("geomed", None), ("centeredclipping", None), ("clippedclustering", None),
                                     ("mean", None)])
def test_simulator_gpu_aggregators(agg, kws, tmp_path):
    from blades_b200 import Simulator
    from blades_b200.datasets import synthetic_fldataset
    from blades_b200.models import MLP
    ds = synthetic_fldataset(10, shape=(28, 28), train_bs=8, seed=3, separation=2.0)
    sim = Simulator(ds, num_byzantine=2, attack="ipm", aggregator=agg, aggregator_kws=kws, use_cuda=True,
                    seed=1, log_path=str(tmp_path / "l"), progress=False)
    m = MLP()
    sim.run(m, global_rounds=2, local_steps=2, server_lr=1.0, client_lr=0.1, validate_interval=2)
    assert all(torch.isfinite(p).all() for p in m.parameters())
