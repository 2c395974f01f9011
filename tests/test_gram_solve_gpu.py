"""On-device Gram solvers (csrc/cuda/gram_solve.cu) against the host solvers (numpy / C++, aggregators/_gramops.py)
on the same Gram matrix, and whole aggregators with the device path on and off."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _matrix(n, d, seed, outliers=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    u = torch.randn(n, d, generator=g) * 0.01
    if outliers:
        u[:outliers] += torch.randn(outliers, d, generator=g) * 0.5
    return u.cuda()


def _dg(u, extra=None):
    from blades_b200.parallel.matrix import LocalMatrix
    dg = LocalMatrix(u).gram_device(extra)
    assert dg is not None
    return dg


@pytest.mark.parametrize("n,f,m,sq", [(20, 5, 1, True), (20, 5, 1, False), (100, 20, 80, False), (200, 40, 160, False),
                                      (37, 3, 5, True), (300, 100, 7, False)])
def test_device_krum_matches_host(n, f, m, sq):
    from blades_b200.aggregators import _gramops as gops
    from blades_b200.ops import gram_solve
    u = _matrix(n, 20000, n + f, outliers=f)
    dg = _dg(u)
    G = dg.dense().double().cpu().numpy()
    want = gops.multi_krum_select(gops.sq_dists(G), f, m, n, squared_twice=sq)
    w = gram_solve.krum_weights(dg, n, f, m, sq, 0.5).cpu().numpy()
    assert sorted(np.nonzero(w)[0].tolist()) == sorted(want)
    assert np.all(w[np.nonzero(w)[0]] == 0.5)


@pytest.mark.parametrize("n,compounding,maxiter", [(20, True, 100), (20, False, 100), (100, True, 50), (100, False, 7),
                                                   (256, True, 20), (500, False, 10)])
def test_device_weiszfeld_matches_host(n, compounding, maxiter):
    from blades_b200.aggregators import _gramops as gops
    from blades_b200.ops import gram_solve
    u = _matrix(n, 8192, 3 * n, outliers=n // 5)
    dg = _dg(u)
    G = dg.dense().double().cpu().numpy()
    alphas = None if n % 2 else np.linspace(0.5, 1.5, n) / n
    want, it_h = gops.weiszfeld_weights(G, alphas, maxiter, 1e-6, 1e-10, compounding=compounding)
    w, it = gram_solve.weiszfeld_weights(dg, alphas, maxiter, 1e-6, 1e-10, compounding)
    np.testing.assert_allclose(w.cpu().numpy(), want, rtol=2e-5, atol=1e-9)
    assert abs(int(it.item()) - it_h) <= 1          # the stopping test sits at 1e-10 relative: fp64 summation order


@pytest.mark.parametrize("n,by_index,compounding,lamb", [(10, True, True, None), (40, False, False, 3.0), (100, True, True, None),
                                                         (100, False, True, 20.0), (300, True, False, None)])
def test_device_autogm_matches_host(n, by_index, compounding, lamb):
    from blades_b200.aggregators import _gramops as gops
    from blades_b200.ops import gram_solve
    u = _matrix(n, 8192, 13 * n, outliers=n // 5)
    dg = _dg(u)
    G = dg.dense().double().cpu().numpy()
    want = gops.autogm_weights(G, lamb, 30, 1e-6, 1e-10, sort_by_index=by_index, compounding=compounding)
    w = gram_solve.autogm_weights(dg, lamb, 30, 1e-6, 1e-10, by_index, compounding).cpu().numpy()
    np.testing.assert_allclose(w, want, rtol=5e-4, atol=1e-7)


@pytest.mark.parametrize("n,tau,iters", [(10, 10.0, 5), (100, 0.05, 5), (100, 1e-3, 1), (300, 0.5, 3)])
def test_device_centered_clip_matches_host(n, tau, iters):
    from blades_b200.aggregators import _gramops as gops
    from blades_b200.ops import gram_solve
    u = _matrix(n, 8192, 7 * n, outliers=n // 4)
    m = (torch.randn(8192) * 0.01).cuda()
    dg = _dg(u, m)
    assert dg.n == n + 1
    G = dg.dense().double().cpu().numpy()
    want = gops.centered_clip_coeffs(G, tau, iters)
    c = gram_solve.centered_clip_coeffs(dg, tau, iters).cpu().numpy()
    np.testing.assert_allclose(c, want, rtol=2e-5, atol=1e-8)


@pytest.mark.parametrize("n,t", [(10, 0), (100, 37), (300, 299)])
def test_device_fltrust_matches_host(n, t):
    from blades_b200.aggregators import _gramops as gops
    from blades_b200.ops import gram_solve
    u = _matrix(n, 8192, 5 * n, outliers=n // 4)
    dg = _dg(u)
    G = dg.dense().double().cpu().numpy()
    want = gops.fltrust_weights(G, t)
    w = gram_solve.fltrust_weights(dg, t).cpu().numpy()
    np.testing.assert_allclose(w, want, rtol=2e-5, atol=1e-9)
    assert w[t] == 0.0


def test_combine_reads_device_weights_and_skips_zero_rows():
    from blades_b200.parallel.matrix import LocalMatrix
    u = _matrix(50, 30001, 5)
    w = torch.zeros(50, device="cuda")
    w[[3, 17, 49]] = torch.tensor([0.5, -2.0, 1.25], device="cuda")
    u[7] = float("nan")                       # a zero-weight row is never read, whatever it holds
    got = LocalMatrix(u.clone()).combine(w)
    want = (w.double()[[3, 17, 49], None] * u.double()[[3, 17, 49]]).sum(0)
    torch.testing.assert_close(got.double(), want, rtol=2e-5, atol=1e-8)
    extra = torch.randn(30001, device="cuda")
    w2 = torch.cat([w, torch.tensor([0.75], device="cuda")])
    got2 = LocalMatrix(u.clone()).combine(w2, extra=extra)
    torch.testing.assert_close(got2.double(), want + 0.75 * extra.double(), rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("name,kws", [("krum", dict(num_clients=30, num_byzantine=6)),
                                      ("multikrum", dict(num_byzantine=6)),
                                      ("geomed", dict(maxiter=50)), ("geomed", dict(maxiter=50, compat=False)),
                                      ("centeredclipping", dict(tau=0.05, n_iter=3)),
                                      ("fltrust", dict(trusted_index=4)), ("autogm", dict(maxiter=20))])
def test_aggregators_device_solve_equals_host_solve(name, kws, monkeypatch):
    import blades_b200.aggregators as A
    cls = {k.lower(): v for k, v in vars(A).items() if isinstance(v, type)}[name]
    u = _matrix(30, 50000, 11, outliers=6)
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("BLADES_DEVICE_SOLVE", mode)
        agg = cls(**kws)
        res = [agg(u.clone()).clone() for _ in range(2)]       # two calls: stateful aggregators (momentum)
        outs[mode] = res[-1]
    torch.testing.assert_close(outs["1"], outs["0"], rtol=2e-4, atol=2e-7)
