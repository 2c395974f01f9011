"""Property tests (hypothesis) of the aggregator oracles and the Gram-domain reformulations."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from blades_b200.aggregators import Centeredclipping, Geomed, Mean, Median, Trimmedmean
from blades_b200.aggregators import _gramops as gops
from blades_b200.parallel.matrix import LocalMatrix, VirtualRows

shapes = st.tuples(st.integers(3, 24), st.integers(1, 40), st.integers(0, 10_000))


def _U(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g, dtype=torch.float64)


@settings(max_examples=40, deadline=None)
@given(shapes, st.integers(0, 11))
def test_trimmed_mean_bounds_and_permutation_invariance(shape, b):
    n, d, seed = shape
    U = _U(n, d, seed)
    agg = Trimmedmean(nb=b)
    out = agg(U)
    assert (out <= U.max(0).values + 1e-12).all() and (out >= U.min(0).values - 1e-12).all()
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(seed + 1))
    assert torch.allclose(agg(U[perm]), out)
    if n - 2 * b > 0 and b > 0:          # b extreme outliers per side cannot move the result past the clean range
        V = U.clone()
        V[:b] = 1e9
        assert (agg(V) <= U[b:].max(0).values + 1e-9).all()


@settings(max_examples=40, deadline=None)
@given(shapes)
def test_median_equivariance(shape):
    n, d, seed = shape
    U = _U(n, d, seed)
    m = Median()(U)
    assert torch.allclose(Median()(3.0 * U + 2.0), 3.0 * m + 2.0)
    assert torch.allclose(Median()(-U), -m)
    assert torch.allclose(Mean()(U), U.mean(0))


@settings(max_examples=25, deadline=None)
@given(st.tuples(st.integers(6, 20), st.integers(2, 30), st.integers(0, 10_000)))
def test_gram_identities(shape):
    n, d, seed = shape
    U = _U(n, d, seed)
    G = (U @ U.T).numpy()
    D = gops.sq_dists(G)
    assert np.allclose(D, (torch.cdist(U, U) ** 2).numpy(), atol=1e-8)
    w = np.random.default_rng(seed).random(n)
    w /= w.sum()
    z = (torch.tensor(w)[:, None] * U).sum(0)
    assert np.allclose(gops.dist_to_combo(G, w), (U - z).norm(dim=1).numpy(), atol=1e-7)
    f = max(0, (n - 3) // 2 - 1)
    sel = gops.multi_krum_select(D, f, 1)
    assert 0 <= sel[0] < n
    # geometric median: translation equivariance of the Gram-domain Weiszfeld solution
    gm = Geomed(compat=False, maxiter=200)
    a = gm(U)
    b = gm(U + 5.0)
    assert torch.allclose(b, a + 5.0, atol=1e-6)


@settings(max_examples=20, deadline=None)
@given(st.tuples(st.integers(6, 16), st.integers(2, 20), st.integers(0, 10_000)), st.sampled_from(["alie", "ipm"]))
def test_virtual_rows_match_materialised(shape, kind):
    n, d, seed = shape
    U = _U(n, d, seed).float()
    f = max(1, n // 4)
    param = 0.37
    honest = U[f:]
    val = honest.mean(0) - param * honest.std(0) if kind == "alie" else -param * honest.mean(0)
    Um = U.clone()
    Um[:f] = val
    v = VirtualRows(kind, param, list(range(f)))
    for agg in (Median(), Trimmedmean(f), Mean()):
        assert torch.allclose(agg(LocalMatrix(U.clone(), v)), agg(Um), atol=1e-5)


def test_centered_clipping_contracts_towards_clean_mean():
    g = torch.Generator().manual_seed(0)
    clean = torch.randn(12, 30, generator=g, dtype=torch.float64) * 0.1
    U = torch.cat([clean, 100 + torch.randn(3, 30, generator=g, dtype=torch.float64)])
    agg = Centeredclipping(tau=1.0, n_iter=20)
    out = None
    for _ in range(5):
        out = agg(U)
    assert (out - clean.mean(0)).norm() < (U.mean(0) - clean.mean(0)).norm() * 0.5
