"""Aggregator oracles vs. brute-force definitions (incl. reference quirks)."""

import numpy as np
import pytest
import torch

from blades_b200.aggregators import (Autogm, Centeredclipping, Clippedclustering, Clustering, Fltrust,
                                     Geomed, Krum, Mean, Median, Multikrum, Trimmedmean)
from blades_b200.aggregators import _gramops as gops
from blades_b200.client import BladesClient
from blades_b200.parallel.matrix import LocalMatrix, VirtualRows


def rand_updates(n=12, d=40, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g, dtype=torch.float64)


def test_mean_median_trimmed():
    U = rand_updates(11, 33)
    assert torch.allclose(Mean()(U), U.mean(0))
    assert torch.allclose(Median()(U), U.median(0).values)
    Ue = rand_updates(10, 33)
    s = Ue.sort(0).values
    assert torch.allclose(Median()(Ue), (s[4] + s[5]) / 2)
    # reference formula: (median(U) - median(-U)) / 2
    ref = (Ue.median(0).values - (-Ue).median(0).values) / 2
    assert torch.allclose(Median()(Ue), ref)
    b = 3
    assert torch.allclose(Trimmedmean(nb=b)(Ue), s[b:10 - b].mean(0))
    # topk formulation of the reference
    largest = torch.topk(Ue, b, 0).values
    neg_small = torch.topk(-Ue, b, 0).values
    ref = torch.cat([Ue, -largest, neg_small]).sum(0) / (10 - 2 * b)
    assert torch.allclose(Trimmedmean(nb=b)(Ue), ref)


def test_trimmedmean_shrinks_b():   # quirk Q4
    U = rand_updates(4, 8)
    s = U.sort(0).values
    assert torch.allclose(Trimmedmean(nb=5)(U), s[1:3].mean(0))


def test_input_conventions():
    U = rand_updates(5, 9).float()
    clients = []
    for i in range(5):
        c = BladesClient(id=i)
        c.save_update(U[i])
        clients.append(c)
    for agg in (Mean(), Median(), Trimmedmean(1)):
        a = agg(U)
        assert torch.allclose(agg(list(U)), a)
        assert torch.allclose(agg(clients), a)


def _ref_krum(U, n, f, m):
    # literal transcription of the reference *semantics* (krum.py:21-25,73-90): stored
    # value is squared distance; score sums the squares of the stored values
    N = len(U)
    D = {(i, j): float((U[i] - U[j]).norm() ** 2) for i in range(N) for j in range(N) if i != j}
    scores = []
    for i in range(n):
        s = sorted(D[(i, j)] ** 2 for j in range(n) if j != i)[: n - f - 2]
        scores.append(sum(s))
    order = sorted(range(n), key=lambda i: scores[i])[:m]
    return sum(U[i] for i in order), order


@pytest.mark.parametrize("seed", range(5))
def test_krum_compat_matches_reference_semantics(seed):
    U = rand_updates(10, 8, seed)
    out = Krum(num_clients=10, num_byzantine=3)(U)
    ref, _ = _ref_krum(U, 10, 3, 1)
    assert torch.allclose(out, ref)


def test_multikrum_textbook():
    U = rand_updates(12, 16, 3)
    f, m = 3, 5
    D = torch.cdist(U, U) ** 2
    scores = []
    for i in range(12):
        row = torch.cat([D[i, :i], D[i, i + 1:]]).sort().values[: 12 - f - 2]
        scores.append(row.sum().item())
    order = np.argsort(scores, kind="stable")[:m]
    out = Multikrum(num_byzantine=f, m=m)(U)
    assert torch.allclose(out, U[order].mean(0))
    with pytest.raises(ValueError):
        Krum(num_clients=6, num_byzantine=3)(U[:6])


def _ref_geomed(U, maxiter=100, eps=1e-6, ftol=1e-10, weights=None):
    # reference algorithm (geomed.py:61-84) on explicit vectors
    n = len(U)
    w = np.ones(n) / n if weights is None else np.asarray(weights, dtype=np.float64)
    dist = lambda z, p: float((z - p).norm())
    median = U.mean(0)
    obj = sum(a * dist(median, p) for a, p in zip(w, U))
    for _ in range(maxiter):
        prev = obj
        w = np.asarray([max(eps, a / max(eps, dist(median, p))) for a, p in zip(w, U)])
        w = w / w.sum()
        median = sum(p * b for p, b in zip(U, w))
        obj = sum(a * dist(median, p) for a, p in zip(w, U))
        if abs(prev - obj) < ftol * obj:
            break
    return median


def test_geomed_gram_domain_matches_direct():
    U = rand_updates(9, 20, 1)
    assert torch.allclose(Geomed()(U), _ref_geomed(U), atol=1e-9)
    U[0] += 50   # outlier
    out = Geomed()(U)
    assert torch.allclose(out, _ref_geomed(U), atol=1e-8)
    assert (out - U[1:].mean(0)).norm() < (U.mean(0) - U[1:].mean(0)).norm()


def test_geomed_textbook_minimises_objective():
    U = rand_updates(15, 6, 2)
    z = Geomed(compat=False, maxiter=500)(U)
    obj = lambda v: (U - v).norm(dim=1).sum()
    for _ in range(20):
        assert obj(z) <= obj(z + 1e-3 * torch.randn_like(z)) + 1e-9


def _ref_autogm(U, lamb=None, maxiter=100, eps=1e-6, ftol=1e-10):
    n = len(U)
    lamb = 1.0 * n if lamb is None else lamb
    alpha = np.ones(n) / n
    median = _ref_geomed(U, maxiter, eps, ftol, alpha)
    dist = lambda z, p: float((z - p).norm())
    obj = sum(a * dist(median, p) for a, p in zip(alpha, U))
    glob = obj + lamb * np.linalg.norm(alpha) ** 2 / 2
    distance = np.zeros(n)
    for _ in range(maxiter):
        prev = glob
        for i, p in enumerate(U):
            distance[i] = dist(p, median)
        idxs = list(range(n))          # Q6: sorted by index
        eta_opt = 1e16
        for p in range(n):
            eta = (sum(distance[i] for i in idxs[:p + 1]) + lamb) / (p + 1)
            if eta - distance[idxs[p]] < 0:
                break
            eta_opt = eta
        alpha = np.array([max(eta_opt - d, 0) / lamb for d in distance])
        median = _ref_geomed(U, maxiter, eps, ftol, alpha)
        gm = sum(a * dist(median, p) for a, p in zip(alpha, U))
        glob = gm + lamb * np.linalg.norm(alpha) ** 2 / 2
        if abs(prev - glob) < ftol * glob:
            break
    return median


def test_autogm_matches_direct():
    U = rand_updates(8, 10, 4)
    U[0] += 10
    assert torch.allclose(Autogm(lamb=2.0)(U), _ref_autogm(U, 2.0), atol=1e-8)
    assert torch.allclose(Autogm()(U), _ref_autogm(U), atol=1e-8)


def test_centered_clipping_stateful():
    U = rand_updates(7, 12, 5)
    agg = Centeredclipping(tau=1.5, n_iter=3)
    m = torch.zeros(12, dtype=torch.float64)
    for rnd in range(3):
        Ur = U + rnd
        for _ in range(3):
            m = sum((v - m) * min(1.0, 1.5 / float((v - m).norm())) for v in Ur) / len(Ur) + m
        out = agg(Ur)
        assert torch.allclose(out, m, atol=1e-9), rnd
    st = agg.state_dict()
    agg2 = Centeredclipping(tau=1.5, n_iter=3)
    agg2.load_state_dict(st)
    assert torch.allclose(agg2(U), agg(U))


def test_clustering_majority():
    g = torch.Generator().manual_seed(0)
    good = torch.randn(7, 30, generator=g, dtype=torch.float64) * 0.1 + 1.0
    bad = torch.randn(3, 30, generator=g, dtype=torch.float64) * 0.1 - 1.0
    U = torch.cat([bad, good])
    out = Clustering(compat=False)(U)
    assert torch.allclose(out, good.mean(0))
    cc = Clippedclustering()
    out = cc(U)
    norms = U.norm(dim=1)
    thr = norms.median()      # numpy median of 10 = mean of middle two
    thr = float(np.median(norms.numpy()))
    scale = torch.where(norms > thr, torch.clamp(thr / (norms + 1e-6), max=1.0), torch.ones_like(norms))
    ref = (good * scale[3:, None]).mean(0)
    assert torch.allclose(out, ref)
    assert len(cc.l2norm_his) == 10
    cc(U)
    assert len(cc.l2norm_his) == 20


def test_complete_linkage_against_sklearn():
    sk = pytest.importorskip("sklearn.cluster")
    rng = np.random.default_rng(0)
    for _ in range(5):
        X = np.concatenate([rng.normal(0, 1, (6, 3)), rng.normal(4, 1, (5, 3))])
        D = np.linalg.norm(X[:, None] - X[None], axis=-1)
        mine = gops.complete_linkage_2(D)
        ref = sk.AgglomerativeClustering(metric='precomputed', linkage='complete', n_clusters=2).fit(D).labels_
        assert (mine == ref).all() or (mine == 1 - ref).all()


def test_fltrust():
    U = rand_updates(6, 15, 7).float()
    clients = []
    for i in range(6):
        c = BladesClient(id=i)
        c.save_update(U[i])
        clients.append(c)
    clients[2].trust()
    out = Fltrust()(clients)
    t = U[2]
    others = [U[i] for i in range(6) if i != 2]
    ts = torch.tensor([max(0.0, float(torch.nn.functional.cosine_similarity(t, u, dim=0))) for u in others])
    pg = torch.stack([u * t.norm() / u.norm() for u in others])
    ref = (pg.T @ ts) / ts.sum()
    assert torch.allclose(out, ref, atol=1e-5)
    assert torch.allclose(Fltrust(trusted_index=2)(U), ref, atol=1e-5)


def test_virtual_rows_equal_materialised():
    U = rand_updates(10, 25, 8).float()
    f = 3
    for kind, param in (("alie", 0.43), ("ipm", 0.5)):
        honest = U[f:]
        val = honest.mean(0) - param * honest.std(0) if kind == "alie" else -param * honest.mean(0)
        Um = U.clone()
        Um[:f] = val
        v = VirtualRows(kind, param, list(range(f)))
        for agg in (Mean(), Median(), Trimmedmean(2), Krum(10, 3), Geomed()):
            a = agg(LocalMatrix(U.clone(), v))
            b = agg(Um)
            assert torch.allclose(a, b, atol=1e-5), (kind, agg)


def test_toy_scene():
    """60 benign N(0, 20I) + 40 outliers N((30,30), 60I) -- examples/plot_comparing_aggregation_schemes."""
    rng = np.random.RandomState(1)
    benign = rng.multivariate_normal([0, 0], 20 * np.eye(2), 60)
    out = rng.multivariate_normal([30, 30], 60 * np.eye(2), 40)
    U = torch.tensor(np.concatenate([benign, out]))
    res = {
        "mean": Mean()(U), "krum": Krum(100, 40)(U), "geomed": Geomed()(U), "median": Median()(U),
        "autogm": Autogm(lamb=1.0)(U), "trimmed": Trimmedmean(nb=40)(U),
        "clippedclustering": Clippedclustering()(U),
    }
    assert res["mean"].norm() > 10          # dragged by outliers
    for k in ("krum", "geomed", "median", "autogm", "trimmed", "clippedclustering"):
        assert res[k].norm() < 10, (k, res[k])


def test_autogm_textbook_mode_is_permutation_invariant_and_robust():
    """Quirk Q6: the reference runs the water-filling in CLIENT order (its sort key is the index).  compat=False
    sorts by distance, which makes the result independent of the order of the clients."""
    from blades_b200.aggregators import Autogm
    g = torch.Generator().manual_seed(4)
    honest = torch.randn(14, 30, generator=g)
    U = torch.cat([honest, honest.mean(0, keepdim=True) + 40.0 + torch.randn(6, 30, generator=g)])
    perm = torch.randperm(len(U), generator=g)
    a = Autogm(lamb=2.0, compat=False)(U)
    b = Autogm(lamb=2.0, compat=False)(U[perm])
    assert torch.allclose(a, b, atol=1e-5)
    assert (a - honest.mean(0)).norm() < 0.2 * (U.mean(0) - honest.mean(0)).norm()
    # the compat result is a valid robust aggregate too, but it depends on the client order for some inputs
    c = Autogm(lamb=2.0, compat=True)(U)
    assert (c - honest.mean(0)).norm() < 0.5 * (U.mean(0) - honest.mean(0)).norm()


@pytest.mark.parametrize("native", [True, False])
def test_complete_linkage_equals_sklearn_label_for_label(native, monkeypatch):
    """600 random matrices -- duplicate rows (an omniscient attacker's clients), distances rounded to integers (masses
    of exact ties), cosine similarity used as a distance (quirk Q7, negative values): the labels equal sklearn's, not
    just the partition.  The reference's majority rule falls back to label 0 when the clusters tie in size."""
    sk = pytest.importorskip("sklearn.cluster")
    monkeypatch.setattr(gops, "USE_NATIVE", native)
    rng = np.random.default_rng(7)
    for trial in range(600):
        n = int(rng.integers(2, 24))
        X = rng.standard_normal((n, 4))
        if trial % 3 == 0 and n > 2:
            X[: int(rng.integers(2, n))] = X[0]
        D = np.linalg.norm(X[:, None] - X[None], axis=-1)
        if trial % 5 == 0:
            D = np.round(D, 0)
        if trial % 7 == 0:
            Xn = X / np.maximum(np.linalg.norm(X, axis=1, keepdims=True), 1e-12)
            D = Xn @ Xn.T
        ref = sk.AgglomerativeClustering(metric='precomputed', linkage='complete', n_clusters=2).fit(D).labels_
        assert np.array_equal(gops.complete_linkage_2(D), ref), trial
