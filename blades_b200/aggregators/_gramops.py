"""Geometry-based aggregation solved on the Gram matrix ``G = U U^T`` (N x N).

Every distance / cosine / norm any reference aggregator computes is a function of
``G``:

    ||u_i - u_j||^2          = G_ii + G_jj - 2 G_ij
    cos(u_i, u_j)            = G_ij / sqrt(G_ii G_jj)
    ||sum_j w_j u_j - u_i||^2 = w^T G w - 2 (G w)_i + G_ii

so Krum, GeoMed (Weiszfeld), AutoGM, centered clipping, (clipped) clustering and
FLTrust each need ONE tensor-core Gram pass (ops.gram, tcgen05) plus ONE weighted
row-combine pass over the 4-50 GB of updates, instead of up to ``maxiter`` x 2
full passes in the reference (geomed.py:71-82).  This file is the host side:
float64 numpy on N x N (N <= 512) -- "K6" in SURVEY 2.7.  The hot loops have a
C++ twin in csrc/host/selectors.cpp (bound through ops.host), used when built.

Each solver returns a weight vector ``w`` such that the aggregate is
``sum_i w_i u_i`` (fed to ``UpdateMatrix.combine``).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

#: set False to force the numpy implementations (tests compare both)
USE_NATIVE = True


def _native():
    if not USE_NATIVE:
        return None
    try:
        from ..ops import host
        return host if host.available() else None
    except Exception:
        return None


__all__ = [
    "sq_dists", "dist_to_combo", "krum_scores", "multi_krum_select", "weiszfeld_weights",
    "autogm_weights", "centered_clip_coeffs", "cosine_matrix", "complete_linkage_2",
    "majority_cluster", "fltrust_weights",
]


def sq_dists(G: np.ndarray) -> np.ndarray:
    """Pairwise squared distances from a Gram matrix (clamped at 0)."""
    g = np.diag(G)
    D = g[:, None] + g[None, :] - 2.0 * G
    np.fill_diagonal(D, 0.0)
    return np.maximum(D, 0.0)


def dist_to_combo(G: np.ndarray, w: np.ndarray) -> np.ndarray:
    """``||sum_j w_j u_j - u_i||`` for every i."""
    Gw = G @ w
    sq = float(w @ Gw) - 2.0 * Gw + np.diag(G)
    return np.sqrt(np.maximum(sq, 0.0))


# ----------------------------------------------------------------------------- Krum
def krum_scores(D: np.ndarray, f: int, n: Optional[int] = None, squared_twice: bool = False) -> np.ndarray:
    """score_i = sum of the ``n-f-2`` smallest entries of row i (self excluded).

    ``squared_twice`` reproduces the reference quirk Q3 (krum.py:21-25,89): the
    stored pairwise value is already a squared distance and gets squared again.
    """
    n = D.shape[0] if n is None else n
    k = n - f - 2
    nat = _native()
    if nat is not None:
        return nat.krum_scores(D[:n, :n], f, squared_twice)
    M = D[:n, :n].copy()
    if squared_twice:
        M = M ** 2
    np.fill_diagonal(M, np.inf)
    part = np.sort(M, axis=1)[:, :max(k, 0)]
    return part.sum(axis=1)


def multi_krum_select(D: np.ndarray, f: int, m: int = 1, n: Optional[int] = None,
                      squared_twice: bool = False) -> List[int]:
    """Indices of the ``m`` best-scoring rows (stable order, like ``sorted``)."""
    n = D.shape[0] if n is None else n
    if n < 1:
        raise ValueError(f"Number of workers should be positive integer. Got {n}.")
    if m < 1 or m > n:
        raise ValueError(f"Number of workers for aggregation should be >=1 and <= {n}. Got {m}.")
    if 2 * f + 2 > n:
        raise ValueError(f"Too many Byzantine workers: 2 * {f} + 2 >= {n}.")
    scores = krum_scores(D, f, n, squared_twice)
    order = np.argsort(scores, kind="stable")
    return [int(i) for i in order[:m]]


# ----------------------------------------------------------------------------- GeoMed
def weiszfeld_weights(G: np.ndarray, alphas: Optional[np.ndarray] = None, maxiter: int = 100,
                      eps: float = 1e-6, ftol: float = 1e-10, compounding: bool = True
                      ) -> Tuple[np.ndarray, int]:
    """Weiszfeld iterations in the Gram domain; returns (weights, iterations).

    The median iterate is always ``z = sum_j w_j u_j``; we track ``w`` only.
    ``compounding=True`` is the reference behaviour Q5 (geomed.py:71-77): new
    weights are derived from the *previous weights* (w <- max(eps, w/max(eps,dist)))
    and the objective uses those running weights.  ``compounding=False`` is the
    textbook algorithm (weights always derived from the original ``alphas``).
    """
    n = G.shape[0]
    alphas = np.full(n, 1.0 / n) if alphas is None else np.asarray(alphas, dtype=np.float64).copy()
    nat = _native()
    if nat is not None:
        return nat.weiszfeld(G, alphas, maxiter, eps, ftol, compounding)
    # starting point: plain mean of the rows (geomed.py:66)
    w = np.full(n, 1.0 / n)
    run = alphas.copy()          # the "weights" variable of the reference
    dist = dist_to_combo(G, w)
    obj = float(run @ dist)
    it = 0
    for it in range(1, maxiter + 1):
        prev_obj = obj
        base = run if compounding else alphas
        new = np.maximum(eps, base / np.maximum(eps, dist))
        new = new / new.sum()
        run = new
        w = new
        dist = dist_to_combo(G, w)
        obj = float(run @ dist)
        if abs(prev_obj - obj) < ftol * obj:
            break
    return w, it


def autogm_weights(G: np.ndarray, lamb: Optional[float] = None, maxiter: int = 100, eps: float = 1e-6,
                   ftol: float = 1e-10, sort_by_index: bool = True, compounding: bool = True) -> np.ndarray:
    """AutoGM (reference autogm.py:36-65) on the Gram matrix.

    ``sort_by_index=True`` reproduces quirk Q6: the water-filling visits clients in
    index order (``sorted(enumerate(d), key=lambda x: x)`` sorts by index).
    """
    n = G.shape[0]
    lamb = float(n) if lamb is None else float(lamb)
    nat = _native()
    if nat is not None:
        return nat.autogm(G, lamb, maxiter, eps, ftol, sort_by_index, compounding)
    alpha = np.full(n, 1.0 / n)
    w, _ = weiszfeld_weights(G, alpha, maxiter, eps, ftol, compounding)
    dist = dist_to_combo(G, w)
    glob = float(alpha @ dist) + lamb * float(alpha @ alpha) / 2.0
    for _ in range(maxiter):
        prev = glob
        order = np.arange(n) if sort_by_index else np.argsort(dist, kind="stable")
        eta_opt = 1e16
        csum = 0.0
        for p, idx in enumerate(order):
            csum += dist[idx]
            eta = (csum + lamb) / (p + 1)
            if eta - dist[idx] < 0:
                break
            eta_opt = eta
        alpha = np.maximum(eta_opt - dist, 0.0) / lamb
        w, _ = weiszfeld_weights(G, alpha, maxiter, eps, ftol, compounding)
        dist = dist_to_combo(G, w)
        glob = float(alpha @ dist) + lamb * float(alpha @ alpha) / 2.0
        if abs(prev - glob) < ftol * glob:
            break
    return w


# ----------------------------------------------------------------------------- centered clipping
def centered_clip_coeffs(G_aug: np.ndarray, tau: float, n_iter: int) -> np.ndarray:
    """Centered clipping (reference centeredclipping.py:30-44) in the Gram domain.

    ``G_aug`` is the Gram matrix of ``[u_0..u_{N-1}, m_prev]`` (N+1 rows).  The
    momentum is represented by coefficients ``c`` over those N+1 vectors, starting
    at ``e_N`` (= m_prev).  Each iteration:
        m <- m + (1/N) sum_i min(1, tau/||u_i - m||) (u_i - m)
    Returns the final coefficient vector (length N+1).
    """
    n = G_aug.shape[0] - 1
    nat = _native()
    if nat is not None:
        return nat.centered_clip(G_aug, tau, n_iter)
    c = np.zeros(n + 1)
    c[n] = 1.0
    for _ in range(n_iter):
        dist = dist_to_combo(G_aug, c)[:n]
        with np.errstate(divide="ignore"):
            scale = np.minimum(1.0, tau / dist)   # dist==0 -> inf -> min gives 1
        scale = np.where(np.isnan(scale), 1.0, scale)
        new = c * (1.0 - scale.sum() / n)
        new[:n] += scale / n
        c = new
    return c


# ----------------------------------------------------------------------------- clustering
def cosine_matrix(G: np.ndarray) -> np.ndarray:
    nrm = np.sqrt(np.maximum(np.diag(G), 0.0))
    with np.errstate(divide="ignore", invalid="ignore"):
        C = G / (nrm[:, None] * nrm[None, :])
    return C


def complete_linkage_2(dist: np.ndarray) -> np.ndarray:
    """Two-cluster complete-linkage agglomeration on a precomputed 'distance' matrix.

    Equivalent to ``AgglomerativeClustering(metric='precomputed', linkage='complete',
    n_clusters=2)`` (reference clustering.py:39-40).  O(N^3) worst case, N <= 512.
    Returns an int array of 0/1 labels in sklearn's label order (``_sklearn_label_order``): callers use the cluster
    sizes (majority vote) and, on a size tie, label 0 like the reference.
    """
    n = dist.shape[0]
    if n == 1:
        return np.zeros(1, dtype=np.int64)
    nat = _native()
    if nat is not None:
        return nat.complete_linkage2(dist)
    D = np.array(dist, dtype=np.float64, copy=True)
    D = np.maximum(D, D.T)            # symmetrise (reference matrices are symmetric)
    np.fill_diagonal(D, np.inf)
    alive = np.ones(n, dtype=bool)
    member = np.arange(n)
    n_clusters = n
    while n_clusters > 2:
        sub = np.where(alive[:, None] & alive[None, :], D, np.inf)
        flat = int(np.argmin(sub))
        i, j = divmod(flat, n)
        if i > j:
            i, j = j, i
        # merge j into i; complete linkage = max of the two rows
        merged = np.maximum(D[i], D[j])
        D[i, :] = merged
        D[:, i] = merged
        D[i, i] = np.inf
        alive[j] = False
        member[member == j] = i
        n_clusters -= 1
    roots = np.unique(member)
    labels = (member != member[0]).astype(np.int64)
    assert len(roots) == 2
    return _sklearn_label_order(np.maximum(dist, dist.T), labels)


def _sklearn_label_order(dist: np.ndarray, labels: np.ndarray) -> np.ndarray:
    """Label convention of sklearn's ``AgglomerativeClustering`` (it matters when the clusters tie in size: the
    reference's majority rule then falls back to label 0): label 0 = the final cluster whose internal complete-linkage
    height (largest pairwise distance inside it) is larger; singletons have none; equal heights keep row 0 in label 0.
    Checked against sklearn on random inputs in tests/test_aggregators.py."""
    def height(mask):
        idx = np.where(mask)[0]
        return -np.inf if len(idx) < 2 else dist[np.ix_(idx, idx)][np.triu_indices(len(idx), 1)].max()
    return 1 - labels if height(labels == 1) > height(labels == 0) else labels


def majority_cluster(labels: np.ndarray) -> np.ndarray:
    """Boolean mask of the larger cluster; on a tie the reference picks label 0
    (``flag = 1 if sum(labels) > num // 2 else 0``, clustering.py:41)."""
    n = len(labels)
    flag = 1 if int(labels.sum()) > n // 2 else 0
    return labels == flag


# ----------------------------------------------------------------------------- FLTrust
def fltrust_weights(G: np.ndarray, trusted: int, eps: float = 1e-6) -> np.ndarray:
    """w_i = relu(cos(u_t,u_i)) * ||u_t||/||u_i|| / sum_j relu(cos(u_t,u_j)), w_t = 0."""
    n = G.shape[0]
    nrm = np.sqrt(np.maximum(np.diag(G), 0.0))
    cos = G[trusted] / np.maximum(nrm[trusted] * nrm, eps)   # torch CosineSimilarity eps semantics
    ts = np.maximum(cos, 0.0)
    ts[trusted] = 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        w = ts * nrm[trusted] / nrm
    w[trusted] = 0.0
    return w / ts.sum()
