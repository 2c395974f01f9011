"""Native C++ selectors (csrc/host) vs their numpy twins."""
import numpy as np
import pytest

from blades_b200.aggregators import _gramops as gops
from blades_b200.ops import host

pytestmark = pytest.mark.skipif(not host.available(), reason="_host.so not built")


def _gram(n=17, d=40, seed=0, outliers=3):
    rng = np.random.default_rng(seed)
    U = rng.standard_normal((n, d))
    U[:outliers] += 6
    return U @ U.T, U


def _both(fn):
    gops.USE_NATIVE = True
    a = fn()
    gops.USE_NATIVE = False
    try:
        b = fn()
    finally:
        gops.USE_NATIVE = True
    return a, b


@pytest.mark.parametrize("seed", range(3))
def test_native_matches_numpy(seed):
    G, U = _gram(seed=seed)
    D = gops.sq_dists(G)
    for sq in (False, True):
        a, b = _both(lambda: gops.krum_scores(D, 3, squared_twice=sq))
        assert np.allclose(a, b)
    for comp in (True, False):
        (wa, ia), (wb, ib) = _both(lambda: gops.weiszfeld_weights(G, None, 100, 1e-6, 1e-10, comp))
        assert np.allclose(wa, wb, atol=1e-12) and ia == ib
        a, b = _both(lambda: gops.autogm_weights(G, 2.0, 50, 1e-6, 1e-10, True, comp))
        assert np.allclose(a, b, atol=1e-10)
    Ga = np.zeros((18, 18))
    Ga[:17, :17] = G
    m = np.random.default_rng(1).standard_normal(40)
    Ga[17, :17] = Ga[:17, 17] = U @ m
    Ga[17, 17] = m @ m
    a, b = _both(lambda: gops.centered_clip_coeffs(Ga, 5.0, 4))
    assert np.allclose(a, b, atol=1e-12)
    dist = 1 - gops.cosine_matrix(G)
    np.fill_diagonal(dist, 0)
    a, b = _both(lambda: gops.complete_linkage_2(dist))
    assert (a == b).all()
