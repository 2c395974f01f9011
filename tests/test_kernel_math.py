"""The arithmetic identities the round-2 kernels rely on, checked on the host with numpy (no GPU):

* select kernels: ``hi = bits(a) + bits(b) - bits(lo)`` in wrapping 32-bit integer arithmetic returns exactly the other
  operand of ``lo = min(a, b)`` -- for every pair of finite floats incl. denormals, +-0, +-FLT_MAX and equal values;
* Gram 3xTF32 with two MMAs: ``hi hi^T + hi (2 lo)^T`` symmetrised equals ``hi hi^T + hi lo^T + lo hi^T`` and recovers
  ~fp32 accuracy where a single tf32 product does not."""
import numpy as np


def _bits(x):
    return x.view(np.uint32)


def test_integer_compare_exchange_is_exact():
    rng = np.random.default_rng(0)
    special = np.array([0.0, -0.0, 1e-45, -1e-45, 1.17549435e-38, np.finfo(np.float32).max, -np.finfo(np.float32).max,
                        1.0, -1.0, 3.0000002, 3.0], dtype=np.float32)
    a = np.concatenate([rng.standard_normal(200000).astype(np.float32) * 1e-3, np.repeat(special, len(special)),
                        (rng.standard_normal(1000) * 1e30).astype(np.float32), (rng.standard_normal(1000) * 1e-41).astype(np.float32)])
    b = np.concatenate([rng.standard_normal(200000).astype(np.float32) * 1e-3, np.tile(special, len(special)),
                        (rng.standard_normal(1000) * 1e30).astype(np.float32), (rng.standard_normal(1000) * 1e-41).astype(np.float32)])
    a[:1000] = b[:1000]                                        # exact ties
    lo = np.minimum(a, b)                                      # IEEE minimum: returns one of the operands bit for bit
    with np.errstate(over="ignore"):
        hi_bits = (_bits(a) + _bits(b) - _bits(lo)).astype(np.uint32)          # wraps modulo 2^32 like the IMAD pair
    hi = hi_bits.view(np.float32)
    # lo is bit-identical to a or to b, hi is then the other one
    is_a = _bits(lo) == _bits(a)
    assert (is_a | (_bits(lo) == _bits(b))).all()
    assert (np.where(is_a, _bits(b), _bits(a)) == hi_bits).all()
    # and as values: {lo, hi} == {a, b}, hi >= lo
    assert (hi >= lo).all() and (np.maximum(a, b) == hi).all()


def _tf32_trunc(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def test_two_mma_3xtf32_gram_identity_and_accuracy():
    rng = np.random.default_rng(1)
    n, d = 24, 4096
    u = (rng.standard_normal((n, d)) * 0.01 + 0.05).astype(np.float32)        # common offset: distances cancel
    hi = _tf32_trunc(u)
    lo2 = (2.0 * (u - hi)).astype(np.float32)                                  # exact: power-of-two scale
    hi64, lo64 = hi.astype(np.float64), (u - hi).astype(np.float64)
    g2 = hi64 @ hi64.T + hi64 @ lo2.astype(np.float64).T                       # what the two MMAs accumulate
    sym = 0.5 * (g2 + g2.T)
    g3 = hi64 @ hi64.T + hi64 @ lo64.T + lo64 @ hi64.T                         # the three-product form
    assert np.allclose(sym, g3, rtol=1e-13, atol=0)
    exact = u.astype(np.float64) @ u.astype(np.float64).T
    g1 = hi64 @ hi64.T                                                         # single tf32 product
    diag = np.diag(exact)
    dist_exact = diag[:, None] + diag[None, :] - 2 * exact
    def dist(g):
        dg = np.diag(g)
        return dg[:, None] + dg[None, :] - 2 * g
    off = ~np.eye(n, dtype=bool)
    err3 = np.abs(dist(sym) - dist_exact)[off].max() / dist_exact[off].mean()
    err1 = np.abs(dist(g1) - dist_exact)[off].max() / dist_exact[off].mean()
    assert err3 < 1e-5 and err1 > 20 * err3, (err1, err3)
