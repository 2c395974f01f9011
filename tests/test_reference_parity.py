"""Parity against the UNMODIFIED reference code (installed under baseline/_ref by baseline/install_ref.sh; ``ray`` and
``torch._six`` shimmed from outside the tree, baseline/ref_arm.py).  Skipped when the reference is not installed."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "blades")),
                                reason="reference not installed under baseline/_ref")


@pytest.fixture(scope="module")
def ref():
    from baseline import ref_arm
    ref_arm.import_reference(0)
    import importlib
    return importlib


def _updates(n=12, d=40, seed=0, outliers=3):
    g = torch.Generator().manual_seed(seed)
    U = torch.randn(n, d, generator=g)
    U[:outliers] += 8.0
    return U


def _ref_agg(ref, name, **kw):
    mod = ref.import_module(f"blades.aggregators.{name}")
    return getattr(mod, name.capitalize())(**kw)


def _our_agg(name, **kw):
    import importlib
    mod = importlib.import_module(f"blades_b200.aggregators.{name}")
    return getattr(mod, name.capitalize())(**kw)


@pytest.mark.parametrize("name,kw", [
    ("mean", {}), ("median", {}), ("trimmedmean", {"nb": 3}), ("trimmedmean", {"nb": 7}),
    ("krum", {"num_clients": 12, "num_byzantine": 3}), ("geomed", {}), ("geomed", {"maxiter": 5}),
    ("autogm", {"lamb": 1.0}), ("autogm", {}),
])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_aggregator_matches_reference(ref, name, kw, seed):
    U = _updates(seed=seed)
    want = _ref_agg(ref, name, **kw)([u.clone() for u in U])
    got = _our_agg(name, **kw)([u.clone() for u in U])
    assert got.shape == want.shape
    assert torch.allclose(got.double(), want.double(), rtol=2e-4, atol=2e-5), (got - want).abs().max()


@pytest.mark.parametrize("name,kw", [("clustering", {}), ("clippedclustering", {}), ("clippedclustering", {"tau": 2.0})])
def test_clustering_matches_reference(ref, name, kw):
    """The reference passes ``affinity=`` to sklearn's AgglomerativeClustering (removed in sklearn 1.4, quirk Q7):
    accept the old kwarg for the duration of the reference call."""
    import sklearn.cluster as skc
    orig = skc.AgglomerativeClustering

    def Compat(*a, affinity=None, **k):
        if affinity is not None:
            k["metric"] = affinity
        return orig(*a, **k)

    mod = ref.import_module(f"blades.aggregators.{name}")
    mod.AgglomerativeClustering = Compat
    try:
        U = _updates(n=14, d=30, seed=3, outliers=4)
        r = getattr(mod, name.capitalize())(**kw)
        o = _our_agg(name, **kw)
        for rnd in range(3):              # Clippedclustering keeps a norm history across rounds
            Ur = U * (1 + 0.3 * rnd)
            want = r([u.clone() for u in Ur])
            got = o([u.clone() for u in Ur])
            assert torch.allclose(got.double(), want.double(), rtol=2e-4, atol=2e-5), (rnd, (got - want).abs().max())
    finally:
        mod.AgglomerativeClustering = orig


class _C:
    """Minimal stand-in for a client as the reference's clients-only aggregators use it (Q16)."""

    def __init__(self, u, trusted=False):
        self.u, self.t = u, trusted

    def get_update(self):
        return self.u

    def is_trusted(self):
        return self.t


def test_centeredclipping_matches_reference_over_rounds(ref):
    r = _ref_agg(ref, "centeredclipping", tau=3.0, n_iter=4)
    o = _our_agg("centeredclipping", tau=3.0, n_iter=4)
    for rnd in range(4):                  # stateful momentum
        U = _updates(seed=10 + rnd)
        want = r([_C(u.clone()) for u in U])
        got = o([_C(u.clone()) for u in U])
        assert torch.allclose(got.double(), want.double(), rtol=2e-4, atol=2e-5), rnd


def test_fltrust_matches_reference(ref):
    U = _updates(seed=5)
    mk = lambda: [_C(u.clone(), trusted=(i == len(U) - 1)) for i, u in enumerate(U)]      # noqa: E731
    want = _ref_agg(ref, "fltrust")(mk())
    got = _our_agg("fltrust")(mk())
    assert torch.allclose(got.double(), want.double(), rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("n,f", [(10, 4), (20, 5), (20, 8), (100, 10), (100, 20), (512, 100)])
def test_alie_zmax_matches_reference(ref, n, f):
    mod = ref.import_module("blades.attackers.alieclient")
    from blades_b200.attackers.alieclient import AlieClient
    want = mod.AlieClient(num_clients=n, num_byzantine=f).z_max
    got = AlieClient(num_clients=n, num_byzantine=f).z_max
    assert abs(float(got) - float(want)) < 1e-6


def _run_both(tmp, attack, attack_kws, agg, agg_kws, rounds, local_steps, n=6, f=2, bs=8, seed=3):
    """Same data (the reference's cache file is read back by our BaseDataset), same seed, same API calls:
    returns (reference parameters, our parameters) after ``rounds`` rounds."""
    import pickle
    from baseline import ref_arm
    rs = ref_arm.import_reference(0)
    from blades.models.mnist import MLP as RefMLP
    ds_ref = ref_arm.make_dataset(n, bs, os.path.join(tmp, "ref"), shape=(28, 28))
    kw = dict(num_byzantine=f if attack else 0, attack=attack, attack_kws=attack_kws, aggregator=agg,
              aggregator_kws=agg_kws, use_cuda=False, seed=seed)
    run_kw = dict(global_rounds=rounds, local_steps=local_steps, validate_interval=1000, server_lr=1.0, client_lr=0.1)
    sim_r = rs.Simulator(dataset=ds_ref, num_actors=1, log_path=os.path.join(tmp, "lr"), **kw)
    m_ref = RefMLP()
    sim_r.run(m_ref, **run_kw)

    from blades_b200 import Simulator
    from blades_b200.datasets import BaseDataset
    from blades_b200.models.mnist import MLP

    class Same(BaseDataset):
        compat = True                     # the reference's batch order (global numpy RNG seeded per stream)

        def generate_datasets(self, path="./data", iid=True, alpha=0.1, num_clients=20, seed=1):
            with open(os.path.join(tmp, "ref", "SyntheticCIFAR10.obj"), "rb") as fh:
                _, a, b, c, d = [pickle.load(fh) for _ in range(5)]
            return a, b, c, d

    ds = Same(data_root=os.path.join(tmp, "ours"), train_bs=bs, num_clients=n, seed=1)
    sim_o = Simulator(dataset=ds, log_path=os.path.join(tmp, "lo"), progress=False, **kw)
    m = MLP()
    sim_o.run(m, **run_kw)
    flat = lambda mod: torch.cat([p.detach().reshape(-1) for p in mod.parameters()])      # noqa: E731
    return flat(m_ref), flat(m)


@pytest.mark.parametrize("attack,attack_kws,agg,agg_kws,rounds,local_steps", [
    (None, None, "mean", None, 2, 1),
    ("ipm", {"epsilon": 0.5}, "median", None, 2, 1),
    ("alie", {"num_clients": 6, "num_byzantine": 2}, "trimmedmean", {"nb": 2}, 2, 1),
    ("labelflipping", None, "krum", {"num_clients": 6, "num_byzantine": 1}, 2, 1),
    ("signflipping", None, "geomed", None, 2, 1),
    ("ipm", {"epsilon": 2.0}, "autogm", {"lamb": 2.0}, 2, 1),
    ("alie", {"num_clients": 6, "num_byzantine": 2}, "centeredclipping", None, 2, 1),
    ("ipm", {"epsilon": 0.5}, "mean", None, 1, 2),            # fedavg: two local steps
    ("ipm", {"epsilon": 0.5}, "median", None, 5, 1),          # crosses epoch boundaries (2 batches per epoch)
    ("alie", {"num_clients": 6, "num_byzantine": 2}, "trimmedmean", {"nb": 2}, 2, 3),
])
def test_simulation_matches_reference_end_to_end(ref, tmp_path, attack, attack_kws, agg, agg_kws, rounds, local_steps):
    """Whole simulations through both public APIs end at the same global model (MLP, CPU)."""
    want, got = _run_both(str(tmp_path), attack, attack_kws, agg, agg_kws, rounds, local_steps)
    assert torch.isfinite(want).all()
    err = (got - want).abs().max().item()
    assert err <= 1e-5 * max(1.0, want.abs().max().item()), err
