"""Engine tests: client-batched fedsgd == time-sliced per-client training; API contract."""
import copy

import pytest
import torch
import torch.nn as nn

from blades_b200 import ByzantineClient, Simulator
from blades_b200.datasets import synthetic_fldataset
from blades_b200.engine import batched as cb
from blades_b200.engine.flat import FlatParams
from blades_b200.models import MLP, CCTNet, resnet18
from blades_b200.models.cifar10.cctnets import cct_2_3x2_32


def per_client_reference(model, X, y, lr, clamp=1e6):
    """Plain autograd: one SGD step per client from the same weights -> delta rows."""
    rows = []
    for c in range(X.shape[0]):
        m = copy.deepcopy(model)
        m.train()
        before = torch.cat([p.detach().reshape(-1) for p in m.parameters() if p.requires_grad])
        opt = torch.optim.SGD(m.parameters(), lr=lr)
        opt.zero_grad()
        loss = torch.clamp(nn.functional.cross_entropy(m(X[c]), y[c]), 0, clamp)
        loss.backward()
        opt.step()
        after = torch.cat([p.detach().reshape(-1) for p in m.parameters() if p.requires_grad])
        rows.append(after - before)
    return torch.stack(rows)


def batched_rows(model, X, y, lr, channels_last=False):
    n, B = X.shape[:2]
    flat = FlatParams(model, dtype=next(model.parameters()).dtype, channels_last=channels_last)
    U = torch.zeros(n, flat.numel, dtype=flat.dtype)
    sink = cb.GradSink(U, flat.specs, n, alpha=-lr)
    model.train()
    with cb.client_batched(model, sink, n * B):
        logits = model(X.reshape((n * B,) + tuple(X.shape[2:])))
        loss, _ = cb.batched_loss(logits, y.reshape(-1), n, torch.full((n,), 1e6, dtype=logits.dtype))
        loss.backward()
    assert sink.written == {s.name for s in flat.specs}
    return flat.to_reference_order(U)


@pytest.mark.parametrize("name", ["mlp", "resnet18", "resnet18_gn", "cct"])
def test_batched_equals_per_client(name):
    torch.manual_seed(0)
    if name == "mlp":
        model, shape, ncls = MLP(), (28, 28), 10
    elif name == "resnet18":
        model, shape, ncls = resnet18(num_classes=10), (3, 32, 32), 10
    elif name == "resnet18_gn":
        model, shape, ncls = resnet18(num_classes=10, norm="group"), (3, 32, 32), 10
    else:
        model = cct_2_3x2_32(attention_dropout=0.0, stochastic_depth=0.0)
        shape, ncls = (3, 32, 32), 10
    n, B = 3, 4
    model = model.double()          # exact comparison: fp32 BN over 4 samples is ill-conditioned
    X = torch.randn(n, B, *shape, dtype=torch.float64)
    y = torch.randint(0, ncls, (n, B))
    ref = per_client_reference(model, X, y, 0.1)
    got = batched_rows(copy.deepcopy(model), X, y, 0.1)
    assert torch.allclose(got, ref, atol=1e-9, rtol=1e-7), float((got - ref).abs().max())
    if name in ("resnet18", "cct"):      # channels_last physical parameter layout (the GPU default)
        got_cl = batched_rows(copy.deepcopy(model), X, y, 0.1, channels_last=True)
        assert torch.allclose(got_cl, ref, atol=1e-9, rtol=1e-7), float((got_cl - ref).abs().max())
    # forward swaps are undone
    assert "forward" not in model.__dict__


def test_unbatchable_model_detected():
    class Odd(nn.Module):
        def __init__(self):
            super().__init__()
            self.w = nn.Parameter(torch.randn(5, 5))

        def forward(self, x):
            return x @ self.w
    assert not cb.is_batchable(Odd())
    assert cb.is_batchable(MLP()) and cb.is_batchable(resnet18()) and cb.is_batchable(CCTNet())


def _run(agg, attack=None, nbyz=0, agg_kws=None, attack_kws=None, rounds=2, local_steps=1, tmp="", n=6,
         model=None, **kw):
    ds = synthetic_fldataset(n, shape=(28, 28), num_classes=10, train_bs=8, train_per_client=32,
                             test_per_client=16, seed=3)
    sim = Simulator(ds, num_byzantine=nbyz, attack=attack, attack_kws=attack_kws, aggregator=agg,
                    aggregator_kws=agg_kws, log_path=tmp, seed=1, progress=False, **kw)
    torch.manual_seed(5)
    m = model or MLP()
    times = sim.run(m, global_rounds=rounds, local_steps=local_steps, server_lr=1.0, client_lr=0.1,
                    validate_interval=rounds)
    return sim, m, times


def test_fedsgd_batched_matches_timesliced_end_to_end(tmp_log):
    # same seed, same data streams: batched engine vs forced time-slicing (custom optimizer factory)
    sim1, m1, t1 = _run("mean", tmp=tmp_log + "a")
    ds = synthetic_fldataset(6, shape=(28, 28), num_classes=10, train_bs=8, train_per_client=32,
                             test_per_client=16, seed=3)
    sim2 = Simulator(ds, aggregator="mean", log_path=tmp_log + "b", seed=1, progress=False)
    torch.manual_seed(5)
    m2 = MLP()
    sim2.run(m2, client_optimizer=lambda params, lr: torch.optim.SGD(params, lr=lr), global_rounds=2,
             local_steps=1, server_lr=1.0, client_lr=0.1, validate_interval=2)
    for p1, p2 in zip(m1.parameters(), m2.parameters()):
        assert torch.allclose(p1, p2, atol=1e-5)
    assert len(t1) == 2


@pytest.mark.parametrize("agg,kws", [("mean", None), ("median", None), ("trimmedmean", {"nb": 1}),
                                     ("krum", {"num_clients": 6, "num_byzantine": 1}), ("geomed", None),
                                     ("autogm", None), ("centeredclipping", None), ("clustering", None),
                                     ("clippedclustering", None), ("multikrum", {"num_byzantine": 1})])
@pytest.mark.parametrize("attack", [None, "noise", "labelflipping", "signflipping", "alie", "ipm"])
def test_all_attack_aggregator_pairs_run(agg, kws, attack, tmp_log):
    akw = {"num_clients": 6, "num_byzantine": 1} if attack == "alie" else None
    sim, m, _ = _run(agg, attack, 1 if attack else 0, kws, akw, rounds=1, tmp=tmp_log)
    assert all(torch.isfinite(p).all() for p in m.parameters())


def test_fused_attack_equals_callback_path(tmp_log):
    out = []
    for fuse in (True, False):
        sim, m, _ = _run("trimmedmean", "alie", 2, {"nb": 2}, {"num_clients": 6, "num_byzantine": 2},
                         rounds=2, tmp=tmp_log + str(fuse), fuse_attack=fuse)
        out.append(torch.cat([p.detach().reshape(-1) for p in m.parameters()]))
    assert torch.allclose(out[0], out[1], atol=1e-5)


def test_fedavg_update_is_param_difference(tmp_log):
    sim, m, _ = _run("mean", rounds=1, local_steps=3, tmp=tmp_log)
    U = sim.engine.U
    assert U.abs().sum() > 0
    # server: theta_new = theta_old + server_lr * mean(U)   (server_lr = 1)
    # reconstruct theta_old from any client row: after - before relation is internal; check shape/finite
    assert U.shape == (6, 59850) and torch.isfinite(U).all()


def test_server_step_semantics(tmp_log):
    from blades_b200.server import BladesServer
    m = MLP()
    flat = FlatParams(m)
    before = flat.theta.clone()
    opt = torch.optim.SGD(m.parameters(), lr=0.5)
    srv = BladesServer(opt, m, aggregator=None, flat=flat)
    upd = torch.randn(flat.numel)
    srv.apply_update(upd)
    assert torch.allclose(flat.theta, before + 0.5 * upd)
    # generic path (momentum) equals torch semantics with grad = -update
    m2 = MLP()
    f2 = FlatParams(m2)
    b2 = f2.theta.clone()
    opt2 = torch.optim.SGD(m2.parameters(), lr=0.5, momentum=0.9)
    srv2 = BladesServer(opt2, m2, aggregator=None, flat=f2)
    srv2.apply_update(upd)
    srv2.apply_update(upd)
    assert torch.allclose(f2.theta, b2 + 0.5 * upd + 0.5 * (1.9 * upd), atol=1e-5)


def test_api_contract(tmp_log):
    ds = synthetic_fldataset(5, shape=(28, 28), train_bs=8)
    with pytest.raises(RuntimeError, match="Unknown keyword"):
        Simulator(ds, log_path=tmp_log, bogus=1)
    sim = Simulator(ds, num_byzantine=2, attack="ipm", aggregator="median", log_path=tmp_log, seed=1,
                    progress=False)
    cl = sim.get_clients()
    assert [c.is_byzantine() for c in cl] == [True, True, False, False, False]
    assert [c.id() for c in cl] == [0, 1, 2, 3, 4]
    assert len(sim.omniscient_callbacks) == 2
    sim.set_trusted_clients([3])
    assert cl[3].is_trusted() and not cl[2].is_trusted()
    import os
    assert os.path.isfile(os.path.join(tmp_log, "stats")) and os.path.isfile(os.path.join(tmp_log, "debug"))


def test_custom_attacker_and_aggregator(tmp_log):
    calls = {"cb": 0, "train": 0}

    class Mal(ByzantineClient):
        def local_training(self, data_batches):
            calls["train"] += 1
            super().local_training(data_batches)

        def omniscient_callback(self, simulator):
            calls["cb"] += 1
            honest = [c.get_update() for c in simulator.get_clients() if not c.is_byzantine()]
            self.save_update(-10 * torch.stack(honest).mean(0))

    def my_agg(clients):
        return torch.stack([c.get_update() for c in clients]).mean(0)

    ds = synthetic_fldataset(5, shape=(28, 28), train_bs=8)
    sim = Simulator(ds, aggregator=my_agg, log_path=tmp_log, seed=1, progress=False)
    sim.register_attackers([Mal(), Mal()])
    m = MLP()
    sim.run(m, global_rounds=2, local_steps=2, validate_interval=1)
    assert calls == {"cb": 4, "train": 4}
    cl = sim.get_clients()
    assert cl[0].is_byzantine() and cl[1].is_byzantine() and not cl[2].is_byzantine()
    honest_mean = torch.stack([c.get_update() for c in cl[2:]]).mean(0)
    assert torch.allclose(cl[0].get_update(), -10 * honest_mean, atol=1e-6)


def test_loss_decreases_and_checkpoint_resume(tmp_log, tmp_path):
    ds_args = dict(shape=(28, 28), num_classes=10, train_bs=16, train_per_client=64, test_per_client=32, seed=3, separation=2.0)
    ck = str(tmp_path / "ck.pt")

    def make():
        ds = synthetic_fldataset(4, **ds_args)
        return Simulator(ds, aggregator="centeredclipping", log_path=tmp_log, seed=1, progress=False)

    sim = make()
    torch.manual_seed(0)
    m = MLP()
    sim.run(m, global_rounds=6, local_steps=2, server_lr=1.0, client_lr=0.1, validate_interval=1,
            checkpoint_path=ck, checkpoint_interval=3)
    recs = [eval(l) for l in open(tmp_log + "/stats") if "'test'" in l]
    assert recs[-1]["Loss"] < recs[0]["Loss"]
    final = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).clone()
    # checkpoint written at round 6 overwrote round 3; take a fresh run to round 3 then resume
    sim_a = make()
    torch.manual_seed(0)
    ma = MLP()
    sim_a.run(ma, global_rounds=3, local_steps=2, server_lr=1.0, client_lr=0.1, validate_interval=1,
              checkpoint_path=ck, checkpoint_interval=3)
    sim_b = make()
    mb = MLP()
    sim_b.run(mb, global_rounds=6, local_steps=2, server_lr=1.0, client_lr=0.1, validate_interval=1, resume=ck)
    resumed = torch.cat([p.detach().reshape(-1) for p in mb.parameters()])
    assert torch.allclose(resumed, final, atol=1e-6)
    sd = torch.load(ck, weights_only=False)
    MLP().load_state_dict(sd["model"])


def test_channels_last_flat_layout_and_server():
    from blades_b200.models import resnet18
    from blades_b200.server import BladesServer
    torch.manual_seed(0)
    m1, m2 = resnet18(10), None
    import copy
    m2 = copy.deepcopy(m1)
    f1 = FlatParams(m1, channels_last=False)
    f2 = FlatParams(m2, channels_last=True)
    assert torch.equal(f2.to_reference_order(f2.theta), f1.theta)
    upd = torch.randn(f1.numel)
    for f, m, u in ((f1, m1, upd), (f2, m2, f2.from_reference_order(upd))):
        opt = torch.optim.SGD(m.parameters(), lr=0.5, momentum=0.9)
        srv = BladesServer(opt, m, None, flat=f)
        srv.apply_update(u)
        srv.apply_update(u)
    for p1, p2 in zip(m1.parameters(), m2.parameters()):
        assert torch.allclose(p1, p2, atol=1e-6)
    x = torch.randn(8, 3, 32, 32)
    assert torch.allclose(m1(x), m2(x), atol=1e-3)


def test_batched_evaluation_equals_per_client_loop(tmp_path):
    """The evaluation fast path produces the per-client records of the reference's client-by-client loop."""
    import torch
    from blades_b200 import Simulator
    from blades_b200.datasets import synthetic_fldataset
    from blades_b200.models.mnist import MLP
    ds = synthetic_fldataset(5, shape=(28, 28), num_classes=10, train_bs=8, train_per_client=16, test_per_client=21, seed=2)
    sim = Simulator(ds, aggregator="mean", use_cuda=False, seed=1, log_path=str(tmp_path / "log"), progress=False)
    model = MLP()
    sim.prepare(model, "SGD", "SGD", "crossentropy", 1.0, 0.1)
    fast = sim._test_batched(3, 8)
    assert fast is not None and len(fast) == 5
    slow = [c.evaluate(round_number=3, test_set=sim.dataset.get_all_test_data(c.id()), batch_size=8,
                       metrics=sim.metrics, model=model) for c in sim.get_clients()]
    for a, b in zip(fast, slow):
        assert a["Length"] == b["Length"] == 21 and a["E"] == b["E"] == 3
        assert abs(a["Loss"] - b["Loss"]) < 1e-5 and abs(a["top1"] - b["top1"]) < 1e-4
    loss, top1 = sim.test_actor(3, 8)
    assert abs(loss - sum(r["Loss"] for r in slow) / 5) < 1e-5


def test_batched_evaluation_declines_when_grouping_matters(tmp_path):
    import torch.nn as nn
    from blades_b200 import Simulator
    from blades_b200.datasets import synthetic_fldataset
    from blades_b200.models.mnist import MLP
    ds = synthetic_fldataset(3, shape=(28, 28), num_classes=10, train_bs=8, train_per_client=16, test_per_client=8, seed=2)
    # custom metric -> per-client loop
    sim = Simulator(ds, aggregator="mean", use_cuda=False, seed=1, log_path=str(tmp_path / "a"), progress=False,
                    metrics={"top1": lambda out, tgt: 0.0})
    sim.prepare(MLP(), "SGD", "SGD", "crossentropy", 1.0, 0.1)
    assert sim._test_batched(1, 8) is None
    # batch-statistics BatchNorm in eval mode -> per-client loop
    sim = Simulator(ds, aggregator="mean", use_cuda=False, seed=1, log_path=str(tmp_path / "b"), progress=False)
    net = nn.Sequential(nn.Flatten(), nn.Linear(784, 16), nn.BatchNorm1d(16, track_running_stats=False), nn.Linear(16, 10))
    sim.prepare(net, "SGD", "SGD", "crossentropy", 1.0, 0.1)
    assert sim._test_batched(1, 8) is None
    loss, top1 = sim.test_actor(1, 8)
    assert loss > 0


@pytest.mark.parametrize("chunk", [0, 2, 3])
def test_prefetcher_with_chunked_requests_matches_synchronous_staging(tmp_path, monkeypatch, chunk):
    """The one-round-ahead prefetcher (worker thread, per-request double buffers) feeds exactly the batches the
    synchronous path would, also when ``max_batched_clients`` splits a round into several requests."""
    def run(prefetch_cpu):
        monkeypatch.setenv("BLADES_PREFETCH_CPU", "1" if prefetch_cpu else "0")
        monkeypatch.setenv("BLADES_MAX_BATCHED_CLIENTS", str(chunk))
        ds = synthetic_fldataset(6, shape=(28, 28), num_classes=10, train_bs=8, train_per_client=24, test_per_client=8, seed=4)
        sim = Simulator(ds, num_byzantine=2, attack="ipm", aggregator="median", log_path=str(tmp_path / f"l{prefetch_cpu}"),
                        seed=1, progress=False)
        torch.manual_seed(0)
        m = MLP()
        sim.run(m, global_rounds=7, local_steps=1, server_lr=1.0, client_lr=0.1, validate_interval=100)
        assert sim.engine.prefetch_on_cpu == prefetch_cpu
        if prefetch_cpu:
            assert len(sim.engine._pf_slot_of) == (1 if chunk == 0 else -(-6 // chunk))
        return torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    assert torch.equal(run(True), run(False))


def test_checkpoint_resume_is_exact_with_prefetch_and_chunking(tmp_path, monkeypatch):
    """The prefetcher has already drawn round r+1's batches when round r is checkpointed: the stored cursors must be
    the ones before those draws, for every chunk's clients."""
    monkeypatch.setenv("BLADES_PREFETCH_CPU", "1")
    monkeypatch.setenv("BLADES_MAX_BATCHED_CLIENTS", "2")
    ck = str(tmp_path / "ck.pt")
    ds_args = dict(shape=(28, 28), num_classes=10, train_bs=8, train_per_client=24, test_per_client=8, seed=5)

    def make(tag):
        return Simulator(synthetic_fldataset(6, **ds_args), aggregator="trimmedmean", aggregator_kws={"nb": 1},
                         log_path=str(tmp_path / tag), seed=1, progress=False)
    kw = dict(local_steps=1, server_lr=1.0, client_lr=0.1, validate_interval=100)
    torch.manual_seed(0)
    m = MLP()
    make("a").run(m, global_rounds=6, **kw)
    final = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).clone()
    torch.manual_seed(0)
    ma = MLP()
    make("b").run(ma, global_rounds=3, checkpoint_path=ck, checkpoint_interval=3, **kw)
    mb = MLP()
    make("c").run(mb, global_rounds=6, resume=ck, **kw)
    assert torch.allclose(torch.cat([p.detach().reshape(-1) for p in mb.parameters()]), final, atol=1e-7)


def test_graph_capture_policy_bounds_the_cache():
    """Learning rates are baked into captured graphs: capture on the second request of a key, keep graphs of at most
    two distinct learning rates (pure host logic of RoundEngine._worth_capturing)."""
    from blades_b200.engine.round import RoundEngine
    eng = RoundEngine.__new__(RoundEngine)
    eng._graph_seen, eng._graph_lr = {}, {}
    cache = {}

    def request(lr, rows=(0, 1)):
        key = (rows, float(lr), (2, 1, 8, 28, 28))
        if key in cache:
            return "replay"
        if eng._worth_capturing(cache, key, lr):
            cache[key] = object()
            return "capture"
        return "eager"
    assert [request(0.1) for _ in range(4)] == ["eager", "capture", "replay", "replay"]
    assert [request(0.01) for _ in range(3)] == ["eager", "capture", "replay"]                 # MultiStepLR drop
    assert request(0.1) == "replay" and len(cache) == 2
    assert [request(0.001) for _ in range(2)] == ["eager", "capture"]                          # third lr: oldest evicted
    assert {k[1] for k in cache} == {0.01, 0.001}
    # a schedule that changes the lr every round never captures anything
    before = len(cache)
    assert all(request(0.1 * 0.99 ** i) == "eager" for i in range(1, 40))
    assert len(cache) == before
    # chunked rounds: several keys share one lr and are all kept
    assert [request(0.001, rows=(2, 3)) for _ in range(2)] == ["eager", "capture"]
    assert len({k[1] for k in cache}) == 2


def test_consecutive_runs_continue_the_data_streams(tmp_path, monkeypatch):
    """run(3 rounds) then run(3 rounds) on one Simulator consumes the same batches as it would without the
    one-round-ahead prefetcher: the batch prefetched for the round after the last one is given back."""
    def go(prefetch):
        monkeypatch.setenv("BLADES_PREFETCH_CPU", prefetch)
        ds = synthetic_fldataset(4, shape=(28, 28), num_classes=10, train_bs=8, train_per_client=40, test_per_client=8, seed=6)
        sim = Simulator(ds, aggregator="mean", log_path=str(tmp_path / f"p{prefetch}"), seed=1, progress=False)
        outs = []
        for _ in range(2):
            torch.manual_seed(0)
            m = MLP()
            sim.run(m, global_rounds=3, local_steps=1, server_lr=1.0, client_lr=0.1, validate_interval=100)
            outs.append(torch.cat([p.detach().reshape(-1) for p in m.parameters()]))
        return outs, sim.dataset.state_dict()
    (a1, a2), sa = go("1")
    (b1, b2), sb = go("0")
    assert torch.equal(a1, b1) and torch.equal(a2, b2)
    assert all(sa[k]["pos"] == sb[k]["pos"] and sa[k]["epoch"] == sb[k]["epoch"] for k in sa)


# ------------------------------------------------------------------ advisor findings (round 1)
def _flat(m):
    return torch.cat([p.detach().reshape(-1) for p in m.parameters()])


@pytest.mark.parametrize("attack,akw", [("ipm", {"epsilon": 100.0}), ("alie", {"num_clients": 6, "num_byzantine": 2})])
def test_custom_callable_aggregator_still_sees_the_fusable_attacks(attack, akw, tmp_log):
    """ALIE / IPM are normally folded into the aggregation kernel as virtual rows; a plain callable aggregator never
    sees an UpdateMatrix, so the attacker callbacks must run for it (the defense was evaluated against NO attack)."""
    def my_mean(clients):
        return torch.stack([c.get_update() for c in clients]).mean(0)
    out = []
    for agg, fuse in ((my_mean, True), ("mean", False)):
        _, m, _ = _run(agg, attack, 2, None, akw, rounds=2, tmp=tmp_log + str(fuse), fuse_attack=fuse)
        out.append(_flat(m))
    assert torch.allclose(out[0], out[1], atol=1e-5), (out[0] - out[1]).abs().max()


def test_reference_style_aggregator_subclass_overriding_only_call(tmp_log):
    """Reference convention (aggregators/mean.py:21-28): subclasses override ``__call__`` and use
    ``self._get_updates(inputs)``; the shipped ByzantineSGD is one of them."""
    from blades_b200.aggregators.base import _BaseAggregator

    class RefMean(_BaseAggregator):
        def __call__(self, inputs):
            return self._get_updates(inputs).mean(0)

    _, m1, _ = _run(RefMean(), "ipm", 2, None, {"epsilon": 100.0}, tmp=tmp_log + "a")
    _, m2, _ = _run("mean", "ipm", 2, None, {"epsilon": 100.0}, tmp=tmp_log + "b", fuse_attack=False)
    assert torch.allclose(_flat(m1), _flat(m2), atol=1e-5)

    from blades_b200.aggregators.byzantinesgd import ByzantineSGD
    torch.manual_seed(5)
    model = MLP()
    opt = torch.optim.SGD(model.parameters(), lr=1.0)
    agg = ByzantineSGD(6, 1e6, 1e6, 1e6, opt)
    _, m3, _ = _run(agg, None, 0, tmp=tmp_log + "c", model=model)
    assert all(torch.isfinite(p).all() for p in m3.parameters())


def test_custom_loss_is_honoured_on_the_fedsgd_path(tmp_log):
    """``run(loss=callable)``: the fused fedsgd pass hard-wires cross-entropy, so such clients train time-sliced."""
    ds = synthetic_fldataset(4, shape=(28, 28), train_bs=8, seed=3)
    res = []
    for steps in (1, 2):
        sim = Simulator(ds, aggregator="mean", log_path=tmp_log + str(steps), seed=1, progress=False)
        torch.manual_seed(5)
        m = MLP()
        sim.run(m, loss=lambda out, target: out.sum() * 0.0, global_rounds=1, local_steps=steps,
                validate_interval=10, server_lr=1.0, client_lr=0.1)
        res.append(sim.last_aggregate.abs().max().item())
    assert res == [0.0, 0.0], res
