"""Multi-process (gloo, world_size 2/3) runs of the SPMD simulator must reproduce the single-process
result: client->shard split, row gathering, replicated server state, distributed evaluation."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(rank, world, port, agg, attack, out_dir, local_steps):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    from blades_b200 import Simulator
    from blades_b200.comm.group import init_world, shutdown
    from blades_b200.datasets import synthetic_fldataset
    from blades_b200.models import MLP
    init_world(use_cuda=False)
    ds = synthetic_fldataset(7, shape=(28, 28), train_bs=8, train_per_client=32, test_per_client=16, seed=3,
                             separation=2.0)
    kws = {"num_clients": 7, "num_byzantine": 2} if attack == "alie" else None
    sim = Simulator(ds, num_byzantine=2 if attack else 0, attack=attack, attack_kws=kws, aggregator=agg,
                    aggregator_kws={"nb": 2} if agg == "trimmedmean" else None,
                    log_path=os.path.join(out_dir, "logs"), seed=1, progress=False)
    torch.manual_seed(5)
    m = MLP()
    sim.run(m, global_rounds=3, local_steps=local_steps, server_lr=1.0, client_lr=0.1, validate_interval=3)
    vec = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    torch.save(vec, os.path.join(out_dir, f"theta_{world}_{rank}.pt"))
    shutdown()


@pytest.mark.parametrize("agg,attack,local_steps", [("trimmedmean", "alie", 1), ("geomed", "noise", 2),
                                                    ("median", "labelflipping", 1)])
def test_world_sizes_agree(agg, attack, local_steps, tmp_path):
    out = str(tmp_path)
    results = {}
    for world in (1, 2, 3):
        port = _free_port()
        if world == 1:
            ctx = mp.get_context("spawn")
            p = ctx.Process(target=_run, args=(0, 1, port, agg, attack, out, local_steps))
            p.start()
            p.join(300)
            assert p.exitcode == 0
        else:
            mp.spawn(_run, args=(world, port, agg, attack, out, local_steps), nprocs=world, join=True)
        results[world] = [torch.load(os.path.join(out, f"theta_{world}_{r}.pt")) for r in range(world)]
    base = results[1][0]
    for world in (2, 3):
        for r, vec in enumerate(results[world]):
            if attack == "noise":     # per-rank RNG streams differ for the noise rows; replicas must still agree
                assert torch.allclose(vec, results[world][0], atol=1e-6)
            else:
                assert torch.allclose(vec, base, atol=1e-5), (world, r, (vec - base).abs().max())


def test_coordinate_shard_emulation_single_process():
    """Single-process emulation of the G coordinate shards (SURVEY 7.6 tests/dist): every rank reduces its own
    coordinate range of ALL rows with the CPU oracle; concatenating the ranges reproduces the global aggregate."""
    import torch
    from blades_b200.comm.symm import coordinate_shards
    from blades_b200.parallel.matrix import LocalMatrix
    g = torch.Generator().manual_seed(0)
    for d in (1, 127, 128, 129, 1000, 59850):
        U = torch.randn(12, d, generator=g)
        want_tm = LocalMatrix(U).trimmed_mean(3)
        want_med = LocalMatrix(U).median()
        w = torch.rand(12, generator=g)
        want_comb = LocalMatrix(U).combine(w)
        for G in (1, 2, 3, 4, 8):
            shards = coordinate_shards(d, G)
            assert shards[0][0] == 0 and shards[-1][1] == d
            assert all(a1 == b0 for (_, a1), (b0, _) in zip(shards, shards[1:]))           # contiguous, disjoint
            assert all(c0 % 128 == 0 for c0, _ in shards[1:] if c0 < d)                  # aligned interior cuts
            parts_tm = [LocalMatrix(U[:, c0:c1]).trimmed_mean(3) for c0, c1 in shards if c1 > c0]
            parts_med = [LocalMatrix(U[:, c0:c1]).median() for c0, c1 in shards if c1 > c0]
            parts_comb = [LocalMatrix(U[:, c0:c1]).combine(w) for c0, c1 in shards if c1 > c0]
            assert torch.equal(torch.cat(parts_tm), want_tm)
            assert torch.equal(torch.cat(parts_med), want_med)
            assert torch.allclose(torch.cat(parts_comb), want_comb)


def _run_ckpt(rank, world, port, out_dir, phase):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    from blades_b200 import Simulator
    from blades_b200.comm.group import init_world, shutdown
    from blades_b200.datasets import synthetic_fldataset
    from blades_b200.models import MLP
    init_world(use_cuda=False)
    ds = synthetic_fldataset(6, shape=(28, 28), train_bs=8, train_per_client=20, test_per_client=8, seed=3)   # 8, 8, 4
    sim = Simulator(ds, num_byzantine=2, attack="noise", aggregator="centeredclipping",
                    log_path=os.path.join(out_dir, "logs" + phase), seed=1, progress=False)
    torch.manual_seed(5)
    m = MLP()
    ck = os.path.join(out_dir, "ck.pt")
    kw = dict(local_steps=1, server_lr=1.0, client_lr=0.1, validate_interval=100)
    if phase == "full":
        sim.run(m, global_rounds=7, **kw)
    elif phase == "first":
        sim.run(m, global_rounds=4, checkpoint_path=ck, checkpoint_interval=4, **kw)
    else:
        sim.run(m, global_rounds=7, resume=ck, **kw)
    torch.save(torch.cat([p.detach().reshape(-1) for p in m.parameters()]), os.path.join(out_dir, f"{phase}_{rank}.pt"))
    shutdown()


def test_distributed_checkpoint_resume_is_exact(tmp_path):
    """Two ranks: each rank owns different clients (their stream cursors) and its own RNG stream (the noise attacker
    lives on rank 0 only); a checkpoint must carry all of them for the resumed run to match the uninterrupted one."""
    out = str(tmp_path)
    for phase in ("full", "first", "resume"):
        mp.spawn(_run_ckpt, args=(2, _free_port(), out, phase), nprocs=2, join=True)
    full = torch.load(os.path.join(out, "full_0.pt"))
    for r in (0, 1):
        assert torch.allclose(torch.load(os.path.join(out, f"resume_{r}.pt")), full, atol=1e-7), r
