"""The number to beat: the reference's round semantics on stock PyTorch
(cuDNN / cuBLAS / ATen) with Ray replaced by torch.distributed NCCL collectives.

Per round, per rank (= one trainer shard, like one Ray actor with a whole GPU):
  for every local client (time-sliced, reference actor.py:23-33):
      load global weights into the client model, k SGD steps with autograd (client.py:178-193),
      update = flat(theta_after) - flat(theta_before)                      (client.py:127-131)
  all_gather_into_tensor -> U[N, d] on every rank                           (simulator.py:235)
  ALIE / IPM / noise on the stacked updates with stock torch ops            (attackers/*.py)
  aggregator with the reference's torch formulation (topk / median / cdist) (aggregators/*.py)
  theta += lr * agg on every rank (replaces the model broadcast)            (server.py:54-75)

None of this repo's kernels or engine is used on this path (only the model definition, the
synthetic data generator and the world bootstrap)."""
from __future__ import annotations

import copy
import math
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F


def _trimmed_mean(U, b):
    largest, _ = torch.topk(U, b, 0)
    neg_smallest, _ = torch.topk(-U, b, 0)
    return torch.cat([U, -largest, neg_smallest]).sum(0) / (U.shape[0] - 2 * b)


def _median(U):
    return (U.median(0).values - (-U).median(0).values) / 2


def _krum_multi(U, f, m):
    D = torch.cdist(U[None], U[None])[0] ** 2
    n = U.shape[0]
    D.fill_diagonal_(float("inf"))
    scores = D.sort(1).values[:, : n - f - 2].sum(1)
    idx = scores.argsort()[:m]
    return U[idx].mean(0)


def _geomed(U, maxiter=100, eps=1e-6, ftol=1e-10):
    n = U.shape[0]
    w = torch.full((n,), 1.0 / n, device=U.device)
    med = U.mean(0)
    dist_ = lambda z: (U - z).norm(dim=1)
    obj = (w * dist_(med)).sum()
    for _ in range(maxiter):
        prev = obj
        w = torch.clamp(w / torch.clamp(dist_(med), min=eps), min=eps)
        w = w / w.sum()
        med = (w[:, None] * U).sum(0)
        obj = (w * dist_(med)).sum()
        if abs(prev - obj) < ftol * obj:
            break
    return med


def run_baseline(args, CONFIGS) -> dict:
    from blades_b200.comm.group import init_world, split_clients
    from blades_b200.datasets import synthetic_fldataset
    from blades_b200.attackers.alieclient import alie_z_max
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import build_model, ClockSampler

    world = init_world(use_cuda=True)
    dev = world.device
    model_name, classes, n_clients, n_byz, attack, agg, agg_kws, local_steps = CONFIGS[args.config]
    if args.clients:
        n_byz = max(1, n_byz * args.clients // n_clients) if n_byz else 0
        n_clients = args.clients
        if "nb" in agg_kws:
            agg_kws = {"nb": n_byz}
    shape = (28, 28) if model_name == "mlp" else (3, 32, 32)
    ds = synthetic_fldataset(n_clients, shape=shape, num_classes=classes, train_bs=args.batch,
                             train_per_client=2 * args.batch, test_per_client=args.batch, seed=1)
    torch.manual_seed(1)
    global_model = build_model(model_name, classes).to(dev)
    if world.distributed:
        for p in global_model.parameters():
            dist.broadcast(p.data, 0)
    client_model = copy.deepcopy(global_model)
    opt = torch.optim.SGD(client_model.parameters(), lr=0.1)
    params = [p for p in global_model.parameters() if p.requires_grad]
    d = sum(p.numel() for p in params)
    parts = split_clients(n_clients, world.size)
    mine = [int(i) for i in parts[world.rank]]
    nmax = max(len(p) for p in parts)
    local = torch.zeros(nmax, d, device=dev)
    gathered = torch.empty(world.size * nmax, d, device=dev) if world.distributed else None
    z = alie_z_max(n_clients, n_byz) if attack == "alie" else 0.0
    server_lr = 1.0

    def flat(m):
        return torch.cat([p.data.view(-1) for p in m.parameters() if p.requires_grad])

    def one_round():
        losses = []
        for r, cid in enumerate(mine):
            client_model.load_state_dict(global_model.state_dict())
            client_model.train()
            before = flat(client_model).clone()
            for data, target in ds.get_train_data(cid, local_steps):
                data, target = data.to(dev), target.to(dev)
                if attack == "labelflipping" and cid < n_byz:
                    target = classes - 1 - target
                opt.zero_grad()
                loss = torch.clamp(F.cross_entropy(client_model(data), target), 0, 1e6)
                loss.backward()
                opt.step()
            losses.append(loss.detach())
            local[r] = torch.nan_to_num(flat(client_model) - before)
        if world.distributed:
            dist.all_gather_into_tensor(gathered, local)
            U = torch.cat([gathered[g * nmax: g * nmax + len(parts[g])] for g in range(world.size)])
        else:
            U = local[: len(mine)]
        if attack == "alie":
            good = U[n_byz:]
            U[:n_byz] = good.mean(0) - z * good.std(0)
        elif attack == "ipm":
            U[:n_byz] = -0.5 * U[n_byz:].mean(0)
        elif attack == "noise":
            U[:n_byz] = torch.normal(0.1, 0.1, size=(n_byz, d), device=dev)
        if agg == "trimmedmean":
            a = _trimmed_mean(U, agg_kws["nb"])
        elif agg == "median":
            a = _median(U)
        elif agg == "multikrum":
            a = _krum_multi(U, n_byz, n_clients - n_byz)
        elif agg == "geomed":
            a = _geomed(U)
        else:
            a = U.mean(0)
        beg = 0
        for p in params:
            p.data.add_(a[beg: beg + p.numel()].view_as(p), alpha=server_lr)
            beg += p.numel()
        return torch.stack(losses).mean()

    for _ in range(args.warmup):
        one_round()
    torch.cuda.synchronize()
    world.barrier()
    sampler = ClockSampler(dev.index or 0)
    if world.rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = one_round()
        loss.cpu()
    e1.record()
    torch.cuda.synchronize()
    world.barrier()
    ms = world.all_reduce_max(e0.elapsed_time(e1))
    wall = world.all_reduce_max((time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop() if world.rank == 0 else {}
    if world.rank != 0:
        return {}
    return {"metric": "FL rounds/sec", "timing": "device (CUDA events around the K rounds, max over ranks)",
            "value": args.steps / (ms / 1e3),
            "unit": "rounds/s", "n_gpus": world.size, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "fp32 (torch defaults: TF32 convs)", "data": "synthetic", "impl": "baseline-nccl-torch",
            "config": {"model": model_name, "clients": n_clients, "byzantine": n_byz, "attack": attack,
                       "aggregator": agg, "local_steps": local_steps, "client_batch": args.batch},
            "clocks": clocks, "gpu_launches": 0,
            "e2e": {"value": args.steps / (wall / 1e3), "unit": "rounds/s", "ms_per_step": wall / args.steps,
                    "h2d_bytes_per_step": len(mine) * local_steps * args.batch * (math.prod(shape) * 4 + 8),
                    "d2h_bytes_per_step": 4}}
