"""Checkpoint / resume (the reference has none -- SURVEY 5.4; format defined here).

One ``torch.save`` dict:

    model            global ``state_dict`` -- keys identical to the torch module (e.g. ``mdoel.*``
                     for CCTNet), loadable with plain ``model.load_state_dict``
    server_opt       server optimizer ``state_dict``
    round            last completed round
    client_lr        current client learning rate
    schedulers       {server, client} scheduler state_dicts (if given)
    rng              torch / cuda / numpy / python RNG states, one entry PER RANK (format 2; format 1 = rank 0 only)
    aggregator_state ``aggregator.state_dict()`` (Centeredclipping momentum, Clippedclustering
                     norm history, ByzantineSGD A/B/good)
    data_cursors     per-client batch-stream cursors, merged from the rank that owns each client
    config           {n_clients, d, world_size, format_version}

Every rank holds identical server state, so rank 0 writes; every rank reads.
"""
from __future__ import annotations

import os
import random
from typing import Optional

import numpy as np
import torch

__all__ = ["save_checkpoint", "load_checkpoint", "FORMAT_VERSION"]

FORMAT_VERSION = 2


def _rng_state(device: torch.device) -> dict:
    st = {"torch": torch.get_rng_state(), "numpy": np.random.get_state(), "python": random.getstate()}
    if device.type == "cuda":
        st["cuda"] = torch.cuda.get_rng_state(device)
    return st


def _set_rng_state(st: dict, device: torch.device) -> None:
    torch.set_rng_state(st["torch"])
    np.random.set_state(st["numpy"])
    random.setstate(st["python"])
    if device.type == "cuda" and "cuda" in st:
        torch.cuda.set_rng_state(st["cuda"], device)


def _merged_cursors(sim) -> dict:
    """Every rank only advances the streams of ITS clients: merge each rank's local cursors."""
    cur = sim.engine.data_cursors()
    if not sim.world.distributed:
        return cur
    eng = sim.engine
    mine = {cid: cur[cid] for cid in (eng.clients[gi].id() for gi in eng.local_idx) if cid in cur}
    merged = {}
    for part in sim.world.all_gather_object(mine):
        merged.update(part)
    return merged


def save_checkpoint(path: str, sim, server_sched=None, client_sched=None) -> Optional[str]:
    agg = sim.aggregator
    rng = _rng_state(sim.device)
    rng_all = sim.world.all_gather_object(rng) if sim.world.distributed else [rng]
    payload = {
        "model": {k: v.detach().cpu().clone() for k, v in sim.server.get_model().state_dict().items()},
        "server_opt": sim.server.get_opt().state_dict(),
        "round": sim.round,
        "client_lr": sim.client_lr,
        "schedulers": {"server": server_sched.state_dict() if server_sched else None,
                       "client": client_sched.state_dict() if client_sched else None},
        "rng": rng_all,
        "aggregator_state": agg.state_dict() if hasattr(agg, "state_dict") else {},
        "data_cursors": _merged_cursors(sim),
        "config": {"n_clients": len(sim.get_clients()), "d": sim.engine.d,
                   "world_size": sim.world.size, "format_version": FORMAT_VERSION},
    }
    if sim.world.rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        tmp = path + ".tmp"
        torch.save(payload, tmp)
        os.replace(tmp, path)
    sim.world.barrier()
    return path


def load_checkpoint(path: str, sim, server_sched=None, client_sched=None) -> int:
    """Restore into a *prepared* simulator (``sim.prepare(model, ...)`` already called).
    Returns the last completed round."""
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert ck["config"]["format_version"] in (1, FORMAT_VERSION)
    model = sim.server.get_model()
    with torch.no_grad():
        own = model.state_dict()
        for k, v in ck["model"].items():
            own[k].copy_(v.to(own[k].device))        # in place: parameters stay views of theta
    sim.server.get_opt().load_state_dict(ck["server_opt"])
    if server_sched and ck["schedulers"]["server"]:
        server_sched.load_state_dict(ck["schedulers"]["server"])
    if client_sched and ck["schedulers"]["client"]:
        client_sched.load_state_dict(ck["schedulers"]["client"])
    rng = ck["rng"]
    if isinstance(rng, list):                  # format 2: one RNG state per rank (rank 0's when the world changed)
        rng = rng[sim.world.rank] if len(rng) == sim.world.size else rng[0]
    _set_rng_state(rng, sim.device)
    if hasattr(sim.aggregator, "load_state_dict"):
        sim.aggregator.load_state_dict(ck["aggregator_state"])
    if hasattr(sim.dataset, "load_state_dict"):
        sim.dataset.load_state_dict(ck["data_cursors"])
    sim.round = ck["round"]
    sim.client_lr = ck["client_lr"]
    return ck["round"]
