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

Format 3 stores only tensors and plain Python containers (numpy arrays / scalars inside the RNG states, data cursors and
aggregator states are converted on save and restored on load), so a checkpoint is read with
``torch.load(weights_only=True)`` -- loading one cannot execute pickled code.  Files of the older formats need
``BLADES_TRUST_CHECKPOINT=1`` (they are unpickled without that restriction).
"""
from __future__ import annotations

import os
import random
from typing import Optional

import numpy as np
import torch

__all__ = ["save_checkpoint", "load_checkpoint", "FORMAT_VERSION"]

FORMAT_VERSION = 3
_ND = "__ndarray__"


def _to_plain(obj):
    """numpy arrays / scalars -> tensors / Python scalars, recursively (tuples are kept as tagged lists)."""
    if isinstance(obj, np.ndarray):
        return {_ND: torch.from_numpy(np.ascontiguousarray(obj).copy()) if obj.dtype != np.uint32
                else torch.from_numpy(obj.astype(np.int64)), "dtype": str(obj.dtype)}
    if isinstance(obj, np.generic):
        return obj.item()
    if isinstance(obj, dict):
        return {k: _to_plain(v) for k, v in obj.items()}
    if isinstance(obj, tuple):
        return {"__tuple__": [_to_plain(v) for v in obj]}
    if isinstance(obj, list):
        return [_to_plain(v) for v in obj]
    return obj


def _from_plain(obj):
    if isinstance(obj, dict):
        if _ND in obj:
            return obj[_ND].numpy().astype(np.dtype(obj["dtype"]))
        if "__tuple__" in obj and len(obj) == 1:
            return tuple(_from_plain(v) for v in obj["__tuple__"])
        return {k: _from_plain(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_from_plain(v) for v in obj]
    return obj


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
        "rng": _to_plain(rng_all),
        "aggregator_state": _to_plain(agg.state_dict() if hasattr(agg, "state_dict") else {}),
        "data_cursors": _to_plain(_merged_cursors(sim)),
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
    try:
        ck = torch.load(path, map_location="cpu", weights_only=True)
    except Exception as e:                      # formats 1-2 pickle numpy objects
        if os.environ.get("BLADES_TRUST_CHECKPOINT", "0") != "1":
            raise RuntimeError(f"{path} is not a format-{FORMAT_VERSION} checkpoint (tensors and plain containers only); "
                               "set BLADES_TRUST_CHECKPOINT=1 to unpickle an older, trusted file") from e
        ck = torch.load(path, map_location="cpu", weights_only=False)
    assert ck["config"]["format_version"] in (1, 2, FORMAT_VERSION)
    if ck["config"]["format_version"] >= 3:
        for k in ("rng", "aggregator_state", "data_cursors"):
            ck[k] = _from_plain(ck[k])
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
