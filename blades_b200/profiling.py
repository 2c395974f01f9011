"""Profiling helpers (the reference only has wall-clock ``time()`` per round).

* ``Simulator(profile=True)`` -> ``sim.engine.timer.records``: CUDA-event milliseconds per phase per round
  (``train`` / ``aggregate`` / ``apply``), also wrapped in NVTX ranges for Nsight captures.
* ``phase_summary(sim)``: mean / min / max per phase.
* ``kernel_launch_count()``: number of this repo's native kernels launched so far (including those replayed
  from CUDA graphs) -- what ``bench.py`` reports as ``gpu_launches``.
* ``roofline(bytes_moved, ms)``: achieved GB/s and fraction of the measured HBM copy bandwidth
  (``MEASURED_PEAKS.json``; fallback 6650 GB/s).
"""
from __future__ import annotations

import contextlib
import json
import os
from typing import Dict

import torch

__all__ = ["phase_summary", "kernel_launch_count", "roofline", "nvtx_range", "measured_peaks"]


def phase_summary(sim) -> Dict[str, Dict[str, float]]:
    recs = sim.engine.timer.records
    out: Dict[str, Dict[str, float]] = {}
    for name in sorted({k for r in recs for k in r}):
        vals = [r[name] for r in recs if name in r]
        out[name] = {"mean_ms": sum(vals) / len(vals), "min_ms": min(vals), "max_ms": max(vals), "rounds": len(vals)}
    return out


def kernel_launch_count() -> int:
    from .ops import _loader
    return _loader.LAUNCHES


def measured_peaks() -> dict:
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}


def roofline(bytes_moved: float, ms: float) -> Dict[str, float]:
    gbs = bytes_moved / ms / 1e6
    peak = measured_peaks().get("hbm_gbs", 6650.0)
    return {"gbs": gbs, "fraction_of_measured_hbm": gbs / peak}


@contextlib.contextmanager
def nvtx_range(name: str):
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()
