"""Small framework-wide helpers: logging setup, metrics, seeding, weight reset.

Public names mirror the reference module ``blades.utils``
(/root/reference/src/blades/utils.py:12-124) so user scripts port unchanged.
Differences (documented in DESIGN.md):
  * ``set_random_seed(None)`` is a no-op instead of a TypeError.
  * ``initialize_logger(root, wipe=True)``: wiping the directory is still the
    default (reference utils.py:70-74) but can be disabled for resume.
"""
from __future__ import annotations

import logging
import os
import random
import shutil
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch
from torch import nn

__all__ = [
    "BColors", "touch", "touch_dir", "accuracy", "top1_accuracy", "log", "log_dict",
    "initialize_logger", "reset_model_weights", "set_random_seed",
]


class BColors:
    """ANSI escape codes used by console printing (reference utils.py:12-21)."""
    HEADER = "\033[95m"
    OK_BLUE = "\033[94m"
    OK_CYAN = "\033[96m"
    OK_GREEN = "\033[92m"
    WARNING = "\033[93m"
    FAIL = "\033[91m"
    END_C = "\033[0m"
    BOLD = "\033[1m"
    UNDERLINE = "\033[4m"


def touch_dir(base_dir: str) -> None:
    os.makedirs(base_dir, exist_ok=True)


def touch(fname: str, times=None, create_dirs: bool = False) -> None:
    if create_dirs:
        parent = os.path.dirname(fname)
        if parent:
            touch_dir(parent)
    with open(fname, "a"):
        os.utime(fname, times)


@torch.no_grad()
def accuracy(output: torch.Tensor, target: torch.Tensor, topk: Sequence[int] = (1,)) -> List[torch.Tensor]:
    """precision@k in percent for each k in ``topk`` (reference utils.py:39-52)."""
    kmax = max(topk)
    n = target.shape[0]
    ranked = output.topk(kmax, dim=1, largest=True, sorted=True).indices  # [n, kmax]
    hits = ranked.eq(target.reshape(-1, 1))
    return [hits[:, :k].any(dim=1).float().sum().mul_(100.0 / n) for k in topk]


def top1_accuracy(output: torch.Tensor, target: torch.Tensor) -> float:
    return accuracy(output, target, topk=(1,))[0].item()


def log(*args, **kwargs) -> None:  # kept for API parity (reference utils.py:59-64: no-ops)
    return None


def log_dict(*args, **kwargs) -> None:
    return None


def _reset_logger(name: str) -> logging.Logger:
    lg = logging.getLogger(name)
    for h in list(lg.handlers):
        lg.removeHandler(h)
        try:
            h.close()
        except Exception:
            pass
    lg.setLevel(logging.INFO)
    lg.propagate = False
    return lg


def initialize_logger(log_root: str, wipe: bool = True, verbose: bool = False) -> None:
    """Create ``<log_root>/stats`` (one dict repr per line) and ``<log_root>/debug``.

    Same two loggers / file names as the reference (utils.py:67-95). Unlike the
    reference we do not ``reload(logging)`` (which invalidates every handler in
    the process); we only reset the two loggers we own.
    """
    if wipe and os.path.isdir(log_root):
        shutil.rmtree(log_root)
    os.makedirs(log_root, exist_ok=True)
    if verbose:
        print(f"Logging files to {log_root}")
    for name in ("stats", "debug"):
        lg = _reset_logger(name)
        fh = logging.FileHandler(os.path.join(log_root, name))
        fh.setLevel(logging.INFO)
        fh.setFormatter(logging.Formatter("%(message)s"))
        lg.addHandler(fh)


def reset_model_weights(model: nn.Module) -> None:
    """Call ``reset_parameters()`` on every submodule that has one (reference utils.py:98-114)."""
    with torch.no_grad():
        for m in model.modules():
            fn = getattr(m, "reset_parameters", None)
            if callable(fn):
                fn()


def set_random_seed(seed_value: Optional[int] = 0, use_cuda: bool = False) -> None:
    """Seed python/numpy/torch. ``None`` means "leave RNGs alone" (fixes SURVEY Q12)."""
    if seed_value is None:
        return
    np.random.seed(seed_value)
    random.seed(seed_value)
    torch.manual_seed(seed_value)
    os.environ["PYTHONHASHSEED"] = str(seed_value)
    if use_cuda and torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed_value)
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False
