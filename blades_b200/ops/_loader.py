"""ctypes loader for the in-tree native libraries (``_cuda.so`` / ``_host.so``).

Policy: on a machine WITH a CUDA device a missing ``_cuda.so`` is a hard error -- the GPU
path must never silently degrade to eager PyTorch.  Without a GPU (CI container) the CUDA
library is simply not needed."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CUDA: Optional[ctypes.CDLL] = None
_HOST: Optional[ctypes.CDLL] = None
_TRIED_HOST = False
LAUNCHES = 0          # number of our kernels launched (bench.py reports it)


def count_launch(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def cuda_lib(optional: bool = False) -> Optional[ctypes.CDLL]:
    global _CUDA
    if _CUDA is not None:
        return _CUDA
    path = os.path.join(PKG, "_cuda.so")
    if not os.path.exists(path):
        if optional and not torch.cuda.is_available():
            return None
        raise RuntimeError(
            f"{path} is missing: build it with `python -m blades_b200.ops.build` "
            "(the CUDA path never falls back to eager PyTorch on a GPU machine)")
    _CUDA = ctypes.CDLL(path)
    return _CUDA


def host_lib() -> Optional[ctypes.CDLL]:
    """CPU-side native helpers; optional (numpy fallbacks exist)."""
    global _HOST, _TRIED_HOST
    if _HOST is None and not _TRIED_HOST:
        _TRIED_HOST = True
        path = os.path.join(PKG, "_host.so")
        if os.path.exists(path):
            try:
                _HOST = ctypes.CDLL(path)
            except OSError:
                _HOST = None
    return _HOST


def stream_ptr(device=None) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def check(code: int, what: str) -> None:
    if code != 0:
        raise RuntimeError(f"{what} failed with CUDA error/code {code}")
