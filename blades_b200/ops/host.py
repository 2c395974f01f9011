"""ctypes bindings of the native host library (csrc/host/selectors.cpp).  Every function has a numpy
twin in aggregators/_gramops.py; ``available()`` tells callers whether the native path can be used."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from . import _loader

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_ready = False


def _lib():
    global _ready
    lib = _loader.host_lib()
    if lib is not None and not _ready:
        lib.bl_krum_scores.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp]
        lib.bl_krum_scores.restype = None
        lib.bl_weiszfeld.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_double, C.c_double, C.c_int, _dp]
        lib.bl_weiszfeld.restype = C.c_int
        lib.bl_autogm.argtypes = [_dp, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _dp]
        lib.bl_autogm.restype = None
        lib.bl_centered_clip.argtypes = [_dp, C.c_int, C.c_double, C.c_int, _dp]
        lib.bl_centered_clip.restype = None
        lib.bl_complete_linkage2.argtypes = [_dp, C.c_int, _ip]
        lib.bl_complete_linkage2.restype = None
        lib.bl_gather_batches.argtypes = [C.c_void_p, C.c_void_p, _ip, C.c_int, C.c_int, C.c_int64, C.c_void_p,
                                          C.c_void_p]
        lib.bl_gather_batches.restype = None
        _ready = True
    return lib


def available() -> bool:
    return _lib() is not None


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def krum_scores(D, f: int, squared_twice: bool) -> np.ndarray:
    D = _c(D)
    out = np.empty(D.shape[0])
    _lib().bl_krum_scores(D, D.shape[0], f, int(squared_twice), out)
    return out


def weiszfeld(G, alphas, maxiter, eps, ftol, compounding) -> Tuple[np.ndarray, int]:
    G = _c(G)
    w = np.empty(G.shape[0])
    it = _lib().bl_weiszfeld(G, G.shape[0], _c(alphas), int(maxiter), float(eps), float(ftol), int(compounding), w)
    return w, it


def autogm(G, lamb, maxiter, eps, ftol, sort_by_index, compounding) -> np.ndarray:
    G = _c(G)
    w = np.empty(G.shape[0])
    _lib().bl_autogm(G, G.shape[0], float(lamb), int(maxiter), float(eps), float(ftol), int(sort_by_index),
                     int(compounding), w)
    return w


def centered_clip(G_aug, tau, n_iter) -> np.ndarray:
    G = _c(G_aug)
    c = np.empty(G.shape[0])
    _lib().bl_centered_clip(G, G.shape[0] - 1, float(tau), int(n_iter), c)
    return c


def complete_linkage2(dist) -> np.ndarray:
    D = _c(dist)
    labels = np.empty(D.shape[0], dtype=np.int64)
    _lib().bl_complete_linkage2(D, D.shape[0], labels)
    return labels
