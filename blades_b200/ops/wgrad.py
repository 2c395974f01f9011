"""Grouped (per-client) weight-gradient GEMM with a fused update-row epilogue:

    out[c] = alpha * a[c] @ b[c]        a: [n, M, K]   b: [n, K, N]   out: [n, M, N]

``out`` is a strided window into the shard's update matrix ``U_g[n, d]`` (batch stride
= d), so the GEMM epilogue IS the client's update write (SGD step + update
diff + save_update fused).  CUDA: tcgen05 kernel (csrc/cuda/wgrad_tcgen05.cu) when
shapes allow, else cuBLAS strided-batched through ``torch.baddbmm``.  CPU: baddbmm.
"""
from __future__ import annotations

import torch

__all__ = ["grouped_wgrad"]

import os

#: BLADES_WGRAD=cublas forces the library fallback (A/B comparisons, debugging)
_USE_KERNEL = os.environ.get("BLADES_WGRAD", "tcgen05") != "cublas"


def grouped_wgrad(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    if a.is_cuda and _USE_KERNEL:
        from . import _loader
        lib = _loader.cuda_lib(optional=True)
        if lib is not None and hasattr(lib, "bl_grouped_wgrad"):
            from ._wgrad_impl import launch
            if launch(lib, a, b, out, alpha):
                return out
    torch.baddbmm(out, a, b, beta=0.0, alpha=alpha, out=out)
    return out
