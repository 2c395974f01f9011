"""Debug probe for the MN-major tcgen05 wgrad kernel: structured inputs, dumps outputs for offline analysis."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blades_b200.ops import wgrad, _wgrad_impl
_wgrad_impl.MIN_OUTPUT_ELEMS = 0
res = {}
for swap in ("0", "1"):
    os.environ["BLADES_WGRAD_SWAP"] = swap
    for (T, M, N) in ((8, 128, 32), (16, 128, 32), (32, 128, 32), (32, 128, 256), (64, 128, 64)):
        for t0 in (0, 3, T - 1):
            a_t = torch.zeros(1, T, M, device="cuda")
            b = torch.zeros(1, T, N, device="cuda")
            a_t[0, t0] = torch.arange(1, M + 1, device="cuda").float()
            b[0, t0] = torch.arange(1, N + 1, device="cuda").float() * 0.001
            out = torch.full((1, M, N), -7.0, device="cuda")
            wgrad.grouped_wgrad(a_t.transpose(1, 2), b, out, 1.0)
            torch.cuda.synchronize()
            ref = (a_t.transpose(1, 2) @ b)
            err = (out - ref).abs().max().item()
            res[(swap, T, M, N, t0)] = out.cpu()
            print(f"swap={swap} T={T} M={M} N={N} t0={t0}: max err {err:.4f}  out[0,:3,:4]={out[0,:3,:4].flatten().tolist()}  nnz={(out != 0).sum().item()} n_minus7={(out == -7).sum().item()}")
        # random
        a_t = torch.randn(1, T, M, device="cuda"); b = torch.randn(1, T, N, device="cuda")
        out = torch.zeros(1, M, N, device="cuda")
        wgrad.grouped_wgrad(a_t.transpose(1, 2), b, out, 1.0)
        ref = a_t.double().transpose(1, 2) @ b.double()
        print(f"swap={swap} T={T} M={M} N={N} random: rel err {((out.double()-ref).norm()/ref.norm()).item():.4f}")
torch.save(res, "gpurun_out/probe_wgrad.pt")
