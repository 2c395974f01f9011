#!/usr/bin/env python
"""GPU check of the tcgen05 implicit-GEMM convolution kernel (ops/conv.py): every case against an fp64 reference and
against the torch emulation of the same plan; full-size ResNet-18 layers are also timed against cuDNN (TF32).

    python scripts/gpu_conv_check.py            # parent: runs the cases in a child, restarts after a fatal CUDA error
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, NB, H, W, Cin, Cout, k, stride, pad, time_it)
CASES = [
    ("l1 3x3 64>64 @8", 64, 8, 8, 64, 64, 3, 1, 1, False),
    ("l2a 3x3s2 64>128 @8", 64, 8, 8, 64, 128, 3, 2, 1, False),
    ("l2 3x3 128>128 @4", 64, 4, 4, 128, 128, 3, 1, 1, False),
    ("l2ds 1x1s2 64>128 @8", 64, 8, 8, 64, 128, 1, 2, 0, False),
    ("l3a 3x3s2 128>256 @4", 96, 4, 4, 128, 256, 3, 2, 1, False),
    ("l3 3x3 256>256 @2", 96, 2, 2, 256, 256, 3, 1, 1, False),
    ("l4a 3x3s2 256>512 @2", 160, 2, 2, 256, 512, 3, 2, 1, False),
    ("l4 3x3 512>512 @1", 160, 1, 1, 512, 512, 3, 1, 1, False),
    ("r50 1x1 64>256 @8", 32, 8, 8, 64, 256, 1, 1, 0, False),
    ("r50 1x1 256>64 @8", 32, 8, 8, 256, 64, 1, 1, 0, False),
    ("odd 3x3 32>48 @7x9", 5, 7, 9, 32, 48, 3, 1, 1, False),
    ("odd 3x3s2 32>40 @9x7", 5, 9, 7, 32, 40, 3, 2, 1, False),
    ("wide 3x3 32>32 @16", 9, 16, 16, 32, 32, 3, 1, 1, False),
    ("5x5 32>64 @8", 7, 8, 8, 32, 64, 5, 1, 2, False),
    ("FULL l1 3x3 64>64 @8", 3200, 8, 8, 64, 64, 3, 1, 1, True),
    ("FULL l2a 3x3s2 64>128", 3200, 8, 8, 64, 128, 3, 2, 1, True),
    ("FULL l2 3x3 128>128 @4", 3200, 4, 4, 128, 128, 3, 1, 1, True),
    ("FULL l2ds 1x1s2 64>128", 3200, 8, 8, 64, 128, 1, 2, 0, True),
    ("FULL l3a 3x3s2 128>256", 3200, 4, 4, 128, 256, 3, 2, 1, True),
    ("FULL l3 3x3 256>256 @2", 3200, 2, 2, 256, 256, 3, 1, 1, True),
    ("FULL l4a 3x3s2 256>512", 3200, 2, 2, 256, 512, 3, 2, 1, True),
    ("FULL l4 3x3 512>512 @1", 3200, 1, 1, 512, 512, 3, 1, 1, True),
]
LINEAR = [("fc 512>10", 3200, 512, 10), ("mlp 784>512", 256, 784, 512), ("small 64>24", 100, 64, 24)]


def child(start: int) -> None:
    import torch
    import torch.nn.functional as F
    from blades_b200.ops import conv as kc
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    dev = torch.device("cuda")

    def timeit(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

    idx = 0
    for name, NB, H, W, Cin, Cout, k, s, p, timed in CASES:
        if idx < start:
            idx += 1
            continue
        print(f"CASE {idx} begin", flush=True)
        torch.manual_seed(idx)
        x = torch.randn(NB, Cin, H, W, device=dev).contiguous(memory_format=torch.channels_last)
        w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5).contiguous(memory_format=torch.channels_last)
        w2d = w.permute(0, 2, 3, 1).reshape(Cout, -1)
        assert w2d.data_ptr() == w.data_ptr()
        ref64 = F.conv2d(x.double(), w.double(), None, s, p)
        y = kc.conv_fprop(x, w2d, (k, k), s, p)
        msg = f"{name:28s} "
        if y is None:
            msg += "fprop UNSUPPORTED "
        else:
            torch.cuda.synchronize()
            msg += f"fprop rel {rel(y, ref64):.2e} "
        gy = torch.randn_like(ref64.float()).contiguous(memory_format=torch.channels_last)
        x64 = x.double().requires_grad_(True)
        (gref,) = torch.autograd.grad(F.conv2d(x64, w.double(), None, s, p), x64, gy.double())
        add = torch.randn(NB, Cin, H, W, device=dev).contiguous(memory_format=torch.channels_last)
        gx = kc.conv_dgrad(gy, w2d, (k, k), s, p, (H, W), Cin)
        if gx is None:
            msg += "dgrad UNSUPPORTED "
        else:
            torch.cuda.synchronize()
            msg += f"dgrad rel {rel(gx, gref):.2e} "
            acc = add.clone(memory_format=torch.channels_last)
            kc.conv_dgrad(gy, w2d, (k, k), s, p, (H, W), Cin, add=acc, out=acc)
            torch.cuda.synchronize()
            msg += f"dgrad+acc rel {rel(acc, gref + add.double()):.2e} "
        if timed and y is not None and gx is not None:
            t_f = timeit(lambda: kc.conv_fprop(x, w2d, (k, k), s, p))
            t_d = timeit(lambda: kc.conv_dgrad(gy, w2d, (k, k), s, p, (H, W), Cin))
            c_f = timeit(lambda: F.conv2d(x, w, None, s, p))
            c_d = timeit(lambda: torch.ops.aten.convolution_backward(gy, x, w, None, [s, s], [p, p], [1, 1], False,
                                                                      [0, 0], 1, [True, False, False]))
            fl = 2.0 * ref64.numel() * Cin * k * k
            msg += (f"| fprop {t_f:7.1f} us ({fl / t_f / 1e6:6.1f} TF/s) cudnn {c_f:7.1f} | dgrad {t_d:7.1f} us "
                    f"cudnn {c_d:7.1f}")
        print("RESULT " + msg, flush=True)
        idx += 1
    for name, M, K, N in LINEAR:
        if idx < start:
            idx += 1
            continue
        print(f"CASE {idx} begin", flush=True)
        torch.manual_seed(idx)
        x = torch.randn(M, K, device=dev)
        w = torch.randn(N, K, device=dev) / K ** 0.5
        b = torch.randn(N, device=dev)
        y = kc.linear_fprop(x, w, b)
        msg = f"{name:28s} "
        msg += "fprop UNSUPPORTED " if y is None else f"fprop rel {rel(y, F.linear(x.double(), w.double(), b.double())):.2e} "
        ld = (N + 3) // 4 * 4
        gyp = torch.zeros(M, ld, device=dev)
        gyp[:, :N] = torch.randn(M, N, device=dev)
        gx = kc.linear_dgrad(gyp[:, :N], w)
        msg += "dgrad UNSUPPORTED" if gx is None else f"dgrad rel {rel(gx, gyp[:, :N].double() @ w.double()):.2e}"
        print("RESULT " + msg, flush=True)
        idx += 1
    print("ALL DONE", flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
        return
    start, total = 0, len(CASES) + len(LINEAR)
    while start < total:
        proc = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(start)], capture_output=True,
                              text=True, timeout=600)
        last = start
        for ln in proc.stdout.splitlines():
            if ln.startswith("CASE "):
                last = int(ln.split()[1])
            if ln.startswith("RESULT "):
                print(ln[7:], flush=True)
        if "ALL DONE" in proc.stdout:
            break
        err = (proc.stderr or "").strip().splitlines()[-3:]
        print(f"case {last} FAILED (rc={proc.returncode}): {' | '.join(err)[:600]}", flush=True)
        start = last + 1


if __name__ == "__main__":
    main()
