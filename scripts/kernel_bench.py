"""Micro-benchmarks of the hand-written kernels at the headline shapes, with roofline fractions
(denominators: MEASURED_PEAKS.json -- hbm_gbs copy bandwidth)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blades_b200.ops import select, combine, gram, wgrad
from blades_b200.parallel.matrix import VirtualRows

peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = peaks.get("hbm_gbs", 6650.0)

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)

d = 11181642
ld = (d + 63) // 64 * 64
for n, f in ((100, 20), (200, 40)):
    store = torch.randn(n, ld, device="cuda") * 0.01
    U = store[:, :d]
    v = VirtualRows("alie", 0.2858, list(range(f)))
    res = []
    if n <= 128:
        res.append(("trimmed_mean b=%d" % f, timeit(lambda: select.trimmed_mean(U, f)), n * d * 4))
        res.append(("trimmed_mean+ALIE(virtual f=%d)" % f, timeit(lambda: select.trimmed_mean(U, f, virtual=v)), (n - f) * d * 4))
        res.append(("median", timeit(lambda: select.median(U)), n * d * 4))
    res.append(("row_combine", timeit(lambda: combine.row_combine(U, [1.0 / n] * n)), n * d * 4))
    for prec in ("tf32", "tf32x3", "fp32"):
        res.append(("gram %s" % prec, timeit(lambda: gram.gram(U, precision=prec), iters=3), n * d * 4))
    for name, ms, nbytes in res:
        gbs = nbytes / ms / 1e6
        print(f"N={n:4d} {name:36s} {ms:8.3f} ms  {gbs:8.1f} GB/s  {gbs / HBM:6.3f} of measured HBM copy ({HBM:.0f} GB/s)")
    del store, U
# grouped wgrad at ResNet-18 layer shapes: (M=Cout, T=B*L, N=Cin*9), 100 clients
for (M, T, N) in ((64, 2048, 576), (128, 512, 1152), (256, 128, 2304), (512, 32, 4608), (512, 32, 2304)):
    n = 100
    a_t = torch.randn(n, T, M, device="cuda")
    b = torch.randn(n, T, N, device="cuda")
    out = torch.empty(n, M, N, device="cuda")
    ms = timeit(lambda: wgrad.grouped_wgrad(a_t.transpose(1, 2), b, out, -0.1))
    os.environ["BLADES_WGRAD_OFF"] = "1"
    ms2 = timeit(lambda: torch.baddbmm(out, a_t.transpose(1, 2), b, beta=0.0, alpha=-0.1, out=out))
    nbytes = (a_t.numel() + b.numel() + out.numel()) * 4
    fl = 2.0 * n * M * T * N
    print(f"wgrad M={M:4d} T={T:5d} N={N:5d}: tcgen05 {ms:7.3f} ms ({nbytes / ms / 1e6:7.1f} GB/s, {fl / ms / 1e9:7.1f} TFLOP/s)   cuBLAS baddbmm {ms2:7.3f} ms")
