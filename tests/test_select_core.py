"""The partition-only trimmed-mean core of the GPU kernel (csrc/cuda/select_part_core.cuh + the generated sorting
networks) is compiled for the HOST and checked against a sort-based double-precision reference: 25 600 cases over
NP = 8..128, f in {0, Q, Q+1, 2Q+5}, ALIE / IPM virtual values, ties, constants and 1e6 outliers.  No GPU needed --
the same source files are what nvcc compiles into coord_select_part_kernel."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_partition_core_on_host(tmp_path):
    src = os.path.join(ROOT, "blades_b200", "csrc", "host", "select_core_check.cu")
    exe = str(tmp_path / "select_core_check")
    build = subprocess.run(["nvcc", "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, src],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout[-2000:]
    assert "0 failures" in run.stdout


def test_sorting_network_generator_is_reproducible(tmp_path):
    """The checked-in generated header is exactly what gen_sortnet.py emits (and every network sorts)."""
    import sys
    gen = os.path.join(ROOT, "blades_b200", "csrc", "gen_sortnet.py")
    out = str(tmp_path / "sortnet_gen.cuh")
    subprocess.run([sys.executable, gen, out], check=True)
    assert open(out).read() == open(os.path.join(ROOT, "blades_b200", "csrc", "cuda", "gen", "sortnet_gen.cuh")).read()
    chk = subprocess.run([sys.executable, gen, "check"], capture_output=True, text=True)
    assert chk.returncode == 0 and "128" in chk.stdout


def test_dispatch_predicate_of_the_select_launcher():
    """Which kernel the launcher picks (asked from the library's own predicate, no CUDA call): the headline round
    (100 clients, 20 ALIE attackers fused as virtual rows, Trimmedmean(nb=20)) takes the partition-only kernel."""
    so = os.path.join(ROOT, "blades_b200", "_cuda.so")
    if not os.path.exists(so):
        pytest.skip("library not built")
    from blades_b200.ops import select
    # (honest rows, other real rows, virtual multiplicity, kind, mode 0 = trimmed mean / 1 = median, b)
    assert select.kernel_choice(80, 0, 20, "alie", 0, 20) == "partition"        # the headline
    assert select.kernel_choice(80, 0, 20, "ipm", 0, 20) == "partition"
    assert select.kernel_choice(80, 0, 0, None, 0, 20) == "partition"           # no attack, n = 4b
    assert select.kernel_choice(40, 0, 10, "alie", 0, 10) == "partition"        # 50 clients, 20 % attackers
    assert select.kernel_choice(45, 3, 12, "alie", 0, 12) == "partition"        # 48 real rows, 3 outside the statistics
    assert select.kernel_choice(80, 0, 20, "alie", 1, 0) == "network"           # median keeps the full network
    assert select.kernel_choice(100, 0, 0, None, 0, 20) == "network"            # 100 real rows: n != 4b
    assert select.kernel_choice(80, 0, 10, "alie", 0, 20) == "network"          # f < b
    assert select.kernel_choice(75, 0, 25, "alie", 0, 20) == "network"          # padded (75 -> 80 slots)
    assert select.kernel_choice(80, 0, 20, "alie", 0, 19) == "network"
    assert select.kernel_choice(160, 0, 40, "alie", 0, 40) == "large"


def test_batchnorm_cluster_plan():
    """The measured dispatch of the per-client BatchNorm kernels (cluster / DSMEM form vs the plain one), asked from
    the library without a GPU: ResNet-18 layers at 100 clients x batch 32."""
    import ctypes as C
    so = os.path.join(ROOT, "blades_b200", "_cuda.so")
    if not os.path.exists(so):
        pytest.skip("library not built")
    from blades_b200.ops import _loader
    lib = _loader.cuda_lib()

    def plan(n, B, Cc, HW, staged):
        q, sm = C.c_int(0), C.c_longlong(0)
        S = lib.bl_client_bn_cluster_plan(n, B, Cc, HW, staged, C.byref(q), C.byref(sm))
        return (S, q.value, sm.value) if S else None
    assert plan(100, 32, 64, 256, 1) == (8, 4, 64 * 1024)        # stem forward: 8-CTA clusters, 16-channel tiles
    assert plan(100, 32, 64, 256, 2) is None                     # stem backward would need 128 KB per CTA: plain kernel
    assert plan(100, 32, 64, 64, 1) == (4, 8, 64 * 1024)         # layer1 forward
    assert plan(100, 32, 64, 64, 2) == (8, 8, 64 * 1024)         # layer1 backward
    assert plan(100, 32, 128, 16, 1) is None                     # R = 512 and shorter: plain kernel measured faster
    assert plan(100, 32, 256, 4, 2) is None
    S, q, sm = plan(1, 32, 64, 256, 1)                           # fedavg visit (one client): still split 8 ways
    assert S == 8 and sm <= 64 * 1024
    assert plan(4, 33, 64, 31, 1) is None or plan(4, 33, 64, 31, 1)[2] <= 64 * 1024      # odd sizes never exceed 64 KB
