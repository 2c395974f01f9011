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
