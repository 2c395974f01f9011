"""bench.py contract pieces that can be checked without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_reports_unavailable():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    assert rec["impl"] == "reference" and "unavailable" in rec


def test_graft_entry_build_is_idempotent():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    g.build()
    assert os.path.exists(os.path.join(ROOT, "blades_b200", "_cuda.so"))
    assert os.path.exists(os.path.join(ROOT, "blades_b200", "_host.so"))
