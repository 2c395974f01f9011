"""Documentation artefacts: the Sphinx configuration loads, the example gallery is up to date with the scripts."""
import os
import runpy
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sphinx_conf_loads_and_points_at_the_examples():
    ns = runpy.run_path(os.path.join(ROOT, "docs", "source", "conf.py"))
    assert "sphinx.ext.autodoc" in ns["extensions"]
    assert os.path.isdir(ns["sphinx_gallery_conf"]["examples_dirs"])
    assert os.path.exists(os.path.join(ROOT, "docs", "source", "index.rst"))


def test_gallery_pages_are_current():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "build_gallery.py"), "--check"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
