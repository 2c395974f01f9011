"""Build the native libraries in-tree (sm_100a only):

    blades_b200/_cuda.so   <- csrc/cuda/*.cu   (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo)
    blades_b200/_host.so   <- csrc/host/*.cpp  (g++ -O3)

``python -m blades_b200.ops.build [--force] [--verbose]``.  nvcc cross-compiles without a GPU.
Objects are cached under ``build/`` keyed by content hash (stale ones are pruned after each link); translation units compile in parallel.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(os.path.dirname(PKG), "build")
CUDA_SO = os.path.join(PKG, "_cuda.so")
HOST_SO = os.path.join(PKG, "_host.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
CXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-Wall"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stamp(paths) -> str:
    h = hashlib.sha1()
    for p in sorted(paths):
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:           # content, not mtime: snapshots copied to a GPU box keep their stamps
            h.update(f.read())
    return h.hexdigest()[:16]


def _headers(d):
    out = []
    for root, _, files in os.walk(d):
        out += [os.path.join(root, f) for f in files if f.endswith((".cuh", ".h", ".hpp"))]
    return out


def _compile(cmd, log):
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"compile failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return r.stderr


def build_cuda(force=False, verbose=False) -> str:
    src_dir = os.path.join(CSRC, "cuda")
    gen = os.path.join(src_dir, "gen", "sortnet_gen.cuh")
    if not os.path.exists(gen):
        os.makedirs(os.path.dirname(gen), exist_ok=True)
        subprocess.check_call([sys.executable, os.path.join(CSRC, "gen_sortnet.py"), gen])
    srcs = sorted(os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith(".cu"))
    hdr_stamp = _stamp(_headers(src_dir))
    stamp_file = CUDA_SO + ".stamp"
    want = _stamp(srcs) + hdr_stamp
    if not force and os.path.exists(CUDA_SO) and os.path.exists(stamp_file) and open(stamp_file).read() == want:
        return CUDA_SO          # up to date (content hash): nothing to do, e.g. on a GPU box snapshot
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    jobs, objs = [], []
    for s in srcs:
        obj = os.path.join(BUILD, os.path.basename(s) + "." + _stamp([s]) + hdr_stamp + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj):
            jobs.append(([nvcc] + NVCC_FLAGS + ["-I", src_dir, "-c", s, "-o", obj], obj + ".log"))
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for out in ex.map(lambda j: _compile(*j), jobs):
            if verbose:
                print(out)
    # same -gencode at link time: otherwise nvcc adds an (empty) device-link stub for its default arch, sm_52
    _compile([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", CUDA_SO] + objs + ["-lcudart"],
             os.path.join(BUILD, "link_cuda.log"))
    with open(stamp_file, "w") as f:
        f.write(want)
    _prune_objects(objs)
    return CUDA_SO


def _prune_objects(current) -> None:
    """Objects are keyed by content hash, so every source edit leaves an orphan behind: keep only the current set."""
    keep = {os.path.basename(o) for o in current}
    for f in os.listdir(BUILD):
        if f.endswith((".o", ".o.log")) and f.replace(".log", "") not in keep:
            os.remove(os.path.join(BUILD, f))


def build_host(force=False, verbose=False) -> str:
    src_dir = os.path.join(CSRC, "host")
    srcs = sorted(os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith(".cpp"))
    if not srcs:
        return ""
    os.makedirs(BUILD, exist_ok=True)
    stamp_file = HOST_SO + ".stamp"
    want = _stamp(srcs + _headers(src_dir))
    if force or not os.path.exists(HOST_SO) or not os.path.exists(stamp_file) or open(stamp_file).read() != want:
        cxx = os.environ.get("CXX", "g++")
        _compile([cxx] + CXX_FLAGS + ["-shared", "-o", HOST_SO] + srcs + ["-lpthread"],
                 os.path.join(BUILD, "host.log"))
        with open(stamp_file, "w") as f:
            f.write(want)
    return HOST_SO


def build_all(force=False, verbose=False):
    return build_cuda(force, verbose), build_host(force, verbose)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build_all(a.force, a.verbose))
