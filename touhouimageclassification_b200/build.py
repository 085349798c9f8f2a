"""In-tree nvcc build of the C-ABI library (``csrc/libtic_b200.so``) and of the oracle's C pieces.

sm_100a only: ``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo``. The built ``.so`` is
git-ignored but travels with the repo snapshot to the GPU box, so nothing is JIT-compiled there.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# Development variants (kernel experiments compiled with extra -D flags next to the shipped library):
#   TIC_LIB_VARIANT=name [TIC_VARIANT_FLAGS="-DFOO -DBAR"] python -m touhouimageclassification_b200.build
# builds csrc/libtic_b200_name.so; a process started with TIC_LIB_VARIANT=name loads that file instead.
VARIANT = os.environ.get("TIC_LIB_VARIANT", "")
LIB_PATH = os.path.join(CSRC, f"libtic_b200_{VARIANT}.so" if VARIANT else "libtic_b200.so")
OBJ_DIR = os.path.join(CSRC, f"build_{VARIANT}" if VARIANT else "build")

NVCC = os.environ.get("TIC_NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
] + (os.environ.get("TIC_VARIANT_FLAGS", "").split() if VARIANT else [])


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths):
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every ``csrc/*.cu`` into ``csrc/libtic_b200.so``. Returns the library path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "tic_b200.h"))
    hdr_digest = _digest([h for h in headers if os.path.exists(h)])
    objs, jobs = [], []
    for src in _sources():
        src_path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        stamp = obj + ".sha"
        digest = _digest([src_path]) + hdr_digest
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
            continue
        jobs.append((src_path, obj, stamp, digest))

    def compile_one(job):
        src_path, obj, stamp, digest = job
        cmd = [NVCC, *NVCC_FLAGS, "-I", os.path.join(os.path.dirname(HERE), "include"), "-c", src_path, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src_path}:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as f:
            f.write(digest)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or force or not os.path.exists(LIB_PATH):
        cmd = [NVCC, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
