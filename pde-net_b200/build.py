"""Build libpdenet_b200.so (hand-written CUDA for sm_100a + the C-ABI) in-tree with nvcc.

    python pde-net_b200/build.py [--force]

One translation unit per (scalar type, filter size) so the heavy kernels compile in parallel.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpdenet_b200.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]
TYPES = [("f64", "double"), ("f32", "float")]
KS = [3, 5, 7]


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [
        os.path.join(HERE, "..", "include", "pdenet_b200.h")]


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(s) <= t for s in _sources())


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    for tname, ctype in TYPES:
        for k in KS:
            o = os.path.join(OBJ, f"inst_{tname}_k{k}.o")
            jobs.append((o, [nvcc, *NVCC_FLAGS, f"-DPDENET_T={ctype}", f"-DPDENET_K={k}", "-c",
                             os.path.join(CSRC, "inst.cu"), "-o", o]))
    o = os.path.join(OBJ, "api.o")
    jobs.append((o, [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, "api.cu"), "-o", o]))
    with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        for out in ex.map(lambda j: _run(j[1]), jobs):
            if verbose and out.strip():
                print(out)
    _run([nvcc, "-shared", "-o", LIB, *[j[0] for j in jobs], "-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
