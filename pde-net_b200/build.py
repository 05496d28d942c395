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
TORCH_LIB = os.path.join(HERE, "libpdenet_b200_torch.so")   # TORCH_LIBRARY shim over the C-ABI (csrc/torch_shim.cpp)

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]
NVCC_FLAGS += os.environ.get("PDENET_EXTRA_NVCC_FLAGS", "").split()   # experiments: -D... switches of the kernels
TYPES = [("f64", "double"), ("f32", "float")]
KS = [3, 5, 7]


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h", ".cpp"))] + [
        os.path.join(HERE, "..", "include", "pdenet_b200.h")]


ID_FILE = os.path.join(OBJ, "build_id.txt")


def source_hash() -> str:
    """Content hash of every source the library is built from (csrc/* and the C-ABI header).  It is
    compiled into the library (``pdenet_build_id()``) and checked by ``_lib.lib()``, so a stale binary
    is rebuilt or refused instead of being called with a mismatching ABI."""
    import hashlib
    h = hashlib.sha256()
    for src in _sources():
        if os.path.isfile(src):
            h.update(os.path.basename(src).encode())
            h.update(open(src, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def up_to_date() -> bool:
    """True if the shared object exists and was built from the present sources (content hash, not
    mtimes: the tree is copied to the GPU box without them)."""
    if not os.path.exists(LIB) or not os.path.exists(TORCH_LIB) or not os.path.exists(ID_FILE):
        return False
    return open(ID_FILE).read().strip() == source_hash()


def _torch_paths():
    import torch
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths()
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(getattr(torch._C, "_GLIBCXX_USE_CXX11_ABI", True))
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    return inc + [cuda_inc], libdir, abi


def _host_cxx():
    """The system g++ -- deliberately NOT $CXX: this image exports CXX=/opt/gcc/bin/g++, a wrapper whose search
    path only has the static libstdc++.a, and a shim carrying its own copy of libstdc++ next to torch's shared
    one crashes the first time it formats a number into an error message (two std::locale worlds)."""
    return os.environ.get("PDENET_CXX", "g++")


def _torch_shim_compile_cmd(obj):
    """g++ (no device code in the shim): torch headers + the C-ABI header."""
    inc, _, abi = _torch_paths()
    cxx = _host_cxx()
    return [cxx, "-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-D_GLIBCXX_USE_CXX11_ABI=%d" % abi,
            *["-isystem" + i for i in inc], "-c", os.path.join(CSRC, "torch_shim.cpp"), "-o", obj]


def _torch_shim_link_cmd(obj):
    _, libdir, _ = _torch_paths()
    cxx = _host_cxx()
    return [cxx, "-shared", "-o", TORCH_LIB, obj, "-L" + libdir, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda",
            "-ltorch", "-L" + HERE, "-l:libpdenet_b200.so", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + libdir]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Build both shared objects in place.  Serialised across processes by a file lock: under torchrun every rank
    may find the binaries stale at the same time, and only one of them may compile."""
    if not force and up_to_date():
        return LIB
    import fcntl
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and up_to_date():      # another process built it while this one waited
                return LIB
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    build_id = source_hash()
    jobs = []
    for tname, ctype in TYPES:
        for k in KS:
            o = os.path.join(OBJ, f"inst_{tname}_k{k}.o")
            jobs.append((o, [nvcc, *NVCC_FLAGS, f"-DPDENET_T={ctype}", f"-DPDENET_K={k}", "-c",
                             os.path.join(CSRC, "inst.cu"), "-o", o]))
    o = os.path.join(OBJ, "api.o")
    jobs.append((o, [nvcc, *NVCC_FLAGS, '-DPDENET_BUILD_ID="%s"' % build_id, "-c", os.path.join(CSRC, "api.cu"),
                     "-o", o]))
    shim_obj = os.path.join(OBJ, "torch_shim.o")
    shim_job = (shim_obj, _torch_shim_compile_cmd(shim_obj))

    def compile_job(job):
        """Compile one object unless an object built from the same inputs (sources it can include + command
        line, content-hashed) is already there: an edit of api.cu does not recompile the six kernel units."""
        import hashlib
        obj, cmd = job
        name = os.path.basename(obj)
        if name.startswith("inst_"):
            deps = [f for f in _sources() if f.endswith(".cuh") or f.endswith("inst.cu")]
        elif name == "api.o":
            deps = [f for f in _sources() if not f.endswith(".cpp") and not f.endswith("inst.cu")]
        else:
            deps = [f for f in _sources() if f.endswith((".cpp", ".h"))]
        h = hashlib.sha256(" ".join(c for c in cmd if "PDENET_BUILD_ID" not in c).encode())
        for f in deps:
            h.update(open(f, "rb").read())
        key = h.hexdigest()
        idf = obj + ".id"
        # api.o embeds the build id of the WHOLE tree: it is recompiled whenever anything changed
        if name != "api.o" and not force and os.path.exists(obj) and os.path.exists(idf) and open(idf).read() == key:
            return ""
        out = _run(cmd)
        with open(idf, "w") as f:
            f.write(key)
        return out

    with cf.ThreadPoolExecutor(max_workers=min(len(jobs) + 1, os.cpu_count() or 4)) as ex:
        for out in ex.map(compile_job, jobs + [shim_job]):
            if verbose and out.strip():
                print(out)
    _run([nvcc, "-shared", "-o", LIB, *[j[0] for j in jobs], "-gencode", "arch=compute_100a,code=sm_100a"])
    _run(_torch_shim_link_cmd(shim_obj))
    syms = _run(["nm", "-D", "--defined-only", TORCH_LIB])
    if "_ZNSo9_M_insertIlEERSoT_" in syms:
        raise RuntimeError("libpdenet_b200_torch.so carries a static copy of libstdc++ (host compiler %s): it must "
                           "link the shared one torch uses" % _host_cxx())
    with open(ID_FILE, "w") as f:
        f.write(build_id + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
