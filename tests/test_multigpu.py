"""Multi-GPU correctness (DIST-1 / DIST-2) on REAL devices: torchrun-launched, one process per
GPU over NCCL.  Skipped when the box has fewer than two GPUs (the gloo world-2 tests in
test_parallel_cpu.py cover the host logic everywhere)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(script, *args, nproc=2, timeout=600):
    if not torch.cuda.is_available() or torch.cuda.device_count() < nproc:
        pytest.skip("needs %d CUDA devices" % nproc)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", script), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, r.stdout[-2000:]
    return json.loads(lines[-1])


@pytest.mark.parametrize("p2p", [False, True])
@pytest.mark.parametrize("keep", ["on", "off"])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_slab_decomposition_two_gpus_equals_single_gpu(p2p, keep, dtype):
    """cfg5 path at a small size: 2 ranks, forward + reverse slab rollout (NCCL halo exchange, and the halo
    exchange fused into the step kernels as peer stores over NVLink) == the single-GPU periodic rollout."""
    steps = 3
    out = _torchrun("slab_multigpu_check.py", "--size", "512", "--steps", str(steps), "--dtype", dtype, "--keep", keep,
                    "--reps", "1", *(["--p2p"] if p2p else []))
    tol = 1e-10 * steps if dtype == "f64" else 1e-3
    assert out["n_gpus"] == 2
    assert max(out["max_rel_err_vs_single_gpu"].values()) < tol, out


def test_batch_sharding_two_gpus_equals_single_gpu():
    """cfg4 path at a small size: shards + ONE in-place all-reduce of the flat gradient buffer == whole batch."""
    out = _torchrun("dp_multigpu_check.py", "--size", "128", "--batch", "6", "--steps", "3")
    assert out["n_gpus"] == 2 and max(out["max_rel_err_vs_single_gpu"].values()) < 3e-10, out
