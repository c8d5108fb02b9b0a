"""Multi-GPU parity where the driver can see it: tests/mgpu_worker.py (slab(halo k) == global, bit for bit)
launched under torch.distributed.run for every power of two <= the number of visible GPUs, and the single-process
multi-GPU entry of the C ABI (hrweno_mgpu_*, examples/example4_multi_gpu.cpp)."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _world_sizes(gpu_lib):
    n = gpu_lib.hrweno_device_count()
    return [w for w in (2, 4, 8) if w <= n]


def test_slabs_under_torchrun_equal_single_domain_oracle(gpu_lib):
    """one process per GPU (the layout bench.py scales with): every rank's slab must equal the same cells of the
    oracle's single-domain run bit for bit in strict mode (1D k=1..3 / rktvd 1..3 / mstvd, 2D mstvd + rktvd3), fast-mode
    slabs must equal the single-domain fast result bit for bit and stay within 1e-12 of the oracle"""
    sizes = _world_sizes(gpu_lib)
    if not sizes:
        pytest.skip("needs at least 2 visible GPUs")
    for world in sizes:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_worker.py")]
        p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
        out = p.stdout + p.stderr
        print(out[-2000:])
        assert p.returncode == 0, out[-4000:]
        assert f"mgpu parity: world={world} failures=0" in out, out[-4000:]
