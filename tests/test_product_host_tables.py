"""Product HOST code against the oracle without a GPU: the per-cell reconstruction tables cnu of `weno(ncells, k, eps, xedges)`
(weno.f90:221-297) are computed on the host by the library in both kinds before they are uploaded; tests/cpp/cnu_host_check.cu
links the library's own objects and calls those functions.  Bar: bit-identical to the oracle (which reproduces the reference's
`weno_calc_cnu` executed from source, tests/golden/ref_exec_reconstruct.npz / ref_exec_f32_reconstruct.npz)."""
import glob
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT

CSRC = os.path.join(ROOT, "hr-weno_b200", "csrc")


@pytest.fixture(scope="module")
def harness(pkg, tmp_path_factory):
    objs = [o for o in sorted(glob.glob(os.path.join(CSRC, "build", "*.o"))) if not o.endswith("real32.o")]  # real32.cu is #included
    if not objs:
        pytest.skip("the library's objects are not in hr-weno_b200/csrc/build (built elsewhere)")
    exe = str(tmp_path_factory.mktemp("cnu_host_check") / "cnu_host_check")
    nvcc = "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC, os.path.join(ROOT, "tests", "cpp", "cnu_host_check.cu"), "-o", exe] + objs + ["-lcuda"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    return exe


def _tables(exe, xe, k):
    kind = xe.dtype.itemsize
    out = subprocess.run([exe], input=struct.pack("qii", xe.size - 1, k, kind) + xe.tobytes(), capture_output=True, check=True).stdout
    return np.frombuffer(out, dtype=xe.dtype).reshape(xe.size - 1, k + 1, k)


@pytest.mark.parametrize("kind", ["real64", "real32"])
def test_product_host_cnu_equals_oracle(harness, ref, kind):
    if kind == "real32":
        from oracle import ref32 as oracle

        dtype = np.float32
    else:
        oracle, dtype = ref, np.float64
    for seed in range(60):
        rng = np.random.default_rng(seed)
        nc, k = int(rng.integers(1, 80)), int(rng.integers(1, 4))
        xe = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 1.5, nc))])
        if seed % 3 == 0:
            xe = xe**2  # strongly stretched
        if seed % 7 == 0:
            xe = np.linspace(-5.0, 5.0, nc + 1)  # uniform edges through the non-uniform formulas (test_hrweno.f90:69-112)
        xe = np.ascontiguousarray(xe, dtype=dtype)
        assert np.array_equal(_tables(harness, xe, k), oracle.calc_cnu(xe, k)), (kind, seed, nc, k)


def test_product_host_cnu_equals_executed_reference_source(harness):
    """directly against the tables the reference's own weno_calc_cnu produced when executed from source"""
    for kind, fixture in ((np.float64, "ref_exec_reconstruct.npz"), (np.float32, "ref_exec_f32_reconstruct.npz")):
        g = np.load(os.path.join(ROOT, "tests", "golden", fixture))
        for gname in ("uniform", "cubic"):
            for k in (1, 2, 3):
                xe = np.ascontiguousarray(g["xe_" + gname], dtype=kind)
                assert np.array_equal(_tables(harness, xe, k), g[f"cnu_{gname}_k{k}"]), (fixture, gname, k)
