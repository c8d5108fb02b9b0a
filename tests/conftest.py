import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds")


def _built():
    need = [os.path.join(ROOT, "hr-weno_b200", "lib", "libhrweno_b200.so"), os.path.join(ROOT, "oracle", "libhrweno_oracle.so")]
    need += [os.path.join(ROOT, "examples", e) for e in
             ("example1_burgers_1d_fv", "example2_pbe_2d_fv", "example3_pbe_2d_growth", "example4_multi_gpu", "grid_dump")]
    return all(os.path.exists(p) for p in need)


@pytest.fixture(scope="session")
def pkg():
    """the product package (hr-weno_b200 loaded as hrweno_b200); does not touch the GPU by itself.
    A fresh checkout has no built artefacts (they are git-ignored): build them once (nvcc cross-compiles without a GPU)."""
    if not _built():
        graft.build()
    return graft.load_package()


@pytest.fixture(scope="session")
def ref(pkg):
    """the CPU oracle (test infrastructure)"""
    return graft.load_oracle()


@pytest.fixture(scope="session")
def npo():
    from oracle import np_oracle

    return np_oracle


@pytest.fixture(scope="session")
def gpu_lib(pkg):
    """the loaded CUDA library; GPU tests must run the real thing, so a missing library is an error"""
    lib = pkg.lib()
    assert lib.hrweno_device_count() >= 1, "no CUDA device visible"
    return lib


def pulse(nc=30):
    """test/test_hrweno.f90:45-46: v = 0; v(nc/3:2*nc/3) = 1 (1-based, inclusive)"""
    v = np.zeros(nc)
    v[nc // 3 - 1 : 2 * nc // 3] = 1.0
    return v


def ex1_ic(x):
    """example1:124-133"""
    xa, xb, va, vb = -4.0, 2.0, 1.0, -0.5
    ic = va + (vb - va) / (xb - xa) * (x - xa)
    return np.maximum(np.minimum(ic, va), vb)


def ex2_ic(c1, c2):
    """example2:49-51,157-168 -> u[(j)*nc1 + i]"""
    m1 = (c1 >= 1.0) & (c1 <= 3.0)
    m2 = (c2 >= 1.0) & (c2 <= 3.0)
    return (m2[:, None] & m1[None, :]).astype(np.float64)


def normwise(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
