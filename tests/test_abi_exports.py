"""CPU-only: the C-ABI library loads and exports every symbol include/hrweno_b200.h declares."""
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "hrweno_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hrweno_[a-z0-9_]+)\s*\(", src)) - {"hrweno_flux_fn", "hrweno_rhs_fn"})


def test_library_exports_every_declared_symbol(pkg):
    lib_path = pkg._abi.LIB_PATH
    assert os.path.exists(lib_path), "build the library first: python __graft_entry__.py"
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (hrweno_[a-z0-9_]+)", out))
    declared = _header_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_ctypes_prototypes_cover_the_header(pkg):
    assert sorted(pkg._abi.PROTOTYPES) == _header_symbols()


def test_library_loads_without_compute(pkg):
    lib = pkg.lib()  # dlopen + prototypes + ABI version; no kernel is launched
    assert lib.hrweno_abi_version() == pkg._abi.ABI_VERSION
    assert lib.hrweno_device_count() >= 0
    assert isinstance(lib.hrweno_last_error(), bytes)


def test_no_cpu_fallback_without_device(pkg):
    """On a box without a GPU every compute entry point must fail loudly (status ECUDA), never compute."""
    lib = pkg.lib()
    if lib.hrweno_device_count() > 0:
        pytest.skip("a device is present")
    with pytest.raises(pkg.HrwenoError) as ei:
        pkg.hrweno_weno.weno(30, 3, 1e-6)
    assert ei.value.status == pkg._abi.ECUDA


def test_input_validation_matches_reference_error_stops(pkg):
    """weno.f90:72-98 / tvdode.f90:80-90: invalid inputs are rejected before any device work"""
    for args in [(0, 3, 1e-6), (10, 0, 1e-6), (10, 4, 1e-6), (10, 3, 1e-17)]:
        with pytest.raises(pkg.HrwenoError) as ei:
            pkg.hrweno_weno.weno(*args)
        assert ei.value.status == pkg._abi.EINVAL
    with pytest.raises(pkg.HrwenoError):
        pkg.hrweno_weno.weno(10, 3, 1e-6, xedges=[0.0, 1.0])  # size(xedges) /= ncells + 1
    with pytest.raises(pkg.HrwenoError) as ei:
        pkg.hrweno_tvdode.rktvd(lambda *a: None, 0, 3)
    assert ei.value.status == pkg._abi.EINVAL
    with pytest.raises(pkg.HrwenoError) as ei:
        pkg.hrweno_tvdode.rktvd(lambda *a: None, 10, 4)
    assert ei.value.status == pkg._abi.EINVAL
