"""The reference's OWN test-drive suites (test/test_hrweno.f90, test_fluxes.f90, test_tvdode.f90, test_grid.f90), executed
over the reference's OWN library source, both through the Fortran-subset translator tools/f90exec/f90py.py.

This closes the loop on the translator: the very procedures whose executed outputs pin the oracle bit for bit
(tests/test_reference_source_exec.py) also satisfy every assertion the reference makes about itself.  Only
`testdrive`'s `check` is supplied here (test-drive is an un-vendored dependency, fpm.toml:22): |actual - expected| <= thr,
relative to |expected| with rel=.true., exact equality without thr.  Needs the reference tree, so it runs where
/root/reference exists (this container) and is skipped elsewhere (the GPU box)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools", "f90exec"))
import f90py  # noqa: E402
from f90py import FArr  # noqa: E402

REFERENCE = os.environ.get("HRWENO_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "test")), reason="the reference tree is not on this machine")


class ErrorBox:
    """type(error_type), allocatable :: error  -- allocated by a failing check"""

    def __init__(self):
        self.msg = None


def check(error, actual, expected=None, message=None, more=None, thr=None, rel=False):
    if error.msg is not None:
        return
    a = actual.a if isinstance(actual, FArr) else np.asarray(f90py.val(actual))
    if expected is None:
        ok = bool(np.all(a))
    else:
        e = expected.a if isinstance(expected, FArr) else np.asarray(f90py.val(expected))
        if a.shape != e.shape and a.size != 1 and e.size != 1:
            ok = False
        elif thr is None:
            ok = bool(np.all(a == e)) if a.dtype.kind in "iub" else bool(np.all(np.abs(a - e) <= np.finfo(f90py.rdtype()).eps))
        else:
            ok = bool(np.all(np.abs(a - e) <= (thr * np.abs(e) if rel else thr)))
    if not ok:
        error.msg = f"check failed: actual={actual!r} expected={expected!r} thr={thr} rel={rel}"


def load(test_file):
    P = f90py.Program(skip_io=True)
    for f in ("src/hrweno_weno.f90", "src/hrweno_fluxes.f90", "src/hrweno_tvdode.f90", "src/hrweno_grids.f90"):
        P.add_source(os.path.join(REFERENCE, f))
    P.add_source(os.path.join(REFERENCE, "test", test_file), skip=[n for n in ("collect_tests_hrweno", "collect_tests_fluxes",
                                                                             "collect_tests_tvdode", "collect_tests_grid")])
    P.ns["check"] = check
    P.ns["allocated"] = lambda x: x is not None and not (isinstance(x, ErrorBox) and x.msg is None)
    P.procs.setdefault("check", {"name": "check"})  # `call check(...)` resolves to the Python implementation above
    del P.procs["check"]
    return P.build()


SUITES = {
    "test_hrweno.f90": ["test_weno_uniform", "test_calc_cnu", "test_weno_nonuniform"],
    "test_fluxes.f90": ["test_allfluxes"],
    "test_tvdode.f90": ["test_rktvd", "test_mstvd"],
    "test_grid.f90": ["test_linear", "test_log", "test_geometric", "test_bilinear"],
}


@pytest.mark.parametrize("suite,test", [(s, t) for s, ts in SUITES.items() for t in ts])
def test_reference_suite_passes_on_the_executed_reference_source(suite, test):
    ns = load(suite)
    assert test in ns, f"{test} not found in {suite}"
    err = ErrorBox()
    ns[test](err)
    assert err.msg is None, err.msg


@pytest.mark.parametrize("suite,test", [(s, t) for s, ts in SUITES.items() for t in ts])
def test_reference_suite_passes_on_the_executed_reference_source_real32(suite, test):
    """the same ten tests on the REAL32 build of the library (hrweno_kinds.F90:9-10): suites and library translated and
    executed with rk = real32 -- the execution that produced tests/golden/ref_exec_f32_*.npz"""
    with f90py.real_kind(4):
        ns = load(suite)
        err = ErrorBox()
        ns[test](err)
    assert err.msg is None, err.msg


def test_a_failing_check_is_reported():
    """the harness is not vacuous: a wrong expectation allocates `error`"""
    err = ErrorBox()
    check(err, FArr(np.array([1.0, 2.0])), FArr(np.array([1.0, 2.1])), thr=1e-3)
    assert err.msg is not None
    err = ErrorBox()
    check(err, FArr(np.array([1.0, 2.0])), FArr(np.array([1.0, 2.0005])), thr=1e-3, rel=True)
    assert err.msg is None
