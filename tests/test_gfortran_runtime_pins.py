"""The translator's reading of Fortran intrinsics against gfortran's OWN RUNTIME.

No Fortran compiler exists in this image, but two pieces of the GNU Fortran implementation do: libgcc (`__powidf2`, what
gfortran calls for `real**integer`) and libgfortran.so.5 (shipped inside the NumPy / SciPy wheels), which holds the library
versions of the array intrinsics.  tools/f90exec/f90py.py -- whose execution of the reference's source pins the oracle --
restates four semantics by hand that the reference's hot path depends on; each is checked here against the real thing,
bit for bit, through hand-built gfortran array descriptors (GCC >= 8 layout):

  * `x**n`, integer n  (example1:120 `v**2`; grids.f90:222-224 `ratio**i`)          libgcc __powidf2, libgfortran pow_r8_i8
  * `sum(a)` accumulates from the first element to the last  (weno.f90:179-214)      libgfortran sum_r8
  * `eoshift(a, shift=-1, dim=2)` moves towards higher indices, zero-fills            libgfortran eoshift0_4
    (the multi-step history shift, tvdode.f90:262-263)
  * `reshape(source, shape, order=)` of the coefficient tables c1 / c2 / c3 (weno.f90:16-21)   libgfortran reshape_r8

What this does not pin: code gfortran generates inline at a given optimisation level (it expands rank-1 `sum` and small
integer powers itself; without -ffast-math it may not re-associate them, which is the standard's and GCC's documented rule,
not something a runtime call can show).  Skipped where the libraries are not found."""
import ctypes as C
import glob
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools", "f90exec"))
import f90py  # noqa: E402


def _libgfortran():
    for pat in ("numpy.libs", "scipy.libs", "*"):
        for base in sys.path:
            hits = sorted(glob.glob(os.path.join(base, pat, "libgfortran*.so.5*")))
            if hits:
                try:
                    return C.CDLL(hits[0])
                except OSError:
                    pass
    for name in ("libgfortran.so.5",):
        try:
            return C.CDLL(name)
        except OSError:
            pass
    return None


GF = _libgfortran()
needs_gf = pytest.mark.skipif(GF is None or not hasattr(GF, "_gfortran_sum_r8"), reason="no libgfortran.so.5 on this machine")


class _Dim(C.Structure):
    _fields_ = [("stride", C.c_ssize_t), ("lbound", C.c_ssize_t), ("ubound", C.c_ssize_t)]


class _DType(C.Structure):
    _fields_ = [("elem_len", C.c_size_t), ("version", C.c_int), ("rank", C.c_byte), ("type", C.c_byte), ("attribute", C.c_short)]


def _descriptor(a):
    """gfc_array_r8 for a column-major float64 array with lower bounds 1 (libgfortran.h, GCC >= 8; BT_REAL = 3)"""
    assert a.dtype == np.float64 and a.flags["F_CONTIGUOUS"]

    class D(C.Structure):
        _fields_ = [("base_addr", C.c_void_p), ("offset", C.c_size_t), ("dtype", _DType), ("span", C.c_ssize_t), ("dim", _Dim * a.ndim)]

    d = D()
    d.base_addr, d.dtype, d.span = a.ctypes.data, _DType(8, 0, a.ndim, 3, 0), 8
    stride, off = 1, 0
    for i, n in enumerate(a.shape):
        d.dim[i] = _Dim(stride, 1, n)
        off -= stride
        stride *= n
    d.offset = off % (1 << 64)
    d._keep = a
    return d


def test_integer_power_equals_libgcc_powidf2():
    """`real**integer` as gfortran evaluates it: the translator's fpow and the host-side grid mirror's _powi"""
    try:
        powi = C.CDLL("libgcc_s.so.1").__powidf2
    except (OSError, AttributeError):
        pytest.skip("libgcc_s.so.1 without __powidf2")
    powi.restype, powi.argtypes = C.c_double, [C.c_double, C.c_int]
    from __graft_entry__ import load_package

    grids = load_package().hrweno_grids
    rng = np.random.default_rng(0)
    xs = list(rng.uniform(0.5, 1.5, 100)) + list(rng.standard_normal(50)) + [1.02, 1.03, 1.1, 1.01, 1.0003, 3.0, -1.7]
    ns = list(range(0, 130)) + [365, 1000]
    for x in xs:
        want = np.array([powi(x, n) for n in ns])
        assert np.array_equal(np.array([f90py.fpow(float(x), n) for n in ns]), want), x
        with np.errstate(over="ignore"):  # 3.0**1000 is +inf for all three
            assert np.array_equal(grids._powi(float(x), np.array(ns)), want), x
        for n in (-1, -2, -7):
            assert f90py.fpow(float(x), n) == powi(x, n)
        assert f90py.fpow(float(x), 2) == x * x  # example1:120


@needs_gf
def test_integer_power_equals_libgfortran_pow_r8_i8():
    f = GF._gfortran_pow_r8_i8
    f.restype, f.argtypes = C.c_double, [C.c_double, C.c_int64]
    rng = np.random.default_rng(1)
    for x in list(rng.uniform(0.9, 1.2, 60)) + [1.02, 1.03, 1.1]:
        for n in list(range(0, 120)) + [365, 1000]:
            assert f90py.fpow(float(x), n) == f(x, n), (x, n)


@needs_gf
def test_sum_accumulates_in_element_order_like_libgfortran():
    rng = np.random.default_rng(2)
    for m, n in ((3, 40), (7, 25), (12, 9), (1, 5)):
        a = np.asfortranarray(rng.standard_normal((m, n)) * 10.0 ** rng.integers(-8, 8, (m, n)))
        out = np.zeros(n)
        dim = C.c_ssize_t(1)
        GF._gfortran_sum_r8(C.byref(_descriptor(out)), C.byref(_descriptor(a)), C.byref(dim))
        mine = np.array([f90py._seq_sum(f90py.FArr(np.ascontiguousarray(a[:, j]))) for j in range(n)])
        assert np.array_equal(out, mine)
        if m >= 7:  # the check can tell orders apart: summing from the last element differs somewhere
            rev = np.array([f90py._seq_sum(f90py.FArr(np.ascontiguousarray(a[::-1, j]))) for j in range(n)])
            assert not np.array_equal(rev, out)


@needs_gf
@pytest.mark.parametrize("shift,dim", [(-1, 2), (1, 2), (-1, 1), (2, 1), (-3, 2), (0, 1)])
def test_eoshift_equals_libgfortran(shift, dim):
    """tvdode.f90:262-263 uses (shift=-1, dim=2): column j moves to column j+1, column 1 is zero-filled"""
    rng = np.random.default_rng(3)
    u = np.asfortranarray(rng.standard_normal((6, 4)))
    ret = np.full((6, 4), 99.0, order="F")
    sh, dm = C.c_int(shift), C.c_int(dim)
    GF._gfortran_eoshift0_4(C.byref(_descriptor(ret)), C.byref(_descriptor(u)), C.byref(sh), None, C.byref(dm))
    assert np.array_equal(ret, f90py._eoshift(f90py.FArr(u), shift, dim).a)
    if (shift, dim) == (-1, 2):
        assert np.array_equal(ret[:, 1:], u[:, :-1]) and np.all(ret[:, 0] == 0.0)


def _int_descriptor(vals):
    """shape_type: a rank-1 array of index_type (BT_INTEGER = 1)"""
    a = np.array(vals, dtype=np.int64)

    class D(C.Structure):
        _fields_ = [("base_addr", C.c_void_p), ("offset", C.c_size_t), ("dtype", _DType), ("span", C.c_ssize_t), ("dim", _Dim * 1)]

    d = D()
    d.base_addr, d.dtype, d.span, d.offset = a.ctypes.data, _DType(8, 0, 1, 1, 0), 8, (-1) % (1 << 64)
    d.dim[0] = _Dim(1, 1, a.size)
    d._keep = a
    return d


@needs_gf
@pytest.mark.parametrize("order", [None, (1, 2), (2, 1)])
def test_reshape_of_the_coefficient_tables_equals_libgfortran(tmp_path, order):
    """weno.f90:16-21 builds c1 / c2 / c3 with `reshape([...], [k, k+1], order=[1, 2])`: element order of the source, the
    column-major fill and the ORDER= permutation as the translator reads a module-level parameter, against reshape_r8"""
    vals = [float(v) for v in np.random.default_rng(4).standard_normal(12)]
    lits = ", ".join(f"{v!r}_rk" for v in vals)
    src = tmp_path / "m.f90"
    src.write_text(f"""
        module m
           real(rk), parameter :: c(0:2, -1:2) = reshape([{lits}], [3, 4]{'' if order is None else f', order=[{order[0]}, {order[1]}]'})
        end module
    """)
    ns = f90py.Program().add_source(str(src)).build()
    got = ns["c"]
    assert got.lb == (0, -1) and got.a.shape == (3, 4)
    ret = np.zeros((3, 4), order="F")
    source = np.array(vals)
    GF._gfortran_reshape_r8(C.byref(_descriptor(ret)), C.byref(_descriptor(source)), C.byref(_int_descriptor([3, 4])), None,
                            C.byref(_int_descriptor(order)) if order else None)
    assert np.array_equal(got.a, ret)
    if order != (2, 1):
        assert np.array_equal(got.a, np.reshape(source, (3, 4), order="F"))  # c(j, r) at source[j + 3*(r+1)]
        # the run-time intrinsic (test_hrweno.f90:105-106 flattens cnu(:, :, i) with it)
        flat = f90py.INTRINSICS["reshape"](got, f90py.FArr.from_list([12]))
        assert np.array_equal(flat.a, source)
