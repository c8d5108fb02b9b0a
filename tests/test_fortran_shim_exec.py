"""fortran/hrweno_b200_shim.f90 EXECUTED (no Fortran compiler exists here): the shim's source runs statement by statement
through tools/f90exec, every `bind(c)` interface body marshalled from its own Fortran declarations into a real call of a
shared library with the product's C ABI -- on this machine tests/cpp/abi_on_oracle.c (the ABI on the CPU oracle), on the
GPU box libhrweno_b200.so (tests/test_zzzz_gpu_fortran_shim_exec.py).

What is held to what:
  * the reference's UNMODIFIED example programs (example1 as shipped; example2 with only its grid-size literal patched
    to 40 x 40) linked against the shim instead of hrweno_weno / hrweno_tvdode: bit-identical, at every output the
    fixtures hold, to the same programs executed over the reference's own modules (tests/golden/ref_exec_*.npz);
  * the reference's own test-drive suites for weno and tvdode: every assertion passes on the shim (incl. `ode%fevals`);
  * the three self-contained programs of fortran/examples/ (fused operators, a caller's own pure rhs): bit-identical to
    the fixtures of the problems they set up;
  * every interface body of the shim against include/hrweno_b200.h (through hrweno_b200._abi.PROTOTYPES, which
    tests/test_abi_exports.py holds to the header): same arguments, by value where C takes a value, a pointer where C
    takes a pointer; the bind(c) callback the shim hands to the library has the C callback type's signature;
  * the reference's `error stop` sites reached through the shim carry the reference's messages.
"""
import ctypes as C
import os

import numpy as np
import pytest

import shim_exec
from shim_exec import GOLDEN, OWN_PROGRAMS, OWN_PROGRAMS_REAL32, REFERENCE, f90py

import f90c  # noqa: E402  (tools/f90exec is on sys.path through shim_exec)

needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "example")), reason="the reference tree is not on this machine")


@pytest.fixture(scope="module")
def abi(ref, tmp_path_factory):
    """the product's C ABI on the CPU oracle (`ref` makes sure the oracle is built)"""
    return shim_exec.build_abi_on_oracle(tmp_path_factory.mktemp("abi_on_oracle"))


@pytest.fixture(scope="module")
def abi32(ref, tmp_path_factory):
    """the REAL32 entry points (hrweno_*_f32) on the REAL32 build of the CPU oracle"""
    from oracle import ref32

    ref32.lib()  # builds oracle/libhrweno_oracle_f32.so if need be
    return shim_exec.build_abi_on_oracle(tmp_path_factory.mktemp("abi_f32_on_oracle"), real32=True)


# ---- the shim's declarations against the C header ----------------------------------------------------------------------


def _cls(ct):
    """class of a ctypes type as an argument-passing convention"""
    if ct is None:
        return "void"
    if ct in (C.c_int, C.c_int32):
        return "i32"
    if ct in (C.c_int64, C.c_longlong):
        return "i64"
    if ct is C.c_double:
        return "f64"
    if ct is C.c_float:
        return "f32"
    if ct in (C.c_void_p, C.c_char_p) or hasattr(ct, "contents") or issubclass(ct, (C._Pointer, C._CFuncPtr)):
        return "ptr"
    raise AssertionError(f"unclassified ctypes type {ct}")


def _shim(defines=None):
    P = f90py.Program(defines=defines)
    P.add_source(shim_exec.SHIM)
    return P


@pytest.mark.parametrize("build", ["REAL64", "REAL32"])
def test_every_interface_body_of_the_shim_matches_the_c_header(pkg, build):
    """both builds of the shim (the reference's kind is a compile-time switch, hrweno_kinds.F90:9-17): with -DREAL32 the same
    interface bodies bind the hrweno_*_f32 entry points with float in every real position"""
    P = _shim({build: "1"})
    P.build()
    protos = pkg._abi.PROTOTYPES
    assert len(P.cprotos) >= (30 if build == "REAL64" else 17)
    for name, proto in P.cprotos.items():
        cname = proto["cname"]
        if build == "REAL64":
            assert cname == name, f"{name}: bind(c, name=) differs from the Fortran name"
        elif name != "hrweno_last_error":
            assert "_f32_" in cname and cname.replace("_f32_", "_") == name, (name, cname)
        assert cname in protos, f"{cname} is not an entry point of include/hrweno_b200.h"
        res, args = protos[cname]
        sig = P.interop._arg_types(proto["args"], proto["decls"])
        got = [_cls(ct) if kind == "value" else "ptr" for _, kind, ct, _ in sig]
        assert got == [_cls(a) for a in args], f"{cname}: Fortran passes {got}, C expects {[_cls(a) for a in args]}"
        fres = f90c.ctype_of(proto["decls"][proto["res"]]["base"], P.ns) if proto["kind"] == "function" else None
        assert _cls(fres) == _cls(res), f"{cname}: result {fres} vs {res}"
        if build == "REAL32":  # no double left in a REAL32 signature
            assert "f64" not in got and _cls(fres) != "f64", cname


@pytest.mark.parametrize("build", ["REAL64", "REAL32"])
def test_interoperable_descriptor_type_has_the_layout_of_the_c_struct(pkg, build):
    P = _shim({build: "1"})
    P.build()
    S = P.interop.struct_type("hrweno_fv_desc")
    D = pkg._abi.FvDesc if build == "REAL64" else pkg._abi.FvDesc32
    assert C.sizeof(S) == C.sizeof(D)
    want = {n: (getattr(D, n).offset, getattr(D, n).size) for n, _ in D._fields_}
    got = {n: (getattr(S, n).offset, getattr(S, n).size) for n, _ in S._fields_}
    assert got == want
    # the defaults a Fortran program starts from are a valid single-GPU descriptor header
    s = P.interop.to_struct("hrweno_fv_desc", P.ns["new_hrweno_fv_desc"]())
    assert (s.abi_version, s.ndim, s.k, s.rank, s.nranks) == (pkg._abi.ABI_VERSION, 1, 3, 0, 1)
    assert s.eps == (1e-6 if build == "REAL64" else float(np.float32(1e-6)))


def test_the_shims_callback_has_the_signature_of_hrweno_rhs_host_fn(pkg):
    P = _shim()
    P.build()
    unit = P.procs["host_trampoline"]
    assert unit["bindc"]
    sig = P.interop._arg_types(unit["args"], unit["cdecls"])
    got = [_cls(ct) if kind == "value" else "ptr" for _, kind, ct, _ in sig]
    want = pkg._abi.RHS_HOST_FN
    assert got == [_cls(a) for a in want._argtypes_] and want._restype_ is None


# ---- self-contained programs (fortran/examples) ----------------------------------------------------------------------------


@pytest.mark.parametrize("name", sorted(OWN_PROGRAMS))
def test_own_fortran_program_on_the_shim_reproduces_the_executed_reference(abi, name):
    fixture, snaps, must_call = OWN_PROGRAMS[name]
    ns, P = shim_exec.run_own_program(abi, name)
    shim_exec.assert_history_equals_fixture(ns, fixture, snaps)
    assert must_call <= set(P.interop.calls)
    assert P.interop.calls[-1] in ("hrweno_fv_destroy", "hrweno_weno_destroy")  # handles released by the explicit destroy()


@pytest.mark.parametrize("name", sorted(OWN_PROGRAMS_REAL32))
def test_own_fortran_program_on_the_real32_build_of_the_shim(abi32, name):
    """shim and program as `-DREAL32` compiles them (rk = real32, src/hrweno_kinds.F90:9-10), executed in binary32: the
    hrweno_*_f32 entry points, against the reference's source executed in real32 for the same problems"""
    fixture, snaps, must_call = OWN_PROGRAMS_REAL32[name]
    ns, P = shim_exec.run_own_program(abi32, name, real32=True)
    shim_exec.assert_history_equals_fixture(ns, fixture, snaps)
    assert must_call <= set(P.interop.calls) and not any("f32" not in c for c in P.interop.calls if c != "hrweno_last_error")


def test_weno_type_of_the_real32_build_of_the_shim(abi32):
    from oracle import ref32

    shim_exec.check_weno_type_real32(abi32, ref32)


# ---- the reference's own programs, unmodified, on the shim ---------------------------------------------------------------------


class _Stop(Exception):
    pass


def _run_reference_program(abi, example, main, npts, snaps, patch=()):
    path = os.path.join(REFERENCE, "example", example)
    if patch:
        text = open(path).read()
        for old, new in patch:
            assert text.count(old) == 1
            text = text.replace(old, new)
        path = os.path.join("/tmp", "shim_exec_" + example)
        open(path, "w").write(text)
    P = shim_exec.program_on_shim(abi, [path], skip=("output", "timer"), reference_modules=("src/hrweno_fluxes.f90", "src/hrweno_grids.f90"))
    ns = P.build()
    rec = {"times": [], "n": 0}

    def output(action):  # the programs' formatted file output, replaced by a recorder (as in make_ref_exec_golden.py)
        if int(f90py.val(action)) != 2:
            return
        ii = rec["n"]
        rec["times"].append(ns["time"])
        if ii in snaps:
            rec[ii] = ns["u"].a.copy()
        rec["n"] += 1
        if ii >= npts:
            raise _Stop

    ns["output"] = output
    try:
        ns[main]()
    except _Stop:
        pass
    return ns, P, rec


@needs_reference
def test_reference_example1_unmodified_runs_on_the_shim_bit_identical(abi):
    """example/example1_burgers_1d_fv.f90 as shipped -- `pure subroutine rhs` calling `myweno%reconstruct`, `godunov` from the
    reference's own hrweno_fluxes, `rktvd(rhs, nc, order=3)`, `call ode%integrate(u, time, time_out, dt)` -- with the shim's
    hrweno_weno / hrweno_tvdode in place of the reference's: all 101 outputs"""
    g = np.load(os.path.join(GOLDEN, "ref_exec_example1.npz"))
    ns, P, rec = _run_reference_program(abi, "example1_burgers_1d_fv.f90", "main_example1_burgers_1d_fv", 100, (0, 1, 50, 100))
    for i in (0, 1, 50, 100):
        assert np.array_equal(rec[i], g[f"u_{i}"]), i
    assert np.array_equal(np.array(rec["times"]), g["times"])
    assert ns["ode"].fevals == int(g["fevals"]) == 3603 and ns["ode"].istate == 2  # public components (tvdode.f90:22-24; 2 after a first integrate, :176)
    calls = P.interop.calls
    assert calls.count("hrweno_weno_reconstruct_s") == 3603 and calls.count("hrweno_ode_integrate") == 101
    assert calls.count("hrweno_rktvd_create_host") == 1 and calls.count("hrweno_weno_create") == 1


@needs_reference
def test_reference_example1_with_a_time_dependent_flux_on_the_shim(abi):
    """the evaluation time crosses the boundary: the library hands t, t + dt, t + dt/2 (tvdode.f90:162-166) to the shim's
    bind(c) trampoline, which hands them to the program's `rhs`, whose flux now depends on t (the fixture's one-line patch)"""
    g = np.load(os.path.join(GOLDEN, "ref_exec_example1_tfactor.npz"))
    ns, P, rec = _run_reference_program(abi, "example1_burgers_1d_fv.f90", "main_example1_burgers_1d_fv", 20, (0, 10, 20),
                                        patch=[("flux = (v**2)/2", "flux = (v**2)/2*(1.0_rk + 0.25_rk*t)")])
    for i in (0, 10, 20):
        assert np.array_equal(rec[i], g[f"u_{i}_o3"]), i
    assert np.array_equal(np.array(rec["times"]), g["times_o3"])


@needs_reference
def test_reference_example2_unmodified_runs_on_the_shim_bit_identical(abi):
    """example/example2_pbe_2d_fv.f90 with only `nc(2) = [250, 250]` patched to 40 x 40 (the fixture's size): `mstvd(rhs,
    size(u))`, an array of two weno objects, row sections and STRIDED column sections `v(i:i+(nc(2)-1)*nc(1):nc(1))` as
    actual arguments of `reconstruct` (copy-in for the contiguous dummy), first 10 outputs"""
    g = np.load(os.path.join(GOLDEN, "ref_exec_example2_40.npz"))
    ns, P, rec = _run_reference_program(abi, "example2_pbe_2d_fv.f90", "main_example_pbe_2d_fv", 10, (0, 1, 10),
                                        patch=[("nc(2) = [250, 250]", "nc(2) = [40, 40]")])
    assert np.array_equal(rec[0], g["u_0"]) and np.array_equal(rec[1], g["u_1"])
    assert np.array_equal(np.array(rec["times"]), g["times"][:11])
    assert P.interop.calls.count("hrweno_mstvd_create_host") == 1 and P.interop.calls.count("hrweno_weno_create") == 2


# ---- the reference's own test suites on the shim -----------------------------------------------------------------------------

SUITES = {"test_hrweno.f90": ["test_weno_uniform", "test_calc_cnu", "test_weno_nonuniform"], "test_tvdode.f90": ["test_rktvd", "test_mstvd"]}


@needs_reference
@pytest.mark.parametrize("suite,test", [(s, t) for s, ts in SUITES.items() for t in ts])
def test_reference_suite_passes_on_the_shim(abi, suite, test):
    """test/test_hrweno.f90 and test/test_tvdode.f90, unmodified, over the shim's modules (test-drive's `check` supplied as in
    tests/test_reference_suites_exec.py): constructor with and without xedges, the public `cnu` component against c1/c2/c3,
    reconstruct, rktvd orders 1-3 and mstvd with a host integrand, and the public `fevals` component"""
    from test_reference_suites_exec import ErrorBox, check

    P = shim_exec.program_on_shim(abi, [os.path.join(REFERENCE, "test", suite)], skip=("collect_tests_hrweno", "collect_tests_tvdode"),
                                  reference_modules=("src/hrweno_grids.f90",), skip_io=True)
    P.ns["check"] = check
    P.ns["allocated"] = lambda x: x is not None and not (isinstance(x, ErrorBox) and x.msg is None)
    ns = P.build()
    err = ErrorBox()
    ns[test](err)
    assert err.msg is None, err.msg
    assert any(c.startswith(("hrweno_weno_create", "hrweno_rktvd_create_host", "hrweno_mstvd_create_host")) for c in P.interop.calls)


def test_multi_gpu_fortran_program_on_the_shim(abi, ref, pkg):
    """hrweno_mgpu_* driven from Fortran (on the CPU stand-in there is one domain; on the GPU box: every visible device)"""
    assert shim_exec.check_multi_gpu_program(abi, ref, pkg) == 1


def test_time_dependent_growth_fortran_program_on_the_shim(abi, ref, pkg):
    shim_exec.check_time_factor_program(abi, ref, pkg)


def test_device_integrand_fortran_program_on_the_shim(abi, ref, pkg):
    shim_exec.check_device_integrand_program(abi, ref, pkg)


def test_weno_type_of_the_shim(abi, ref):
    shim_exec.check_weno_type(abi, ref)


# ---- error stops --------------------------------------------------------------------------------------------------------------


def test_error_stops_carry_the_reference_messages(abi):
    P = shim_exec.program_on_shim(abi, [])
    ns = P.build()
    with pytest.raises(f90py.FortranStop, match="Invalid input 'ncells'. Valid range: ncells > 0."):  # weno.f90:75
        ns["weno"](0)
    with pytest.raises(f90py.FortranStop, match="Invalid input 'k'. Valid range: 1 <= k <= 3."):  # weno.f90:84
        ns["weno"](10, 4)
    with pytest.raises(f90py.FortranStop, match="size\\(xedges\\) /= ncells \\+ 1"):  # weno.f90:103
        ns["weno"](10, 3, 1e-6, f90py.FArr(np.linspace(0.0, 1.0, 10), (0,)))
    with pytest.raises(f90py.FortranStop, match="Invalid input 'order' in 'rktvd'"):  # tvdode.f90:89
        ns["rktvd"](lambda t, u, udot: None, 10, 4)


def test_an_error_stop_inside_the_users_integrand_propagates_through_the_c_call(abi):
    """the library calls the integrand back through the shim's bind(c) trampoline; a Fortran `error stop` in there ends the
    program -- here: surfaces as FortranStop after the C call returns, not swallowed by the foreign frame"""
    P = shim_exec.program_on_shim(abi, [])
    ns = P.build()

    def fu(t, u, udot):
        raise f90py.FortranStop("stop inside rhs")

    ode = ns["rktvd"](fu, 4, 1)
    with pytest.raises(f90py.FortranStop, match="stop inside rhs"):
        f90py.callm(ode, "integrate", f90py.FArr(np.ones(4)), f90py.Ref(0.0), 1.0, 0.1)


def test_the_shim_on_the_product_library_without_a_gpu_stops_with_the_librarys_message(pkg):
    """no CPU fallback anywhere: on a machine without a CUDA device the shim's constructors `error stop` with the library's text
    (validation of the arguments comes first, as in the reference: weno.f90:72-98)"""
    lib = C.CDLL(os.path.join(os.path.dirname(pkg.__file__), "lib", "libhrweno_b200.so"))  # own handle: argtypes come from the Fortran side
    lib.hrweno_device_count.restype = C.c_int
    if lib.hrweno_device_count() > 0:
        pytest.skip("a CUDA device is present")
    ns = shim_exec.program_on_shim(lib, []).build()
    with pytest.raises(f90py.FortranStop, match="Invalid input 'ncells'. Valid range: ncells > 0."):
        ns["weno"](0)
    with pytest.raises(f90py.FortranStop, match="no CUDA device \\(this library has no CPU fallback\\)"):
        ns["weno"](10)
    with pytest.raises(f90py.FortranStop, match="no CUDA device"):
        shim_exec.run_own_program(lib, "burgers_fused")


# ---- the marshalling layer itself ------------------------------------------------------------------------------------------------


def test_interop_rejects_what_a_compiler_would_reject(abi):
    P = shim_exec.program_on_shim(abi, [])
    ns = P.build()
    h = f90py.Ref(0)
    assert ns["hrweno_weno_create"](h, 8, 3, 1e-6, 0) == 0 and h.v != 0
    v = np.ones(8)
    with pytest.raises(f90c.CInteropError):  # a real(c_float) / integer array where real(c_double) v(*) is declared
        ns["hrweno_weno_reconstruct"](h.v, v.astype(np.float32), v, v)
    with pytest.raises(f90c.CInteropError):  # a real where integer(c_int), value is declared
        ns["hrweno_weno_create"](f90py.Ref(0), 8, 2.5, 1e-6, 0)
    with pytest.raises(f90c.CInteropError):  # intent(out) dummy with an actual that is not a variable
        ns["hrweno_weno_create"](0, 8, 3, 1e-6, 0)
    with pytest.raises(f90c.CInteropError):  # too few arguments
        ns["hrweno_weno_create"](h, 8, 3)
    # non-contiguous actual for an assumed-size dummy: copy-in / copy-out
    big, vl, vr = np.arange(16.0), np.zeros(16), np.zeros(16)
    st = f90py.Ref(-1)
    ns["hrweno_weno_reconstruct_s"](h.v, f90py.FArr(big[::2]), f90py.FArr(vl[::2]), f90py.FArr(vr[::2]), st)
    assert st.v == 0 and np.all(vl[1::2] == 0) and np.any(vl[::2] != 0)
    ns["hrweno_weno_destroy"](h.v)
