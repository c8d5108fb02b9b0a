"""Shared harness of the Fortran-shim execution tests (TEST INFRASTRUCTURE).

fortran/hrweno_b200_shim.f90 is the binding a Fortran host uses; no Fortran compiler exists in this image, so the tests
EXECUTE its source through tools/f90exec (f90py.py translates the statements, f90c.py turns every `bind(c)` interface body
into a real call of a shared library, marshalled from the Fortran declarations alone).  The library is
  * on a machine without a GPU: tests/cpp/abi_on_oracle.c -- the product's C ABI implemented on the CPU oracle, built here
    into a temporary directory (the tests are the only thing that ever loads it);
  * on the GPU box: hr-weno_b200/lib/libhrweno_b200.so itself.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools", "f90exec"))
import f90py  # noqa: E402

SHIM = os.path.join(ROOT, "fortran", "hrweno_b200_shim.f90")
PROGRAMS = os.path.join(ROOT, "fortran", "examples")
REFERENCE = os.environ.get("HRWENO_REFERENCE", "/root/reference")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def build_abi_on_oracle(outdir, real32=False):
    """gcc tests/cpp/abi_on_oracle.c against oracle/libhrweno_oracle.so -> <outdir>/libhrweno_abi_on_oracle.so
    (real32: tests/cpp/abi_f32_on_oracle.c against oracle/libhrweno_oracle_f32.so, the hrweno_*_f32 entry points)"""
    tag = "abi_f32_on_oracle" if real32 else "abi_on_oracle"
    so = os.path.join(str(outdir), f"libhrweno_{tag}.so")
    odir = os.path.join(ROOT, "oracle")
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", odir,
           os.path.join(ROOT, "tests", "cpp", tag + ".c"), "-o", so, "-L", odir,
           "-l:libhrweno_oracle_f32.so" if real32 else "-lhrweno_oracle", f"-Wl,-rpath,{odir}",
           "-Wl,-Bsymbolic"]  # the product library is loaded RTLD_GLOBAL in the same process and exports the same names: this
    #                           library's own calls between its entry points must bind to its own definitions
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert proc.returncode == 0, proc.stdout
    return ctypes.CDLL(so)


def program_on_shim(lib, sources, skip=(), reference_modules=(), skip_io=False, real32=False):
    """translate: [reference modules that stay on the host] + the shim + `sources`; bind(c) calls go to `lib`.
    real32: the shim preprocessed with -DREAL32 (translate AND run inside `with f90py.real_kind(4)`)"""
    P = f90py.Program(skip_io=skip_io, defines={"REAL32": "1"} if real32 else None)
    for f in reference_modules:
        P.add_source(os.path.join(REFERENCE, f))
    P.add_source(SHIM)
    for path in sources:
        P.add_source(path, skip=skip)
    P.clib = lib
    return P


def run_own_program(lib, name, real32=False):
    """one of fortran/examples/*.f90 (self-contained programs written for this repository); returns its variables.
    real32: program and shim as the reference's REAL32 build compiles them (rk = real32, -DREAL32)"""
    if real32:
        with f90py.real_kind(4):
            P = program_on_shim(lib, [os.path.join(PROGRAMS, name + ".f90")], real32=True)
            ns = P.build()
            ns["main_" + name]()
        return ns, P
    P = program_on_shim(lib, [os.path.join(PROGRAMS, name + ".f90")])
    ns = P.build()
    ns["main_" + name]()
    return ns, P


def assert_history_equals_fixture(ns, fixture, snaps):
    """`history(:, io)`, the output times and the fevals counter of a program against an executed-reference-source fixture
    (a fixture may hold fewer outputs than the program produces: the common ones are compared)"""
    g = np.load(os.path.join(GOLDEN, fixture))
    h = ns["history"].a
    assert h.dtype == g["u_0"].dtype
    for i in snaps:
        assert np.array_equal(h[:, i], g[f"u_{i}"]), f"{fixture}: output {i} differs (max {np.max(np.abs(h[:, i] - g[f'u_{i}'])):.3e})"
    nt = len(g["times"])
    assert np.array_equal(ns["tgrid"].a[:nt], g["times"]), "output times differ"
    if nt == ns["tgrid"].a.size:
        assert int(ns["nfev"]) == int(g["fevals"])


# what each self-contained program reproduces: (fixture, outputs held by the fixture, C entry points it must have gone through)
OWN_PROGRAMS = {
    "burgers_fused": ("ref_exec_example1.npz", (0, 1, 50, 100), {"hrweno_fv_create", "hrweno_rktvd_create_fused", "hrweno_ode_integrate"}),
    "burgers_host_rhs": ("ref_exec_example1_lf.npz", (0, 50, 100),
                         {"hrweno_weno_create", "hrweno_rktvd_create_host", "hrweno_ode_integrate", "hrweno_weno_reconstruct_s"}),
    "pbe2d_growth_fused": ("ref_exec_example2_growth.npz", (0, 10, 20),
                           {"hrweno_fv_create", "hrweno_fv_set_xedges", "hrweno_fv_set_flux_coef", "hrweno_mstvd_create_fused"}),
    "pbe2d_fused": ("ref_exec_example2_40.npz", (0, 1, 50, 100), {"hrweno_fv_create", "hrweno_mstvd_create_fused", "hrweno_ode_integrate"}),
}


# the same programs compiled as the REAL32 build (fused constructors only): fixtures of the reference's source executed in real32
OWN_PROGRAMS_REAL32 = {
    "burgers_fused": ("ref_exec_f32_example1.npz", (0, 1, 10, 50, 100), {"hrweno_fv_f32_create", "hrweno_rktvd_f32_create_fused", "hrweno_ode_f32_integrate"}),
    "pbe2d_fused": ("ref_exec_f32_example2_40.npz", (0, 1, 5, 10), {"hrweno_fv_f32_create", "hrweno_mstvd_f32_create_fused", "hrweno_ode_f32_integrate"}),
    "pbe2d_growth_fused": ("ref_exec_f32_example2_growth.npz", (0, 10, 20),
                           {"hrweno_fv_f32_create", "hrweno_fv_f32_set_xedges", "hrweno_fv_f32_set_flux_coef", "hrweno_mstvd_f32_create_fused"}),
}


def check_weno_type_real32(lib, ref32):
    """type(weno) of the REAL32 build of the shim against the REAL32 oracle"""
    F = np.float32
    with f90py.real_kind(4):
        P = program_on_shim(lib, [], real32=True)
        ns = P.build()
        nc, k = 37, 3
        rng = np.random.default_rng(7)
        xe = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 1.5, nc))]).astype(F)
        w = ns["weno"](nc, k, F(1e-6), f90py.FArr(xe, (0,)))
        assert w.cnu.a.dtype == F and w.cnu.a.shape == (k, k + 1, nc)
        cnu = ref32.calc_cnu(xe, k)
        assert np.array_equal(np.transpose(w.cnu.a, (2, 1, 0)), cnu)
        big = rng.standard_normal(3 * nc).astype(F)
        for v in (np.ascontiguousarray(big[:nc]), big[::3]):
            vl, vr = np.zeros(nc, dtype=F), np.zeros(nc, dtype=F)
            f90py.callm(w, "reconstruct", f90py.FArr(v), f90py.FArr(vl), f90py.FArr(vr))
            rl, rr = ref32.reconstruct(np.ascontiguousarray(v), k, 1e-6, cnu=cnu)
            assert np.array_equal(vl, rl) and np.array_equal(vr, rr)
        f90py.callm(w, "destroy")
        assert "rktvd_init" not in ns and "host_trampoline" not in ns  # the host-integrand constructors are REAL64 only
    assert {"hrweno_weno_f32_create", "hrweno_weno_f32_get_cnu", "hrweno_weno_f32_reconstruct_s", "hrweno_weno_f32_destroy"} <= set(P.interop.calls)


def check_multi_gpu_program(lib, ref, pkg):
    """fortran/examples/burgers_multi_gpu.f90 (hrweno_mgpu_* from Fortran, resident state) against the oracle on the program's
    OWN widths and initial state; returns the number of GPUs the program found"""
    ns, P = run_own_program(lib, "burgers_multi_gpu")
    n = 40000
    dx, u0 = ns["dx"].a.copy(), ns["q0"].a.copy()
    dt = 0.2 * 10.0 / n
    rode = ref.rktvd(ref.FV(pkg.fv.make_desc(n, k=3, width=[dx])), 3)
    u, t = u0.copy(), 0.0
    for io, tout in enumerate((0.0, 10 * dt, 25 * dt)):
        t = rode.integrate(u, t, tout, dt)
        assert ns["tgrid"].a[io] == t
        assert np.array_equal(ns["history"].a[:, io], u), f"output {io}: max diff {np.max(np.abs(ns['history'].a[:, io] - u)):.3e}"
    assert int(ns["nfev"]) == rode.fevals
    assert {"hrweno_mgpu_create", "hrweno_mgpu_rktvd", "hrweno_mgpu_upload", "hrweno_mgpu_integrate_resident", "hrweno_mgpu_download",
            "hrweno_mgpu_destroy"} <= set(P.interop.calls)
    return int(ns["ngpu"])


def check_time_factor_program(lib, ref, pkg):
    """fortran/examples/pbe2d_growth_time_factor.f90 (x- and t-dependent fluxes, g(t) a bind(c) Fortran function the library
    calls back) against the oracle on the program's OWN grids and initial state with the same g in Python"""
    ns, P = run_own_program(lib, "pbe2d_growth_time_factor")
    n1, n2, dt = 130, 97, 1e-4
    e1, e2, c1 = ns["e1"].a.copy(), ns["e2"].a.copy(), ns["c1x"].a.copy()
    rfv = ref.FV(pkg.fv.make_desc((n1, n2), flux_model=1, bc=1, width=[ns["dx1"].a.copy(), ns["dx2"].a.copy()]))
    rfv.set_xedges(0, e1)
    rfv.set_flux_coef(0, ns["g1face"].a.copy(), None)
    rfv.set_flux_coef(1, e2, c1)
    rfv.set_flux_time_fn(lambda t: 1.0 + 0.25 * t)
    rode = ref.rktvd(rfv, 3)
    u, t = ns["q0"].a.copy(), 0.0
    assert u.sum() > 0
    for io, tout in enumerate((0.0, 4.5 * dt, 10.5 * dt)):
        t = rode.integrate(u, t, tout, dt)
        assert ns["tgrid"].a[io] == t
        assert np.array_equal(ns["history"].a[:, io], u), f"output {io}: max diff {np.max(np.abs(ns['history'].a[:, io] - u)):.3e}"
    assert int(ns["nfev"]) == rode.fevals == 33
    assert {"hrweno_fv_set_xedges", "hrweno_fv_set_flux_coef", "hrweno_fv_set_flux_time_fn", "hrweno_rktvd_create_fused"} <= set(P.interop.calls)
    # and the time factor matters: the same run without it differs
    rfv.set_flux_time_fn(None)
    u2 = ns["q0"].a.copy()
    ref.rktvd(rfv, 3).integrate(u2, 0.0, 10.5 * dt, dt)
    assert not np.array_equal(u2, u)


def check_device_integrand_program(lib, ref, pkg):
    """fortran/examples/burgers_device_rhs.f90: rktvd_dev / mstvd_dev with a bind(c) integrand on device pointers that applies
    hrweno_fv_rhs_dev; against the oracle's fused integrators on the program's own widths and initial state"""
    ns, P = run_own_program(lib, "burgers_device_rhs")
    n, dt = 200, 5e-3
    dx, u0 = ns["dx"].a.copy(), ns["q0"].a.copy()
    for which, make in enumerate((lambda fv: ref.rktvd(fv, 3), lambda fv: ref.mstvd(fv))):
        rode = make(ref.FV(pkg.fv.make_desc(n, k=3, width=[dx])))
        u, t = u0.copy(), 0.0
        for io, tout in enumerate((0.0, 0.1, 0.3)):
            t = rode.integrate(u, t, tout, dt)
            assert ns["tgrid"].a[io, which] == t
            assert np.array_equal(ns["history"].a[:, io, which], u), (which, io)
        assert int(ns["nfev"].a[which]) == rode.fevals
    assert list(ns["asked"].a) == [0.0, 0.005, 0.0025]  # stage times t, t + dt, t + dt/2 (tvdode.f90:162-166)
    assert {"hrweno_rktvd_create", "hrweno_mstvd_create", "hrweno_fv_rhs_dev"} <= set(P.interop.calls)


def check_weno_type(lib, ref):
    """type(weno) of the shim against the oracle (shared by the CPU and the GPU test)"""
    P = program_on_shim(lib, [])
    ns = P.build()
    nc, k = 37, 3
    rng = np.random.default_rng(7)
    xe = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 1.5, nc))])
    w = ns["weno"](nc, k, 1e-6, f90py.FArr(xe, (0,)))
    assert w.ierr == 0 and w.ncells == nc and w.k == k and w.eps == 1e-6
    assert w.cnu.a.shape == (k, k + 1, nc) and w.cnu.lb == (0, -1, 1)  # cnu(0:k-1, -1:k-1, 1:ncells), weno.f90:41
    cnu = ref.calc_cnu(xe, k)  # [i-1, r+1, j]
    assert np.array_equal(np.transpose(w.cnu.a, (2, 1, 0)), cnu)
    big = rng.standard_normal(3 * nc)
    for v in (np.ascontiguousarray(big[:nc]), big[::3]):
        vl, vr = np.zeros(nc), np.zeros(nc)
        f90py.callm(w, "reconstruct", f90py.FArr(v), f90py.FArr(vl), f90py.FArr(vr))
        rl, rr = ref.reconstruct(np.ascontiguousarray(v), k, 1e-6, cnu=cnu)
        assert np.array_equal(vl, rl) and np.array_equal(vr, rr)
    f90py.callm(w, "destroy")
    assert w.handle == 0
    # uniform tables (no xedges): cnu stays unallocated, weno.f90:110-112
    w = ns["weno"](nc)
    assert w.cnu is None and w.k == 3
    vl, vr = np.zeros(nc), np.zeros(nc)
    f90py.callm(w, "reconstruct", f90py.FArr(big[:nc].copy()), f90py.FArr(vl), f90py.FArr(vr))
    rl, rr = ref.reconstruct(big[:nc].copy(), 3, 1e-6)
    assert np.array_equal(vl, rl) and np.array_equal(vr, rr)
    f90py.callm(w, "destroy")
