"""State kept inside the integrator between output times (hrweno_ode_attach / _integrate_attached / _fetch): the same
steps as integrate_dev on the caller's dense vector, bit for bit, without the dense <-> padded copies of every call."""
import numpy as np
import pytest

from conftest import ex1_ic, ex2_ic

pytestmark = pytest.mark.gpu


def _steps_to(t, dt, k):
    for _ in range(k - 1):
        t = t + dt
    return t


@pytest.mark.parametrize("case", ["rk1_rows", "rk3_1d", "rk3_small", "ms_1d", "ms_2d"])
@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_attached_state_equals_integrate_dev(gpu_lib, pkg, case, mode):
    import torch

    m = pkg._abi.MODE_FAST if mode == "fast" else pkg._abi.MODE_STRICT
    stream = torch.cuda.current_stream().cuda_stream
    if case == "ms_2d":
        n1, n2 = 200, 150
        g1, g2 = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n1), pkg.hrweno_grids.grid1().linear(0.0, 10.0, n2)
        u0 = (ex2_ic(g1.center, g2.center) + 1e-3 * np.random.default_rng(1).standard_normal((n2, n1))).reshape(-1)
        mk = lambda: pkg.hrweno_tvdode.mstvd(pkg.fv.FV(pkg.fv.make_desc((n1, n2), flux_model=1, bc=1, width=[g1.width, g2.width], mode=m)), n1 * n2)
        dt = 5e-3
    else:
        rows = 37 if case == "rk1_rows" else 1
        nc = 100 if case == "rk3_small" else 4096
        g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
        u0 = (ex1_ic(g.center)[None, :] * np.linspace(0.5, 1.5, rows)[:, None]).reshape(-1)
        fvk = lambda: pkg.fv.FV(pkg.fv.make_desc(nc, k=3, rows=rows, linear=(-5.0, 5.0), mode=m))
        if case.startswith("ms"):
            mk = lambda: pkg.hrweno_tvdode.mstvd(fvk(), nc * rows)
        else:
            mk = lambda: pkg.hrweno_tvdode.rktvd(fvk(), nc * rows, 1 if case.startswith("rk1") else 3)
        dt = 0.1 * 10.0 / nc
    a, b = mk(), mk()
    ua, ub = torch.from_numpy(u0).cuda(), torch.from_numpy(u0).cuda()
    ta = tb = 0.0
    b.attach(ub.data_ptr(), stream)
    for nsteps in (1, 6, 9):
        ta = a.integrate_dev(ua.data_ptr(), ta, _steps_to(ta, dt, nsteps), dt, 1, stream)
        tb = b.integrate_attached(tb, _steps_to(tb, dt, nsteps), dt, 1, stream)
        out = torch.empty_like(ub)
        b.fetch(out.data_ptr(), stream)
        torch.cuda.synchronize()
        assert ta == tb and torch.equal(ua, out), f"{case} {mode} after {nsteps} steps"
    assert a.fevals == b.fevals
    # the attached path launches fewer kernels (no pack / unpack per call)
    if case != "rk3_small":
        assert b.launches < a.launches


def test_integrate_attached_needs_attach(gpu_lib, pkg):
    ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(4096, k=3, linear=(-5.0, 5.0))), 4096, 3)
    with pytest.raises(pkg.HrwenoError) as ei:
        ode.integrate_attached(0.0, 1.0, 1e-3)
    assert ei.value.status == pkg._abi.ESTATE
