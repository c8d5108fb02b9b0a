// Compile-and-link check of the C++ host mirror (hr-weno_b200/host/hrweno.hpp): every wrapper of the C ABI is instantiated,
// so a signature that drifts from include/hrweno_b200.h or a symbol missing from libhrweno_b200.so fails the build.
// Nothing is executed unless a CUDA device exists AND the program is started with an argument (tests/ start it without).
#include <cstdio>
#include <vector>

#include "hrweno.hpp"

using namespace hrweno;

static int exercise() {
   // fp64: grid, weno, fluxes, fused operator, integrators (host, device-resident, attached), one-process multi-GPU
   hrweno_grids::grid1 gx;
   gx.linear(-5.0, 5.0, 64);
   hrweno_weno::weno w(64, 3, 1e-6);
   std::vector<double> v(64, 1.0), vl(64), vr(64);
   w.reconstruct(v.data(), vl.data(), vr.data());
   hrweno_fluxes::flux burgers = [](double u, const std::vector<double> &, double) { return 0.5 * u * u; };
   double f = hrweno_fluxes::godunov(burgers, 1.0, 0.5, {0.0}, 0.0) + hrweno_fluxes::lax_friedrichs(burgers, 1.0, 0.5, {0.0}, 0.0, 1.0);
   hrweno_fv_desc d = hrweno_fv::fv::desc1d(64, 3, 1e-6, gx.width.data());
   hrweno_fv::fv op(d);
   op.set_flux_time_fn([](double t) { return 1.0 + t; });
   op.set_flux_time_fn(nullptr);
   op.set_alpha(1.0);
   op.rhs(0.0, v.data(), vl.data());
   op.rhs_dev(0.0, nullptr, nullptr);
   op.max_wavespeed_dev(nullptr, nullptr);
   hrweno_tvdode::rktvd ode(op, op.neq(), 3);
   double t = 0.0;
   ode.integrate(v.data(), t, 0.1, 1e-2);
   ode.integrate_dev(nullptr, t, 0.2, 1e-2);
   ode.attach(nullptr);
   ode.integrate_attached(t, 0.3, 1e-2);
   ode.fetch(nullptr);
   hrweno_tvdode::mstvd ms(op, op.neq());
   ms.integrate(v.data(), t, 0.4, 1e-2);
   hrweno_multi::mgpu m(d, 0);
   hrweno_multi::mgpu::slab_info s = m.slab(0);
   m.set_xedges(0, gx.edges.data());
   m.set_flux_coef(0, gx.edges.data());
   m.set_flux_time_fn([](double) { return 1.0; });
   m.rktvd(3);
   m.mstvd();
   m.integrate(v.data(), t, 0.5, 1e-2);
   m.upload(v.data());
   m.integrate_resident(t, 0.6, 1e-2);
   m.download(v.data());
   f += m.max_wavespeed(true) + (double)(m.fevals() + m.launches() + m.ngpus() + s.count + ode.launches() + ode.fevals() + ode.istate());
   m.set_alpha(1.0);
   // rk = real32
   std::vector<float> wf(64, 10.0f / 64), vf(64, 1.0f), lf(64), rf(64);
   real32::weno w32(64, 3, 1e-6f);
   w32.reconstruct(vf.data(), lf.data(), rf.data());
   f += (double)w32.cnu().size();
   hrweno_fv_desc_f32 d32 = real32::fv::desc1d(64, 3, 1e-6f, wf.data());
   real32::fv op32(d32);
   op32.set_xedges(0, wf.data());
   op32.set_flux_coef(0, wf.data());
   op32.set_flux_time_fn([](float tt) { return 1.0f + tt; });
   op32.rhs(0.0f, vf.data(), lf.data());
   op32.rhs_dev(0.0f, nullptr, nullptr);
   real32::rktvd ode32(op32, op32.neq(), 3);
   float t32 = 0.0f;
   ode32.integrate(vf.data(), t32, 0.1f, 1e-2f);
   ode32.integrate_dev(nullptr, t32, 0.2f, 1e-2f);
   real32::mstvd ms32(op32, op32.neq());
   ms32.integrate(vf.data(), t32, 0.3f, 1e-2f);
   hrweno_fv_desc_f32 d2 = real32::fv::desc2d(8, 8, 3, 1e-6f, wf.data(), wf.data());
   f += (double)(ode32.fevals() + ode32.istate() + ode32.launches() + d2.ndim);
   std::printf("%g\n", f);
   return 0;
}

int main(int argc, char **) {
   std::printf("abi %d, devices %d\n", hrweno_abi_version(), hrweno_device_count());
   if (argc > 1 && hrweno_device_count() > 0) return exercise();
   return 0;
}
