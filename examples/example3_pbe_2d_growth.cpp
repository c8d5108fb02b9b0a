// example3_pbe_2d_growth.cpp -- the reference's example/example2_pbe_2d_fv.f90 with what its comments point at:
// geometric grids (grids.f90:182-230), `weno(nc, k, eps, xedges)` per axis (weno.f90:100-112) and the growth fluxes
// `flux1 = v*x(1)**2`, `flux2 = v*x(1)*x(2)` (example2:140,153), on the fused B200 path (general stage kernel).
// Same driver loop as example2:25-69; writes the grids and the final state for the parity test.
#include <cstdio>

#include "hrweno.hpp"

using namespace hrweno;

static void dump(const std::string &path, const std::vector<double> &v) {
   FILE *f = std::fopen(path.c_str(), "wb");
   std::fwrite(v.data(), sizeof(double), v.size(), f);
   std::fclose(f);
}

int main(int argc, char **argv) {
   const std::string folder = argc > 1 ? argv[1] : ".";
   const int64_t nc[2] = {80, 60};
   const int k = 3;
   hrweno_grids::grid1 gx[2];
   gx[0].geometric(0.0, 10.0, 1.02, nc[0]);
   gx[1].geometric(0.0, 10.0, 1.03, nc[1]);
   hrweno::hrweno_fv::fv rhs(hrweno::hrweno_fv::fv::desc2d(nc[0], nc[1], k, 1e-6, gx[0].width.data(), gx[1].width.data()));
   rhs.set_xedges(0, gx[0].edges.data()); // was: myweno(1) = weno(nc(1), k, eps, xedges=gx(1)%edges)
   rhs.set_xedges(1, gx[1].edges.data());
   std::vector<double> face1(gx[0].edges.size());
   for (size_t i = 0; i < face1.size(); ++i) face1[i] = gx[0].edges[i] * gx[0].edges[i]; // x(1)**2, x(1) = gx(1)%right(i)
   rhs.set_flux_coef(0, face1.data());                                                   // flux1 = v*x(1)**2
   rhs.set_flux_coef(1, gx[1].edges.data(), gx[0].center.data());                        // flux2 = v*x(1)*x(2)
   std::vector<double> u((size_t)(nc[0] * nc[1]));
   for (int64_t jj = 0; jj < nc[1]; ++jj) // example2:49-51, 157-168
      for (int64_t ii = 0; ii < nc[0]; ++ii) {
         const double x1 = gx[0].center[(size_t)ii], x2 = gx[1].center[(size_t)jj];
         u[(size_t)(jj * nc[0] + ii)] = (x1 >= 1.0 && x1 <= 3.0 && x2 >= 1.0 && x2 <= 3.0) ? 1.0 : 0.0;
      }
   hrweno_tvdode::mstvd ode(rhs, (int64_t)u.size()); // example2:54
   const double time_end = 0.1, dt = 2.5e-4; // max speed x1*x2 = 100 at the far corner: CFL 0.1 on the 0.25-wide cells there
   double time = 0.0;
   const int num_time_points = 20;
   for (int ii = 0; ii <= num_time_points; ++ii) {
      const double time_out = time_end * ii / num_time_points; // example2:62
      ode.integrate(u.data(), time, time_out, dt);             // example2:63
   }
   double mass = 0.0;
   for (int64_t jj = 0; jj < nc[1]; ++jj)
      for (int64_t ii = 0; ii < nc[0]; ++ii) mass += u[(size_t)(jj * nc[0] + ii)] * gx[0].width[(size_t)ii] * gx[1].width[(size_t)jj];
   std::printf(" fevals = %lld  t_end = %.17g  mass = %.17g\n", (long long)ode.fevals(), time, mass);
   dump(folder + "/edges1.bin", gx[0].edges);
   dump(folder + "/edges2.bin", gx[1].edges);
   dump(folder + "/u_final.bin", u);
   return 0;
}
