// example1_burgers_1d_fv.cpp -- the reference's example/example1_burgers_1d_fv.f90 on the B200 path.
// Same set-up, same driver loop (example1:31-65); the rhs (example1:72-109) runs as fused CUDA stages.
// Writes x.txt / u.txt in the reference's formats (example1:156-177) and u_final.bin (raw fp64) for parity checks.
#include <algorithm>
#include <chrono>
#include <cstdio>

#include "hrweno.hpp"

using namespace hrweno; // the module namespaces hrweno_grids, hrweno_weno, hrweno_fv, hrweno_tvdode

static double ic(double x) { // example1:124-133
   const double xa = -4.0, xb = 2.0, va = 1.0, vb = -0.5;
   double v = va + (vb - va) / (xb - xa) * (x - xa);
   return std::max(std::min(v, va), vb);
}

int main(int argc, char **argv) {
   const char *folder = argc > 1 ? argv[1] : ".";
   const int nc = 100;
   hrweno_grids::grid1 gx;
   gx.linear(-5.0, 5.0, nc);                                             // example1:41
   hrweno::hrweno_weno::weno myweno = hrweno::hrweno_weno::weno(nc, 3, 1e-6);            // example1:44 (kept: shows the constructor)
   hrweno::hrweno_fv::fv rhs(hrweno::hrweno_fv::fv::desc1d(nc, myweno.k, myweno.eps, gx.width.data()));
   std::vector<double> u((size_t)nc);
   for (int i = 0; i < nc; ++i) u[(size_t)i] = ic(gx.center[(size_t)i]);  // example1:50
   hrweno_tvdode::rktvd ode(rhs, nc, 3);                                 // example1:53
   const double time_start = 0.0, time_end = 12.0, dt = 1e-2;
   double time = time_start;
   const int num_time_points = 100;

   std::printf(" Running example1 ...\n");
   const auto t0 = std::chrono::steady_clock::now();
   std::string fx = std::string(folder) + "/x.txt", fu = std::string(folder) + "/u.txt";
   FILE *funit_x = std::fopen(fx.c_str(), "w"), *funit_u = std::fopen(fu.c_str(), "w");
   if (!funit_x || !funit_u) { std::fprintf(stderr, "cannot open output files in %s\n", folder); return 2; }
   std::fprintf(funit_x, "%5s %15s %15s\n", "i", "x(i)", "dx(i)");
   for (int i = 0; i < nc; ++i) std::fprintf(funit_x, "%5d %15.5E %15.5E\n", i + 1, gx.center[(size_t)i], gx.width[(size_t)i]);
   std::fprintf(funit_u, "%16s", "t");
   for (int i = 0; i < nc; ++i) { char h[32]; std::snprintf(h, sizeof h, "u(%d)", i + 1); std::fprintf(funit_u, " %16s", h); }
   std::fprintf(funit_u, "\n");
   for (int ii = 0; ii <= num_time_points; ++ii) {
      const double time_out = time_end * ii / num_time_points;           // example1:62
      ode.integrate(u.data(), time, time_out, dt);                       // example1:63
      std::fprintf(funit_u, "%16.5E", time);
      for (int i = 0; i < nc; ++i) std::fprintf(funit_u, " %16.5E", u[(size_t)i]);
      std::fprintf(funit_u, "\n");
   }
   std::fclose(funit_x);
   std::fclose(funit_u);
   const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
   std::printf(" Elapsed time (ms) : %6.1f\n", ms);
   std::printf(" fevals = %lld  t_end = %.17g\n", (long long)ode.fevals(), time);
   FILE *fb = std::fopen((std::string(folder) + "/u_final.bin").c_str(), "wb");
   std::fwrite(u.data(), sizeof(double), u.size(), fb);
   std::fclose(fb);
   return 0;
}
