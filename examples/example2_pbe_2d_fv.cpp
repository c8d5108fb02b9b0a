// example2_pbe_2d_fv.cpp -- the reference's example/example2_pbe_2d_fv.f90 on the B200 path (driver: example2:25-69).
#include <chrono>
#include <cstdio>

#include "hrweno.hpp"

using namespace hrweno; // the module namespaces hrweno_grids, hrweno_weno, hrweno_fv, hrweno_tvdode

int main(int argc, char **argv) {
   const char *folder = argc > 1 ? argv[1] : ".";
   const int64_t nc[2] = {250, 250};
   const int k = 3;
   hrweno_grids::grid1 gx[2];
   gx[0].linear(0.0, 10.0, nc[0]);                                       // example2:38-39
   gx[1].linear(0.0, 10.0, nc[1]);
   hrweno::hrweno_fv::fv rhs(hrweno::hrweno_fv::fv::desc2d(nc[0], nc[1], k, 1e-6, gx[0].width.data(), gx[1].width.data()));
   std::vector<double> u((size_t)(nc[0] * nc[1]));
   for (int64_t jj = 0; jj < nc[1]; ++jj)                                 // example2:49-51, 157-168
      for (int64_t ii = 0; ii < nc[0]; ++ii) {
         const double x1 = gx[0].center[(size_t)ii], x2 = gx[1].center[(size_t)jj];
         u[(size_t)(jj * nc[0] + ii)] = (x1 >= 1.0 && x1 <= 3.0 && x2 >= 1.0 && x2 <= 3.0) ? 1.0 : 0.0;
      }
   hrweno_tvdode::mstvd ode(rhs, (int64_t)u.size());                     // example2:54
   const double time_end = 5.0, dt = 5e-3;
   double time = 0.0;
   const int num_time_points = 100;
   std::printf(" Running example2 ...\n");
   const auto t0 = std::chrono::steady_clock::now();
   double mass = 0.0;
   for (int ii = 0; ii <= num_time_points; ++ii) {
      const double time_out = time_end * ii / num_time_points;           // example2:62
      ode.integrate(u.data(), time, time_out, dt);                       // example2:63
   }
   for (int64_t jj = 0; jj < nc[1]; ++jj)
      for (int64_t ii = 0; ii < nc[0]; ++ii) mass += u[(size_t)(jj * nc[0] + ii)] * gx[0].width[(size_t)ii] * gx[1].width[(size_t)jj];
   const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
   std::printf(" Elapsed time (s) : %5.2f\n", s);
   std::printf(" fevals = %lld  t_end = %.17g  mass = %.17g\n", (long long)ode.fevals(), time, mass);
   FILE *fb = std::fopen((std::string(folder) + "/u_final.bin").c_str(), "wb");
   std::fwrite(u.data(), sizeof(double), u.size(), fb);
   std::fclose(fb);
   return 0;
}
