// example4_multi_gpu.cpp -- BASELINE.json configs[2] (1D Burgers, WENO5 + rktvd order 3, slabs over the GPUs of one box)
// from ONE host process through the C ABI alone: no launcher, no Python, no out-of-band handle exchange.
// The program has the shape of the reference's example1 (example/example1_burgers_1d_fv.f90:31-65): build the grid,
// set the initial condition at the cell centres, create the integrator, call integrate in an output loop.
//   usage: example4_multi_gpu <folder> [log2_cells = 20] [ngpus = all] [steps = 20] [mode: 0 strict | 1 fast] [lax_friedrichs: 0|1]
// Writes u_final.bin (raw fp64, global vector) for the parity test and prints cell-updates/s.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "hrweno.hpp"

using namespace hrweno;

static double ic(double x) { // example1:124-133
   const double xa = -4.0, xb = 2.0, va = 1.0, vb = -0.5;
   double v = va + (vb - va) / (xb - xa) * (x - xa);
   return std::max(std::min(v, va), vb);
}

int main(int argc, char **argv) {
   const char *folder = argc > 1 ? argv[1] : ".";
   const int log2n = argc > 2 ? std::atoi(argv[2]) : 20;
   const int ngpus = argc > 3 ? std::atoi(argv[3]) : 0;
   const int steps = argc > 4 ? std::atoi(argv[4]) : 20;
   const int mode = argc > 5 ? std::atoi(argv[5]) : HRWENO_MODE_STRICT;
   const int lf = argc > 6 ? std::atoi(argv[6]) : 0;
   const int64_t nc = (int64_t)1 << log2n;
   hrweno_grids::grid1 gx;
   gx.linear(-5.0, 5.0, nc);
   hrweno_fv_desc d = hrweno_fv::fv::desc1d(nc, 3, 1e-6, gx.width.data());
   d.mode = mode;
   if (lf) d.flux_scheme = HRWENO_SCHEME_LAX_FRIEDRICHS;
   hrweno_mgpu *m = nullptr;
   hrweno::check(hrweno_mgpu_create(&m, &d, ngpus, nullptr));
   const int n = hrweno_mgpu_ngpus(m);
   std::printf(" example4: %lld cells on %d GPU(s), %s mode, %s flux\n", (long long)nc, n, mode ? "fast" : "strict", lf ? "Lax-Friedrichs" : "Godunov");
   for (int r = 0; r < n; ++r) {
      int dev;
      int64_t off, cnt;
      hrweno::check(hrweno_mgpu_slab(m, r, &dev, &off, &cnt));
      std::printf("   slab %d: device %d, cells [%lld, %lld)\n", r, dev, (long long)off, (long long)(off + cnt));
   }
   std::vector<double> u((size_t)nc);
   unsigned long long s = 12345; // small deterministic perturbation (splitmix-like), keeps the Godunov branch mixed
   for (int64_t i = 0; i < nc; ++i) {
      s = s * 6364136223846793005ULL + 1442695040888963407ULL;
      u[(size_t)i] = ic(gx.center[(size_t)i]) + 1e-3 * ((double)(s >> 11) / 9007199254740992.0 - 0.5);
   }
   hrweno::check(hrweno_mgpu_rktvd(m, 3));
   const double dt = 0.1 * 10.0 / (double)nc;
   double t = 0.0;
   hrweno::check(hrweno_mgpu_upload(m, u.data()));
   if (lf) { // extension: alpha = max|f'(u)| over all slabs, reduced inside the library
      double alpha = 0.0;
      hrweno::check(hrweno_mgpu_max_wavespeed(m, &alpha, 1));
      std::printf(" alpha = %.17g\n", alpha);
   }
   const auto t0 = std::chrono::steady_clock::now();
   double tout = t;
   for (int i = 0; i < steps - 1; ++i) tout = tout + dt; // the strict is_done test then takes exactly `steps` steps
   hrweno::check(hrweno_mgpu_integrate_resident(m, &t, tout, dt, 1));
   const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
   hrweno::check(hrweno_mgpu_download(m, u.data()));
   std::printf(" %d steps in %.3f ms: %.3e cell-updates/s (device-resident state, wall clock incl. launch)\n", steps, sec * 1e3,
               (double)nc * 3.0 * steps / sec);
   std::printf(" fevals = %lld  t_end = %.17g  launches = %lld\n", (long long)hrweno_mgpu_fevals(m), t, (long long)hrweno_mgpu_launches(m));
   FILE *fb = std::fopen((std::string(folder) + "/u_final.bin").c_str(), "wb");
   if (!fb) { std::fprintf(stderr, "cannot write to %s\n", folder); return 2; }
   std::fwrite(u.data(), sizeof(double), u.size(), fb);
   std::fclose(fb);
   hrweno_mgpu_destroy(m);
   return 0;
}
