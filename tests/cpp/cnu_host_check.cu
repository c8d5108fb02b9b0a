// cnu_host_check.cu -- TEST HARNESS (no GPU needed).  The product computes the per-cell reconstruction tables cnu of
// weno(ncells, k, eps, xedges) (weno.f90:221-297) on the HOST, in both kinds (weno.cu: weno_calc_cnu_host for real64,
// real32.cu: calc_cnu_t<float> for the REAL32 build), and uploads them.  This program calls those very functions (real32.cu
// is included for its file-local template, everything else is linked from the library's objects) on edges read from stdin and
// writes the tables to stdout; tests/test_product_host_tables.py compares them with the oracle bit for bit.
//   stdin : int64 nc, int32 k, int32 kind (4 | 8), then nc+1 edges of that kind        stdout : nc*k*(k+1) values
#include <cstdio>

#include "real32.cu"

int main() {
   long long nc;
   int k, kind;
   if (fread(&nc, sizeof nc, 1, stdin) != 1 || fread(&k, sizeof k, 1, stdin) != 1 || fread(&kind, sizeof kind, 1, stdin) != 1) return 1;
   if (kind == 4) {
      std::vector<float> xe((size_t)nc + 1), cnu;
      if (fread(xe.data(), sizeof(float), xe.size(), stdin) != xe.size()) return 2;
      hrw::calc_cnu_t<float>(nc, k, xe.data(), cnu);
      fwrite(cnu.data(), sizeof(float), cnu.size(), stdout);
   } else {
      std::vector<double> xe((size_t)nc + 1), cnu;
      if (fread(xe.data(), sizeof(double), xe.size(), stdin) != xe.size()) return 2;
      hrw::weno_calc_cnu_host(nc, k, xe.data(), cnu);
      fwrite(cnu.data(), sizeof(double), cnu.size(), stdout);
   }
   return 0;
}
