// grid_dump.cpp -- prints the edges of every grid1 kind of the C++ host mirror (hrweno.hpp) as hex floats, one grid per
// line, for the CPU test that holds the C++ and Python mirrors of src/hrweno_grids.f90 to the same bits.
// Settings: test/test_grid.f90 (linear :31-60, log :62-92, geometric :94-126, bilinear :128-160).
#include <cstdio>

#include "hrweno.hpp"

static void dump(const char *name, const hrweno::hrweno_grids::grid1 &g) {
   std::printf("%s %lld", name, (long long)g.ncells);
   for (double e : g.edges) std::printf(" %a", e);
   std::printf("\n");
}

int main() {
   hrweno::hrweno_grids::grid1 g;
   g.linear(-5.0, 5.0, 100); // example1:41
   dump("linear", g);
   g.log(1e-1, 1e3, 10000);
   dump("log", g);
   g.geometric(1e1, 1e3, 1.1, 100);
   dump("geometric", g);
   const int64_t nc[2] = {124, 365};
   g.bilinear(0.0, 1e1, 1e3, nc);
   dump("bilinear", g);
   return 0;
}
