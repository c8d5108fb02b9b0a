// tmap.hpp -- TMA tensor maps (cuTensorMapEncodeTiled) for the padded 2D states, cached per (buffer, box).
// The driver entry point is resolved at run time (cudaGetDriverEntryPoint), so the library does not link libcuda.
#pragma once

#include <cuda.h>

#include "internal.hpp"

namespace hrw {

// tensor map of a padded 2D state whose cell (0,0) is at `cell0` (rows -PAD2.., columns -PAD..; fv2d.cuh) for boxes of
// box0 x box1 doubles; coordinates are (column + PAD, row + PAD2).  Out-of-bounds parts of a box arrive as zeros.
int fv_tmap_2d(Fv *fv, const double *cell0, uint32_t box0, uint32_t box1, CUtensorMap *out);

} // namespace hrw
