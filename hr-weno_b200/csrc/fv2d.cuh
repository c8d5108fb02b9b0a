// fv2d.cuh -- fused 2D dimension-split finite-volume stage (example2:73-129).
#pragma once

#include "internal.hpp"
#include "weno_core.cuh"

namespace hrw {

constexpr int PAD2 = 4; // ghost rows above and below a 2D state (>= k)

int fv2d_stage(Fv *fv, int combine, const StageArgs &args, cudaStream_t st);

} // namespace hrw
