// fv2d.cu -- placeholder until the tile kernel lands
#include "fv2d.cuh"

namespace hrw {
int fv2d_stage(Fv *, int, const StageArgs &, cudaStream_t) { return fail(HRWENO_EINVAL, "2D fused stage not built yet"); }
} // namespace hrw
