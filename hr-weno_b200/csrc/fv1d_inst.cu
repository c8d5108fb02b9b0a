// fv1d_inst.cu -- instantiations of the 1D stage kernel for one (k, mode) pair; the Makefile compiles this file
// six times (-DHRW_INST_K=1..3 -DHRW_INST_MODE=0|1) so the 144 specialisations build in parallel.
#include "fv1d.cuh"

// cells per thread run / threads per CTA, per arithmetic mode (profiles/r1_variant_sweeps.txt): the fast kernel has the
// registers for runs of 8 (fewer shared differences recomputed at run ends, half the per-tile overhead per cell)
#ifndef HRW_R1_STRICT
#define HRW_R1_STRICT 4
#endif
#ifndef HRW_NT1_STRICT
#define HRW_NT1_STRICT 256
#endif
#ifndef HRW_R1_FAST
#define HRW_R1_FAST 10 // 80-B thread stride: the 16-B window loads of a quarter warp hit eight distinct bank groups (runs of 8: four-way conflicts)
#endif
#ifndef HRW_NT1_FAST
#define HRW_NT1_FAST 128
#endif

namespace hrw {

constexpr int IK = HRW_INST_K;
#if HRW_INST_MODE == 0
using IM = Strict;
constexpr int R1 = HRW_R1_STRICT, NT1 = HRW_NT1_STRICT;
#else
using IM = Fast;
constexpr int R1 = HRW_R1_FAST, NT1 = HRW_NT1_FAST;
#endif

template <int COMBINE, int FK, int WK, int NT = NT1>
static int launch(const Fv1dGeom &g, const StageArgs &a, cudaStream_t st) {
   auto kern = fv1d_stage_kernel<IK, COMBINE, IM, FK, WK, R1, NT>;
   // persistent grid: SMs x resident CTAs of this specialisation (queried once), never more than there are tiles
   static int resident = 0;
   if (resident == 0) {
      int dev = 0, sms = 0, per_sm = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, 0);
      resident = (sms > 0 ? sms : 148) * (per_sm > 0 ? per_sm : 1);
   }
   int64_t blocks = g.tile_end - g.tile_begin;
   if (blocks > resident) blocks = resident;
   if (blocks < 1) return HRWENO_OK;
   // programmatic dependent launch: this grid may be scheduled while the previous kernel of the stream drains; the kernel
   // itself waits (griddepcontrol.wait) before it touches anything an earlier kernel wrote
   cudaLaunchConfig_t cfg = {};
   cfg.gridDim = dim3((unsigned)blocks);
   cfg.blockDim = dim3(NT);
   cfg.dynamicSmemBytes = 0;
   cfg.stream = st;
   cudaLaunchAttribute attr[1];
   attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
   attr[0].val.programmaticStreamSerializationAllowed = 1;
   cfg.attrs = attr;
   cfg.numAttrs = 1;
   HRW_CUDA(cudaLaunchKernelEx(&cfg, kern, g, a));
   return HRWENO_OK;
}

template <int COMBINE>
static int launch_c(int fk, int wk, int half_tile, const Fv1dGeom &g, const StageArgs &a, cudaStream_t st) {
   if (fk == FK_BURGERS_GODUNOV) {
      // short rows (batched ensembles): the half-size tile wastes fewer thread runs at the row ends
      if (wk == WK_DICT && half_tile) return launch<COMBINE, FK_BURGERS_GODUNOV, WK_DICT, NT1 / 2>(g, a, st);
      if (wk == WK_DICT) return launch<COMBINE, FK_BURGERS_GODUNOV, WK_DICT>(g, a, st);
      return launch<COMBINE, FK_BURGERS_GODUNOV, WK_ARRAY>(g, a, st);
   }
   if (wk == WK_DICT) return launch<COMBINE, FK_GENERIC, WK_DICT>(g, a, st);
   return launch<COMBINE, FK_GENERIC, WK_ARRAY>(g, a, st);
}

#define HRW_CAT2(a, b, c, d) a##b##c##d
#define HRW_CAT(a, b, c, d) HRW_CAT2(a, b, c, d)
int HRW_CAT(fv1d_launch_k, HRW_INST_K, _m, HRW_INST_MODE)(int combine, int fk, int wk, int half_tile, const Fv1dGeom &g, const StageArgs &a, cudaStream_t st) {
   switch (combine) {
   case C_RHS: return launch_c<C_RHS>(fk, wk, half_tile, g, a, st);
   case C_EULER: return launch_c<C_EULER>(fk, wk, half_tile, g, a, st);
   case C_RK2_FINAL: return launch_c<C_RK2_FINAL>(fk, wk, half_tile, g, a, st);
   case C_RK3_S2: return launch_c<C_RK3_S2>(fk, wk, half_tile, g, a, st);
   case C_RK3_S3: return launch_c<C_RK3_S3>(fk, wk, half_tile, g, a, st);
   default: return launch_c<C_MS>(fk, wk, half_tile, g, a, st);
   }
}

#if HRW_INST_K == 3 && HRW_INST_MODE == 0
// defined by one of the six objects
int fv1d_tile_cells(int mode, int half_tile) {
   const int r = mode == HRWENO_MODE_STRICT ? HRW_R1_STRICT : HRW_R1_FAST;
   int nt = mode == HRWENO_MODE_STRICT ? HRW_NT1_STRICT : HRW_NT1_FAST;
   if (half_tile) nt /= 2;
   return (HRW_WARP_TILES ? (nt / 32) * 30 : nt - 2) * r;
}
int fv1d_tile_slots(int mode, int half_tile) { // thread runs x cells per run, including the overlap runs
   const int r = mode == HRWENO_MODE_STRICT ? HRW_R1_STRICT : HRW_R1_FAST;
   const int nt = mode == HRWENO_MODE_STRICT ? HRW_NT1_STRICT : HRW_NT1_FAST;
   return (half_tile ? nt / 2 : nt) * r;
}
#endif

} // namespace hrw
