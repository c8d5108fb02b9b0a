// fvgen.cu -- general fused finite-volume stage (kernel K7): non-uniform grids and x-dependent fluxes.
//
// The tuned stage kernels (fv1d.cuh, fv2d.cu) are specialised for what the two shipped examples run: the uniform-grid
// tables c1/c2/c3 (weno.f90:12-21) and fluxes that do not depend on x.  The reference API reaches further:
//   * `weno(ncells, k, eps, xedges)` selects the per-cell tables cnu(:,:,i) (weno.f90:100-112, 177, 221-297), which
//     is what a population balance on a geometric or logarithmic grid1 (grids.f90:138-230) needs;
//   * the abstract flux `f(u, x(:), t)` (fluxes.f90:12-18) receives the face coordinates -- example2 passes
//     `[gx(1)%right(i), gx(2)%center(j)]` / `[gx(1)%center(i), gx(2)%right(j)]` (example2:100-101,109-110) and hints at
//     the growth terms `v*x(1)**2`, `v*x(1)*x(2)` (example2:140,153).
// This kernel covers both behind the same fused stage interface (StageArgs, Combine): one pass reads the stage input,
// reconstructs along x1 (and x2) with per-cell or uniform tables, evaluates the numerical flux with per-face and
// per-cross-cell coefficients, forms -(df1)/w1 [- (df2)/w2] and applies the RK / multistep combination.  Arithmetic is
// ALWAYS the reference's operation order (Strict policy: separately rounded IEEE operations, __ddiv_rn), i.e.
// bit-identical to the oracle; the reconstruction of a cell is the code path of recon_kernel (weno.cu).
//
// Decomposition: CTA = 256 threads = one tile of 254x1 cells (1D rows) or 32x16 cells (2D).  Phase A reconstructs the
// tile plus a one-cell frame along the sweep direction (vl, vr -> shared memory; windows are read straight from global
// memory with the cell index clamped to the row, which IS the edge replication of weno.f90:171-173, so no ghost cell
// of the padded layout is read or written); the tile shapes make it whole passes over the 256 threads (1D: one pass,
// 2D: 1120 items, 2.19 reconstructions per cell).  Phase B: a thread owns one (1D) or two (2D) cells: two faces per
// direction from shared memory, boundary rule, divergence, combination, store.  HBM-bound by the 8k(k+1) B/cell of coefficient traffic per
// non-uniform axis (96 B/cell at k=3) on top of the stage's 16-40 B/cell.
// Slabs: the frame cells beyond a slab interface are the neighbour's edge cells (ghost cells of the padded state, filled by
// fv_exchange after every stage); their tables come with the slab's own (fv.cu: fv_set_xedges takes the global edges).
#include "fvgen.cuh"

namespace hrw {

template <int K>
static void fvgen_launch(bool two_d, unsigned blocks, const GenGeomT<double> &g, const StageArgs &a, int combine, cudaStream_t st) {
   (void)two_d; // 2D general operators run on the tile kernel of fv2d.cu (GEN = 1) since round 2: 1D rows only here
   fvgen_stage_kernel<K, false, double, StageArgs><<<blocks, GEN_NT, 0, st>>>(g, a, combine);
}

int fvgen_stage(Fv *fv, int combine, const StageArgs &args, cudaStream_t st) {
   const hrweno_fv_desc &d = fv->d;
   if (d.ndim == 2) return fail(HRWENO_EINVAL, "general stage: 2D operators take the tile kernel (fv2d.cu)");
   const bool two_d = false;
   GenGeomT<double> g{};
   g.n0 = fv->n0;
   g.n1 = two_d ? fv->n1 : fv->rows;
   g.ld = fv->pitch;
   g.bc = d.bc;
   g.cnu0 = fv->d_cnu[0];
   g.cnu1 = fv->d_cnu[1];
   g.w0 = fv->d_width[0];
   g.w1 = fv->d_width[1];
   g.fc0 = fv->d_fcoef[0];
   g.fc1 = fv->d_fcoef[1];
   g.cc0 = fv->d_ccoef[0];
   g.cc1 = fv->d_ccoef[1];
   g.kc = make_wenok(d.eps);
   g.eps = d.eps;
   g.fx0 = FluxCfgT<double>{d.flux_model, d.flux_scheme, d.flux_coef[0], d.alpha};
   g.fx1 = FluxCfgT<double>{d.flux_model, d.flux_scheme, d.flux_coef[1], d.alpha};
   g.phys_l = d.rank == 0;
   g.phys_r = d.rank == d.nranks - 1;
   g.has_ts = fv->tfn != nullptr;
   g.ts = fv->tfn ? fv->tfn(fv->tfn_ctx, args.t) : 1.0; // the caller's g(t) at this evaluation's time
   if (!g.w0 || (two_d && !g.w1)) return fail(HRWENO_EINVAL, "general stage: width arrays missing");
   const int64_t tx = two_d ? GenTile<true>::TX : GenTile<false>::TX, ty = two_d ? GenTile<true>::TY : GenTile<false>::TY;
   const int64_t tiles = ((g.n0 + tx - 1) / tx) * ((g.n1 + ty - 1) / ty);
   const int64_t cap = 148 * (two_d ? gen_minb<true>() : gen_minb<false>()); // resident CTAs of 256 threads per SM
   const unsigned blocks = (unsigned)(tiles < cap ? tiles : cap);
   if (d.k == 1)
      fvgen_launch<1>(two_d, blocks, g, args, combine, st);
   else if (d.k == 2)
      fvgen_launch<2>(two_d, blocks, g, args, combine, st);
   else
      fvgen_launch<3>(two_d, blocks, g, args, combine, st);
   HRW_CUDA(cudaGetLastError());
   return HRWENO_OK;
}

} // namespace hrw
