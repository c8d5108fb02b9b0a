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
#include "fv2d.cuh"
#include "internal.hpp"
#include "weno_core.cuh"

namespace hrw {

struct GenGeom {
   int64_t n0, n1;  // cells along x1; rows (1D) or cells along x2 (2D)
   int64_t ld;      // pitch of the padded state vectors (vin, a, b, out2)
   int bc;
   const double *cnu0, *cnu1; // per-cell tables or nullptr (uniform tables)
   const double *w0, *w1;     // cell widths
   const double *fc0, *fc1;   // face coefficient along the axis, index 0..n like edges(0:n), or nullptr
   const double *cc0, *cc1;   // cross coefficient: cc0[j] for x1 faces of row j, cc1[i] for x2 faces of column i, or nullptr
   WenoK kc;
   FluxCfg fx0, fx1;
   int phys_l, phys_r; // the row ends are physical boundaries (else slab interfaces: ghost cells hold the neighbour's cells)
   int has_ts;  // the flux carries a time factor g(t) (fluxes.f90:12-18 passes t to the flux)
   double ts;   // its value at this evaluation's time, computed on the host
};

// reconstruct one cell (the arithmetic of recon_kernel, weno.cu).  `cell` points at the cell inside its row, the row is
// strided by `inc` (UNIT: inc == 1); `i` / `n` = its index / the row length, read only when CLAMP (the tile touches a
// domain edge); `has_tab` (the same for every thread of the launch, so the branch stays uniform) selects the cell's
// table cnu(:,:,i) at `ctab` or the uniform tables.
template <int K, bool UNIT, bool CLAMP>
__device__ __forceinline__ void gen_recon(const double *cell, int64_t inc, int64_t i, int64_t n, bool has_tab, const double *ctab,
                                          const WenoK &kc, double &l, double &r, bool phys_l = true, bool phys_r = true) {
   int lo = -(K - 1), hi = K - 1;
   if constexpr (CLAMP) { // edge replicas (weno.f90:171-173): offsets clamped to the row at PHYSICAL ends, in 32 bits
      if (phys_l) lo = -(int)(i < K - 1 ? i : K - 1);
      if (phys_r) hi = (int)(n - 1 - i < K - 1 ? n - 1 - i : K - 1);
   }
   double w[2 * K - 1];
#pragma unroll
   for (int o = -(K - 1); o <= K - 1; ++o) {
      const int oo = CLAMP ? (o < lo ? lo : (o > hi ? hi : o)) : o;
      w[o + K - 1] = UNIT ? cell[oo] : cell[(int64_t)oo * inc];
   }
   if (has_tab) {
      double ci[K * (K + 1)]; // K(K+1) is even and the table is cudaMalloc'ed: 16-B loads
      const double2 *c2 = reinterpret_cast<const double2 *>(ctab);
#pragma unroll
      for (int q = 0; q < K * (K + 1) / 2; ++q) {
         const double2 t = __ldg(c2 + q);
         ci[2 * q] = t.x;
         ci[2 * q + 1] = t.y;
      }
      weno_cell_nonuniform<K, Strict>(ci, w + (K - 1), kc.eps, l, r);
   } else {
      weno_run<K, 1, Strict>(w, kc, &l, &r);
   }
}

// f(v, x, t) = ((model(v)*cross)*face)*g(t), left to right like `v*x(1)*x(2)` (example2:153); an absent factor is not multiplied in
struct GenTs {
   bool has;
   double v;
};
__device__ __forceinline__ double gen_phys(const FluxCfg &c, double v, bool has_cc, double cc, bool has_fc, double fc, const GenTs &ts) {
   double f = phys_flux<Strict>(c, v);
   if (has_cc) f = __dmul_rn(f, cc);
   if (has_fc) f = __dmul_rn(f, fc);
   if (ts.has) f = __dmul_rn(f, ts.v);
   return f;
}

__device__ __forceinline__ double gen_face_flux(const FluxCfg &c, double vm, double vp, bool has_cc, double cc, bool has_fc, double fc,
                                                const GenTs &ts) {
   const double fm = gen_phys(c, vm, has_cc, cc, has_fc, fc, ts);
   const double fp = gen_phys(c, vp, has_cc, cc, has_fc, fc, ts);
   if (c.scheme == HRWENO_SCHEME_LAX_FRIEDRICHS) // (f(vm) + f(vp) - alpha*(vp - vm))/2      fluxes.f90:43
      return __dmul_rn(__dsub_rn(__dadd_rn(fm, fp), __dmul_rn(c.alpha, __dsub_rn(vp, vm))), 0.5);
   const double lo = fm < fp ? fm : fp; // fluxes.f90:70-74
   const double hi = fm > fp ? fm : fp;
   return vm <= vp ? lo : hi;
}

// boundary rule on the two faces of cell i of a row of n cells: interior faces are given in fl (i > 0) and fr (i < n-1)
__device__ __forceinline__ void gen_bc(int bc, int64_t i, int64_t n, double &fl, double &fr, bool phys_l = true, bool phys_r = true) {
   const bool zero = bc == HRWENO_BC_ZERO_FLUX;
   if (i == 0 && phys_l) fl = zero ? 0.0 : fr;     // fedges(0) = fedges(1) (example1:103) | 0 (example2:117,119); copy needs n >= 2
   if (i == n - 1 && phys_r) fr = zero ? 0.0 : fl; // fedges(nc) = fedges(nc-1) (example1:104) | 0 (example2:118,120)
}

// tile of one CTA: 2D 32x16 cells (two per thread in phase B), 1D 254 cells (threads 1..254 own one; 0 and 255 only
// reconstruct the frame cells), so that phase A is a whole number of full passes over the 256 threads
constexpr int GEN_NT = 256;
// resident CTAs per SM the register allocation is capped for.  The kernel is latency-bound (long dependent fp64 chains
// of the exact divisions), so warps beat registers: measured 4 / 5 / 6 per SM -> 1D 3.10 / 3.16 / 2.82e10 cell-stages/s,
// 2D 1.02 / 1.19 / 1.31e10 cell-steps/s (profiles/r1_variant_sweeps.txt)
template <bool TWO_D>
constexpr int gen_minb() { return TWO_D ? 6 : 5; }
template <bool TWO_D>
struct GenTile {
   static constexpr int TX = TWO_D ? 32 : GEN_NT - 2, TY = TWO_D ? 16 : 1;
   static constexpr int SX = TX + 2;                    // x1 sweep: cells i0-1 .. i0+TX
   static constexpr int N1 = SX * TY;                   // items of the x1 sweep (2D: 544 = 17 warps; 1D: 256)
   static constexpr int N2 = TWO_D ? TX * (TY + 2) : 0; // items of the x2 sweep: cells j0-1 .. j0+TY
};

// One tile.  INTERIOR: the tile, its frame and their stencils lie inside the domain -- no bounds tests, no clamped
// offsets, no boundary rule (the common case on large grids; the arithmetic per cell is the same code).
template <int K, bool TWO_D, bool INTERIOR>
__device__ __forceinline__ void gen_tile(const GenGeom &g, const StageArgs &a, const int combine, const int64_t i0, const int64_t j0,
                                         double *s_l1, double *s_r1, double *s_l2, double *s_r2) {
   using T = GenTile<TWO_D>;
   constexpr int TX = T::TX, TY = T::TY, SX = T::SX, N1 = T::N1, N2 = T::N2, KK = K * (K + 1);
   const double *vt = a.vin + j0 * g.ld + i0; // cell (i0, j0)
   const bool has0 = g.cnu0 != nullptr, has1 = g.cnu1 != nullptr;
   const double *t0 = g.cnu0 + i0 * KK, *t1 = g.cnu1 + j0 * KK; // tables of cell i0 / j0 (dereferenced only when present)
   // ---- phase A: reconstruct the tile and its frame --------------------------------------------------------
   for (int q = threadIdx.x; q < N1 + N2; q += GEN_NT) {
      if (q < N1) { // N1 is a multiple of 32: the branch is warp-uniform
         const int iy = q / SX, ix = q - iy * SX - 1; // cell (i0 + ix, j0 + iy), ix = -1 .. TX
         const int64_t i = i0 + ix, j = j0 + iy;
         // at a slab interface the frame cell beyond the row end is the neighbour's edge cell (a ghost cell of the padded state)
         if (INTERIOR || (i >= (g.phys_l ? 0 : -1) && i < g.n0 + (g.phys_r ? 0 : 1) && j < g.n1)) {
            double l, r;
            gen_recon<K, true, !INTERIOR>(vt + (int64_t)iy * g.ld + ix, 1, i, g.n0, has0, t0 + ix * KK, g.kc, l, r, g.phys_l != 0,
                                          g.phys_r != 0); // example1:93, example2:98 (contiguous row)
            s_l1[q] = l;
            s_r1[q] = r;
         }
      } else if constexpr (TWO_D) {
         const int p = q - N1;
         const int iy = p / TX - 1, ix = p - (iy + 1) * TX; // cell (i0 + ix, j0 + iy), iy = -1 .. TY
         const int64_t i = i0 + ix, j = j0 + iy;
         if (INTERIOR || (i < g.n0 && j >= 0 && j < g.n1)) {
            double l, r;
            gen_recon<K, false, !INTERIOR>(vt + (int64_t)iy * g.ld + ix, g.ld, j, g.n1, has1, t1 + iy * KK, g.kc, l,
                                           r); // example2:107 (stride-nc1 column)
            s_l2[p] = l;
            s_r2[p] = r;
         }
      }
   }
   __syncthreads();
   // ---- phase B: faces, divergence, combination ------------------------------------------------------------
   for (int c = threadIdx.x; c < (TWO_D ? TX * TY : GEN_NT); c += GEN_NT) {
      int lx, ly;
      if constexpr (TWO_D) {
         ly = c / TX;
         lx = c - ly * TX;
      } else {
         ly = 0;
         lx = c - 1; // thread t reconstructed cell i0-1+t and owns it when 1 <= t <= TX
         if (lx < 0 || lx >= TX) continue;
      }
      const int64_t i = i0 + lx, j = j0 + ly;
      if (!INTERIOR && !(i < g.n0 && j < g.n1)) continue;
      // x1: face f lies between cells f-1 and f: godunov(flux, vr(f-1), vl(f), [right(f-1), center2(j)])  (example1:99, example2:100)
      const GenTs ts{g.has_ts != 0, g.ts};
      const bool hc0 = g.cc0 != nullptr, hf0 = g.fc0 != nullptr;
      const double cc0 = hc0 ? g.cc0[j] : 1.0;
      const int c1 = lx + 1 + ly * SX; // shared index of cell (i, j) in the x1 arrays
      double fl = 0.0, fr = 0.0;
      if (INTERIOR || i > 0 || !g.phys_l) fl = gen_face_flux(g.fx0, s_r1[c1 - 1], s_l1[c1], hc0, cc0, hf0, hf0 ? g.fc0[i] : 1.0, ts);
      if (INTERIOR || i < g.n0 - 1 || !g.phys_r) fr = gen_face_flux(g.fx0, s_r1[c1], s_l1[c1 + 1], hc0, cc0, hf0, hf0 ? g.fc0[i + 1] : 1.0, ts);
      if constexpr (!INTERIOR) gen_bc(g.bc, i, g.n0, fl, fr, g.phys_l != 0, g.phys_r != 0);
      double L = -__ddiv_rn(__dsub_rn(fr, fl), g.w0[i]); // -(fedges(i) - fedges(i-1))/width(i)   example1:107, example2:125
      if constexpr (TWO_D) {
         const bool hc1 = g.cc1 != nullptr, hf1 = g.fc1 != nullptr;
         const double cc1 = hc1 ? g.cc1[i] : 1.0;
         const int c2 = lx + (ly + 1) * TX; // shared index of cell (i, j) in the x2 arrays
         double gl = 0.0, gr = 0.0;
         if (INTERIOR || j > 0) gl = gen_face_flux(g.fx1, s_r2[c2 - TX], s_l2[c2], hc1, cc1, hf1, hf1 ? g.fc1[j] : 1.0, ts);
         if (INTERIOR || j < g.n1 - 1) gr = gen_face_flux(g.fx1, s_r2[c2], s_l2[c2 + TX], hc1, cc1, hf1, hf1 ? g.fc1[j + 1] : 1.0, ts);
         if constexpr (!INTERIOR) gen_bc(g.bc, j, g.n1, gl, gr);
         L = __dsub_rn(L, __ddiv_rn(__dsub_rn(gr, gl), g.w1[j])); // ... - (fedges2(j,i) - fedges2(j-1,i))/width2(j)   example2:126
      }
      // stage combination (tvdode.f90:141,149-167,257), the expressions of combine_kernel (ode.cu)
      const int64_t off = j * g.ld + i;
      const double x = a.vin[off];
      double o;
      switch (combine) {
      case C_RHS: o = L; break;
      case C_EULER: o = __dadd_rn(x, __dmul_rn(a.c0, L)); break;
      case C_RK2_FINAL: o = __dmul_rn(__dadd_rn(__dadd_rn(a.a[off], x), __dmul_rn(a.c0, L)), 0.5); break;
      case C_RK3_S2: o = __dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(3.0, a.a[off]), x), __dmul_rn(a.c0, L)), 0.25); break;
      case C_RK3_S3: o = div3<Strict>(__dadd_rn(__fma_rn(2.0, x, a.a[off]), __dmul_rn(a.c0, L))); break; // /3: the exact division of K2/K3
      default: // C_MS
         o = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(25.0, x), __dmul_rn(a.c0, L)), __dmul_rn(7.0, a.a[off])),
                                 __dmul_rn(a.c1, a.b[off])),
                       0.03125);
         a.out2[off] = L; // out2 aliases b element for element: b[off] was read above
      }
      a.out[j * a.ld_out + i] = o; // out may alias a element for element (read above); it never aliases vin
   }
   __syncthreads(); // the next tile overwrites the shared arrays
}

template <int K, bool TWO_D>
__global__ void __launch_bounds__(GEN_NT, gen_minb<TWO_D>()) fvgen_stage_kernel(const GenGeom g, const StageArgs a, const int combine) {
   using T = GenTile<TWO_D>;
   constexpr int TX = T::TX, TY = T::TY, N1 = T::N1, N2 = T::N2;
   __shared__ double s_l1[N1], s_r1[N1];
   __shared__ double s_l2[TWO_D ? N2 : 1], s_r2[TWO_D ? N2 : 1];
   const int64_t tiles_x = (g.n0 + TX - 1) / TX, tiles_y = (g.n1 + TY - 1) / TY;
   for (int64_t tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
      const int64_t tj = tile / tiles_x;
      const int64_t i0 = (tile - tj * tiles_x) * TX, j0 = tj * TY;
      // frame cell i0-1 reaches down to i0-K, frame cell i0+TX up to i0+TX+K-1.  The interior specialisation is used
      // in 1D only: measured +2.5 % there, -10 % in 2D, where the second inlined body costs registers the 40-register cap
      // (6 CTAs/SM) does not have (profiles/r1_variant_sweeps.txt)
      bool interior = false;
      if constexpr (!TWO_D) interior = i0 >= K && i0 + TX + K <= g.n0;
      if (interior)
         gen_tile<K, false, true>(g, a, combine, i0, j0, s_l1, s_r1, s_l2, s_r2);
      else
         gen_tile<K, TWO_D, false>(g, a, combine, i0, j0, s_l1, s_r1, s_l2, s_r2);
   }
}

template <int K>
static void fvgen_launch(bool two_d, unsigned blocks, const GenGeom &g, const StageArgs &a, int combine, cudaStream_t st) {
   (void)two_d; // 2D general operators run on the tile kernel of fv2d.cu (GEN = 1) since round 2: 1D rows only here
   fvgen_stage_kernel<K, false><<<blocks, GEN_NT, 0, st>>>(g, a, combine);
}

int fvgen_stage(Fv *fv, int combine, const StageArgs &args, cudaStream_t st) {
   const hrweno_fv_desc &d = fv->d;
   if (d.ndim == 2) return fail(HRWENO_EINVAL, "general stage: 2D operators take the tile kernel (fv2d.cu)");
   const bool two_d = false;
   GenGeom g{};
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
   g.fx0 = FluxCfg{d.flux_model, d.flux_scheme, d.flux_coef[0], d.alpha};
   g.fx1 = FluxCfg{d.flux_model, d.flux_scheme, d.flux_coef[1], d.alpha};
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
