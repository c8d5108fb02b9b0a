// fv2d.cu -- fused 2D dimension-split finite-volume stage (example2:73-129).
//
// One CTA owns a TX x TY tile of cells.  The tile plus a 4-cell frame is staged in shared memory once
// (coalesced 16-B loads along x1); both sweeps read it from there, so the stride-nc1 column gathers of
// the reference (example2:107) never touch HBM.
//   phase A: reconstruction along x1 for every tile row and along x2 for every tile column.  Work items
//            are runs of R=4 consecutive cells of a line (row or column) -- the same register-window
//            routine as the 1D kernel, addressed by (base, stride) so x1- and x2-items share one code
//            path.  The faces on the tile edge need vr / vl of the neighbouring tile's cells: single-cell
//            "halo items".  x1-results go to shared memory (the consumer is a different thread); an
//            x2-item (column lx, run ry) is owned by the thread that later does phase B for the same
//            cells, so its vl / vr stay in registers and only the run-end values are exchanged.
//   phase B: thread = (x, run of R cells along x2): Godunov / Lax-Friedrichs flux at the x1- and x2-faces,
//            boundary constraints (example2:117-120), vdot = -(df1)/w1 - (df2)/w2 (example2:123-127),
//            stage combination, coalesced stores along x1.
#include "fv2d.cuh"
#include "fv1d.cuh" // TMA bulk copy + mbarrier helpers
#include "tmap.hpp"

namespace hrw {

struct Fv2dGeom {
   int64_t n0, n1;   // cells along x1 (contiguous) and x2 (local slab)
   int64_t pitch;    // padded row pitch
   int tiles_x, tiles_y;
   int small_tiles;  // 32x16 tiles instead of 64x32
   const double *w1, *w2;   // widths (device, padded)
   const double *rw1, *rw2; // their refined reciprocals (exact_recip)
   WenoK kc;
   FluxCfg flux1, flux2;
   int bc;
   int phys_lo, phys_hi; // x2 ends are physical boundaries (x1 ends always are)
   // general operator (GEN = 1): per-cell reconstruction tables cnu(:,:,i) per axis (weno.f90:100-112,177) and the flux
   // coefficients f = ((model(v)*cross)*face)*g(t) (fluxes.f90:12-18; example2:100-101,109-110,140,153); nullptr = absent
   const double *cnu0, *cnu1;
   const double *fc0, *fc1, *cc0, *cc1;
   int has_ts;
   double ts;
};

// reconstruction of a run of R cells of a line whose first cell has index i0 on its axis (n cells): uniform tables, or the
// per-cell tables of weno(ncells, k, eps, xedges) at `cnu` (indices clamped to the axis: cells beyond it are never used)
// `tab` = the table of the run's first cell inside the per-tile copy of the axis tables in shared memory (cells -1 .. T of
// the tile, staged by TMA with the tile; consecutive cells KK doubles apart), or nullptr for the uniform tables.
template <int K, int R, class M>
__device__ __forceinline__ void run2d(const double *w /* window, cell j at w[2 + j] */, const WenoK &kc, const double *tab, double *vl,
                                      double *vr) {
   if (tab == nullptr) {
      weno_run<K, R, M>(w + (2 - (K - 1)), kc, vl, vr);
   } else {
      constexpr int KK = K * (K + 1);
#pragma unroll
      for (int j = 0; j < R; ++j) {
         double ci[KK]; // K(K+1) is even, the staged tables are 16-B aligned
         const double2 *c2 = reinterpret_cast<const double2 *>(tab + j * KK);
#pragma unroll
         for (int q = 0; q < KK / 2; ++q) {
            const double2 t = c2[q];
            ci[2 * q] = t.x;
            ci[2 * q + 1] = t.y;
         }
         weno_cell_nonuniform<K, M>(ci, w + 2 + j, kc.eps, vl[j], vr[j]);
      }
   }
}

// numerical flux at one face with the general operator's coefficients, reference operation order (fvgen.cu: same expressions)
__device__ __forceinline__ double gen2d_phys(const FluxCfg &c, double v, const double *cc, const double *fc, int has_ts, double ts) {
   double f = phys_flux<Strict>(c, v);
   if (cc) f = __dmul_rn(f, *cc);
   if (fc) f = __dmul_rn(f, *fc);
   if (has_ts) f = __dmul_rn(f, ts);
   return f;
}
__device__ __forceinline__ double gen2d_face(const FluxCfg &c, double vm, double vp, const double *cc, const double *fc, int has_ts, double ts) {
   const double fm = gen2d_phys(c, vm, cc, fc, has_ts, ts), fp = gen2d_phys(c, vp, cc, fc, has_ts, ts);
   if (c.scheme == HRWENO_SCHEME_LAX_FRIEDRICHS) return __dmul_rn(__dsub_rn(__dadd_rn(fm, fp), __dmul_rn(c.alpha, __dsub_rn(vp, vm))), 0.5);
   const double lo = fm < fp ? fm : fp, hi = fm > fp ? fm : fp; // fluxes.f90:70-74
   return vm <= vp ? lo : hi;
}

template <int TX, int TY, int UPW = 0>
struct Tile2d {
   static constexpr int R = 4, H = 4;
   static constexpr int SP = TX + 2 * H;         // tile pitch
   static constexpr int SROWS = TY + 2 * H;
   static constexpr int XP = TX + 2 * R;         // pitch of the x1-sweep vl/vr arrays (cells x0-R .. x0+TX+R-1)
   static constexpr int RY = TY / R;             // x2-runs per column
   // two staged-tile buffers: the next tile is fetched by TMA row copies while this one is being computed
   static constexpr int OFF_V = 0;
   static constexpr int TILE_DOUBLES = SROWS * SP;
   static constexpr int OFF_VLX = OFF_V + 2 * TILE_DOUBLES;
   // upwind specialisation: the left-side arrays are never touched and get no storage (more CTAs per SM)
   static constexpr int OFF_VRX = OFF_VLX + (UPW ? 0 : TY * XP);
   // x2-sweep exchange: vr of the cell below each run (slot 0: the cell below the tile) and vl of the cell above it
   // (slot RY: the cell above the tile); slot s of column lx at [s * TX + lx]
   static constexpr int OFF_VRE = OFF_VRX + TY * XP;
   static constexpr int OFF_VLE = OFF_VRE + (RY + 1) * TX;
   static constexpr int TOTAL = OFF_VLE + (UPW ? 0 : (RY + 1) * TX);
   // per-tile copies of the widths and their reciprocals along x1 (TX) and x2 (TY)
   static constexpr int OFF_W1 = TOTAL, OFF_RW1 = OFF_W1 + TX, OFF_W2 = OFF_RW1 + TX, OFF_RW2 = OFF_W2 + TY;
   // pointwise operands of the tile (a, b), fetched by TMA at the top of the tile iteration: 128-B aligned boxes of TX x TY
   static constexpr int OFF_OPS = ((OFF_RW2 + TY + 15) / 16) * 16;
   static constexpr int OP_DOUBLES = TX * TY;
   static constexpr size_t bytes(int nstaged) { return (size_t)(OFF_OPS + nstaged * OP_DOUBLES) * sizeof(double); }
};

// pointwise operands the combination reads (a; b for the multistep stage) and how many of them are staged in shared memory
// by TMA: all of them for the upwind variant; the two-sided variant has room for one next to its vl/vr arrays if two
// CTAs are to share an SM.
__host__ __device__ constexpr int fv2d_nops(int combine) {
   return combine == C_MS ? 2 : ((combine == C_RK2_FINAL || combine == C_RK3_S2 || combine == C_RK3_S3) ? 1 : 0);
}
__host__ __device__ constexpr int fv2d_nstaged(int combine, int upw) { return upw ? fv2d_nops(combine) : (fv2d_nops(combine) > 0 ? 1 : 0); }

// 2D tensor-map TMA load (SASS: UTMALDG) of one box whose first element is (c0, c1), completing on an mbarrier
__device__ __forceinline__ void tma_tensor_2d_g2s(uint32_t dst_smem, const CUtensorMap *tm, int c0, int c1, uint32_t bar) {
   asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
                "l"(tm), "r"(c0), "r"(c1), "r"(bar)
                : "memory");
}

// UPW = 1: linear flux f = a*v with a >= 0 in both directions and the Godunov rule.  Then vm <= vp implies
// a*vm <= a*vp (rounding is monotone), so godunov (fluxes.f90:70-74) returns f(vm) = a*vr(i) in every case and
// vl never influences the result (SURVEY 3.2 notes this for example2): the left-side reconstruction is skipped.
// The result is bit-identical to evaluating both sides.
template <int UPW, class M>
__device__ __forceinline__ double face2d(const FluxCfg &c, double vm, double vp) {
   if constexpr (UPW) {
      // a separately rounded product in both modes: a plain a*vm could be contracted with the flux difference that follows
      // in one variant of phase B (interior / edge tiles) and not in the other, and fast-mode results would depend on
      // where the tile and slab edges fall (see fv1d.cuh: face_flux_k)
      return __dmul_rn(c.coef, vm);
   } else {
      return face_flux<M>(c, vm, vp);
   }
}

// phase B for one thread = (column lx, run of R rows).  INTERIOR: the tile touches no domain edge and is complete.
template <int K, int COMBINE, class M, int UPW, int TX, int TY, bool INTERIOR, int GEN = 0>
__device__ __forceinline__ void fv2d_phase_b(const Fv2dGeom &g, const StageArgs &s, const double *s_v, const double *s_vlx,
                                             const double *s_vrx, const double *s_wd /* staged widths */, const double *s_ops,
                                             const double *vmY /* vr below face j, j = 0..R */,
                                             const double *vpY /* vl above face j */, int64_t x0, int64_t y0, int lx, int ly0) {
   using T = Tile2d<TX, TY, UPW>;
   constexpr int R = T::R, H = T::H;
   constexpr int NSTG = fv2d_nstaged(COMBINE, UPW);
   const int64_t gx = x0 + lx, gy0 = y0 + ly0;
   if constexpr (!INTERIOR) {
      if (gx >= g.n0 || gy0 >= g.n1) return;
   }
   const bool copy = g.bc == HRWENO_BC_COPY_NEIGHBOUR;
   // x2-faces gy0 .. gy0+R of column gx: face f lies between rows f-1 and f
   double F2[R + 1];
#pragma unroll
   for (int j = 0; j <= R; ++j) {
      if constexpr (GEN) { // x2 face gy0 + j of column gx: [center1(i), right2(j)] (example2:109-110)
         const int64_t f = gy0 + j > g.n1 ? g.n1 : gy0 + j, c = gx < g.n0 ? gx : g.n0 - 1;
         F2[j] = gen2d_face(g.flux2, vmY[j], vpY[j], g.cc1 ? g.cc1 + c : nullptr, g.fc1 ? g.fc1 + f : nullptr, g.has_ts, g.ts);
      } else {
         F2[j] = face2d<UPW, M>(g.flux2, vmY[j], UPW ? 0.0 : vpY[j]);
      }
   }
   if constexpr (!INTERIOR) {
      if (g.phys_lo && gy0 == 0) F2[0] = copy ? F2[1] : 0.0;
      if (g.phys_hi) {
#pragma unroll
         for (int j = 1; j <= R; ++j)
            if (gy0 + j == g.n1) F2[j] = copy ? F2[j - 1] : 0.0;
      }
   }
   // widths of the tile's columns and rows come from the per-tile copies in shared memory
   const double w1 = s_wd[lx], rw1 = s_wd[TX + lx];
   // pointwise operands first: out may alias a (and out2 alias b) element for element in the multistep stage, so
   // loads issued after the first store could not be hoisted by the compiler and would serialise on DRAM latency.
   // Staged operands were fetched by TMA at the top of the tile iteration (s_ops: a, then b, TX x TY boxes).
   constexpr bool NEED_A = COMBINE == C_RK2_FINAL || COMBINE == C_RK3_S2 || COMBINE == C_RK3_S3 || COMBINE == C_MS;
   double av[R], bv[R], w2v[R], rw2v[R];
#pragma unroll
   for (int j = 0; j < R; ++j) {
      const int64_t gy = gy0 + j;
      const bool in = INTERIOR || gy < g.n1;
      const int64_t off = gy * g.pitch + gx;
      av[j] = 0.0;
      bv[j] = 0.0;
      if constexpr (NEED_A) av[j] = NSTG >= 1 ? s_ops[(ly0 + j) * TX + lx] : (in ? s.a[off] : 0.0);
      if constexpr (COMBINE == C_MS) bv[j] = NSTG >= 2 ? s_ops[T::OP_DOUBLES + (ly0 + j) * TX + lx] : (in ? s.b[off] : 0.0);
      w2v[j] = s_wd[2 * TX + ly0 + j];
      rw2v[j] = s_wd[2 * TX + TY + ly0 + j];
   }
   double res[R], lres[R];
#pragma unroll
   for (int j = 0; j < R; ++j) {
      const int ly = ly0 + j;
      // x1-faces gx and gx+1 of row gy
      const double *vlx = s_vlx + ly * T::XP + R + lx, *vrx = s_vrx + ly * T::XP + R + lx;
      double Fl, Fr;
      if constexpr (GEN) { // x1 faces gx, gx+1 of row gy: [right1(i), center2(j)] (example2:100-101)
         const int64_t gyc = gy0 + j < g.n1 ? gy0 + j : g.n1 - 1, f0 = gx < g.n0 ? gx : g.n0 - 1;
         const double *cc = g.cc0 ? g.cc0 + gyc : nullptr;
         Fl = gen2d_face(g.flux1, vrx[-1], vlx[0], cc, g.fc0 ? g.fc0 + f0 : nullptr, g.has_ts, g.ts);
         Fr = gen2d_face(g.flux1, vrx[0], vlx[1], cc, g.fc0 ? g.fc0 + f0 + 1 : nullptr, g.has_ts, g.ts);
      } else {
         Fl = face2d<UPW, M>(g.flux1, vrx[-1], UPW ? 0.0 : vlx[0]);
         Fr = face2d<UPW, M>(g.flux1, vrx[0], UPW ? 0.0 : vlx[1]);
      }
      if constexpr (!INTERIOR) {
         if (copy) { // fedges(0) = fedges(1), fedges(nc) = fedges(nc-1)
            if (gx == 0) Fl = Fr;
            if (gx == g.n0 - 1) Fr = Fl;
         } else {
            if (gx == 0) Fl = 0.0;
            if (gx == g.n0 - 1) Fr = 0.0;
         }
      }
      // vdot = -(f1(i)-f1(i-1))/w1(i) - (f2(j)-f2(j-1))/w2(j)   (example2:123-127)
      double L;
      if constexpr (M::strict) {
         // IEEE quotients from the precomputed refined reciprocals (common.cuh: exact_div_q), cold fallback
         bool ok = true;
         const double d1 = M::sub(Fr, Fl), d2 = M::sub(F2[j + 1], F2[j]);
         L = M::sub(-exact_div_q(d1, w1, rw1, ok), exact_div_q(d2, w2v[j], rw2v[j], ok));
         if (!ok) L = M::sub(-M::div(d1, w1), M::div(d2, w2v[j]));
      } else {
         L = fma(-(F2[j + 1] - F2[j]), rw2v[j], -((Fr - Fl) * rw1));
      }
      const double v = s_v[(ly + H) * T::SP + lx + H];
      double o;
      if constexpr (COMBINE == C_RHS) {
         o = L;
      } else if constexpr (COMBINE == C_EULER) {
         o = M::add(v, M::mul(s.c0, L));
      } else if constexpr (COMBINE == C_RK2_FINAL) {
         o = M::mul(M::add(M::add(av[j], v), M::mul(s.c0, L)), 0.5);
      } else if constexpr (COMBINE == C_RK3_S2) {
         o = M::mul(M::add(M::add(M::mul(3.0, av[j]), v), M::mul(s.c0, L)), 0.25);
      } else if constexpr (COMBINE == C_RK3_S3) {
         o = div3<M>(M::add(M::fma_exact(2.0, v, av[j]), M::mul(s.c0, L)));
      } else {
         o = M::mul(M::add(M::add(M::add(M::mul(25.0, v), M::mul(s.c0, L)), M::mul(7.0, av[j])), M::mul(s.c1, bv[j])), 0.03125);
         lres[j] = L;
      }
      res[j] = o;
   }
#pragma unroll
   for (int j = 0; j < R; ++j) {
      const int64_t gy = gy0 + j;
      if constexpr (!INTERIOR) {
         if (gy >= g.n1) break;
      }
      const int64_t off = gy * g.pitch + gx;
      const double o = res[j];
      if constexpr (COMBINE == C_MS) s.out2[off] = lres[j];
      if constexpr (INTERIOR) {
         s.out[off] = o;
      } else if (s.out_dense) {
         s.out[gy * s.ld_out + gx] = o;
      } else {
         s.out[off] = o;
         if (COMBINE != C_RHS) {
            // ghost cells of the result at physical boundaries (edge replicas, weno.f90:172-173)
            if (gx == 0) {
#pragma unroll
               for (int q = 1; q <= K; ++q) s.out[off - q] = o;
            }
            if (gx == g.n0 - 1) {
#pragma unroll
               for (int q = 1; q <= K; ++q) s.out[off + q] = o;
            }
            if (g.phys_lo && gy == 0) {
#pragma unroll
               for (int q = 1; q <= K; ++q) s.out[off - q * g.pitch] = o;
            }
            if (g.phys_hi && gy == g.n1 - 1) {
#pragma unroll
               for (int q = 1; q <= K; ++q) s.out[off + q * g.pitch] = o;
            }
         }
      }
   }
}

#ifndef HRW_MINB2_BIG
#define HRW_MINB2_BIG 2 // resident 512-thread CTAs per SM the registers are capped for (64 registers)
#endif
// resident CTAs per SM the registers are capped for: 3 tiles of the upwind fast variant fit in shared memory, 2 otherwise
template <class M, int UPW, int NT>
constexpr int fv2d_min_blocks() {
   return NT > 256 ? HRW_MINB2_BIG : ((UPW && !M::strict) ? 3 : 2) * (256 / NT);
}

template <int K, int COMBINE, class M, int UPW, int TX, int TY, int NT, int GEN = 0>
__global__ void __launch_bounds__(NT, fv2d_min_blocks<M, UPW, NT>())
   fv2d_stage_kernel(const Fv2dGeom g, const StageArgs s, const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_a,
                     const __grid_constant__ CUtensorMap tm_b) {
   using T = Tile2d<TX, TY, UPW>;
   constexpr int R = T::R, H = T::H;
   constexpr int NSTG = fv2d_nstaged(COMBINE, UPW);
   extern __shared__ __align__(128) double smem[];
   double *s_v = smem + T::OFF_V;
   double *s_vlx = smem + T::OFF_VLX, *s_vrx = smem + T::OFF_VRX;
   double *s_vre = smem + T::OFF_VRE, *s_vle = smem + T::OFF_VLE;
   double *s_wd = smem + T::OFF_W1;
   double *s_ops = smem + T::OFF_OPS;
   // GEN: per-tile copies of the axis tables cnu(:,:,i), cells -1 .. TX (x1) and -1 .. TY (x2) of the tile
   constexpr int KK = K * (K + 1);
   double *s_tab0 = smem + T::OFF_OPS + NSTG * T::OP_DOUBLES, *s_tab1 = s_tab0 + (TX + 2) * KK;
   (void)s_tab1;

   __shared__ __align__(8) unsigned long long s_bar[4]; // [0], [1]: tile buffer full; [2]: staged operands; [3]: staged tables (GEN)
   asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); // the next kernel may be scheduled as SMs free up
   const int tid = threadIdx.x;
   const int total_tiles = g.tiles_x * g.tiles_y;

   // one thread fetches a tile + frame (rows y0-H .. y0+TY+H-1, columns x0-H .. x0+TX+H-1) with ONE tensor-map TMA load;
   // the parts of the box beyond the padded state arrive as zeros
   auto issue = [&](int tile_id, int buf) {
      const int ty = tile_id / g.tiles_x, tx = tile_id - ty * g.tiles_x;
      mbar_expect_tx(&s_bar[buf], (uint32_t)(T::TILE_DOUBLES * sizeof(double)));
      tma_tensor_2d_g2s(smem_u32(s_v + buf * T::TILE_DOUBLES), &tm_v, tx * TX - H + PAD, ty * TY - H + PAD2, smem_u32(&s_bar[buf]));
   };
   // the pointwise operands of a tile (TX x TY boxes of a and b), one tensor-map TMA load each
   auto issue_ops = [&](int tile_id) {
      if constexpr (NSTG > 0) {
         const int ty = tile_id / g.tiles_x, tx = tile_id - ty * g.tiles_x;
         mbar_expect_tx(&s_bar[2], (uint32_t)(NSTG * T::OP_DOUBLES * sizeof(double)));
         tma_tensor_2d_g2s(smem_u32(s_ops), &tm_a, tx * TX + PAD, ty * TY + PAD2, smem_u32(&s_bar[2]));
         if constexpr (NSTG > 1) tma_tensor_2d_g2s(smem_u32(s_ops + T::OP_DOUBLES), &tm_b, tx * TX + PAD, ty * TY + PAD2, smem_u32(&s_bar[2]));
      }
   };

   // GEN: the tables of the tile's columns / rows (one ghost cell on either side), two TMA bulk copies; issued for the NEXT
   // tile right behind the phase barrier of this one (phase B does not read them), so they never sit in front of their use
   const bool has_tab = GEN && (g.cnu0 != nullptr || g.cnu1 != nullptr);
   auto issue_tabs = [&](int tile_id) {
      if constexpr (GEN) {
         const int ty = tile_id / g.tiles_x, tx = tile_id - ty * g.tiles_x;
         const int64_t x0 = (int64_t)tx * TX, y0 = (int64_t)ty * TY;
         const int64_t hx = x0 + TX < g.n0 ? x0 + TX : g.n0, hy = y0 + TY < g.n1 ? y0 + TY : g.n1; // last staged cell (<= n: the ghost table)
         const uint32_t b0 = g.cnu0 ? (uint32_t)(hx - x0 + 2) * KK * 8u : 0u, b1 = g.cnu1 ? (uint32_t)(hy - y0 + 2) * KK * 8u : 0u;
         mbar_expect_tx(&s_bar[3], b0 + b1);
         if (b0) tma_bulk_g2s(s_tab0, g.cnu0 + (x0 - 1) * KK, b0, &s_bar[3]);
         if (b1) tma_bulk_g2s(s_tab1, g.cnu1 + (y0 - 1) * KK, b1, &s_bar[3]);
      }
   };

   if (tid == 0) {
      mbar_init(&s_bar[0], 1);
      mbar_init(&s_bar[1], 1);
      mbar_init(&s_bar[2], 1);
      mbar_init(&s_bar[3], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __syncthreads();
   asm volatile("griddepcontrol.wait;" ::: "memory"); // state vectors written by earlier kernels are complete from here on
   if (tid == 0 && (int)blockIdx.x < total_tiles) {
      issue(blockIdx.x, 0);
      if (has_tab) issue_tabs(blockIdx.x);
   }

   int it_n = 0;
   for (int tile_id = blockIdx.x; tile_id < total_tiles; tile_id += gridDim.x, ++it_n) {
   const int buf = it_n & 1;
   // every thread is past phase B of the previous tile: its staged tile, operands, widths and vl/vr arrays may be overwritten
   __syncthreads();
   if (tid == 0) {
      issue_ops(tile_id);
      if (tile_id + (int)gridDim.x < total_tiles) issue(tile_id + gridDim.x, buf ^ 1);
   }
   const int ty_ = tile_id / g.tiles_x, tx_ = tile_id - ty_ * g.tiles_x;
   const int64_t x0 = (int64_t)tx_ * TX;
   const int64_t y0 = (int64_t)ty_ * TY;
   // per-tile copies of the widths (read by every thread in phase B, after the barrier below)
   for (int i = tid; i < TX + TY; i += NT) {
      if (i < TX) {
         const int64_t gx = x0 + i < g.n0 ? x0 + i : g.n0 - 1;
         s_wd[i] = __ldg(g.w1 + gx);
         s_wd[TX + i] = __ldg(g.rw1 + gx);
      } else {
         const int64_t gy = y0 + (i - TX) < g.n1 ? y0 + (i - TX) : g.n1 - 1;
         s_wd[2 * TX + (i - TX)] = __ldg(g.w2 + gy);
         s_wd[2 * TX + TY + (i - TX)] = __ldg(g.rw2 + gy);
      }
   }
   double *s_vt = s_v + buf * T::TILE_DOUBLES; // this tile
   mbar_wait(&s_bar[buf], (uint32_t)((it_n >> 1) & 1));
   if (has_tab) mbar_wait(&s_bar[3], (uint32_t)(it_n & 1));
   const double *tabx = (GEN && g.cnu0) ? s_tab0 + KK : nullptr; // table of the tile's cell 0 along x1 / x2
   const double *taby = (GEN && g.cnu1) ? s_tab1 + KK : nullptr;

   // ---- phase A: reconstruction items ------------------------------------------------------------------
   // x1-sweep: runs of R cells of a tile row -> shared memory (phase B reads them from other threads)
   constexpr int NXR = (TX / R) * TY, NYR = TX * (TY / R);
   constexpr int YI = NYR / NT; // x2-items (= phase-B items) per thread
   static_assert(NYR % NT == 0, "every thread owns the same number of x2-runs");
   if constexpr (GEN) {
      // general operators: a thread takes R rows of ONE column, so the column's table is read once per R cells (and by
      // consecutive lanes for consecutive columns); results land where phase B expects them
      for (int it = tid; it < TX * (TY / R); it += NT) {
         const int rb = it / TX, lx = it - rb * TX;
         double ci[KK];
         if (tabx) {
            const double2 *c2 = reinterpret_cast<const double2 *>(tabx + lx * KK);
#pragma unroll
            for (int q = 0; q < KK / 2; ++q) {
               const double2 t = c2[q];
               ci[2 * q] = t.x;
               ci[2 * q + 1] = t.y;
            }
         }
#pragma unroll
         for (int j = 0; j < R; ++j) {
            const int ly = rb * R + j;
            const double *cell = s_vt + (ly + H) * T::SP + (H + lx);
            double w5[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) w5[q] = cell[q - 2];
            double l, r;
            if (tabx)
               weno_cell_nonuniform<K, M>(ci, w5 + 2, g.kc.eps, l, r);
            else
               weno_run<K, 1, M>(w5 + (2 - (K - 1)), g.kc, &l, &r);
            s_vlx[ly * T::XP + R + lx] = l;
            s_vrx[ly * T::XP + R + lx] = r;
         }
      }
   }
   for (int it = tid; it < (GEN ? 0 : NXR); it += NT) {
      const int ly = it / (TX / R);
      const int rx = it - ly * (TX / R);
      const double *base = s_vt + (ly + H) * T::SP + (H + rx * R); // first cell of the run
      const int oidx = ly * T::XP + (rx + 1) * R;
      double w[R + 4];
#pragma unroll
      for (int j = 0; j < R + 4; j += 2) { // base - 2 is 16-B aligned: even tile pitch, even frame, runs of 4
         const double2 t = *reinterpret_cast<const double2 *>(base + j - 2);
         w[j] = t.x;
         w[j + 1] = t.y;
      }
      double vl[R], vr[R];
      run2d<K, R, M>(w, g.kc, nullptr, vl, vr);
#pragma unroll
      for (int j = 0; j < R; ++j) {
         if constexpr (!UPW) s_vlx[oidx + j] = vl[j]; // upwind: the left side is never used and is eliminated
         s_vrx[oidx + j] = vr[j];
      }
   }
   // halo items: the faces on the tile edge need vr of the cell just below/left of the tile and (unless upwind) vl of the
   // cell just above/right of it: single-cell reconstructions, grouped so that whole warps take them
   constexpr int NHALO = UPW ? (TY + TX) : 2 * (TY + TX);
   for (int it = tid; it < NHALO; it += NT) {
      const bool high = it >= TY + TX; // high side: vl of cell index T (only when !UPW)
      const int q = high ? it - (TY + TX) : it;
      const double *base;
      int stride;
      double *dst;
      const double *tab = nullptr; // GEN: the staged table of this cell
      if (q < TY) { // x1 line q
         const int cx = high ? TX : -1;
         base = s_vt + (q + H) * T::SP + (H + cx);
         stride = 1;
         dst = (high ? s_vlx : s_vrx) + q * T::XP + R + cx;
         if constexpr (GEN) tab = tabx ? tabx + cx * KK : nullptr;
      } else { // x2 line
         const int lx = q - TY;
         const int cy = high ? TY : -1;
         base = s_vt + (H + cy) * T::SP + (lx + H);
         stride = T::SP;
         dst = high ? s_vle + T::RY * TX + lx : s_vre + lx;
         if constexpr (GEN) tab = taby ? taby + cy * KK : nullptr;
      }
      double w[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) w[j] = base[(j - 2) * stride];
      double vl1[1], vr1[1];
      run2d<K, 1, M>(w, g.kc, tab, vl1, vr1);
      *dst = high ? vl1[0] : vr1[0];
   }
   // x2-sweep: item (column lx, run ry) is also this thread's phase-B item, so vl / vr stay in registers; the values the
   // neighbouring run needs (vr of its cell below, vl of its cell above) go through the exchange arrays
   double vlY[YI][R], vrY[YI][R];
#pragma unroll
   for (int q = 0; q < YI; ++q) {
      const int it = tid + q * NT;
      const int ry = it / TX;
      const int lx = it - ry * TX;
      const double *base = s_vt + (H + ry * R) * T::SP + (lx + H);
      double w[R + 4];
#pragma unroll
      for (int j = 0; j < R + 4; ++j) w[j] = base[(j - 2) * T::SP];
      run2d<K, R, M>(w, g.kc, taby ? taby + ry * R * KK : nullptr, vlY[q], vrY[q]);
      s_vre[(ry + 1) * TX + lx] = vrY[q][R - 1];
      if constexpr (!UPW) s_vle[ry * TX + lx] = vlY[q][0];
   }
   __syncthreads();
   if (has_tab && tid == 0 && tile_id + (int)gridDim.x < total_tiles) issue_tabs(tile_id + gridDim.x);
   if constexpr (NSTG > 0) mbar_wait(&s_bar[2], (uint32_t)(it_n & 1));

   // ---- phase B: fluxes, divergence, combination ---------------------------------------------------------
   const bool interior = !s.out_dense && x0 > 0 && x0 + TX < g.n0 && y0 > 0 && y0 + TY < g.n1;
#pragma unroll
   for (int q = 0; q < YI; ++q) {
      const int it = tid + q * NT;
      const int ry = it / TX;
      const int lx = it - ry * TX;
      double vmY[R + 1], vpY[R + 1];
      vmY[0] = s_vre[ry * TX + lx];
#pragma unroll
      for (int j = 0; j < R; ++j) {
         vmY[j + 1] = vrY[q][j];
         vpY[j] = UPW ? 0.0 : vlY[q][j];
      }
      vpY[R] = UPW ? 0.0 : s_vle[(ry + 1) * TX + lx];
      if (interior)
         fv2d_phase_b<K, COMBINE, M, UPW, TX, TY, true, GEN>(g, s, s_vt, s_vlx, s_vrx, s_wd, s_ops, vmY, vpY, x0, y0, lx, ry * R);
      else
         fv2d_phase_b<K, COMBINE, M, UPW, TX, TY, false, GEN>(g, s, s_vt, s_vlx, s_vrx, s_wd, s_ops, vmY, vpY, x0, y0, lx, ry * R);
   }
   } // tile loop
}

#ifndef HRW_TX2
#define HRW_TX2 64
#define HRW_TY2 32
#define HRW_NT2 256
#endif
constexpr int TX2 = HRW_TX2, TY2 = HRW_TY2, NT2 = HRW_NT2;   // large grids
constexpr int NT2F = 2 * HRW_NT2;                            // threads per tile in fast mode
constexpr int TX2G = 32, TY2G = 32;                          // general operators on large grids
constexpr int TX2S = 32, TY2S = 16, NT2S = 128; // small grids (e.g. example2's 250x250): enough tiles to occupy every SM

template <int K, int COMBINE, class M, int UPW, int TX2, int TY2, int NT2, int GEN = 0>
static int launch2d_t(Fv *fv, const Fv2dGeom &g, const StageArgs &a, cudaStream_t st) {
   using T = Tile2d<TX2, TY2, UPW>;
   constexpr int NSTG = fv2d_nstaged(COMBINE, UPW);
   constexpr size_t BYTES = T::bytes(NSTG) + (GEN ? (size_t)((TX2 + 2) + (TY2 + 2)) * K * (K + 1) * sizeof(double) : 0);
   auto kern = fv2d_stage_kernel<K, COMBINE, M, UPW, TX2, TY2, NT2, GEN>;
   static bool configured[64] = {}; // one flag per instantiation and device (function attributes are per device)
   int dev = 0;
   cudaGetDevice(&dev);
   if (!configured[dev & 63]) {
      HRW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BYTES));
      configured[dev & 63] = true;
   }
   // tensor maps of the stage input (tile + frame boxes) and of the staged operands (tile boxes)
   CUtensorMap tm_v, tm_a, tm_b;
   HRW_TRY(fv_tmap_2d(fv, a.vin, T::SP, T::SROWS, &tm_v));
   tm_a = tm_v;
   tm_b = tm_v;
   if (NSTG >= 1) HRW_TRY(fv_tmap_2d(fv, a.a, TX2, TY2, &tm_a));
   if (NSTG >= 2) HRW_TRY(fv_tmap_2d(fv, a.b, TX2, TY2, &tm_b));
   // persistent grid: SMs x resident CTAs of this instantiation, never more than there are tiles
   static int resident = 0;
   if (resident == 0) {
      int sms = 0, per_sm = 0;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT2, BYTES);
      resident = (sms > 0 ? sms : 148) * (per_sm > 0 ? per_sm : 1);
   }
   const int64_t tiles = (int64_t)g.tiles_x * g.tiles_y;
   // programmatic dependent launch, as for the 1D kernel (fv1d_inst.cu)
   cudaLaunchConfig_t cfg = {};
   cfg.gridDim = dim3((unsigned)(tiles < resident ? tiles : resident));
   cfg.blockDim = dim3(NT2);
   cfg.dynamicSmemBytes = BYTES;
   cfg.stream = st;
   cudaLaunchAttribute attr[1];
   attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
   attr[0].val.programmaticStreamSerializationAllowed = 1;
   cfg.attrs = attr;
   cfg.numAttrs = 1;
   HRW_CUDA(cudaLaunchKernelEx(&cfg, kern, g, a, tm_v, tm_a, tm_b));
   return HRWENO_OK;
}

template <int K, int COMBINE, class M, int UPW>
static int launch2d_u(Fv *fv, const Fv2dGeom &g, const StageArgs &a, cudaStream_t st) {
   if (g.small_tiles) return launch2d_t<K, COMBINE, M, UPW, TX2S, TY2S, NT2S>(fv, g, a, st);
   // fast mode: 512 threads per tile (one x1-run and one x2-run each, 64 registers, 32 resident warps per SM instead of 24):
   // +6 % on cfg4; strict mode needs its 112-128 registers and keeps 256 threads (profiles/r1_variant_sweeps.txt)
   if constexpr (!M::strict) return launch2d_t<K, COMBINE, M, UPW, TX2, TY2, NT2F>(fv, g, a, st);
   return launch2d_t<K, COMBINE, M, UPW, TX2, TY2, NT2>(fv, g, a, st);
}

// general operator: two-sided fluxes, per-cell tables / coefficients where present.  STRICT: the reference's operation order
// (bit-identical to the oracle).  FAST: division-light weights and FMA contraction (1e-12 normwise, like the uniform fast path).
template <int K, int COMBINE, class M>
static int launch2d_gen(Fv *fv, const Fv2dGeom &g, const StageArgs &a, cudaStream_t st) {
   if (g.small_tiles) return launch2d_t<K, COMBINE, M, 0, TX2S, TY2S, NT2S, 1>(fv, g, a, st);
   // 32x32 tiles, 256 threads: one column item and one x2-run per thread, ~66 KB of shared memory with the staged tables
   // (the 64x32 tile of the uniform path would need 119 KB and leave one CTA per SM)
   return launch2d_t<K, COMBINE, M, 0, TX2G, TY2G, NT2, 1>(fv, g, a, st);
}

template <int K, int COMBINE, class M>
static int launch2d(Fv *fv, const Fv2dGeom &g, const StageArgs &a, cudaStream_t st) {
   if (fv->general) return launch2d_gen<K, COMBINE, M>(fv, g, a, st);
   const bool upw = g.flux1.model == HRWENO_FLUX_LINEAR && g.flux1.scheme == HRWENO_SCHEME_GODUNOV && g.flux1.coef >= 0.0 && g.flux2.coef >= 0.0;
   return upw ? launch2d_u<K, COMBINE, M, 1>(fv, g, a, st) : launch2d_u<K, COMBINE, M, 0>(fv, g, a, st);
}

template <int K, class M>
static int launch2d_c(Fv *fv, int combine, const Fv2dGeom &g, const StageArgs &a, cudaStream_t st) {
   switch (combine) {
   case C_RHS: return launch2d<K, C_RHS, M>(fv, g, a, st);
   case C_EULER: return launch2d<K, C_EULER, M>(fv, g, a, st);
   case C_RK2_FINAL: return launch2d<K, C_RK2_FINAL, M>(fv, g, a, st);
   case C_RK3_S2: return launch2d<K, C_RK3_S2, M>(fv, g, a, st);
   case C_RK3_S3: return launch2d<K, C_RK3_S3, M>(fv, g, a, st);
   default: return launch2d<K, C_MS, M>(fv, g, a, st);
   }
}

template <class M>
static int launch2d_k(Fv *fv, int k, int combine, const Fv2dGeom &g, const StageArgs &a, cudaStream_t st) {
   if (k == 1) return launch2d_c<1, M>(fv, combine, g, a, st);
   if (k == 2) return launch2d_c<2, M>(fv, combine, g, a, st);
   return launch2d_c<3, M>(fv, combine, g, a, st);
}

int fv2d_stage(Fv *fv, int combine, const StageArgs &args, cudaStream_t st) {
   const hrweno_fv_desc &d = fv->d;
   Fv2dGeom g{};
   g.n0 = fv->n0;
   g.n1 = fv->n1;
   g.pitch = fv->pitch;
   const int txl = fv->general ? TX2G : TX2, tyl = fv->general ? TY2G : TY2; // large-grid tile of this operator
   g.tiles_x = (int)((fv->n0 + txl - 1) / txl);
   g.tiles_y = (int)((fv->n1 + tyl - 1) / tyl);
   g.small_tiles = (int64_t)g.tiles_x * g.tiles_y < 2 * 148; // fewer large tiles than two per SM: use the small ones
   if (g.small_tiles) {
      g.tiles_x = (int)((fv->n0 + TX2S - 1) / TX2S);
      g.tiles_y = (int)((fv->n1 + TY2S - 1) / TY2S);
   }
   g.w1 = fv->d_width[0];
   g.w2 = fv->d_width[1];
   g.rw1 = fv->d_rwidth[0];
   g.rw2 = fv->d_rwidth[1];
   g.kc = make_wenok(d.eps);
   g.flux1 = FluxCfg{d.flux_model, d.flux_scheme, d.flux_coef[0], d.alpha};
   g.flux2 = FluxCfg{d.flux_model, d.flux_scheme, d.flux_coef[1], d.alpha};
   g.bc = d.bc;
   g.phys_lo = d.rank == 0;
   g.phys_hi = d.rank == d.nranks - 1;
   if (fv->general) { // non-uniform grids / x- or t-dependent fluxes: the same tile kernel in the reference's operation order
      g.cnu0 = fv->d_cnu[0];
      g.cnu1 = fv->d_cnu[1];
      g.fc0 = fv->d_fcoef[0];
      g.fc1 = fv->d_fcoef[1];
      g.cc0 = fv->d_ccoef[0];
      g.cc1 = fv->d_ccoef[1];
      g.has_ts = fv->tfn != nullptr;
      g.ts = fv->tfn ? fv->tfn(fv->tfn_ctx, args.t) : 1.0;
   }
   if ((int64_t)g.tiles_x * g.tiles_y > 2000000000LL) return fail(HRWENO_EINVAL, "2D grid too large for one launch");
   if (d.mode == HRWENO_MODE_STRICT) return launch2d_k<Strict>(fv, d.k, combine, g, args, st);
   return launch2d_k<Fast>(fv, d.k, combine, g, args, st);
}

} // namespace hrw
