// fv1d.cuh -- fused 1D finite-volume stage kernel:
//   reconstruct (weno.f90:174-216) -> face fluxes (example1:97-104) -> -(df)/width (example1:107)
//   -> RK / multistep stage combination (tvdode.f90:141,149-167,257), one HBM pass per stage.
//
// Work decomposition.  A CTA of NT threads covers NT*R consecutive cells of one row; thread t owns
// the R cells i0 = c0 + (t-1)*R .. i0+R-1 and reconstructs vl, vr for them from a register window.
// Face f (between cells f-1 and f) needs vr(f-1) and vl(f); the two values that belong to the
// neighbouring threads travel through shared memory (one __syncthreads).  Threads 0 and NT-1 only
// provide those neighbour values (tiles overlap by one thread-run on each side: 2/NT redundant work)
// so no thread ever reconstructs a cell twice.
//
// The kernel is persistent (grid = SMs x resident CTAs); the tile of the next iteration is fetched by one
// TMA bulk copy (cp.async.bulk + mbarrier) into the other shared-memory buffer while this one computes.
//
// Measured on B200 (tools/microbench/fp64_issue.cu, profiles/): a warp-wide DFMA/DMUL/DADD holds the SMSP
// issue port for 2 cycles and every other instruction costs ~0.8 cycles on top, so this kernel is bound by
// instruction issue, not by HBM or latency.  Hence: the hot path is specialised at compile time (flux, width
// source), boundary handling lives in a cold path only edge threads take, the 1/2 of Burgers' flux and the
// sign of the divergence are folded into the (exact, power-of-two scaled) stage coefficient, and widths come
// from a small dictionary (value + refined reciprocal) indexed by one byte per cell.
#pragma once

#include "internal.hpp"
#include "weno_core.cuh"

#ifndef HRW_MINB
#define HRW_MINB 2 // resident CTAs per SM the register allocation is tuned for
#endif
#ifndef HRW_STAGE_A
#define HRW_STAGE_A 0 // 1: the pointwise operand a travels with the tile's bulk copies (measured slower: profiles/r1_variant_sweeps.txt)
#endif
#ifndef HRW_STAGE_W
#define HRW_STAGE_W 1 // 1: the width indices of the tile travel with the tile's bulk copies (0: global loads per thread)
#endif
#ifndef HRW_NBUF
#define HRW_NBUF 2 // staged-tile buffers of the TMA pipeline: tile it + NBUF - 1 is in flight while tile it is computed.
                   // Measured with 3 (profiles/r2o_k2_three_tile_buffers_ab.txt): no gain for k = 3 (0.794 vs 0.797) nor for the
                   // pure streaming k = 1 / rktvd1 stage (0.59 vs 0.60): the prefetch depth is not what limits either
#endif
#ifndef HRW_CST
// results leave through a warp-private staging area as coalesced stores (see CST in fv1d_stage_kernel): 0 never, 1 the
// k = 1 instantiations, 2 every k.  Measured on one box (profiles/r2ac_k1_coalesced_stores_ab.txt, r2ad_coalesced_stores_all_k.txt):
// k = 1 ensemble rktvd1 / 2 / 3 0.72 / 0.80 / 0.84 -> 0.86 / 0.92 / 0.93 of the HBM roofline; k = 3 (issue-bound) 0.801 -> 0.797 on one
// long row and 0.66 -> 0.64 on the ensemble, k = 2 no gain either: the ten extra shared-memory instructions per run cost what
// the store path gives back.  Results are bit-identical with and without (480 cases, tools/lib_checksums.py).
#define HRW_CST 1
#endif
#ifndef HRW_WIDX_LDS
#define HRW_WIDX_LDS HRW_CST // staged width indices are read with ld.shared instead of generic loads: 0 / 1 (k = 1) / 2 (every k)
#endif
#ifndef HRW_K1_NOX
// 1: the k = 1 instantiations run without the exchange barrier (see NOX in fv1d_stage_kernel).  Measured
// (profiles/r2aa_k1_no_exchange_barrier_ab.txt): rktvd1 +2 %, rktvd3 -1 % -- the barrier is not what limits k = 1; off.
#define HRW_K1_NOX 0
#endif
#ifndef HRW_WARP_TILES
#define HRW_WARP_TILES 0 // 1: every warp overlaps its neighbours by one thread run and exchanges by shuffles (no CTA barrier per
                         // tile).  Measured (profiles/r1_variant_sweeps.txt): the barrier stall disappears but 11 % more instructions
                         // (30 of 32 lanes useful) and late tile prefetches eat the gain; kept as an option, off.
#endif

namespace hrw {

enum FluxKind { FK_BURGERS_GODUNOV = 0, FK_GENERIC = 1 };
enum WidthKind { WK_DICT = 0, WK_ARRAY = 1 };

struct Fv1dGeom {
   int64_t n;             // cells per row
   int64_t ld;            // padded row pitch of vin / a / b / out2
   int64_t tiles_per_row;
   int64_t rows;          // independent rows (batched ensemble)
   int tile_begin, tile_end; // linear tile range (row-major over (row, tile)) this launch covers
   const double *width;   // WK_ARRAY: device array (padded, >= n + 16)
   const double2 *wtab;   // WK_DICT: 256 entries {width, refined reciprocal (exact_recip)}
   const unsigned char *widx; // WK_DICT: one byte per cell (padded, >= n + 16)
   WenoK kc;              // eps and the WENO tables (parameter-bank operands)
   FluxCfg flux;
   int bc;
   int phys_left, phys_right; // the row ends are physical boundaries (not slab interfaces)
   HaloIO halo;               // slab-interface halo traffic fused into this launch (all null on one GPU)
};

// ---- TMA bulk copy + mbarrier (PTX; SASS: UBLKCP / SYNCS) ----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// non-blocking phase test
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
   uint32_t done;
   asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
   return done != 0;
}
// arrival on a consumer-release barrier (release semantics), waited for with mbar_wait / mbar_test
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// overloads on 32-bit shared-window addresses: a persistent loop converts its barrier / buffer pointers once
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
   uint32_t done;
   do {
      asm volatile(
         "{\n"
         ".reg .pred p;\n"
         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
         "selp.u32 %0, 1, 0, p;\n"
         "}\n"
         : "=r"(done)
         : "r"(bar), "r"(parity)
         : "memory");
   } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src_gmem),
                "r"(bytes), "r"(bar)
                : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
   uint32_t done;
   do {
      asm volatile(
         "{\n"
         ".reg .pred p;\n"
         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
         "selp.u32 %0, 1, 0, p;\n"
         "}\n"
         : "=r"(done)
         : "r"(smem_u32(bar)), "r"(parity)
         : "memory");
   } while (!done);
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-B aligned), completion on `bar`
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar) {
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                : "memory");
}

// ---- numerical flux at one face ---------------------------------------------------------------------------
// FK_BURGERS_GODUNOV returns TWICE the reference's value: godunov (fluxes.f90:67-74) of f = (v**2)/2 (example1:120)
// is  vm<=vp ? min(fm,fp) : max(fm,fp)  and  min(x/2, y/2) = min(x,y)/2 exactly, so the selection is done on the
// squares and the exact factor 1/2 is folded into the stage coefficient by the caller.
template <int FK, class M>
__device__ __forceinline__ double face_flux_k(const FluxCfg &c, double vm, double vp) {
   if constexpr (FK == FK_BURGERS_GODUNOV) {
      // up: the smaller square, else the larger one.  Squaring is monotone in |v| (rounding included), so the selection
      // is made on |vm| < |vp| and only the chosen value is squared: one fp64 multiply per face instead of two, the
      // same bits (where the two squares round to the same value either choice gives it)
      const bool up = vm <= vp, lt = fabs(vm) < fabs(vp);
      const double q = (up == lt) ? vm : vp;
      // a separately rounded product in BOTH modes: left to the compiler, fast mode would contract q*q with the flux
      // difference into an FMA in the interior instantiation but not in the boundary one, and a cell's value would depend
      // on where the slab / tile edges fall (caught by the slab == single-domain bit-identity test of tests/mgpu_worker.py)
      return __dmul_rn(q, q);
   } else {
      return face_flux<M>(c, vm, vp);
   }
}

// one thread's share of a tile after the reconstruction: fluxes, divergence, combination, stores.
// EDGE = false: interior thread -- all R cells exist, no physical boundary touches its faces (hot path).
// EDGE = true : everything else (row ends, partial runs): boundary constraints, ghost cells, scalar accesses.
// operands an interior thread fetches from global memory at the top of the iteration, so that their latency
// hides behind the reconstruction
template <int R>
struct Prefetched {
   double av[R], bv[R];
   uint32_t idx4[(R + 3) / 4];
   double wd[R];
};

// the R one-byte width indices of a run starting at p (global or shared), packed four per word.  Runs of a multiple of
// four cells start on a 4-B boundary; other (even) run lengths start on a 2-B boundary and are assembled from the aligned
// words around them (the index arrays are padded, fv.cu).
// shared-memory accesses with the state space spelled out (a pointer that went through an integer cast compiles to a
// generic LD, which takes the long-scoreboard path: profiles/r2ab_k2_k1_euler_hotspots.txt)
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
   uint32_t v;
   asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
   return v;
}
__device__ __forceinline__ void sts_f64x2(uint32_t addr, double a, double b) {
   asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
   double2 v;
   asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
   return v;
}

// load_widx for a run whose indices were staged in shared memory (byte address `a` in the shared window)
template <int R>
__device__ __forceinline__ void load_widx_shared(uint32_t a, uint32_t *out) {
   constexpr int NW = (R + 3) / 4;
   constexpr int NL = R % 4 == 0 ? NW : (R + 2 + 3) / 4;
   const uint32_t q = a & ~3u, sh = (a & 3u) * 8u;
   uint32_t wd[NL + 1];
#pragma unroll
   for (int i = 0; i < NL; ++i) wd[i] = lds_u32(q + 4u * i);
   wd[NL] = 0u;
#pragma unroll
   for (int i = 0; i < NW; ++i) out[i] = R % 4 == 0 ? wd[i] : __funnelshift_r(wd[i], wd[i + 1 < NL ? i + 1 : NL], sh);
}

template <int R>
__device__ __forceinline__ void load_widx(const unsigned char *p, uint32_t *out) {
   constexpr int NW = (R + 3) / 4;
   if constexpr (R % 4 == 0) {
#pragma unroll
      for (int q = 0; q < NW; ++q) out[q] = reinterpret_cast<const uint32_t *>(p)[q];
   } else {
      static_assert(R % 2 == 0, "runs have an even number of cells");
      constexpr int NL = (R + 2 + 3) / 4;
      const uintptr_t a = reinterpret_cast<uintptr_t>(p);
      const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
      const uint32_t sh = ((uint32_t)a & 3u) * 8u;
      uint32_t wd[NL + 1];
#pragma unroll
      for (int i = 0; i < NL; ++i) wd[i] = q[i];
      wd[NL] = 0u;
#pragma unroll
      for (int i = 0; i < NW; ++i) out[i] = __funnelshift_r(wd[i], wd[i + 1 < NL ? i + 1 : NL], sh);
   }
}

// STAGE_OUT (interior threads only): the R results go to the shared-memory address `stage` (this thread's slot of its
// warp's staging area) instead of global memory; the warp writes them out together afterwards.
template <int K, int COMBINE, class M, int FK, int WK, int R, bool EDGE, bool STAGE_OUT = false>
__device__ __forceinline__ void fv1d_finish(const Fv1dGeom &g, const StageArgs &s, const double2 *s_wtab, int64_t row, int i0,
                                            const double *w /* window, cell j at w[2+j] */, const double *vl, const double *vr,
                                            double vr_left, double vl_right, double cL, double lscale, const Prefetched<R> &pf,
                                            uint32_t stage = 0u) {
   const int n = (int)g.n;
   // numerical flux at the R+1 faces i0 .. i0+R (face f lies between cells f-1 and f)
   double F[R + 1];
   F[0] = face_flux_k<FK, M>(g.flux, vr_left, vl[0]);
#pragma unroll
   for (int j = 1; j < R; ++j) F[j] = face_flux_k<FK, M>(g.flux, vr[j - 1], vl[j]);
   F[R] = face_flux_k<FK, M>(g.flux, vr[R - 1], vl_right);
   if constexpr (EDGE) {
      // problem-specific constraints at the domain boundaries (example1:103-104, example2:117-120)
      const bool copy = g.bc == HRWENO_BC_COPY_NEIGHBOUR;
      if (g.phys_left && i0 == 0) F[0] = copy ? F[1] : 0.0;
      if (g.phys_right) {
#pragma unroll
         for (int j = 1; j <= R; ++j)
            if (i0 + j == n) F[j] = copy ? F[j - 1] : 0.0;
      }
   }

   // pointwise operands and widths {w, refined reciprocal of w}
   double av[R], bv[R], wd[R], wr[R];
   constexpr bool NEED_A = COMBINE == C_RK2_FINAL || COMBINE == C_RK3_S2 || COMBINE == C_RK3_S3 || COMBINE == C_MS;
   uint32_t idx4[(R + 3) / 4];
   if constexpr (!EDGE) {
#pragma unroll
      for (int j = 0; j < R; ++j) {
         av[j] = pf.av[j];
         bv[j] = pf.bv[j];
         wd[j] = pf.wd[j];
      }
#pragma unroll
      for (int j = 0; j < (R + 3) / 4; ++j) idx4[j] = pf.idx4[j];
   } else {
      if constexpr (NEED_A) {
         const double *ap = s.a + row * g.ld + i0;
#pragma unroll
         for (int j = 0; j < R; ++j) av[j] = (i0 + j < n) ? ap[j] : 0.0;
         if constexpr (COMBINE == C_MS) {
            const double *bp = s.b + row * g.ld + i0;
#pragma unroll
            for (int j = 0; j < R; ++j) bv[j] = (i0 + j < n) ? bp[j] : 0.0;
         }
      }
      if constexpr (WK == WK_DICT) {
         load_widx<R>(g.widx + i0, idx4);
      } else {
#pragma unroll
         for (int j = 0; j < R; ++j) wd[j] = __ldg(g.width + i0 + j);
      }
   }
   if constexpr (WK == WK_DICT) {
#pragma unroll
      for (int j = 0; j < R; ++j) {
         const double2 e = s_wtab[(idx4[j / 4] >> (8 * (j % 4))) & 0xffu];
         wd[j] = e.x;
         wr[j] = e.y;
      }
   } else {
#pragma unroll
      for (int j = 0; j < R; ++j) wr[j] = M::strict ? exact_recip(wd[j]) : fast_rcp(wd[j]);
   }

   double res[R], lres[R];
   bool ok = true;
#pragma unroll
   for (int j = 0; j < R; ++j) {
      // vdot = -(fedges(i) - fedges(i-1))/width (example1:107); q = (df)/width here, sign and the flux's 1/2 are in cL/lscale
      const double dF = M::sub(F[j + 1], F[j]);
      double q;
      if constexpr (M::strict)
         q = exact_div_q(dF, wd[j], wr[j], ok);
      else
         q = dF * wr[j];
      const double v = w[2 + j];
      double o;
      constexpr bool FOLD = !M::strict && WK == WK_DICT && (COMBINE == C_EULER || COMBINE == C_RK2_FINAL || COMBINE == C_RK3_S2 || COMBINE == C_RK3_S3);
      if constexpr (FOLD) {
         // wr already carries the stage coefficient (dF*wr == dt*L), so the product rides in the combination's FMA
         if constexpr (COMBINE == C_EULER) o = fma(dF, wr[j], v);
         if constexpr (COMBINE == C_RK2_FINAL) o = fma(dF, wr[j], av[j] + v) * 0.5;
         if constexpr (COMBINE == C_RK3_S2) o = fma(0.75, av[j], 0.25 * fma(dF, wr[j], v));
         if constexpr (COMBINE == C_RK3_S3) o = div3<M>(fma(dF, wr[j], fma(2.0, v, av[j])));
      } else if constexpr (COMBINE == C_RHS) {
         o = M::mul(lscale, q);
      } else if constexpr (COMBINE == C_EULER) {
         o = M::add(v, M::mul(cL, q)); // u + dt*udot
      } else if constexpr (COMBINE == C_RK2_FINAL) {
         o = M::mul(M::add(M::add(av[j], v), M::mul(cL, q)), 0.5); // (u + ui + dt*udot)/2
      } else if constexpr (COMBINE == C_RK3_S2) {
         o = M::mul(M::add(M::add(M::mul(3.0, av[j]), v), M::mul(cL, q)), 0.25); // (3*u + ui + dt*udot)/4
      } else if constexpr (COMBINE == C_RK3_S3) {
         // (u + 2*ui + 2*dt*udot)/3 ; 2*ui is exact
         o = div3<M>(M::add(M::fma_exact(2.0, v, av[j]), M::mul(cL, q)));
      } else {
         // (25*u + 50*dt*udot + 7*uold(:,4) + 10*dt*udotold(:,4))/32 ; /32 is exact
         o = M::mul(M::add(M::add(M::add(M::mul(25.0, v), M::mul(cL, q)), M::mul(7.0, av[j])), M::mul(s.c1, bv[j])), 0.03125);
         lres[j] = M::mul(lscale, q);
      }
      res[j] = o;
   }
   if constexpr (M::strict) {
      if (!ok) { // a quotient in the denormal range: redo this run with the compiler's division (cold)
#pragma unroll
         for (int j = 0; j < R; ++j) {
            const double q = __ddiv_rn(__dsub_rn(F[j + 1], F[j]), wd[j]);
            const double v = w[2 + j];
            if constexpr (COMBINE == C_RHS) res[j] = __dmul_rn(lscale, q);
            if constexpr (COMBINE == C_EULER) res[j] = __dadd_rn(v, __dmul_rn(cL, q));
            if constexpr (COMBINE == C_RK2_FINAL) res[j] = __dmul_rn(__dadd_rn(__dadd_rn(av[j], v), __dmul_rn(cL, q)), 0.5);
            if constexpr (COMBINE == C_RK3_S2) res[j] = __dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(3.0, av[j]), v), __dmul_rn(cL, q)), 0.25);
            if constexpr (COMBINE == C_RK3_S3) res[j] = __ddiv_rn(__dadd_rn(__fma_rn(2.0, v, av[j]), __dmul_rn(cL, q)), 3.0);
            if constexpr (COMBINE == C_MS) {
               res[j] = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(25.0, v), __dmul_rn(cL, q)), __dmul_rn(7.0, av[j])), __dmul_rn(s.c1, bv[j])), 0.03125);
               lres[j] = __dmul_rn(lscale, q);
            }
         }
      }
   }

   double *orow = s.out + row * s.ld_out + i0;
   if constexpr (!EDGE) {
      if constexpr (STAGE_OUT) {
#pragma unroll
         for (int j = 0; j < R; j += 2) sts_f64x2(stage + 8u * j, res[j], res[j + 1]);
      } else {
#pragma unroll
         for (int j = 0; j < R; j += 2) *reinterpret_cast<double2 *>(orow + j) = make_double2(res[j], res[j + 1]);
      }
      if constexpr (COMBINE == C_MS) {
         double *lrow = s.out2 + row * g.ld + i0;
#pragma unroll
         for (int j = 0; j < R; j += 2) *reinterpret_cast<double2 *>(lrow + j) = make_double2(lres[j], lres[j + 1]);
      }
   } else {
#pragma unroll
      for (int j = 0; j < R; ++j)
         if (i0 + j < n) orow[j] = res[j];
      if constexpr (COMBINE == C_MS) {
         double *lrow = s.out2 + row * g.ld + i0;
#pragma unroll
         for (int j = 0; j < R; ++j)
            if (i0 + j < n) lrow[j] = lres[j];
      }
      // ghost cells of the result at physical boundaries: edge replicas (weno.f90:172-173)
      if (!s.out_dense && COMBINE != C_RHS) {
         if (g.phys_left && i0 == 0) {
#pragma unroll
            for (int q = 1; q <= K; ++q) orow[-q] = res[0];
         }
         if (g.phys_right) {
#pragma unroll
            for (int j = 0; j < R; ++j)
               if (i0 + j == n - 1) {
#pragma unroll
                  for (int q = 1; q <= K; ++q) orow[j + q] = res[j];
               }
         }
      }
   }
}

template <int K, int COMBINE, class M, int FK, int WK, int R, int NT>
__global__ void __launch_bounds__(NT, HRW_MINB) fv1d_stage_kernel(const Fv1dGeom g, const StageArgs s) {
   constexpr int P = 4;              // tile halo in shared memory (>= K-1, keeps 32-B alignment)
   // Thread runs per tile.  HRW_WARP_TILES: each warp covers 30 runs; its lanes 0 and 31 recompute the neighbouring warps'
   // edge runs (as threads 0 and NT-1 do for the neighbouring tiles), so vr/vl of the adjacent run arrive by warp shuffles
   // and the tile loop needs no CTA barrier: warps drift apart by up to one tile and hide each other's stalls.
   constexpr bool WT = HRW_WARP_TILES != 0;
   constexpr int WARPS = NT / 32;
   constexpr int RUNS = WT ? WARPS * 30 : NT - 2; // runs that write cells
   constexpr int TILE = RUNS * R;                 // cells written per CTA and tile
   constexpr int SM_N = (RUNS + 2) * R + 2 * P;
   constexpr int WN = R + 4;         // aligned register window (superset for K < 3)
   constexpr bool NEED_A = COMBINE == C_RK2_FINAL || COMBINE == C_RK3_S2 || COMBINE == C_RK3_S3 || COMBINE == C_MS;
   // the width indices of the tile travel with the tile's bulk copies (one iteration ahead), so interior threads read them
   // from shared memory instead of waiting for global loads (the same for operand a measured slower: HRW_STAGE_A)
   constexpr bool STAGE_A = NEED_A && HRW_STAGE_A;
   constexpr bool STAGE_W = WK == WK_DICT && HRW_STAGE_W;
   constexpr int WROW = ((TILE + 15) / 16) * 16 + 16 + (R % 4 ? 16 : 0); // staged index bytes per tile: from the 16-B boundary at or below its first cell
   constexpr int NBUF = WT ? 2 : HRW_NBUF; // (the warp-tile variant keeps its two "consumed" barriers)
   // k = 1: vl = vr = v (weno.f90:186), so the neighbouring run's edge values are cells of this thread's own register
   // window and nothing has to be exchanged between threads.  Without the exchange barrier the tile loop has no CTA
   // barrier at all: each warp releases the staged tile with one arrival on the buffer's "consumed" barrier once its
   // window is in registers (the protocol of the warp-tile variant), and the producer waits on that barrier before it
   // refills the buffer.  Warps then drift apart by up to one tile and cover each other's load and store phases, which is
   // all a stage without arithmetic has to hide.
   constexpr bool NOX = !WT && K == 1 && NBUF == 2 && HRW_K1_NOX != 0;
   constexpr bool REL = WT || NOX; // per-warp release of the tile buffers instead of a CTA barrier
   // k = 1 has no arithmetic to hide the store path under, and the direct stores are its worst part: a thread owns R
   // consecutive cells, so one warp-wide 16-B store touches 32 pieces 80 B apart (20 lines, 32 half-written sectors) and
   // the five of them keep the LSU busy five times longer than the 2560 contiguous bytes need; the next iteration then
   // waits for the stores to take their source registers (profiles/r2ab_k2_k1_euler_hotspots.txt: lg/mio throttle, 7.7 % of the
   // samples on that register hand-over).  CST: interior threads put their run into a warp-private staging area (16-B
   // shared stores at the same 80-B stride are conflict-free) and the warp writes its 2560 B as five fully coalesced 16-B
   // stores per lane.  Lanes on the edge path (row ends, dense output) keep their own scalar stores; their cells are
   // masked out of the warp's stores.
   constexpr bool CST = !WT && (HRW_CST == 2 || (HRW_CST == 1 && K == 1));
   static_assert(NBUF >= 2 && NBUF <= 4, "two to four staged tiles");
   __shared__ __align__(128) double s_v[NBUF][SM_N];
   __shared__ __align__(128) double s_a[STAGE_A ? NBUF : 1][STAGE_A ? TILE : 2];
   __shared__ __align__(16) unsigned char s_wi[STAGE_W ? NBUF : 1][STAGE_W ? WROW : 16];
   __shared__ double s_vr[REL ? 1 : 2][REL ? 1 : NT];
   __shared__ double s_vl[REL ? 1 : 2][REL ? 1 : NT];
   (void)s_vr;
   (void)s_vl;
   __shared__ __align__(16) double s_st[CST ? NT * R : 2]; // staging area: warp w owns [32*w*R, 32*(w+1)*R)
   (void)s_st;
   __shared__ __align__(16) double2 s_wtab[WK == WK_DICT ? 256 : 1];
   __shared__ __align__(8) unsigned long long s_bar[NBUF + 2]; // [0 .. NBUF): tile buffer full (TMA complete_tx); then two "consumed" barriers (warp tiles)

   // let the next kernel of the stream be scheduled as soon as SM resources free up (it waits for this grid to
   // complete before reading or writing global data, see griddepcontrol.wait below)
   asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
   int tid = threadIdx.x;
   asm volatile("" : "+r"(tid)); // keep the thread index in a register: re-reading SR_TID.X inside the tile loop stalls
   // Loop-invariant launch parameters the tile loop tests every iteration are pinned in registers: left to itself the
   // compiler re-reads them from the parameter bank right in front of each use, and those LDCU round trips showed up as
   // ~8 % of the stall samples (loop test, halo tests; profiles/r2b_k2_fast_source_hotspots.txt)
   const int n = (int)g.n;                     // cells per row (< 2^31, validated at creation)
   const int tpr = (int)g.tiles_per_row;
   int tile_end = g.tile_end;
   int nblk = (int)gridDim.x;
   int lflags = (g.halo.seq_in != 0 ? 1 : 0) | (g.halo.seq_out != 0 ? 2 : 0) | (s.out_dense ? 4 : 0);
   asm volatile("" : "+r"(tile_end), "+r"(nblk), "+r"(lflags));
   const bool halo_in = lflags & 1, halo_out = lflags & 2, out_dense = lflags & 4;

   const uint32_t bar_u32 = smem_u32(&s_bar[0]), sv_u32 = smem_u32(&s_v[0][0]);
   const uint32_t sa_u32 = smem_u32(&s_a[0][0]), sw_u32 = smem_u32(&s_wi[0][0]);
   const int lane = tid & 31;
   const int slot = WT ? (tid >> 5) * 30 + lane : tid; // run index within the staged tile (0 = the run left of the tile)
   // one elected thread issues the bulk copy of a tile: cells [c0-R-P, c0-R-P+SM_N) clipped to the padded row
   auto issue = [&](int row, int tcol, int buf) {
      const int ts = tcol * TILE - R - P;
      const int lo = ts < -PAD ? -PAD : ts;
      int hi = ts + SM_N;
      if (hi > (int)g.ld - PAD) hi = (int)g.ld - PAD;
      const uint32_t bytes = (uint32_t)(hi - lo) * (uint32_t)sizeof(double);
      const int c0 = tcol * TILE; // first cell the tile writes
      uint32_t abytes = 0;
      if constexpr (STAGE_A) {
         int ahi = c0 + TILE;
         if (ahi > (int)g.ld - PAD) ahi = (int)g.ld - PAD;
         abytes = (uint32_t)(ahi - c0) * (uint32_t)sizeof(double);
      }
      mbar_expect_tx(bar_u32 + 8u * buf, bytes + abytes + (STAGE_W ? (uint32_t)WROW : 0u));
      tma_bulk_g2s(sv_u32 + (uint32_t)(buf * SM_N + (lo - ts)) * 8u, s.vin + (int64_t)row * g.ld + lo, bytes, bar_u32 + 8u * buf);
      if constexpr (STAGE_A) tma_bulk_g2s(sa_u32 + (uint32_t)(buf * TILE) * 8u, s.a + (int64_t)row * g.ld + c0, abytes, bar_u32 + 8u * buf);
      if constexpr (STAGE_W) tma_bulk_g2s(sw_u32 + (uint32_t)(buf * WROW), g.widx + (c0 & ~15), (uint32_t)WROW, bar_u32 + 8u * buf);
   };

   for (int idx = tid; idx < NBUF * SM_N; idx += NT) (&s_v[0][0])[idx] = 0.0; // parts a clipped copy never writes
   // stage coefficient with the sign of the divergence and, for Burgers/Godunov, the flux's exact 1/2 folded in:
   // dt*L = dt*(-(1/2) q) = (-dt/2)*q bit for bit (power-of-two scaling commutes with rounding)
   const double lscale = FK == FK_BURGERS_GODUNOV ? -0.5 : -1.0;
   const double cL = lscale * s.c0;
   // fast mode with a width dictionary: the table carries cL/w, so (dt*L) is one multiply of the flux difference
   constexpr bool FOLD = !M::strict && WK == WK_DICT && (COMBINE == C_EULER || COMBINE == C_RK2_FINAL || COMBINE == C_RK3_S2 || COMBINE == C_RK3_S3);
   if constexpr (WK == WK_DICT)
      for (int idx = tid; idx < 256; idx += NT) {
         double2 e = g.wtab[idx];
         if constexpr (FOLD) e.y *= cL;
         s_wtab[idx] = e;
      }
   asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // order the generic-proxy zero fill before async-proxy writes
   if (tid == 0) {
      for (int b = 0; b < NBUF; ++b) mbar_init(&s_bar[b], 1);
      mbar_init(&s_bar[NBUF], WARPS);
      mbar_init(&s_bar[NBUF + 1], WARPS);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __syncthreads();

   // tiles are walked as (row, tile-in-row) pairs advanced by gridDim.x without any division in the loop
   const int step_rows = (int)((unsigned)nblk / (unsigned)tpr), step_cols = (int)((unsigned)nblk % (unsigned)tpr);
   // everything above touched only shared memory and creation-time constants; from here on this grid reads and writes
   // state vectors produced by earlier kernels of the stream
   asm volatile("griddepcontrol.wait;" ::: "memory");
   int lin = g.tile_begin + (int)blockIdx.x; // linear tile id; this launch owns [tile_begin, tile_end)
   int row = lin / tpr, tcol = lin % tpr;
   // slabs: the two edge tiles are walked first (logical tile 1 <-> last tile), so the boundary cells reach the
   // neighbour GPU while the bulk of the stage is still being computed
   auto remap = [&](int tc) { return (!g.halo.edge_first || tpr < 3) ? tc : (tc == 1 ? tpr - 1 : (tc == tpr - 1 ? 1 : tc)); };
   // (row, tile) of the tile PD iterations ahead: the prefetch cursor.  The first PD tiles are issued before the loop.
   constexpr int PD = NBUF - 1;
   int prow = row, ptcol = tcol;
   auto advance = [&](int &r, int &c) {
      r += step_rows;
      c += step_cols;
      if (c >= tpr) {
         c -= tpr;
         ++r;
      }
   };
   if (tid == 0) {
#pragma unroll
      for (int d = 0; d < PD; ++d) {
         if (lin + d * nblk < tile_end) issue(prow, remap(ptcol), d);
         advance(prow, ptcol);
      }
   } else {
#pragma unroll
      for (int d = 0; d < PD; ++d) advance(prow, ptcol);
   }

   int buf = 0, pbuf = PD % NBUF; // buffer of this iteration's tile / of the tile the prefetch of this iteration fills
   uint32_t bphase = 0;           // phase parity of this iteration's buffer (flips every time the ring wraps)
   for (int it = 0; lin < tile_end; ++it, lin += nblk) {
      int nrow = row, ntcol = tcol;
      advance(nrow, ntcol);
      // prefetch the next tile into the other buffer once every warp has taken the previous tile out of it: with the
      // exchange barrier, every thread passed last iteration's barrier after its loads; with warp tiles, each warp
      // arrives on the buffer's "consumed" barrier after its loads
      // The producer (thread 0) never blocks its warp early: it tests the "consumed" barrier here, again after the
      // reconstruction, and waits only at the end of the iteration.
      bool pending = tid == 0 && lin + PD * nblk < tile_end;
      auto try_issue = [&](bool block) {
         if (!pending) return;
         if (REL && it > 0) {
            const uint32_t eb = bar_u32 + 8u * NBUF + 8u * (buf ^ 1), ep = (uint32_t)(((it - 1) >> 1) & 1);
            if (block)
               mbar_wait(eb, ep);
            else if (!mbar_test(eb, ep))
               return;
         }
         issue(prow, remap(ptcol), pbuf);
         pending = false;
      };
      try_issue(false);

      const int ptc = remap(tcol);                // physical tile of this iteration
      const int i0 = ptc * TILE + (slot - 1) * R; // first owned cell (slot 0: the run left of the tile)
      // a caller's dense output array carries no alignment guarantee: scalar stores (edge path) for every thread
      const bool skip = (WT ? (lane == 0 || lane == 31) : (tid == 0 || tid == NT - 1)) || i0 >= n;
      const bool edge = out_dense || (i0 + R >= n) || (i0 == 0);

      // interior threads: start the global loads of the operands that are not staged with the tile now
      Prefetched<R> pf;
#pragma unroll
      for (int q = 0; q < (R + 3) / 4; ++q) pf.idx4[q] = 0u;
      if constexpr (STAGE_A) {
#pragma unroll
         for (int j = 0; j < R; ++j) pf.av[j] = 0.0;
      }
      if (!skip && !edge) {
         if constexpr (NEED_A && !STAGE_A) {
            const double *ap = s.a + (int64_t)row * g.ld + i0;
#pragma unroll
            for (int j = 0; j < R; j += 2) {
               const double2 t = __ldg(reinterpret_cast<const double2 *>(ap + j));
               pf.av[j] = t.x;
               pf.av[j + 1] = t.y;
            }
         }
         if constexpr (COMBINE == C_MS) {
            const double *bp = s.b + (int64_t)row * g.ld + i0;
#pragma unroll
            for (int j = 0; j < R; j += 2) {
               const double2 t = __ldg(reinterpret_cast<const double2 *>(bp + j));
               pf.bv[j] = t.x;
               pf.bv[j + 1] = t.y;
            }
         }
         if constexpr (WK == WK_DICT && !STAGE_W) load_widx<R>(g.widx + i0, pf.idx4);
         if constexpr (WK != WK_DICT) {
#pragma unroll
            for (int j = 0; j < R; j += 2) {
               const double2 t = __ldg(reinterpret_cast<const double2 *>(g.width + i0 + j));
               pf.wd[j] = t.x;
               pf.wd[j + 1] = t.y;
            }
         }
      }

      mbar_wait(bar_u32 + 8u * buf, bphase);

      // staged operands of this thread's run (overlap runs own no cells: their slots are never used)
      if (!skip && !edge) {
         if constexpr (STAGE_A) {
#pragma unroll
            for (int j = 0; j < R; j += 2) {
               const double2 t = *reinterpret_cast<const double2 *>(&s_a[buf][(slot - 1) * R + j]);
               pf.av[j] = t.x;
               pf.av[j + 1] = t.y;
            }
         }
         if constexpr (STAGE_W) {
            const int woff = (ptc * TILE) & 15;
            if constexpr (HRW_WIDX_LDS == 2 || (HRW_WIDX_LDS == 1 && K == 1))
               load_widx_shared<R>(sw_u32 + (uint32_t)(buf * WROW + woff + (slot - 1) * R), pf.idx4);
            else
               load_widx<R>(&s_wi[buf][woff + (slot - 1) * R], pf.idx4);
         }
      }

      // slab interface: the ghost cells of an edge tile come from this GPU's mailbox (stored there by the neighbour's
      // previous stage); wait for the sequence number, then patch them into the staged tile
      if (halo_in) {
         const bool rl = ptc == 0 && g.halo.recv_left, rr = ptc == tpr - 1 && g.halo.recv_right;
         if (rl || rr) { // CTA-uniform
            if (tid < 32) {
               if (tid == 0) {
                  const long long t0 = clock64();
                  for (int side = 0; side < 2; ++side) {
                     const unsigned long long *f = side == 0 ? (rl ? g.halo.rflag_left : nullptr) : (rr ? g.halo.rflag_right : nullptr);
                     if (!f) continue;
                     while (*reinterpret_cast<const volatile unsigned long long *>(f) < g.halo.seq_in) {
                        if (*reinterpret_cast<volatile unsigned int *>(g.halo.err) != 0) break;
                        if (clock64() - t0 > g.halo.timeout_cycles) {
                           atomicExch(g.halo.err, 1u);
                           break;
                        }
                     }
                  }
                  __threadfence_system();
               }
               __syncwarp();
               const int ts = ptc * TILE - R - P;
               if (rl && tid < K) s_v[buf][(-K + tid) - ts] = __ldcv(g.halo.recv_left + tid);
               if (rr && tid < K) s_v[buf][(n + tid) - ts] = __ldcv(g.halo.recv_right + tid);
            }
            __syncthreads();
         }
      }

      // register window: cells i0-2 .. i0+R+1
      double w[WN];
#pragma unroll
      for (int j = 0; j < WN; j += 2) {
         const double2 t = *reinterpret_cast<const double2 *>(&s_v[buf][slot * R + P - 2 + j]);
         w[j] = t.x;
         w[j + 1] = t.y;
      }
      if constexpr (REL) {
         // this warp has its part of the tile in registers (the empty asm makes the loads' results a dependency of what
         // follows): release the buffer, one arrival per warp
#pragma unroll
         for (int j = 0; j < WN; ++j) asm volatile("" : "+d"(w[j]));
         if constexpr (STAGE_W) {
#pragma unroll
            for (int q = 0; q < (R + 3) / 4; ++q) asm volatile("" : "+r"(pf.idx4[q]));
         }
         if constexpr (STAGE_A) {
#pragma unroll
            for (int j = 0; j < R; ++j) asm volatile("" : "+d"(pf.av[j]));
         }
         __syncwarp();
         if (lane == 0) mbar_arrive(bar_u32 + 8u * NBUF + 8u * buf);
      }
      double vl[R], vr[R];
      weno_run<K, R, M>(w + (2 - (K - 1)), g.kc, vl, vr);
      try_issue(false);

      double vr_left, vl_right;
      if constexpr (WT) {
         vr_left = __shfl_up_sync(0xffffffffu, vr[R - 1], 1);
         vl_right = __shfl_down_sync(0xffffffffu, vl[0], 1);
      } else if constexpr (NOX) {
         vr_left = w[1];      // cell i0-1: vr of the last cell of the run to the left
         vl_right = w[R + 2]; // cell i0+R: vl of the first cell of the run to the right
      } else {
         // exchange arrays are double buffered by iteration parity: a slot is rewritten two iterations later, after
         // the barrier of the iteration in between, which every reader of the old value has passed
         const int xb = it & 1;
         s_vr[xb][tid] = vr[R - 1];
         s_vl[xb][tid] = vl[0];
         __syncthreads();
         vr_left = s_vr[xb][tid > 0 ? tid - 1 : 0];
         vl_right = s_vl[xb][tid < NT - 1 ? tid + 1 : tid];
      }
      if constexpr (CST) {
         // (the ballot is also the point every lane passes before the staging area is written again: the reads of the
         // previous iteration are complete)
         const unsigned okm = __ballot_sync(0xffffffffu, !skip && !edge); // lanes that stage their run
         const uint32_t st_w = smem_u32(&s_st[(tid & ~31) * R]);
         if (!skip) {
            if (!edge)
               fv1d_finish<K, COMBINE, M, FK, WK, R, false, true>(g, s, s_wtab, row, i0, w, vl, vr, vr_left, vl_right, cL, lscale, pf,
                                                                  st_w + (uint32_t)(lane * R) * 8u);
            else
               fv1d_finish<K, COMBINE, M, FK, WK, R, true>(g, s, s_wtab, row, i0, w, vl, vr, vr_left, vl_right, cL, lscale, pf);
         }
         if (okm != 0u) { // warp-uniform
            __syncwarp();
            // the warp's runs are consecutive: lane l owns cells [i0w + l*R, i0w + (l+1)*R) of the row
            const int i0w = __shfl_sync(0xffffffffu, i0, 0);
            double *ow = s.out + (int64_t)row * s.ld_out + i0w;
#pragma unroll
            for (int q = 0; q < R / 2; ++q) {
               const int c = lane + 32 * q; // 16-B piece of the warp's span; it belongs to the run of lane c / (R/2)
               if ((okm >> (c / (R / 2))) & 1u) {
                  const double2 t = lds_f64x2(st_w + 16u * c);
                  *reinterpret_cast<double2 *>(ow + 2 * c) = t;
               }
            }
            __syncwarp(); // the staging area is rewritten in the next iteration
         }
      } else if (!skip) {
         if (!edge)
            fv1d_finish<K, COMBINE, M, FK, WK, R, false>(g, s, s_wtab, row, i0, w, vl, vr, vr_left, vl_right, cL, lscale, pf);
         else
            fv1d_finish<K, COMBINE, M, FK, WK, R, true>(g, s, s_wtab, row, i0, w, vl, vr, vr_left, vl_right, cL, lscale, pf);
      }
      try_issue(true);
      // slab interface: the CTA that just wrote the first / last k cells of the slab stores them into the neighbour's
      // mailbox over NVLink and publishes the sequence number (release: fence, then flag)
      if (halo_out) {
         const bool sl = ptc == 0 && g.halo.send_left, sr = ptc == tpr - 1 && g.halo.send_right;
         if (sl || sr) { // CTA-uniform
            __syncthreads(); // every store of this tile has been issued
            if (tid == 0) {
               const double *orow = s.out + (int64_t)row * s.ld_out;
               if (sl) {
#pragma unroll
                  for (int q = 0; q < K; ++q) g.halo.send_left[q] = __ldcg(orow + q);
               }
               if (sr) {
#pragma unroll
                  for (int q = 0; q < K; ++q) g.halo.send_right[q] = __ldcg(orow + (n - K + q));
               }
               __threadfence_system();
               if (sl) *reinterpret_cast<volatile unsigned long long *>(g.halo.sflag_left) = g.halo.seq_out;
               if (sr) *reinterpret_cast<volatile unsigned long long *>(g.halo.sflag_right) = g.halo.seq_out;
            }
         }
      }
      row = nrow;
      tcol = ntcol;
      advance(prow, ptcol);
      if (++buf == NBUF) {
         buf = 0;
         bphase ^= 1u;
      }
      if (++pbuf == NBUF) pbuf = 0;
   }
}

} // namespace hrw
