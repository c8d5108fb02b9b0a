// fvgen.cuh -- the general fused finite-volume stage (kernel K7) as a template on the value type.
// T = double: the fp64 library path (fvgen.cu).  T = float: the REAL32 build of the reference (src/hrweno_kinds.F90:9-17,
// `rk = real32`), where every real of the path -- state, widths, tables, eps, dt -- is single precision (real32.cu).
// Arithmetic is the reference's operation order in separately rounded IEEE operations of T (RT<T>), never contracted.
#pragma once

#include <type_traits>

#include "fv2d.cuh"
#include "internal.hpp"
#include "weno_core.cuh"

namespace hrw {

// reference-order arithmetic of the value type
template <class T>
struct RT;
template <>
struct RT<double> {
   __device__ __forceinline__ static double add(double a, double b) { return __dadd_rn(a, b); }
   __device__ __forceinline__ static double sub(double a, double b) { return __dsub_rn(a, b); }
   __device__ __forceinline__ static double mul(double a, double b) { return __dmul_rn(a, b); }
   __device__ __forceinline__ static double div(double a, double b) { return __ddiv_rn(a, b); }
   __device__ __forceinline__ static double fma_exact(double a, double b, double c) { return __fma_rn(a, b, c); }
};
template <>
struct RT<float> {
   __device__ __forceinline__ static float add(float a, float b) { return __fadd_rn(a, b); }
   __device__ __forceinline__ static float sub(float a, float b) { return __fsub_rn(a, b); }
   __device__ __forceinline__ static float mul(float a, float b) { return __fmul_rn(a, b); }
   __device__ __forceinline__ static float div(float a, float b) { return __fdiv_rn(a, b); }
   __device__ __forceinline__ static float fma_exact(float a, float b, float c) { return __fmaf_rn(a, b, c); }
};

template <class T>
struct FluxCfgT {
   int model, scheme;
   T coef, alpha;
};

// stage arguments of the value type (field for field StageArgs, internal.hpp)
template <class T>
struct StageArgsT {
   const T *vin, *a, *b;
   T *out, *out2;
   int64_t ld_out;
   int out_dense;
   T c0, c1;
};

// the uniform-grid tables c(j,r) (weno.f90:12-21) in the layout of one cell's cnu(:,:,i): ci[j + K*(r+1)], each entry the
// correctly rounded quotient of T (11.0_rk/6 etc.)
template <int K, class T>
__device__ __forceinline__ void uniform_table(T *ci) {
   if constexpr (K == 1) {
      ci[0] = T(1), ci[1] = T(1);
   } else if constexpr (K == 2) {
      const T c[6] = {T(3) / T(2), T(-1) / T(2), T(1) / T(2), T(1) / T(2), T(-1) / T(2), T(3) / T(2)};
      for (int q = 0; q < 6; ++q) ci[q] = c[q];
   } else {
      const T c[12] = {T(11) / T(6), T(-7) / T(6), T(1) / T(3),  T(1) / T(3),  T(5) / T(6),  T(-1) / T(6),
                       T(-1) / T(6), T(5) / T(6),  T(1) / T(3),  T(1) / T(3),  T(-7) / T(6), T(11) / T(6)};
      for (int q = 0; q < 12; ++q) ci[q] = c[q];
   }
}

// one cell in the reference's order with true divisions of T (weno.f90:174-216); ci[j + K*(r+1)] = c(j,r), ve at vext(i)
template <int K, class T>
__device__ __forceinline__ void weno_cell_reference(const T *ci, const T *ve, T eps, T &vl, T &vr) {
   using A = RT<T>;
   T vrr[K], vlr[K], beta[K];
#pragma unroll
   for (int r = 0; r < K; ++r) {
      T sr = A::mul(ci[K * (r + 1)], ve[-r]);
      T sl = A::mul(ci[K * r], ve[-r]);
#pragma unroll
      for (int j = 1; j < K; ++j) {
         sr = A::add(sr, A::mul(ci[j + K * (r + 1)], ve[-r + j]));
         sl = A::add(sl, A::mul(ci[j + K * r], ve[-r + j]));
      }
      vrr[r] = sr;
      vlr[r] = sl;
   }
   if constexpr (K == 1) {
      beta[0] = T(0);
   } else if constexpr (K == 2) {
      const T a = A::sub(ve[1], ve[0]), b = A::sub(ve[0], ve[-1]);
      beta[0] = A::mul(a, a);
      beta[K - 1] = A::mul(b, b);
   } else {
      const T c1312 = T(13) / T(12), c14 = T(1) / T(4);
      T a = A::add(A::sub(ve[0], A::mul(T(2), ve[1])), ve[2]);
      T b = A::add(A::sub(A::mul(T(3), ve[0]), A::mul(T(4), ve[1])), ve[2]);
      beta[0] = A::add(A::mul(c1312, A::mul(a, a)), A::mul(c14, A::mul(b, b)));
      a = A::add(A::sub(ve[-1], A::mul(T(2), ve[0])), ve[1]);
      b = A::sub(ve[-1], ve[1]);
      beta[1] = A::add(A::mul(c1312, A::mul(a, a)), A::mul(c14, A::mul(b, b)));
      a = A::add(A::sub(ve[-(K - 1)], A::mul(T(2), ve[-1])), ve[0]);
      b = A::add(A::sub(ve[-(K - 1)], A::mul(T(4), ve[-1])), A::mul(T(3), ve[0]));
      beta[K - 1] = A::add(A::mul(c1312, A::mul(a, a)), A::mul(c14, A::mul(b, b)));
   }
   // d1 = [1], d2 = [2/3, 1/3], d3 = [0.3, 0.6, 0.1] (weno.f90:12-14), literals of kind rk
   T d[K];
   if constexpr (K == 1) d[0] = T(1);
   if constexpr (K == 2) d[0] = T(2) / T(3), d[K - 1] = T(1) / T(3);
   if constexpr (K == 3) {
      if constexpr (std::is_same<T, float>::value)
         d[0] = 0.3f, d[1] = 0.6f, d[K - 1] = 0.1f;
      else
         d[0] = 0.3, d[1] = 0.6, d[K - 1] = 0.1;
   }
   T al[K], at[K];
#pragma unroll
   for (int r = 0; r < K; ++r) {
      const T e = A::add(eps, beta[r]);
      const T den = A::mul(e, e);
      al[r] = A::div(d[r], den);
      at[r] = A::div(d[K - 1 - r], den);
   }
   T s = al[0], st = at[0];
#pragma unroll
   for (int r = 1; r < K; ++r) {
      s = A::add(s, al[r]);
      st = A::add(st, at[r]);
   }
   T xr = A::mul(A::div(al[0], s), vrr[0]);
   T xl = A::mul(A::div(at[0], st), vlr[0]);
#pragma unroll
   for (int r = 1; r < K; ++r) {
      xr = A::add(xr, A::mul(A::div(al[r], s), vrr[r]));
      xl = A::add(xl, A::mul(A::div(at[r], st), vlr[r]));
   }
   vr = xr;
   vl = xl;
}

template <class T>
__device__ __forceinline__ T phys_flux_t(const FluxCfgT<T> &c, T v) {
   if (c.model == HRWENO_FLUX_BURGERS) return RT<T>::mul(RT<T>::mul(v, v), T(0.5)); // (v**2)/2, /2 is exact
   return RT<T>::mul(c.coef, v);
}

template <class T>
struct GenGeomT {
   int64_t n0, n1;  // cells along x1; rows (1D) or cells along x2 (2D)
   int64_t ld;      // pitch of the padded state vectors (vin, a, b, out2)
   int bc;
   const T *cnu0, *cnu1; // per-cell tables or nullptr (uniform tables)
   const T *w0, *w1;     // cell widths
   const T *fc0, *fc1;   // face coefficient along the axis, index 0..n like edges(0:n), or nullptr
   const T *cc0, *cc1;   // cross coefficient: cc0[j] for x1 faces of row j, cc1[i] for x2 faces of column i, or nullptr
   WenoK kc;             // fp64 path: eps and the uniform tables of weno_core.cuh
   T eps;
   FluxCfgT<T> fx0, fx1;
   int phys_l, phys_r; // the row ends are physical boundaries (else slab interfaces: ghost cells hold the neighbour's cells)
   int has_ts;  // the flux carries a time factor g(t) (fluxes.f90:12-18 passes t to the flux)
   T ts;        // its value at this evaluation's time, computed on the host
};

// reconstruct one cell (the arithmetic of recon_kernel, weno.cu).  `cell` points at the cell inside its row, the row is
// strided by `inc` (UNIT: inc == 1); `i` / `n` = its index / the row length, read only when CLAMP (the tile touches a
// domain edge); `has_tab` (the same for every thread of the launch, so the branch stays uniform) selects the cell's
// table cnu(:,:,i) at `ctab` or the uniform tables.
template <int K, bool UNIT, bool CLAMP, class T>
__device__ __forceinline__ void gen_recon(const T *cell, int64_t inc, int64_t i, int64_t n, bool has_tab, const T *ctab,
                                          const WenoK &kc, T eps, T &l, T &r, bool phys_l = true, bool phys_r = true) {
   int lo = -(K - 1), hi = K - 1;
   if constexpr (CLAMP) { // edge replicas (weno.f90:171-173): offsets clamped to the row at PHYSICAL ends, in 32 bits
      if (phys_l) lo = -(int)(i < K - 1 ? i : K - 1);
      if (phys_r) hi = (int)(n - 1 - i < K - 1 ? n - 1 - i : K - 1);
   }
   T w[2 * K - 1];
#pragma unroll
   for (int o = -(K - 1); o <= K - 1; ++o) {
      const int oo = CLAMP ? (o < lo ? lo : (o > hi ? hi : o)) : o;
      w[o + K - 1] = UNIT ? cell[oo] : cell[(int64_t)oo * inc];
   }
   if constexpr (std::is_same<T, double>::value) {
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
   } else { // REAL32: the reference's loop body with true divisions of T; uniform grids use c1/c2/c3 as every cell's table
      T ci[K * (K + 1)];
      if (has_tab) {
#pragma unroll
         for (int q = 0; q < K * (K + 1); ++q) ci[q] = __ldg(ctab + q);
      } else {
         uniform_table<K, T>(ci);
      }
      weno_cell_reference<K, T>(ci, w + (K - 1), eps, l, r);
   }
}

// f(v, x, t) = ((model(v)*cross)*face)*g(t), left to right like `v*x(1)*x(2)` (example2:153); an absent factor is not multiplied in
template <class T>
struct GenTs {
   bool has;
   T v;
};
template <class T>
__device__ __forceinline__ T gen_phys(const FluxCfgT<T> &c, T v, bool has_cc, T cc, bool has_fc, T fc, const GenTs<T> &ts) {
   T f = phys_flux_t<T>(c, v);
   if (has_cc) f = RT<T>::mul(f, cc);
   if (has_fc) f = RT<T>::mul(f, fc);
   if (ts.has) f = RT<T>::mul(f, ts.v);
   return f;
}

template <class T>
__device__ __forceinline__ T gen_face_flux(const FluxCfgT<T> &c, T vm, T vp, bool has_cc, T cc, bool has_fc, T fc, const GenTs<T> &ts) {
   using A = RT<T>;
   const T fm = gen_phys(c, vm, has_cc, cc, has_fc, fc, ts);
   const T fp = gen_phys(c, vp, has_cc, cc, has_fc, fc, ts);
   if (c.scheme == HRWENO_SCHEME_LAX_FRIEDRICHS) // (f(vm) + f(vp) - alpha*(vp - vm))/2      fluxes.f90:43
      return A::mul(A::sub(A::add(fm, fp), A::mul(c.alpha, A::sub(vp, vm))), T(0.5));
   const T lo = fm < fp ? fm : fp; // fluxes.f90:70-74
   const T hi = fm > fp ? fm : fp;
   return vm <= vp ? lo : hi;
}

// boundary rule on the two faces of cell i of a row of n cells: interior faces are given in fl (i > 0) and fr (i < n-1)
template <class T>
__device__ __forceinline__ void gen_bc(int bc, int64_t i, int64_t n, T &fl, T &fr, bool phys_l = true, bool phys_r = true) {
   const bool zero = bc == HRWENO_BC_ZERO_FLUX;
   if (i == 0 && phys_l) fl = zero ? T(0) : fr;     // fedges(0) = fedges(1) (example1:103) | 0 (example2:117,119); copy needs n >= 2
   if (i == n - 1 && phys_r) fr = zero ? T(0) : fl; // fedges(nc) = fedges(nc-1) (example1:104) | 0 (example2:118,120)
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
template <int K, bool TWO_D, bool INTERIOR, class T, class SA>
__device__ __forceinline__ void gen_tile(const GenGeomT<T> &g, const SA &a, const int combine, const int64_t i0, const int64_t j0,
                                         T *s_l1, T *s_r1, T *s_l2, T *s_r2) {
   using A = RT<T>;
   using GT = GenTile<TWO_D>;
   constexpr int TX = GT::TX, TY = GT::TY, SX = GT::SX, N1 = GT::N1, N2 = GT::N2, KK = K * (K + 1);
   const T *vt = a.vin + j0 * g.ld + i0; // cell (i0, j0)
   const bool has0 = g.cnu0 != nullptr, has1 = g.cnu1 != nullptr;
   const T *t0 = g.cnu0 + i0 * KK, *t1 = g.cnu1 + j0 * KK; // tables of cell i0 / j0 (dereferenced only when present)
   // ---- phase A: reconstruct the tile and its frame --------------------------------------------------------
   for (int q = threadIdx.x; q < N1 + N2; q += GEN_NT) {
      if (q < N1) { // N1 is a multiple of 32: the branch is warp-uniform
         const int iy = q / SX, ix = q - iy * SX - 1; // cell (i0 + ix, j0 + iy), ix = -1 .. TX
         const int64_t i = i0 + ix, j = j0 + iy;
         // at a slab interface the frame cell beyond the row end is the neighbour's edge cell (a ghost cell of the padded state)
         if (INTERIOR || (i >= (g.phys_l ? 0 : -1) && i < g.n0 + (g.phys_r ? 0 : 1) && j < g.n1)) {
            T l, r;
            gen_recon<K, true, !INTERIOR, T>(vt + (int64_t)iy * g.ld + ix, 1, i, g.n0, has0, t0 + ix * KK, g.kc, g.eps, l, r, g.phys_l != 0,
                                             g.phys_r != 0); // example1:93, example2:98 (contiguous row)
            s_l1[q] = l;
            s_r1[q] = r;
         }
      } else if constexpr (TWO_D) {
         const int p = q - N1;
         const int iy = p / TX - 1, ix = p - (iy + 1) * TX; // cell (i0 + ix, j0 + iy), iy = -1 .. TY
         const int64_t i = i0 + ix, j = j0 + iy;
         if (INTERIOR || (i < g.n0 && j >= 0 && j < g.n1)) {
            T l, r;
            gen_recon<K, false, !INTERIOR, T>(vt + (int64_t)iy * g.ld + ix, g.ld, j, g.n1, has1, t1 + iy * KK, g.kc, g.eps, l,
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
      const GenTs<T> ts{g.has_ts != 0, g.ts};
      const bool hc0 = g.cc0 != nullptr, hf0 = g.fc0 != nullptr;
      const T cc0 = hc0 ? g.cc0[j] : T(1);
      const int c1 = lx + 1 + ly * SX; // shared index of cell (i, j) in the x1 arrays
      T fl = T(0), fr = T(0);
      if (INTERIOR || i > 0 || !g.phys_l) fl = gen_face_flux(g.fx0, s_r1[c1 - 1], s_l1[c1], hc0, cc0, hf0, hf0 ? g.fc0[i] : T(1), ts);
      if (INTERIOR || i < g.n0 - 1 || !g.phys_r) fr = gen_face_flux(g.fx0, s_r1[c1], s_l1[c1 + 1], hc0, cc0, hf0, hf0 ? g.fc0[i + 1] : T(1), ts);
      if constexpr (!INTERIOR) gen_bc(g.bc, i, g.n0, fl, fr, g.phys_l != 0, g.phys_r != 0);
      T L = -A::div(A::sub(fr, fl), g.w0[i]); // -(fedges(i) - fedges(i-1))/width(i)   example1:107, example2:125
      if constexpr (TWO_D) {
         const bool hc1 = g.cc1 != nullptr, hf1 = g.fc1 != nullptr;
         const T cc1 = hc1 ? g.cc1[i] : T(1);
         const int c2 = lx + (ly + 1) * TX; // shared index of cell (i, j) in the x2 arrays
         T gl = T(0), gr = T(0);
         if (INTERIOR || j > 0) gl = gen_face_flux(g.fx1, s_r2[c2 - TX], s_l2[c2], hc1, cc1, hf1, hf1 ? g.fc1[j] : T(1), ts);
         if (INTERIOR || j < g.n1 - 1) gr = gen_face_flux(g.fx1, s_r2[c2], s_l2[c2 + TX], hc1, cc1, hf1, hf1 ? g.fc1[j + 1] : T(1), ts);
         if constexpr (!INTERIOR) gen_bc(g.bc, j, g.n1, gl, gr);
         L = A::sub(L, A::div(A::sub(gr, gl), g.w1[j])); // ... - (fedges2(j,i) - fedges2(j-1,i))/width2(j)   example2:126
      }
      // stage combination (tvdode.f90:141,149-167,257), the expressions of combine_kernel (ode.cu)
      const int64_t off = j * g.ld + i;
      const T x = a.vin[off];
      T o;
      switch (combine) {
      case C_RHS: o = L; break;
      case C_EULER: o = A::add(x, A::mul(a.c0, L)); break;
      case C_RK2_FINAL: o = A::mul(A::add(A::add(a.a[off], x), A::mul(a.c0, L)), T(0.5)); break;
      case C_RK3_S2: o = A::mul(A::add(A::add(A::mul(T(3), a.a[off]), x), A::mul(a.c0, L)), T(0.25)); break;
      case C_RK3_S3: { // (u + 2*ui + 2*dt*udot)/3: 2*ui exact; /3 a true division (fp64: the exact division of K2/K3, same bits)
         const T num = A::add(A::fma_exact(T(2), x, a.a[off]), A::mul(a.c0, L));
         if constexpr (std::is_same<T, double>::value)
            o = div3<Strict>(num);
         else
            o = A::div(num, T(3));
         break;
      }
      default: // C_MS
         o = A::mul(A::add(A::add(A::add(A::mul(T(25), x), A::mul(a.c0, L)), A::mul(T(7), a.a[off])), A::mul(a.c1, a.b[off])), T(0.03125));
         a.out2[off] = L; // out2 aliases b element for element: b[off] was read above
      }
      a.out[j * a.ld_out + i] = o; // out may alias a element for element (read above); it never aliases vin
   }
   __syncthreads(); // the next tile overwrites the shared arrays
}

template <int K, bool TWO_D, class T, class SA>
__global__ void __launch_bounds__(GEN_NT, gen_minb<TWO_D>()) fvgen_stage_kernel(const GenGeomT<T> g, const SA a, const int combine) {
   using GT = GenTile<TWO_D>;
   constexpr int TX = GT::TX, TY = GT::TY, N1 = GT::N1, N2 = GT::N2;
   __shared__ T s_l1[N1], s_r1[N1];
   __shared__ T s_l2[TWO_D ? N2 : 1], s_r2[TWO_D ? N2 : 1];
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
         gen_tile<K, false, true, T, SA>(g, a, combine, i0, j0, s_l1, s_r1, s_l2, s_r2);
      else
         gen_tile<K, TWO_D, false, T, SA>(g, a, combine, i0, j0, s_l1, s_r1, s_l2, s_r2);
   }
}


} // namespace hrw
