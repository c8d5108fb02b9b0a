// weno_core.cuh -- WENO reconstruction of a run of R consecutive cells held in registers.
//
// Restates weno_reconstruct (src/hrweno_weno.f90:174-216) for the uniform-grid tables
// (weno.f90:12-21).  A thread owns R consecutive cells and holds the window
//     w[0 .. R+2(K-1)-1] = vext(i0-(K-1)) .. vext(i0+R-1+(K-1))
// so that everything the reference recomputes per cell but that is *the same expression* for
// neighbouring cells is evaluated once (bit-identical by construction):
//   * the coefficient*value products c(j,r)*vext(m) (5 distinct coefficients for k=3),
//   * vlr(r)_i == vrr(r-1)_{i-1},
//   * the second-difference term 13/12*(a-2b+c)**2 shared by beta0_{i-1}, beta1_i, beta2_{i+1},
//   * (eps+beta_r)**2 shared by alfa_r and alfatilde_r, and alfa_1 == alfatilde_1 (d(1) both).
// sum() in the reference accumulates from 0; "0 + x" is dropped here (it can only turn -0 into +0).
#pragma once

#include "common.cuh"

namespace hrw {

template <int K, int R>
struct Window {
   static constexpr int G = K - 1;       // reconstruct ghost width (weno.f90:163)
   static constexpr int N = R + 2 * G;   // window length
};

// fallbacks with the compiler's own division (taken only when the range test of the strict branches fails:
// eps+beta >= 2^256); same formulas.  Returned by value {vl, vr} so nothing is forced onto the stack.
static __device__ __noinline__ double2 weno_weights_slow_k2(double den0, double den1, double vrr0, double vrr1, double vlr0, double vlr1) {
   const double d0 = 2.0 / 3, d1 = 1.0 / 3;
   const double al0 = __ddiv_rn(d0, den0), al1 = __ddiv_rn(d1, den1);
   const double at0 = __ddiv_rn(d1, den0), at1 = __ddiv_rn(d0, den1);
   const double s = __dadd_rn(al0, al1), st = __dadd_rn(at0, at1);
   double2 r;
   r.y = __dadd_rn(__dmul_rn(__ddiv_rn(al0, s), vrr0), __dmul_rn(__ddiv_rn(al1, s), vrr1));
   r.x = __dadd_rn(__dmul_rn(__ddiv_rn(at0, st), vlr0), __dmul_rn(__ddiv_rn(at1, st), vlr1));
   return r;
}

static __device__ __noinline__ double2 weno_weights_slow_k3(double den0, double den1, double den2, double vrr0, double vrr1, double vrr2,
                                                            double vlr0, double vlr1, double vlr2) {
   const double al0 = __ddiv_rn(0.3, den0), al1 = __ddiv_rn(0.6, den1), al2 = __ddiv_rn(0.1, den2);
   const double at0 = __ddiv_rn(0.1, den0), at2 = __ddiv_rn(0.3, den2);
   const double s = __dadd_rn(__dadd_rn(al0, al1), al2);
   const double st = __dadd_rn(__dadd_rn(at0, al1), at2);
   double2 r;
   r.y = __dadd_rn(__dadd_rn(__dmul_rn(__ddiv_rn(al0, s), vrr0), __dmul_rn(__ddiv_rn(al1, s), vrr1)), __dmul_rn(__ddiv_rn(al2, s), vrr2));
   r.x = __dadd_rn(__dadd_rn(__dmul_rn(__ddiv_rn(at0, st), vlr0), __dmul_rn(__ddiv_rn(al1, st), vlr1)), __dmul_rn(__ddiv_rn(at2, st), vlr2));
   return r;
}

// strict branches: every quotient of the weights stays in the range where the compiler's division takes its
// fast path (normal quotient, divisor < 2^1016) as long as eps+beta < 2^256 (eps > 2^-52 is validated at
// creation): den < 2^512, alfa > 2^-516, sum(alfa) <= 1/eps^2 < 2^105, w > 2^-621.  One integer test per cell.
__device__ __forceinline__ bool weights_in_range(double e0, double e1, double e2) {
   const uint32_t h0 = (uint32_t)__double2hiint(e0), h1 = (uint32_t)__double2hiint(e1), h2 = (uint32_t)__double2hiint(e2);
   return max(max(h0, h1), h2) < 0x4ff00000u;
}

// ---------------------------------------------------------------------------------------- k = 1
template <int R, class M>
__device__ __forceinline__ void weno_run_k1(const double *w, const WenoK &, double *vl, double *vr) {
   // vrr = 1*v, beta = 0, alfa = 1/eps**2, w = alfa/alfa = 1  ->  vl = vr = v   (weno.f90:186)
#pragma unroll
   for (int j = 0; j < R; ++j) {
      vl[j] = w[j];
      vr[j] = w[j];
   }
}

// ---------------------------------------------------------------------------------------- k = 2
template <int R, class M>
__device__ __forceinline__ void weno_run_k2(const double *w, const WenoK &kc, double *vl, double *vr) {
   const double eps = kc.eps;
   constexpr int N = R + 2;
   // c2 (weno.f90:17-18): c(:,-1) = [3/2,-1/2], c(:,0) = [1/2,1/2], c(:,1) = [-1/2,3/2]
   double h[N], q[N]; // h = v/2 (exact), q = 3/2*v
#pragma unroll
   for (int j = 0; j < N; ++j) {
      h[j] = M::mul(0.5, w[j]);
      q[j] = M::mul(1.5, w[j]);
   }
   double dsq[N - 1]; // (v[j+1]-v[j])**2   (weno.f90:190-191)
#pragma unroll
   for (int j = 0; j < N - 1; ++j) {
      const double d = M::sub(w[j + 1], w[j]);
      dsq[j] = M::mul(d, d);
   }
   double a0[R + 1]; // a0[j] = vrr(0) of cell j-1 (window index j) = vlr(1) of cell j
#pragma unroll
   for (int j = 0; j < R + 1; ++j) a0[j] = M::add(h[j], h[j + 1]);
   const double d0 = kc.d23, d1 = kc.d13;
#pragma unroll
   for (int j = 0; j < R; ++j) {
      const int c = j + 1; // window index of the cell
      const double vrr0 = a0[c];
      const double vrr1 = M::sub(q[c], h[c - 1]); // -1/2*v[c-1] + 3/2*v[c]
      const double vlr0 = M::sub(q[c], h[c + 1]); //  3/2*v[c] - 1/2*v[c+1]
      const double vlr1 = a0[c - 1];
      const double e0 = M::add(eps, dsq[c]), e1 = M::add(eps, dsq[c - 1]);
      const double den0 = M::mul(e0, e0), den1 = M::mul(e1, e1);
      if constexpr (M::strict) {
         const double r0 = exact_recip(den0), r1 = exact_recip(den1);
         const double al0 = exact_div_nc(d0, den0, r0), al1 = exact_div_nc(d1, den1, r1);
         const double at0 = exact_div_nc(d1, den0, r0), at1 = exact_div_nc(d0, den1, r1);
         const double s = M::add(al0, al1), st = M::add(at0, at1);
         const double rs = exact_recip(s), rst = exact_recip(st);
         vr[j] = M::add(M::mul(exact_div_nc(al0, s, rs), vrr0), M::mul(exact_div_nc(al1, s, rs), vrr1));
         vl[j] = M::add(M::mul(exact_div_nc(at0, st, rst), vlr0), M::mul(exact_div_nc(at1, st, rst), vlr1));
         if (!weights_in_range(e0, e1, e1)) {
            const double2 r = weno_weights_slow_k2(den0, den1, vrr0, vrr1, vlr0, vlr1);
            vl[j] = r.x;
            vr[j] = r.y;
         }
      } else {
         // alfa_r ~ d_r * prod_{s != r} den_s ; one reciprocal per side
         const double al0 = d0 * den1, al1 = d1 * den0;
         const double at0 = d1 * den1, at1 = d0 * den0;
         vr[j] = fma(al0, vrr0, al1 * vrr1) * rcp3(al0 + al1);
         vl[j] = fma(at0, vlr0, at1 * vlr1) * rcp3(at0 + at1);
      }
   }
}

// ---------------------------------------------------------------------------------------- k = 3
template <int R, class M>
__device__ __forceinline__ void weno_run_k3(const double *w, const WenoK &kc, double *vl, double *vr) {
   const double eps = kc.eps;
   constexpr int N = R + 4;
   // c3 (weno.f90:19-21): c(:,-1) = [11/6,-7/6,1/3], c(:,0) = [1/3,5/6,-1/6],
   //                      c(:,1) = [-1/6,5/6,1/3],   c(:,2) = [1/3,-7/6,11/6]
   const double C13 = kc.c13, C56 = kc.c56, C16 = kc.c16m, C76 = kc.c76m, C116 = kc.c116;
   double p13[N], p56[N], p16[N], p76[N], p116[N], t3[N];
#pragma unroll
   for (int j = 0; j < N; ++j) {
      p13[j] = M::mul(C13, w[j]);
      p56[j] = M::mul(C56, w[j]);
      p16[j] = M::mul(C16, w[j]);
      p76[j] = M::mul(C76, w[j]);
      p116[j] = M::mul(C116, w[j]);
      t3[j] = M::mul(3.0, w[j]);
   }
   // m2[j] = 13/12 * ((v[j-1] - 2 v[j]) + v[j+1])**2, j = 1..N-2   (weno.f90:195,198,201)
   double m2[N];
#pragma unroll
   for (int j = 1; j < N - 1; ++j) {
      const double d2 = M::add(M::fma_exact(-2.0, w[j], w[j - 1]), w[j + 1]);
      m2[j] = M::mul(kc.k1312, M::mul(d2, d2));
   }
   // A0[j] = vrr(0) of the cell at window index j = (1/3 v[j] + 5/6 v[j+1]) - 1/6 v[j+2]; also vlr(1) of cell j+1
   // A1[j] = vrr(1) of the cell at window index j = (-1/6 v[j-1] + 5/6 v[j]) + 1/3 v[j+1]; also vlr(2) of cell j+1
   double A0[N], A1[N];
#pragma unroll
   for (int j = 1; j < R + 2; ++j) {
      A0[j] = M::add(M::add(p13[j], p56[j + 1]), p16[j + 2]);
      A1[j] = M::add(M::add(p16[j - 1], p56[j]), p13[j + 1]);
   }
#pragma unroll
   for (int j = 0; j < R; ++j) {
      const int c = j + 2; // window index of the cell
      const double vrr0 = A0[c], vrr1 = A1[c];
      const double vrr2 = M::add(M::add(p13[c - 2], p76[c - 1]), p116[c]);
      const double vlr0 = M::add(M::add(p116[c], p76[c + 1]), p13[c + 2]);
      const double vlr1 = A0[c - 1], vlr2 = A1[c - 1];
      // beta (weno.f90:195-202); 1/4*x is exact so the final add may be fused
      const double b0 = M::add(M::fma_exact(-4.0, w[c + 1], t3[c]), w[c + 2]);
      const double b1 = M::sub(w[c - 1], w[c + 1]);
      const double b2 = M::add(M::fma_exact(-4.0, w[c - 1], w[c - 2]), t3[c]);
      const double beta0 = M::fma_exact(0.25, M::mul(b0, b0), m2[c + 1]);
      const double beta1 = M::fma_exact(0.25, M::mul(b1, b1), m2[c]);
      const double beta2 = M::fma_exact(0.25, M::mul(b2, b2), m2[c - 1]);
      const double e0 = M::add(eps, beta0), e1 = M::add(eps, beta1), e2 = M::add(eps, beta2);
      const double den0 = M::mul(e0, e0), den1 = M::mul(e1, e1), den2 = M::mul(e2, e2);
      if constexpr (M::strict) {
         // alfa = d/(eps+beta)**2, alfatilde = d(k-1:0:-1)/(eps+beta)**2   (weno.f90:207-208), d3 = [0.3,0.6,0.1]
         // IEEE quotients with the reciprocal refinement shared per denominator (common.cuh: exact_div)
         const double r0 = exact_recip(den0), r1 = exact_recip(den1), r2 = exact_recip(den2);
         const double al0 = exact_div_nc(kc.d03, den0, r0), al1 = exact_div_nc(kc.d06, den1, r1), al2 = exact_div_nc(kc.d01, den2, r2);
         const double at0 = exact_div_nc(kc.d01, den0, r0), at2 = exact_div_nc(kc.d03, den2, r2); // at1 == al1
         const double s = M::add(M::add(al0, al1), al2);
         const double st = M::add(M::add(at0, al1), at2);
         const double rs = exact_recip(s), rst = exact_recip(st);
         // w = alfa/sum(alfa); vr = sum(w*vrr)   (weno.f90:209-214)
         vr[j] = M::add(M::add(M::mul(exact_div_nc(al0, s, rs), vrr0), M::mul(exact_div_nc(al1, s, rs), vrr1)),
                        M::mul(exact_div_nc(al2, s, rs), vrr2));
         vl[j] = M::add(M::add(M::mul(exact_div_nc(at0, st, rst), vlr0), M::mul(exact_div_nc(al1, st, rst), vlr1)),
                        M::mul(exact_div_nc(at2, st, rst), vlr2));
         if (!weights_in_range(e0, e1, e2)) {
            const double2 r = weno_weights_slow_k3(den0, den1, den2, vrr0, vrr1, vrr2, vlr0, vlr1, vlr2);
            vl[j] = r.x;
            vr[j] = r.y;
         }
      } else {
         // division-light weights: alfa_r ~ d_r * prod_{s != r} den_s, one reciprocal per side
         const double p0 = den1 * den2, p1 = den0 * den2, p2 = den0 * den1;
         const double al0 = 0.3 * p0, al1 = 0.6 * p1, al2 = 0.1 * p2;
         const double at0 = 0.1 * p0, at2 = 0.3 * p2;
         const double s = (al0 + al1) + al2, st = (at0 + al1) + at2;
         vr[j] = fma(al2, vrr2, fma(al1, vrr1, al0 * vrr0)) * fast_rcp(s);
         vl[j] = fma(at2, vlr2, fma(al1, vlr1, at0 * vlr0)) * fast_rcp(st);
      }
   }
}

// ------------------------------------------------------------------------------------ k = 3, fast mode
// Same scheme, fewest fp64 instructions (the kernel is bound by instruction issue).
// Everything is written in differences of the cell averages, d1[j] = v[j+1] - v[j], d2[j] = d1[j] - d1[j-1],
// D3[j] = d2[j+1] - d2[j] (each one instruction, shared by neighbouring cells):
//   * smoothness indicators scaled by 4 (beta' = 4*beta, eps' = 4*eps leaves every weight ratio unchanged):
//       beta'_0 = (d1[c+1] - 3 d1[c])^2 + (13/3) d2[c+1]^2,   beta'_1 = (d1[c-1] + d1[c])^2 + (13/3) d2[c]^2,
//       beta'_2 = (d1[c-2] - 3 d1[c-1])^2 + (13/3) d2[c-1]^2                                    (weno.f90:195-202)
//   * the three candidates of a side differ by multiples of a third difference (c3, weno.f90:19-21):
//       vrr1 = v[c] + (d1[c-1] + 2 d1[c])/6,   vrr0 = vrr1 - D3[c]/6,   vrr2 = vrr1 - D3[c-1]/3,
//       vlr1 = vrr0 of cell c-1,   vlr2 = vlr1 + D3[c-1]/6,   vlr0 = vlr1 + D3[c]/3
//   * division-light weights alfa_r ~ d_r * P_r with P_r = prod_{s != r} (eps'+beta'_s)^2, d = (3,6,1)/10
//     (weno.f90:207-214), so the convex combination is the central candidate plus a weighted sum of two third
//     differences, one reciprocal per side (MUFU seed + one cubic step):
//       vr = vrr1 - (1.5 X + Y)/(9 P0 + 18 P1 + 3 P2),   vl = vlr1 + (X + 1.5 Y)/(3 P0 + 18 P1 + 9 P2),
//       X = P0*D3[c],  Y = P2*D3[c-1].
// Not bit-identical to the reference order (and less cancellation than it): parity is the north-star tolerance
// (1e-12 normwise per output time, a few ULP per reconstruction; oracle/np_oracle.py: reconstruct_fast restates it).
template <int R>
__device__ __forceinline__ void weno_run_k3_fast(const double *w, const WenoK &kc, double *vl, double *vr) {
   constexpr int N = R + 4;
   const double C13 = kc.c13, C16 = kc.c16;
   const double eps4 = kc.eps4;
   double d1[N], d2[N], m2e[N], D3[N];
#pragma unroll
   for (int j = 0; j < N - 1; ++j) d1[j] = w[j + 1] - w[j];
#pragma unroll
   for (int j = 1; j < N - 1; ++j) {
      d2[j] = d1[j] - d1[j - 1];
      m2e[j] = fma(kc.k133 * d2[j], d2[j], eps4);
   }
#pragma unroll
   for (int j = 1; j < N - 2; ++j) D3[j] = d2[j + 1] - d2[j];
   // V1[j] = vrr1 of the cell at window index j, V0[j] = its vrr0 (= vlr1 of cell j+1)
   double V1[N], V0[N];
#pragma unroll
   for (int j = 1; j < R + 2; ++j) {
      V1[j] = fma(C13, d1[j], fma(C16, d1[j - 1], w[j]));
      V0[j] = fma(-C16, D3[j], V1[j]);
   }
#pragma unroll
   for (int j = 0; j < R; ++j) {
      const int c = j + 2;
      const double b0 = fma(-3.0, d1[c], d1[c + 1]);
      const double b1 = d1[c - 1] + d1[c];
      const double b2 = fma(-3.0, d1[c - 1], d1[c - 2]);
      const double e0 = fma(b0, b0, m2e[c + 1]);
      const double e1 = fma(b1, b1, m2e[c]);
      const double e2 = fma(b2, b2, m2e[c - 1]);
      const double f0 = e1 * e2, f1 = e0 * e2, f2 = e0 * e1;
      const double P0 = f0 * f0, P2 = f2 * f2;
      const double Q = (18.0 * f1) * f1;
      const double X = P0 * D3[c], Y = P2 * D3[c - 1];
      const double A = fma(3.0, P0 + P2, Q);
      const double sr = fma(6.0, P0, A), sl = fma(6.0, P2, A);
      vr[j] = fma(-fma(1.5, X, Y), rcp3(sr), V1[c]);
      vl[j] = fma(fma(1.5, Y, X), rcp3(sl), V0[c - 1]);
   }
}

// ------------------------------------------------------------------------------------ k = 2, fast mode
// Difference form of the third-order scheme (c2, d2 of weno.f90:13,17-18):
//   vrr0 = v[c] + d1[c]/2,  vrr1 = vrr0 - d2[c]/2,  alfa ~ (2 den1, den0):  vr = vrr0 - (den0*d2[c]/2)/(2 den1 + den0)
//   vlr1 = v[c] - d1[c-1]/2, vlr0 = vlr1 - d2[c]/2, alfat ~ (den1, 2 den0): vl = vlr1 - (den1*d2[c]/2)/(den1 + 2 den0)
// with den0 = (eps + d1[c]^2)^2, den1 = (eps + d1[c-1]^2)^2 (weno.f90:190-191,207-214).
template <int R>
__device__ __forceinline__ void weno_run_k2_fast(const double *w, const WenoK &kc, double *vl, double *vr) {
   constexpr int N = R + 2;
   double d1[N], den[N];
#pragma unroll
   for (int j = 0; j < N - 1; ++j) {
      d1[j] = w[j + 1] - w[j];
      const double e = fma(d1[j], d1[j], kc.eps);
      den[j] = e * e;
   }
#pragma unroll
   for (int j = 0; j < R; ++j) {
      const int c = j + 1;
      const double h = 0.5 * (d1[c] - d1[c - 1]); // d2[c]/2
      const double den0 = den[c], den1 = den[c - 1];
      vr[j] = fma(-(den0 * h), rcp3(fma(2.0, den1, den0)), fma(0.5, d1[c], w[c]));
      vl[j] = fma(-(den1 * h), rcp3(fma(2.0, den0, den1)), fma(-0.5, d1[c - 1], w[c]));
   }
}

// cold path of the magnitude guard below: the same fast formulas on the window scaled to magnitude ~1.  Not inlined and
// called with scalars only (window in, results out through shared... no: through registers by value in small structs), so
// the hot path keeps its registers.
template <int N>
struct WinPack {
   double v[N];
};
template <int R>
struct ResPack {
   double vl[R], vr[R];
};
template <int K, int R>
static __device__ __noinline__ ResPack<R> weno_run_fast_scaled_impl(WinPack<R + 2 * (K - 1)> win, double eps, int hi_max) {
   constexpr int N = R + 2 * (K - 1);
   // s = 2^(40 - exponent of the largest |v|): exact scaling to magnitude ~2^40 (>= 2^100 before), which keeps the products
   // of the fast formulas below 2^400 and the scaled eps = eps*s^2 representable (not clamped) for |v| < ~1e45
   const int e = ((hi_max >> 20) & 0x7ff) - 1023 - 40;
   const double sdn = __hiloint2double((1023 - e) << 20, 0), sup = __hiloint2double((1023 + e) << 20, 0);
   WenoK kc = make_wenok(fmax(eps * sdn * sdn, 0x1p-240));
   double w[N];
#pragma unroll
   for (int j = 0; j < N; ++j) w[j] = win.v[j] * sdn;
   ResPack<R> r;
   if constexpr (K == 2)
      weno_run_k2_fast<R>(w, kc, r.vl, r.vr);
   else
      weno_run_k3_fast<R>(w, kc, r.vl, r.vr);
#pragma unroll
   for (int j = 0; j < R; ++j) {
      r.vl[j] *= sup;
      r.vr[j] *= sup;
   }
   return r;
}
template <int K, int R>
__device__ __forceinline__ void weno_run_fast_scaled(const double *w, const WenoK &kc, float mhi, double *vl, double *vr) {
   constexpr int N = R + 2 * (K - 1);
   WinPack<N> win;
#pragma unroll
   for (int j = 0; j < N; ++j) win.v[j] = w[j];
   const ResPack<R> r = weno_run_fast_scaled_impl<K, R>(win, kc.eps, __float_as_int(mhi));
#pragma unroll
   for (int j = 0; j < R; ++j) {
      vl[j] = r.vl[j];
      vr[j] = r.vr[j];
   }
}

// ---- magnitude guard of the fast formulas ---------------------------------------------------------------------------
// The division-light weights multiply smoothness terms: P = ((eps'+beta'_a)(eps'+beta'_b))^2 grows like v^8 and X = P*D3 like
// v^9, where the reference only forms (eps+beta)^2 ~ v^4.  With |v| < 2^100 every intermediate stays below 2^930 (e <= 85 v^2,
// P <= 5.2e7 v^8, X <= 4.2e8 v^9, the sums <= 1.6e9 v^8), far from overflow and from the flush-to-zero range of the
// reciprocal seed.  A run whose window holds a larger value is reconstructed from the window scaled by a power of two
// (exact; the scheme is homogeneous: vl, vr(s*v; s^2*eps) = s*vl, vr(v; eps)) in a second pass through the same code, so
// fast mode stays finite wherever the reference does (the reference overflows from |v| ~ 1e77 on).  At those magnitudes
// the window is scaled to ~2^40 and eps by the square of the factor; the scaled eps is kept away from underflow
// (>= 2^-240, which only binds for |v| > ~1e45 and only matters for cells whose differences are 2^-160 of that).
// The test compares the high words of the doubles as floats (monotone in |v|; fp32 min/max, off the fp64 pipe): the float
// patterns of finite doubles below 2^1017 are ordinary floats, beyond that the reference is not finite either.
template <int N>
__device__ __forceinline__ float fast_window_maxhi(const double *w) {
   float m = 0.0f;
#pragma unroll
   for (int j = 0; j < N; ++j) m = fmaxf(m, fabsf(__int_as_float(__double2hiint(w[j]))));
   return m;
}
constexpr int FAST_RANGE_HI = 0x46300000; // high word of 2^100 = (1023 + 100) << 20

template <int K, int R, class M>
__device__ __forceinline__ void weno_run(const double *w_in, const WenoK &kc, double *vl, double *vr) {
   if constexpr (K == 1) {
      weno_run_k1<R, M>(w_in, kc, vl, vr);
   } else if constexpr (!M::strict) {
      constexpr int N = R + 2 * (K - 1);
      const float mhi = fast_window_maxhi<N>(w_in);
#ifdef HRW_NO_FAST_GUARD // tuning A/B only
      if (true) {
#else
      if (mhi < __int_as_float(FAST_RANGE_HI)) { // always, for data of ordinary magnitude
#endif
         if constexpr (K == 2)
            weno_run_k2_fast<R>(w_in, kc, vl, vr);
         else
            weno_run_k3_fast<R>(w_in, kc, vl, vr);
      } else {
         weno_run_fast_scaled<K, R>(w_in, kc, mhi, vl, vr);
      }
   } else if constexpr (K == 2) {
      weno_run_k2<R, M>(w_in, kc, vl, vr);
   } else {
      weno_run_k3<R, M>(w_in, kc, vl, vr);
   }
}

// ------------------------------------------------------------------------------------------------
// One cell with per-cell coefficients (non-uniform grid: ci = cnu(:,:,i), weno.f90:177).  Reference order; in strict
// mode the weight quotients share their reciprocal refinements (common.cuh), which leaves every bit unchanged.
// ci[j + K*(r+1)] = c(j,r); ve points at vext(i).
// ------------------------------------------------------------------------------------------------
template <int K, class M>
__device__ __forceinline__ void weno_cell_nonuniform_core(const double *ci, const double *ve /* [-(K-1)..K-1] */,
                                                          double eps, double &vl, double &vr) {
   double vrr[K], vlr[K], beta[K];
#pragma unroll
   for (int r = 0; r < K; ++r) {
      double sr = M::mul(ci[K * (r + 1)], ve[-r]);
      double sl = M::mul(ci[K * r], ve[-r]);
#pragma unroll
      for (int j = 1; j < K; ++j) {
         sr = M::mad(ci[j + K * (r + 1)], ve[-r + j], sr);
         sl = M::mad(ci[j + K * r], ve[-r + j], sl);
      }
      vrr[r] = sr;
      vlr[r] = sl;
   }
   if constexpr (K == 1) {
      beta[0] = 0.0;
   } else if constexpr (K == 2) {
      const double a = M::sub(ve[1], ve[0]), b = M::sub(ve[0], ve[-1]);
      beta[0] = M::mul(a, a);
      beta[K - 1] = M::mul(b, b);
   } else {
      double a = M::add(M::fma_exact(-2.0, ve[1], ve[0]), ve[2]);
      double b = M::add(M::fma_exact(-4.0, ve[1], M::mul(3.0, ve[0])), ve[2]);
      beta[0] = M::fma_exact(0.25, M::mul(b, b), M::mul(13.0 / 12, M::mul(a, a)));
      a = M::add(M::fma_exact(-2.0, ve[0], ve[-1]), ve[1]);
      b = M::sub(ve[-1], ve[1]);
      beta[1] = M::fma_exact(0.25, M::mul(b, b), M::mul(13.0 / 12, M::mul(a, a)));
      a = M::add(M::fma_exact(-2.0, ve[-1], ve[-(K - 1)]), ve[0]);
      b = M::add(M::fma_exact(-4.0, ve[-1], ve[-(K - 1)]), M::mul(3.0, ve[0]));
      beta[K - 1] = M::fma_exact(0.25, M::mul(b, b), M::mul(13.0 / 12, M::mul(a, a)));
   }
   const double d1[1] = {1.0}, d2[2] = {2.0 / 3, 1.0 / 3}, d3[3] = {0.3, 0.6, 0.1};
   const double *d = K == 1 ? d1 : (K == 2 ? d2 : d3);
   double xr, xl;
   if constexpr (M::strict && K == 1) {
      // alfa = 1/eps**2, w = alfa/alfa = 1 exactly: vr = 1*vrr(0), vl = 1*vlr(0)   (weno.f90:186,207-214)
      xr = vrr[0];
      xl = vlr[0];
   } else if constexpr (M::strict) {
      // the eleven (k=3) IEEE quotients of weno.f90:207-210 from five refined reciprocals, as in weno_run_k2/_k3:
      // bit-identical to the divisions while eps+beta < 2^256 (one integer range test), else the literal slow path
      double e[K], den[K], al[K], at[K];
#pragma unroll
      for (int r = 0; r < K; ++r) {
         e[r] = M::add(eps, beta[r]);
         den[r] = M::mul(e[r], e[r]);
         const double rc = exact_recip(den[r]);
         al[r] = exact_div_nc(d[r], den[r], rc);
         at[r] = exact_div_nc(d[K - 1 - r], den[r], rc);
      }
      double s = al[0], st = at[0];
#pragma unroll
      for (int r = 1; r < K; ++r) {
         s = M::add(s, al[r]);
         st = M::add(st, at[r]);
      }
      const double rs = exact_recip(s), rst = exact_recip(st);
      xr = M::mul(exact_div_nc(al[0], s, rs), vrr[0]);
      xl = M::mul(exact_div_nc(at[0], st, rst), vlr[0]);
#pragma unroll
      for (int r = 1; r < K; ++r) {
         xr = M::add(xr, M::mul(exact_div_nc(al[r], s, rs), vrr[r]));
         xl = M::add(xl, M::mul(exact_div_nc(at[r], st, rst), vlr[r]));
      }
      if (!weights_in_range(e[0], e[K - 1], e[K / 2])) {
         double2 f;
         if constexpr (K == 2)
            f = weno_weights_slow_k2(den[0], den[1], vrr[0], vrr[1], vlr[0], vlr[1]);
         else
            f = weno_weights_slow_k3(den[0], den[1], den[K - 1], vrr[0], vrr[1], vrr[K - 1], vlr[0], vlr[1], vlr[K - 1]);
         xl = f.x;
         xr = f.y;
      }
   } else {
      // FAST mode with per-cell tables (2D general operators, fv2d.cu GEN): division-light weights as in the uniform fast
      // path -- alfa_r ~ d_r * prod_{s != r} (eps+beta_s)^2, one reciprocal per side -- with FMA contraction; the candidates
      // and smoothness indicators above are the reference's expressions (contracted).  Within the north-star tolerance of
      // the reference order (a few ULP per reconstruction), not bit-identical.  Products grow like v^8..v^9: cells whose
      // stencil holds |v| >= 2^100 are evaluated on the stencil scaled by an exact power of two (as weno_run does).
      double den[K];
#pragma unroll
      for (int r = 0; r < K; ++r) {
         const double e = eps + beta[r];
         den[r] = e * e;
      }
      if constexpr (K == 1) { // w = alfa/alfa = 1 (weno.f90:186,207-214)
         xr = vrr[0];
         xl = vlr[0];
      } else if constexpr (K == 2) {
         const double a0 = d[0] * den[1], a1 = d[1] * den[0];
         const double t0 = d[1] * den[1], t1 = d[0] * den[0];
         xr = fma(a1, vrr[1], a0 * vrr[0]) * rcp3(a0 + a1);
         xl = fma(t1, vlr[1], t0 * vlr[0]) * rcp3(t0 + t1);
      } else {
         const double p0 = den[1] * den[K - 1], p1 = den[0] * den[K - 1], p2 = den[0] * den[1];
         const double a0 = 0.3 * p0, a1 = 0.6 * p1, a2 = 0.1 * p2;
         const double t0 = 0.1 * p0, t2 = 0.3 * p2; // alfatilde_1 == alfa_1
         xr = fma(a2, vrr[K - 1], fma(a1, vrr[1], a0 * vrr[0])) * fast_rcp((a0 + a1) + a2);
         xl = fma(t2, vlr[K - 1], fma(a1, vlr[1], t0 * vlr[0])) * fast_rcp((t0 + a1) + t2);
      }
   }
   vr = xr;
   vl = xl;
}

template <int K, class M>
__device__ __forceinline__ void weno_cell_nonuniform(const double *ci, const double *ve /* [-(K-1)..K-1] */, double eps, double &vl,
                                                     double &vr) {
   if constexpr (M::strict || K == 1) {
      weno_cell_nonuniform_core<K, M>(ci, ve, eps, vl, vr);
   } else {
      // magnitude guard of the division-light weights (see weno_run): rare, predicated rescaling of the stencil in registers
      constexpr int N = 2 * K - 1;
      const float mhi = fast_window_maxhi<N>(ve - (K - 1));
      double vv[N];
#pragma unroll
      for (int j = 0; j < N; ++j) vv[j] = ve[j - (K - 1)];
      double epsv = eps, sup = 1.0;
      const bool big = !(mhi < __int_as_float(FAST_RANGE_HI));
      if (big) {
         const int e = ((__float_as_int(mhi) >> 20) & 0x7ff) - 1023 - 40;
         const double sdn = __hiloint2double((1023 - e) << 20, 0);
         sup = __hiloint2double((1023 + e) << 20, 0);
#pragma unroll
         for (int j = 0; j < N; ++j) vv[j] *= sdn;
         epsv = fmax(eps * sdn * sdn, 0x1p-240);
      }
      weno_cell_nonuniform_core<K, M>(ci, vv + (K - 1), epsv, vl, vr);
      if (big) {
         vl *= sup;
         vr *= sup;
      }
   }
}

} // namespace hrw
