// fv1d_small.cu -- one-launch integrator for small 1D problems (the shipped example1: 100 cells).
//
// A problem of <= 1024 cells is launch-bound on the tiled path (3 kernels per RK3 step, ~3.5 us each, for ~0.1 us of
// work).  Here ONE CTA keeps the whole state on chip and runs every stage of every step of an `integrate` call:
// thread i owns cell i; per stage the cell values and the reconstructed vl/vr travel through shared memory
// (two __syncthreads), everything else stays in registers.  The arithmetic is the tiled kernel's, function for
// function (weno_run, face_flux_k, exact_div_q, div3, the stage formulas of fv1d_finish), so strict mode stays
// bit-identical to the oracle.  The host driver (ode.cu) counts the steps with the reference's own t accumulation and
// is_done test before the launch.
#include "fv1d.cuh"

namespace hrw {

struct SmallArgs {
   double *u; // dense state, in/out (device)
   int n;
   long long nsteps;
   double dt;
   WenoK kc;
   FluxCfg flux;
   int bc;
   const double2 *wtab; // width dictionary (or nullptr)
   const unsigned char *widx;
   const double *width; // width array (when no dictionary)
};

template <int K, int ORDER, class M, int FK>
__global__ void __launch_bounds__(1024) fv1d_small_kernel(const SmallArgs a) {
   extern __shared__ __align__(16) double sm[];
   const int n = a.n;
   double *sv = sm + 4;           // sv[-4 .. n+3]: cell values with replicated ghosts
   double *svl = sm + (n + 8) + 1; // svl[-1 .. n]
   double *svr = svl + (n + 2);    // svr[-1 .. n]
   const int i = threadIdx.x;
   const bool act = i < n;
   const bool dict = a.wtab != nullptr;

   double wd = 1.0, wr = 1.0;
   if (act) {
      if (dict) {
         const double2 e = a.wtab[a.widx[i]];
         wd = e.x;
         wr = e.y; // exact_recip(w), computed at creation (fv.cu: wtab_recip_kernel)
      } else {
         wd = a.width[i];
         wr = M::strict ? exact_recip(wd) : fast_rcp(wd);
      }
   }
   const double lscale = FK == FK_BURGERS_GODUNOV ? -0.5 : -1.0; // fv1d.cuh: sign of the divergence, Burgers' exact 1/2
   const bool copy = a.bc == HRWENO_BC_COPY_NEIGHBOUR;
   double un = act ? a.u[i] : 0.0;
   if (i == 0) {
      svl[-1] = svr[-1] = 0.0;
      svl[n] = svr[n] = 0.0;
   }

   for (long long step = 0; step < a.nsteps; ++step) {
      double v = un;
#pragma unroll
      for (int j = 0; j < ORDER; ++j) {
         // stage input with edge-replicated ghost cells (weno.f90:171-173)
         if (act) sv[i] = v;
         if (i == 0) sv[-1] = sv[-2] = sv[-3] = v;
         if (i == n - 1) sv[n] = sv[n + 1] = sv[n + 2] = v;
         __syncthreads();
         double w5[5], vl, vr;
#pragma unroll
         for (int q = 0; q < 5; ++q) w5[q] = act ? sv[i - 2 + q] : 0.0;
         weno_run<K, 1, M>(w5 + (2 - (K - 1)), a.kc, &vl, &vr);
         if (act) {
            svl[i] = vl;
            svr[i] = vr;
         }
         __syncthreads();
         if (act) {
            // faces i (left) and i+1 (right) of cell i; boundary rules of example1:103-104 / example2:117-120
            double Fl = face_flux_k<FK, M>(a.flux, svr[i - 1], vl);
            double Fr = face_flux_k<FK, M>(a.flux, vr, svl[i + 1]);
            if (i == 0) Fl = copy ? Fr : 0.0;
            if (i == n - 1) Fr = copy ? Fl : 0.0;
            const double c0 = (ORDER == 3 && j == 2) ? 2 * a.dt : a.dt; // tvdode.f90:163-167
            const double cL = lscale * c0;
            const double dF = M::sub(Fr, Fl);
            double o;
            if constexpr (M::strict) {
               bool ok = true;
               double q = exact_div_q(dF, wd, wr, ok);
               if (!ok) q = __ddiv_rn(dF, wd);
               if (ORDER == 1 || j == 0)
                  o = M::add(v, M::mul(cL, q)); // u + dt*udot
               else if (ORDER == 2)
                  o = M::mul(M::add(M::add(un, v), M::mul(cL, q)), 0.5); // (u + ui + dt*udot)/2
               else if (j == 1)
                  o = M::mul(M::add(M::add(M::mul(3.0, un), v), M::mul(cL, q)), 0.25); // (3*u + ui + dt*udot)/4
               else {
                  const double sum = M::add(M::fma_exact(2.0, v, un), M::mul(cL, q)); // (u + 2*ui + 2*dt*udot)/3
                  bool ok3 = true;
                  o = exact_div(sum, 3.0, 1.0 / 3, ok3);
                  if (!ok3) o = __ddiv_rn(sum, 3.0);
               }
            } else {
               const double q = dict ? dF * (wr * cL) : cL * (dF * wr); // dict: the tiled kernel's table carries cL/w
               if (ORDER == 1 || j == 0)
                  o = v + q;
               else if (ORDER == 2)
                  o = ((un + v) + q) * 0.5;
               else if (j == 1)
                  o = (fma(3.0, un, v) + q) * 0.25;
               else
                  o = div3<M>(fma(2.0, v, un) + q);
            }
            v = o;
         }
         // the next stage overwrites sv/svl/svr: every thread must be done reading them
         __syncthreads();
      }
      un = v;
   }
   if (act) a.u[i] = un;
}

template <int K, int ORDER, class M>
static int launch_small_f(const SmallArgs &a, int fk, cudaStream_t st) {
   const int nt = ((a.n + 31) / 32) * 32;
   const size_t smem = (size_t)((a.n + 8) + 2 * (a.n + 2)) * sizeof(double);
   if (fk == FK_BURGERS_GODUNOV)
      fv1d_small_kernel<K, ORDER, M, FK_BURGERS_GODUNOV><<<1, nt, smem, st>>>(a);
   else
      fv1d_small_kernel<K, ORDER, M, FK_GENERIC><<<1, nt, smem, st>>>(a);
   HRW_CUDA(cudaGetLastError());
   return HRWENO_OK;
}

template <int K, class M>
static int launch_small_o(const SmallArgs &a, int order, int fk, cudaStream_t st) {
   if (order == 1) return launch_small_f<K, 1, M>(a, fk, st);
   if (order == 2) return launch_small_f<K, 2, M>(a, fk, st);
   return launch_small_f<K, 3, M>(a, fk, st);
}

template <class M>
static int launch_small_k(const SmallArgs &a, int k, int order, int fk, cudaStream_t st) {
   if (k == 1) return launch_small_o<1, M>(a, order, fk, st);
   if (k == 2) return launch_small_o<2, M>(a, order, fk, st);
   return launch_small_o<3, M>(a, order, fk, st);
}

bool fv_small_eligible(const Fv *fv) {
   return !fv->general && fv->d.ndim == 1 && fv->rows == 1 && fv->d.nranks <= 1 && fv->n0 >= 2 && fv->n0 <= 1024;
}

// nsteps RK steps of order `order` on the dense device vector u, in one launch
int fv_small_integrate(Fv *fv, double *u_dev, int order, long long nsteps, double dt, cudaStream_t st) {
   const hrweno_fv_desc &d = fv->d;
   SmallArgs a{};
   a.u = u_dev;
   a.n = (int)fv->n0;
   a.nsteps = nsteps;
   a.dt = dt;
   a.kc = make_wenok(d.eps);
   a.flux = FluxCfg{d.flux_model, d.flux_scheme, d.flux_coef[0], d.alpha};
   a.bc = d.bc;
   a.wtab = fv->width_dict ? fv->d_wtab : nullptr;
   a.widx = fv->d_widx;
   a.width = fv->d_width[0];
   const int fk = (d.flux_model == HRWENO_FLUX_BURGERS && d.flux_scheme == HRWENO_SCHEME_GODUNOV) ? FK_BURGERS_GODUNOV : FK_GENERIC;
   fv->launches++;
   if (d.mode == HRWENO_MODE_STRICT) return launch_small_k<Strict>(a, d.k, order, fk, st);
   return launch_small_k<Fast>(a, d.k, order, fk, st);
}

} // namespace hrw
