// ode.cu -- rktvd / mstvd drivers (src/hrweno_tvdode.f90).
//
// The time loop, `t = t + dt` accumulation, the strict is_done test and the fevals/istate
// bookkeeping run on the host in fp64 exactly as the reference does (tvdode.f90:126-176, 228-268,
// 282); every stage is one fused device kernel (fused path) or one user rhs call plus one combine
// kernel (callback path).  mstvd keeps its history in rings addressed by the step number instead of
// the reference's two eoshift copies per step (tvdode.f90:262-263).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <new>

#include "internal.hpp"

namespace hrw {

static inline bool is_done(double t, double tout, double dt) { return (t - tout) * std::copysign(1.0, dt) > 0.0; }

Ode::~Ode() {
   for (double *p : ext_bufs) cudaFree(p);
   delete fv_ext;
   if (ev_edge) cudaEventDestroy(ev_edge);
   cudaFreeHost(h_u);
   cudaFreeHost(h_udot);
   for (double *p : bufs) cudaFree(p);
   for (cudaEvent_t e : ev_in) cudaEventDestroy(e);
   for (cudaEvent_t e : ev_fin) cudaEventDestroy(e);
   if (s_in) cudaStreamDestroy(s_in);
   if (s_out) cudaStreamDestroy(s_out);
   if (stream) cudaStreamDestroy(stream);
}

// ------------------------------------------------------------------------------------------------
// K6: stage combinations on dense vectors for the callback path (tvdode.f90:141,149-167,257)
// ------------------------------------------------------------------------------------------------
template <int COMBINE>
__global__ void combine_kernel(int64_t n, const double *__restrict__ v, const double *__restrict__ L,
                               const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ out,
                               double c0, double c1) {
   for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const double l = L[i], x = v[i];
      double o;
      if (COMBINE == C_EULER)
         o = __dadd_rn(x, __dmul_rn(c0, l));
      else if (COMBINE == C_RK2_FINAL)
         o = __dmul_rn(__dadd_rn(__dadd_rn(a[i], x), __dmul_rn(c0, l)), 0.5);
      else if (COMBINE == C_RK3_S2)
         o = __dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(3.0, a[i]), x), __dmul_rn(c0, l)), 0.25);
      else if (COMBINE == C_RK3_S3)
         o = __ddiv_rn(__dadd_rn(__fma_rn(2.0, x, a[i]), __dmul_rn(c0, l)), 3.0);
      else
         o = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(25.0, x), __dmul_rn(c0, l)), __dmul_rn(7.0, a[i])),
                                 __dmul_rn(c1, b[i])),
                       0.03125);
      out[i] = o;
   }
}

static int launch_combine(int combine, int64_t n, const double *v, const double *L, const double *a, const double *b,
                          double *out, double c0, double c1, cudaStream_t st) {
   int64_t blocks = (n + 255) / 256;
   if (blocks > 148 * 16) blocks = 148 * 16;
   const unsigned g = (unsigned)blocks;
   switch (combine) {
   case C_EULER: combine_kernel<C_EULER><<<g, 256, 0, st>>>(n, v, L, a, b, out, c0, c1); break;
   case C_RK2_FINAL: combine_kernel<C_RK2_FINAL><<<g, 256, 0, st>>>(n, v, L, a, b, out, c0, c1); break;
   case C_RK3_S2: combine_kernel<C_RK3_S2><<<g, 256, 0, st>>>(n, v, L, a, b, out, c0, c1); break;
   case C_RK3_S3: combine_kernel<C_RK3_S3><<<g, 256, 0, st>>>(n, v, L, a, b, out, c0, c1); break;
   default: combine_kernel<C_MS><<<g, 256, 0, st>>>(n, v, L, a, b, out, c0, c1); break;
   }
   HRW_CUDA(cudaGetLastError());
   return HRWENO_OK;
}

// ------------------------------------------------------------------------------------------------
// creation
// ------------------------------------------------------------------------------------------------
static int ode_alloc(Ode *o) {
   size_t n_bufs, doubles;
   if (o->fused) {
      n_bufs = o->is_ms ? 5 + 4 + 2 : 3;
      doubles = o->fv->state_doubles();
   } else {
      // callback path: dense vectors.  RK: u, ui, udot.  MS: u ring (5), udot ring (4), ui, udot.
      n_bufs = o->is_ms ? 5 + 4 + 2 : 3;
      doubles = (size_t)o->neq;
   }
   for (size_t i = 0; i < n_bufs; ++i) {
      double *p = nullptr;
      HRW_CUDA(cudaMalloc(&p, doubles * sizeof(double)));
      o->bufs.push_back(p);
      HRW_CUDA(cudaMemset(p, 0, doubles * sizeof(double)));
   }
   // dense staging vector for the host-pointer entry point
   double *p = nullptr;
   HRW_CUDA(cudaMalloc(&p, (size_t)o->neq * sizeof(double)));
   o->bufs.push_back(p);
   HRW_CUDA(cudaStreamCreateWithFlags(&o->stream, cudaStreamNonBlocking));
   return HRWENO_OK;
}

int ode_create(Ode **out, bool is_ms, Fv *fv, hrweno_rhs_fn fu, void *ctx, int64_t neq, int order) {
   if (!out) return fail(HRWENO_EINVAL, "ode create: null argument");
   if (!fv && !fu) return fail(HRWENO_EINVAL, "ode create: no integrand");
   if (fv) neq = fv->neq;
   if (!(neq > 0)) return fail(HRWENO_EINVAL, "Invalid input 'neq'. Valid range: neq >= 1."); // tvdode.f90:83,192
   if (!(order >= 1 && order <= 3))
      return fail(HRWENO_EINVAL, "Invalid input 'order' in 'rktvd'. Valid range: 1 <= k <= 3."); // :89
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
      return fail(HRWENO_ECUDA, "no CUDA device (this library has no CPU fallback)");
   Ode *o = new (std::nothrow) Ode();
   if (!o) return fail(HRWENO_ENOMEM, "out of host memory");
   o->is_ms = is_ms;
   o->fused = fv != nullptr;
   o->fv = fv;
   o->fu = fu;
   o->ctx = ctx;
   o->neq = neq;
   o->order = is_ms ? 3 : order; // tvdode.f90:195
   const int st = ode_alloc(o);
   if (st != HRWENO_OK) {
      delete o;
      return st;
   }
   o->istate = 1; // tvdode.f90:93,199
   *out = o;
   return HRWENO_OK;
}

// ------------------------------------------------------------------------------------------------
// one RK step  src -> dst  (dst may equal src); t1, t2 are temporaries.  Fused: padded states.
// ------------------------------------------------------------------------------------------------
static int rk_step_fused(Ode *o, int order, double t, double *src, double *dst, double *t1, double *t2, double dt, cudaStream_t st) {
   Fv *fv = o->fv;
   StageArgs a{};
   a.ld_out = fv->pitch;
   a.out_dense = 0;
   a.t = t; // fu(t, u, udot)  tvdode.f90:138,149,162
   auto c0 = [&](double *p) { return fv->cell0(p); };
   if (order == 1) { // u = u + dt*udot  (needs dst != src: the stencil reads src)
      a.vin = c0(src);
      a.out = c0(dst);
      a.c0 = dt;
      HRW_TRY(fv_stage_halo(fv, C_EULER, a, true, st));
      o->launches += 1;
      return HRWENO_OK;
   }
   a.vin = c0(src);
   a.out = c0(t1);
   a.c0 = dt;
   HRW_TRY(fv_stage_halo(fv, C_EULER, a, true, st)); // ui = u + dt*udot
   a.t = t + dt; // fu(t + dt, ui, udot)  tvdode.f90:151,164
   if (order == 2) {
      a.vin = c0(t1);
      a.a = c0(src);
      a.out = c0(dst);
      HRW_TRY(fv_stage_halo(fv, C_RK2_FINAL, a, true, st)); // u = (u + ui + dt*udot)/2
      o->launches += 2;
      return HRWENO_OK;
   }
   a.vin = c0(t1);
   a.a = c0(src);
   a.out = c0(t2);
   HRW_TRY(fv_stage_halo(fv, C_RK3_S2, a, true, st)); // ui = (3*u + ui + dt*udot)/4
   a.t = t + dt / 2; // fu(t + dt/2, ui, udot)  tvdode.f90:166
   a.vin = c0(t2);
   a.a = c0(src);
   a.out = c0(dst);
   a.c0 = 2 * dt;
   HRW_TRY(fv_stage_halo(fv, C_RK3_S3, a, true, st)); // u = (u + 2*ui + 2*dt*udot)/3
   o->launches += 3;
   return HRWENO_OK;
}

// callback path: dense vectors; ui and udot are work vectors (tvdode.f90:30-31)
static int rk_step_cb(Ode *o, int order, double t, double *u, double *ui, double *udot, double dt, cudaStream_t st) {
   const int64_t n = o->neq;
   o->fu(o->ctx, t, n, u, udot, st);
   if (order == 1) {
      HRW_TRY(launch_combine(C_EULER, n, u, udot, nullptr, nullptr, u, dt, 0, st));
      o->launches += 1;
      return HRWENO_OK;
   }
   HRW_TRY(launch_combine(C_EULER, n, u, udot, nullptr, nullptr, ui, dt, 0, st));
   o->fu(o->ctx, t + dt, n, ui, udot, st);
   if (order == 2) {
      HRW_TRY(launch_combine(C_RK2_FINAL, n, ui, udot, u, nullptr, u, dt, 0, st));
      o->launches += 2;
      return HRWENO_OK;
   }
   HRW_TRY(launch_combine(C_RK3_S2, n, ui, udot, u, nullptr, ui, dt, 0, st));
   o->fu(o->ctx, t + dt / 2, n, ui, udot, st);
   HRW_TRY(launch_combine(C_RK3_S3, n, ui, udot, u, nullptr, u, 2 * dt, 0, st));
   o->launches += 3;
   return HRWENO_OK;
}

// ------------------------------------------------------------------------------------------------
// rktvd_integrate (tvdode.f90:97-178)
// ------------------------------------------------------------------------------------------------
static int rk_integrate(Ode *o, double *u_dev, double *t, double tout, double dt, int itask, cudaStream_t st) {
   if (o->istate < 1) return HRWENO_OK;          // :126
   if (is_done(*t, tout, dt)) return HRWENO_OK;   // :127
   const bool resident = u_dev == nullptr; // state attached to the integrator (hrweno_ode_attach): no dense <-> padded copies
   if (resident) u_dev = o->bufs.back();   // small problems keep their (dense) state in the staging vector
   if (o->fused && fv_small_eligible(o->fv)) {
      // small problem: count the steps exactly as the loop below would take them, then one single-CTA launch
      long long nsteps = 0;
      double tt = *t;
      for (;;) {
         tt = tt + dt;
         ++nsteps;
         if (is_done(tt, tout, dt) || itask == 2) break;
      }
      HRW_TRY(fv_small_integrate(o->fv, u_dev, o->order, nsteps, dt, st));
      *t = tt;
      o->fevals += (int64_t)o->order * nsteps;
      if (o->istate == 1) o->istate = 2;
      return HRWENO_OK;
   }
   if (o->fused) {
      Fv *fv = o->fv;
      double *U = o->bufs[0], *T1 = o->bufs[1], *T2 = o->bufs[2];
      if (!resident) {
         HRW_TRY(fv_pack(fv, u_dev, fv->cell0(U), st));
         HRW_TRY(fv_exchange(fv, fv->cell0(U), st));
         o->launches++;
      }
      for (;;) {
         if (o->order == 1) {
            HRW_TRY(rk_step_fused(o, 1, *t, U, T1, nullptr, nullptr, dt, st));
            std::swap(U, T1);
         } else {
            HRW_TRY(rk_step_fused(o, o->order, *t, U, U, T1, T2, dt, st));
         }
         *t = *t + dt;
         o->fevals += o->order;
         if (is_done(*t, tout, dt) || itask == 2) break;
      }
      if (!resident) {
         HRW_TRY(fv_unpack(fv, fv->cell0(U), u_dev, st));
         o->launches++;
      }
      o->bufs[0] = U;
      o->bufs[1] = T1;
   } else {
      for (;;) {
         HRW_TRY(rk_step_cb(o, o->order, *t, u_dev, o->bufs[1], o->bufs[2], dt, st));
         *t = *t + dt;
         o->fevals += o->order;
         if (is_done(*t, tout, dt) || itask == 2) break;
      }
   }
   if (o->istate == 1) o->istate = 2; // :176
   return HRWENO_OK;
}

// ------------------------------------------------------------------------------------------------
// mstvd_integrate (tvdode.f90:203-271).  Step number n = o->ring: u^n lives in uring[n % 5],
// L(u^n) in lring[n % 4]; u^{n-4} and L^{n-4} are read and overwritten in place by u^{n+1}, L^n.
// ------------------------------------------------------------------------------------------------
static int ms_integrate(Ode *o, double *u_dev, double *t, double tout, double dt, cudaStream_t st) {
   if (o->istate < 1) return HRWENO_OK;        // :228
   if (is_done(*t, tout, dt)) return HRWENO_OK; // :229
   double **uring = &o->bufs[0], **lring = &o->bufs[5];
   double *T1 = o->bufs[9], *T2 = o->bufs[10];
   Fv *fv = o->fv;
   const int64_t n = o->neq;
   int &step = o->ring;
   const bool resident = u_dev == nullptr; // state attached to the integrator (hrweno_ode_attach)
   // the caller's u is the current state (it may have been edited between calls, like the reference's inout u)
   if (resident) {
      // the ring slot of the current step already holds the state, ghost cells included
   } else if (o->fused) {
      HRW_TRY(fv_pack(fv, u_dev, fv->cell0(uring[step % 5]), st));
      HRW_TRY(fv_exchange(fv, fv->cell0(uring[step % 5]), st));
      o->launches++;
   } else {
      HRW_CUDA(cudaMemcpyAsync(uring[step % 5], u_dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
   }
   if (o->istate == 1) { // :236-249: four RK3 single steps; uold(:,i) = u, udotold(:,i) = fu(t,u)
      for (int i = 0; i < 4; ++i) {
         double *U = uring[step % 5], *Un = uring[(step + 1) % 5];
         if (o->fused) {
            StageArgs a{};
            a.vin = fv->cell0(U);
            a.out = fv->cell0(lring[step % 4]);
            a.ld_out = fv->pitch;
            a.t = *t; // fu(t, u, udotold(:,i))  tvdode.f90:242
            HRW_TRY(fv_stage_halo(fv, C_RHS, a, false, st));
            o->launches++;
            HRW_TRY(rk_step_fused(o, 3, *t, U, Un, T1, T2, dt, st));
         } else {
            o->fu(o->ctx, *t, n, U, lring[step % 4], st);
            HRW_CUDA(cudaMemcpyAsync(Un, U, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
            HRW_TRY(rk_step_cb(o, 3, *t, Un, T1, T2, dt, st));
         }
         *t = *t + dt;
         step++;
      }
      o->fevals = 12; // :246: fevals = ode_start%fevals (the four extra fu calls are not counted)
      o->istate = 2;
   }
   const double c50 = 50 * dt, c10 = 10 * dt;
   for (;;) { // :252-268
      if (is_done(*t, tout, dt)) break;
      double *U = uring[step % 5], *Uo4 = uring[(step + 1) % 5], *Lo4 = lring[step % 4];
      if (o->fused) {
         StageArgs a{};
         a.vin = fv->cell0(U);
         a.a = fv->cell0(Uo4);
         a.b = fv->cell0(Lo4);
         a.out = fv->cell0(Uo4);
         a.out2 = fv->cell0(Lo4);
         a.ld_out = fv->pitch;
         a.c0 = c50;
         a.c1 = c10;
         a.t = *t; // fu(t, u, udot)  tvdode.f90:256
         HRW_TRY(fv_stage_halo(fv, C_MS, a, true, st));
         o->launches++;
      } else {
         o->fu(o->ctx, *t, n, U, T1, st); // udot
         HRW_TRY(launch_combine(C_MS, n, U, T1, Uo4, Lo4, Uo4, c50, c10, st));
         HRW_CUDA(cudaMemcpyAsync(Lo4, T1, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
         o->launches++;
      }
      *t = *t + dt;
      o->fevals += 1;
      step++;
   }
   if (resident) {
      // stays where it is
   } else if (o->fused) {
      HRW_TRY(fv_unpack(fv, fv->cell0(uring[step % 5]), u_dev, st));
      o->launches++;
   } else {
      HRW_CUDA(cudaMemcpyAsync(u_dev, uring[step % 5], (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
   }
   // keep the step counter small: only its residues mod 5 and mod 4 matter
   step %= 20;
   return HRWENO_OK;
}

// host integrand behind the device-callback machinery: D2H u, the user's procedure, H2D udot.  The stage combinations
// (K6) and the state stay on the device; what runs on the host is the caller's own code, as in the reference.
static void host_integrand_adapter(void *ctx, double t, int64_t neq, const double *u_dev, double *udot_dev, void *stream) {
   Ode *o = static_cast<Ode *>(ctx);
   cudaStream_t st = static_cast<cudaStream_t>(stream);
   const size_t bytes = (size_t)neq * sizeof(double);
   cudaError_t e = cudaMemcpyAsync(o->h_u, u_dev, bytes, cudaMemcpyDeviceToHost, st);
   if (e == cudaSuccess) e = cudaStreamSynchronize(st);
   if (e == cudaSuccess) {
      o->fu_host(o->ctx_host, t, neq, o->h_u, o->h_udot);
      e = cudaMemcpyAsync(udot_dev, o->h_udot, bytes, cudaMemcpyHostToDevice, st);
   }
   if (e != cudaSuccess && o->cb_status == HRWENO_OK) o->cb_status = cuda_fail(e, "host integrand staging", __FILE__, __LINE__);
}

int ode_create_host(Ode **out, bool is_ms, hrweno_rhs_host_fn fu, void *ctx, int64_t neq, int order) {
   if (!fu) return fail(HRWENO_EINVAL, "ode create: no integrand");
   HRW_TRY(ode_create(out, is_ms, nullptr, host_integrand_adapter, nullptr, neq, order));
   Ode *o = *out;
   o->ctx = o; // the adapter's context is the object itself
   o->fu_host = fu;
   o->ctx_host = ctx;
   const size_t bytes = (size_t)neq * sizeof(double);
   cudaError_t e = cudaMallocHost(&o->h_u, bytes);
   if (e == cudaSuccess) e = cudaMallocHost(&o->h_udot, bytes);
   if (e != cudaSuccess) {
      delete o;
      *out = nullptr;
      return cuda_fail(e, "pinned staging for the host integrand", __FILE__, __LINE__);
   }
   return HRWENO_OK;
}

// Device callers that keep their state inside the integrator between output times (fused path): attach packs the dense
// vector into the library-owned padded state ONCE, integrate_attached advances it with no dense <-> padded copies (they are
// 2 x 16 B per cell and call: 20 % of an RK1 call's traffic), fetch writes the current state to a dense vector.
int ode_attach(Ode *o, const double *u_dev, cudaStream_t st) {
   if (!o || !u_dev) return fail(HRWENO_EINVAL, "hrweno_ode_attach: null argument");
   if (!o->fused) return fail(HRWENO_EINVAL, "hrweno_ode_attach: only the fused integrators keep a padded state");
   Fv *fv = o->fv;
   if (fv_small_eligible(fv) && !o->is_ms) {
      HRW_CUDA(cudaMemcpyAsync(o->bufs.back(), u_dev, (size_t)o->neq * sizeof(double), cudaMemcpyDeviceToDevice, st));
   } else {
      double *cur = o->is_ms ? o->bufs[o->ring % 5] : o->bufs[0];
      HRW_TRY(fv_pack(fv, u_dev, fv->cell0(cur), st));
      HRW_TRY(fv_exchange(fv, fv->cell0(cur), st));
   }
   o->attached = true;
   return HRWENO_OK;
}

int ode_fetch(Ode *o, double *u_dev, cudaStream_t st) {
   if (!o || !u_dev) return fail(HRWENO_EINVAL, "hrweno_ode_fetch: null argument");
   if (!o->attached) return fail(HRWENO_ESTATE, "hrweno_ode_fetch: no attached state (hrweno_ode_attach first)");
   Fv *fv = o->fv;
   if (fv_small_eligible(fv) && !o->is_ms) {
      HRW_CUDA(cudaMemcpyAsync(u_dev, o->bufs.back(), (size_t)o->neq * sizeof(double), cudaMemcpyDeviceToDevice, st));
   } else {
      double *cur = o->is_ms ? o->bufs[o->ring % 5] : o->bufs[0];
      HRW_TRY(fv_unpack(fv, fv->cell0(cur), u_dev, st));
   }
   return HRWENO_OK;
}

int ode_integrate_dev(Ode *o, double *u_dev, double *t, double tout, double dt, int itask, cudaStream_t st) {
   if (!o || !t) return fail(HRWENO_EINVAL, "integrate: null argument");
   if (!u_dev && !(o->fused && o->attached)) return fail(HRWENO_ESTATE, "integrate: null u and no attached state (hrweno_ode_attach first)");
   const int s = o->is_ms ? ms_integrate(o, u_dev, t, tout, dt, st) : rk_integrate(o, u_dev, t, tout, dt, itask, st);
   if (s == HRWENO_OK && o->cb_status != HRWENO_OK) return o->cb_status; // a host-integrand staging copy failed
   return s;
}

// ------------------------------------------------------------------------------------------------
// Host-pointer rktvd_integrate for one large 1D row: a time-skewed chunk pipeline.
// The row is cut into C chunks of whole tiles.  Chunk c is copied host->device on its own stream straight into the
// padded state; stage g of chunk c (a launch restricted to that chunk's tiles) needs stage g-1 of chunks c-1, c, c+1,
// hence chunk c+g+1 to have arrived.  Launching the tasks (c, g) diagonal by diagonal (d = c + g, g ascending inside a
// diagonal) on ONE compute stream satisfies every read-after-write and write-after-read dependency of the RK buffers by
// stream order alone, and lets the early chunks run many stages ahead while the late ones are still on the PCIe bus;
// chunk c goes back to the host as soon as its last stage is done.  Arithmetic and results are those of the plain path.
// ------------------------------------------------------------------------------------------------
__global__ void fill_ghost_kernel(double *cell0, int64_t n, int k, int left, int right) {
   const int q = threadIdx.x;
   if (q < k) {
      if (left) cell0[-1 - q] = cell0[0];      // weno.f90:172
      if (right) cell0[n + q] = cell0[n - 1];  // weno.f90:173
   }
}

// Slabs (one row cut over several GPUs): the per-stage halo traffic of the resident path would chain every slab's pipeline
// to its neighbours' (the first chunk of a slab needs, stage by stage, the LAST chunk of the slab to its left).  Instead the
// WIDE_H cells next to each interface are exchanged ONCE per call (halo.cu: fv_exchange_wide, peer stores + flags) and the
// pipeline runs on an operator over the slab extended by those cells (same global grid indices, hence the same widths and
// the same arithmetic per cell).  The extension's outer end has no valid neighbour data: what is computed there is wrong
// from the first stage on, and the error moves inwards k cells per stage -- after G stages it has covered k*G <= WIDE_H
// cells, i.e. it stops short of the slab's own cells, which therefore equal the single-domain result bit for bit.
static int pipeline_operator(Ode *o, int64_t G, Fv **P, double **B, int64_t *HL, int64_t *HR) {
   Fv *fv = o->fv;
   *P = nullptr;
   *HL = *HR = 0;
   if (fv->d.nranks <= 1) {
      *P = fv;
      for (int i = 0; i < 3; ++i) B[i] = fv->cell0(o->bufs[i]);
      return HRWENO_OK;
   }
   constexpr int64_t H = Halo::WIDE_H;
   if (fv->d.grid_kind != HRWENO_GRID_LINEAR || !fv->halo.ready || !fv->halo.wide_off) return HRWENO_OK;
   if ((int64_t)fv->d.k * G > H) return HRWENO_OK; // (slab sizes were checked by the caller on the smallest slab)
   const int64_t hl = fv->d.rank > 0 ? H : 0, hr = fv->d.rank < fv->d.nranks - 1 ? H : 0;
   if (!o->fv_ext) {
      hrweno_fv_desc d = fv->d; // same rank / nranks / global grid: physical-boundary flags and widths follow the global indices
      d.n[0] = fv->n0 + hl + hr;
      d.global_offset = fv->d.global_offset - hl;
      if (d.n[0] > 2000000000LL) return HRWENO_OK;
      HRW_TRY(fv_create(&o->fv_ext, &d));
      for (int i = 0; i < 3; ++i) {
         double *p = nullptr;
         HRW_TRY(o->fv_ext->alloc_state(&p));
         o->ext_bufs.push_back(p);
      }
      HRW_CUDA(cudaEventCreateWithFlags(&o->ev_edge, cudaEventDisableTiming));
   }
   *P = o->fv_ext;
   for (int i = 0; i < 3; ++i) B[i] = o->fv_ext->cell0(o->ext_bufs[i]);
   *HL = hl;
   *HR = hr;
   return HRWENO_OK;
}

static int rk_integrate_pipelined(Ode *o, double *u, double *t, double tout, double dt, int itask, bool *done) {
   *done = false;
   Fv *fv = o->fv;
   if (!o->fused || o->is_ms || fv->general || fv->d.ndim != 1 || fv->rows != 1) return HRWENO_OK;
   int chunk_tiles = 148 * 48; // a multiple of every resident-CTA count in use (148 x 1,2,3,4,6,8): no ragged last wave;
                               // 7.2 M cells per chunk measured best on B200 (profiles/r1_variant_sweeps.txt)
   if (const char *e = std::getenv("HRWENO_PIPE_CHUNK_TILES")) chunk_tiles = std::atoi(e);
   if (chunk_tiles < 1) return HRWENO_OK; // pipeline disabled
   // number of steps this call takes (tvdode.f90:161-171: step, then test)
   int64_t nsteps = 0;
   double tt = *t;
   for (;;) {
      tt = tt + dt;
      ++nsteps;
      if (is_done(tt, tout, dt) || itask == 2) break;
      if (nsteps > 4000000) return HRWENO_OK;
   }
   const int order = o->order;
   const int64_t G = (int64_t)order * nsteps;
   if (G > 1000000) return HRWENO_OK; // (the same on every rank)
   // the operator the pipeline runs on: this one (single GPU) or the slab extended by the wide halos
   Fv *P = nullptr;
   double *B[3] = {nullptr, nullptr, nullptr}; // U, T1, T2 at cell 0 of P
   int64_t HL = 0, HR = 0;
   {
      // slabs: every rank must take the same decision (a rank on the plain path would wait for per-stage halos its
      // neighbours never send), so the size tests use the smallest slab, floor(global_n / nranks), not this rank's own
      int tile0 = 0, tpr0 = 0;
      fv_tiling_1d(fv, &tile0, &tpr0);
      const int64_t nmin = fv->d.nranks > 1 ? fv->d.global_n / fv->d.nranks : fv->n0;
      const int64_t tiles_min = (nmin + tile0 - 1) / tile0;
      if ((tiles_min + chunk_tiles - 1) / chunk_tiles < 4) return HRWENO_OK; // fewer than four chunks: nothing to overlap
      if (fv->d.nranks > 1 && nmin < 4 * (int64_t)Halo::WIDE_H + 1) return HRWENO_OK;
   }
   HRW_TRY(pipeline_operator(o, G, &P, B, &HL, &HR));
   if (!P) return HRWENO_OK;
   int tile = 0, tpr = 0;
   fv_tiling_1d(P, &tile, &tpr);
   // chunk c covers tiles [cb[c], cb[c+1])  (smaller first chunks to shorten the ramp were tried: no measurable gain)
   std::vector<int> cb(1, 0);
   while (cb.back() < tpr) cb.push_back(std::min(tpr, cb.back() + chunk_tiles));
   const int C = (int)cb.size() - 1;
   if (C < 2) return HRWENO_OK; // cannot happen after the size tests above
   const int64_t n = fv->n0;      // this slab's cells: u[0..n) <-> cells [HL, HL+n) of P
   const int64_t np = P->n0;      // = HL + n + HR
   const int64_t launches0 = P->launches;
   if (!o->s_in) HRW_CUDA(cudaStreamCreateWithFlags(&o->s_in, cudaStreamNonBlocking));
   if (!o->s_out) HRW_CUDA(cudaStreamCreateWithFlags(&o->s_out, cudaStreamNonBlocking));
   while ((int)o->ev_in.size() < C) {
      cudaEvent_t e;
      HRW_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      o->ev_in.push_back(e);
      HRW_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      o->ev_fin.push_back(e);
   }
   const int k = fv->d.k;
   auto c_lo = [&](int c) { return (int64_t)cb[c] * tile; };
   auto c_hi = [&](int c) { return std::min<int64_t>(np, (int64_t)cb[c + 1] * tile); };
   cudaStream_t cs = o->stream;
   if (P != fv) {
      // the cells next to the slab interfaces first, then the wide exchange with the neighbours on the compute stream
      constexpr int64_t H = Halo::WIDE_H;
      HRW_CUDA(cudaMemcpyAsync(B[0] + HL, u, (size_t)H * sizeof(double), cudaMemcpyHostToDevice, o->s_in));
      HRW_CUDA(cudaMemcpyAsync(B[0] + HL + n - H, u + n - H, (size_t)H * sizeof(double), cudaMemcpyHostToDevice, o->s_in));
      HRW_CUDA(cudaEventRecord(o->ev_edge, o->s_in));
      HRW_CUDA(cudaStreamWaitEvent(cs, o->ev_edge, 0));
      HRW_TRY(fv_exchange_wide(fv, B[0] + HL, n, cs));
   }
   // host -> device, chunk by chunk, straight into the padded state (one row: dense offset == padded offset)
   for (int c = 0; c < C; ++c) {
      const int64_t lo = std::max<int64_t>(c_lo(c), HL), hi = std::min<int64_t>(c_hi(c), HL + n);
      if (hi > lo) HRW_CUDA(cudaMemcpyAsync(B[0] + lo, u + (lo - HL), (size_t)(hi - lo) * sizeof(double), cudaMemcpyHostToDevice, o->s_in));
      HRW_CUDA(cudaEventRecord(o->ev_in[c], o->s_in));
   }
   // Time-skewed chunks: stage g of chunk c covers tiles [L(c,g), L(c+1,g)) with L(c,g) = cb[c] - (g+1) for the inner
   // boundaries (one tile further left per stage; L(0,g) = 0, L(C,g) = tpr).  A stage needs its input k cells beyond
   // either end of its range: on the right that stays inside the range the same chunk wrote one stage earlier, on the
   // left inside what chunk c-1 wrote.  So chunk c depends on chunks <= c only: all of its stages run as soon as it has
   // arrived, in plain chunk-major stream order, while later chunks are still on the bus, and its final range goes back
   // to the host right after its last stage.  Nothing a later task reads is overwritten early: a buffer is rewritten
   // `order` stages later, i.e. at least two tiles further left (one for RK1/RK2, still more than k cells).
   auto Lb = [&](int c, int64_t g) -> int {
      if (c <= 0) return 0;
      if (c >= C) return tpr;
      const int64_t v = (int64_t)cb[c] - (g + 1);
      return (int)(v < 0 ? 0 : v);
   };
   const int fin_buf = order == 1 ? (int)(nsteps & 1) : 0; // RK1 ping-pongs U <-> T1
   const int phys_l = P->d.rank == 0, phys_r = P->d.rank == P->d.nranks - 1;
   for (int c = 0; c < C; ++c) {
      HRW_CUDA(cudaStreamWaitEvent(cs, o->ev_in[c], 0));
      if (c == 0 && phys_l) fill_ghost_kernel<<<1, 32, 0, cs>>>(B[0], np, k, 1, 0);
      if (c == C - 1 && phys_r) fill_ghost_kernel<<<1, 32, 0, cs>>>(B[0], np, k, 0, 1);
      if ((c == 0 && phys_l) || (c == C - 1 && phys_r)) P->launches++;
      for (int64_t g = 0; g < G; ++g) {
         const int64_t step = g / order;
         const int j = (int)(g - step * order);
         StageArgs a{};
         a.ld_out = P->pitch;
         a.tile_begin = Lb(c, g);
         a.tile_end = Lb(c + 1, g);
         if (a.tile_end <= a.tile_begin) continue; // the skew has moved this chunk's range out of the domain
         int combine;
         if (order == 1) {
            combine = C_EULER;
            a.vin = B[step & 1];
            a.out = B[(step & 1) ^ 1];
            a.c0 = dt;
         } else if (j == 0) {
            combine = C_EULER; // ui = u + dt*udot
            a.vin = B[0];
            a.out = B[1];
            a.c0 = dt;
         } else if (order == 2) {
            combine = C_RK2_FINAL; // u = (u + ui + dt*udot)/2
            a.vin = B[1];
            a.a = B[0];
            a.out = B[0];
            a.c0 = dt;
         } else if (j == 1) {
            combine = C_RK3_S2; // ui = (3*u + ui + dt*udot)/4
            a.vin = B[1];
            a.a = B[0];
            a.out = B[2];
            a.c0 = dt;
         } else {
            combine = C_RK3_S3; // u = (u + 2*ui + 2*dt*udot)/3
            a.vin = B[2];
            a.a = B[0];
            a.out = B[0];
            a.c0 = 2 * dt;
         }
         HRW_TRY(fv_stage(P, combine, a, cs));
      }
      // the cells chunk c finished with its last stage: device -> host on the output stream (this slab's own cells only)
      const int64_t f_lo = std::max<int64_t>((int64_t)Lb(c, G - 1) * tile, HL);
      const int64_t f_hi = std::min<int64_t>(std::min<int64_t>(np, (int64_t)Lb(c + 1, G - 1) * tile), HL + n);
      if (f_hi > f_lo) {
         HRW_CUDA(cudaEventRecord(o->ev_fin[c], cs));
         HRW_CUDA(cudaStreamWaitEvent(o->s_out, o->ev_fin[c], 0));
         HRW_CUDA(cudaMemcpyAsync(u + (f_lo - HL), B[fin_buf] + f_lo, (size_t)(f_hi - f_lo) * sizeof(double), cudaMemcpyDeviceToHost, o->s_out));
      }
   }
   HRW_CUDA(cudaStreamSynchronize(o->s_out));
   HRW_CUDA(cudaStreamSynchronize(cs));
   if (P != fv) {
      fv->launches += P->launches - launches0;
      HRW_TRY(fv_halo_status(fv));
      if (order == 1 && (nsteps & 1)) std::swap(o->ext_bufs[0], o->ext_bufs[1]);
   } else if (order == 1 && (nsteps & 1)) {
      std::swap(o->bufs[0], o->bufs[1]);
   }
   *t = tt;
   o->fevals += (int64_t)order * nsteps;
   if (o->istate == 1) o->istate = 2;
   *done = true;
   return HRWENO_OK;
}

int ode_integrate_host(Ode *o, double *u, double *t, double tout, double dt, int itask) {
   if (!o || !u || !t) return fail(HRWENO_EINVAL, "integrate: null argument");
   if (o->istate < 1) return HRWENO_OK;
   if (is_done(*t, tout, dt)) return HRWENO_OK;
   if (o->fused) {
      bool done = false;
      HRW_TRY(rk_integrate_pipelined(o, u, t, tout, dt, itask, &done));
      if (done) return HRWENO_OK;
   }
   double *d = o->bufs.back();
   const size_t bytes = (size_t)o->neq * sizeof(double);
   HRW_CUDA(cudaMemcpyAsync(d, u, bytes, cudaMemcpyHostToDevice, o->stream));
   HRW_TRY(ode_integrate_dev(o, d, t, tout, dt, itask, o->stream));
   HRW_CUDA(cudaMemcpyAsync(u, d, bytes, cudaMemcpyDeviceToHost, o->stream));
   HRW_CUDA(cudaStreamSynchronize(o->stream));
   if (o->fused) HRW_TRY(fv_halo_status(o->fv));
   return HRWENO_OK;
}

} // namespace hrw
