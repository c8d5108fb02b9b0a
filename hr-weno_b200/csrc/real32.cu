// real32.cu -- the REAL32 build of the path (src/hrweno_kinds.F90:9-17: `rk` = real32 is a compile-time switch of the whole
// reference, so every real -- state, widths, tables, eps, t, dt -- is single precision).
//
// Scope: one GPU, the reference's operation order in separately rounded float operations (the general kernels of
// fvgen.cuh instantiated for float: uniform or per-cell tables, Godunov / Lax-Friedrichs, x- and t-dependent flux
// factors, 1D rows and 2D, rktvd 1-3 and mstvd), bit-identical to the REAL32 build of the oracle
// (oracle/libhrweno_oracle_f32.so).  The state is the caller's dense layout (these kernels clamp at the domain ends and
// neither read nor write ghost cells), so the integrators work in place on the caller's device vector.  The tuned TMA
// kernels (fv1d.cuh, fv2d.cu) exist in fp64 only: their tile geometry and exact-division sequences are fp64-specific.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <new>

#include "fvgen.cuh"

namespace hrw {

static inline bool is_done32(float t, float tout, float dt) { return (t - tout) * std::copysign(1.0f, dt) > 0.0f; } // tvdode.f90:282

struct Fv32 {
   hrweno_fv_desc_f32 d{};
   int64_t n0 = 0, n1 = 1, rows = 1, neq = 0;
   float *d_w[2] = {nullptr, nullptr};
   float *d_cnu[2] = {nullptr, nullptr};
   float *d_fc[2] = {nullptr, nullptr}, *d_cc[2] = {nullptr, nullptr};
   hrweno_time_fn_f32 tfn = nullptr;
   void *tfn_ctx = nullptr;
   float *d_scratch = nullptr; // host-pointer rhs
   cudaStream_t stream = nullptr;
   int64_t launches = 0;
   ~Fv32() {
      for (int a = 0; a < 2; ++a) {
         cudaFree(d_w[a]);
         cudaFree(d_cnu[a]);
         cudaFree(d_fc[a]);
         cudaFree(d_cc[a]);
      }
      cudaFree(d_scratch);
      if (stream) cudaStreamDestroy(stream);
   }
};

static int upload32(const float *h, int64_t n, float **dev) {
   HRW_CUDA(cudaMalloc(dev, (size_t)n * sizeof(float)));
   HRW_CUDA(cudaMemcpy(*dev, h, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
   return HRWENO_OK;
}

// weno_calc_cnu (weno.f90:221-297, Shu eq. 2.20) in the arithmetic of T (the fp64 copy lives in weno.cu)
template <class T>
static void calc_cnu_t(int64_t nc, int k, const T *xedges, std::vector<T> &cnu) {
   const int ng = k + 1;
   std::vector<T> buf((size_t)(nc + 2 * ng + 1));
   T *xext = buf.data() + ng;
   for (int64_t i = 0; i <= nc; ++i) xext[i] = xedges[i];
   volatile T dx = xext[1] - xext[0]; // volatile: no contraction / excess precision on the host
   for (int64_t i = -1; i >= -ng; --i) {
      volatile T v = xext[i + 1] - dx;
      xext[i] = v;
   }
   dx = xext[nc] - xext[nc - 1];
   for (int64_t i = nc + 1; i <= nc + ng; ++i) {
      volatile T v = xext[i - 1] + dx;
      xext[i] = v;
   }
   auto xl = [&](int64_t m) { return xext[m - 1]; };
   auto xr = [&](int64_t m) { return xext[m]; };
   cnu.assign((size_t)nc * k * (k + 1), T(0));
   for (int64_t i = 1; i <= nc; ++i)
      for (int r = -1; r <= k - 1; ++r)
         for (int j = 0; j <= k - 1; ++j) {
            volatile T sum2 = T(0);
            for (int m = j + 1; m <= k; ++m) {
               volatile T prod2 = T(1);
               for (int l = 0; l <= k; ++l) {
                  if (l == m) continue;
                  volatile T diff = xl(i - r + m) - xl(i - r + l);
                  prod2 = prod2 * diff;
               }
               volatile T sum1 = T(0);
               for (int l = 0; l <= k; ++l) {
                  if (l == m) continue;
                  volatile T prod1 = T(1);
                  for (int q = 0; q <= k; ++q) {
                     if (q == m || q == l) continue;
                     volatile T diff = xr(i) - xl(i - r + q);
                     prod1 = prod1 * diff;
                  }
                  sum1 = sum1 + prod1;
               }
               volatile T quot = sum1 / prod2;
               sum2 = sum2 + quot;
            }
            volatile T wdt = xr(i - r + j) - xl(i - r + j);
            volatile T val = sum2 * wdt;
            cnu[(size_t)(j + k * ((r + 1) + (k + 1) * (i - 1)))] = val;
         }
}

static int fv32_create(Fv32 **out, const hrweno_fv_desc_f32 *desc) {
   if (!out || !desc) return fail(HRWENO_EINVAL, "hrweno_fv_f32_create: null argument");
   if (desc->abi_version != HRWENO_ABI_VERSION) return fail(HRWENO_EINVAL, "hrweno_fv_f32_create: abi_version mismatch");
   if (desc->ndim != 1 && desc->ndim != 2) return fail(HRWENO_EINVAL, "Invalid input 'ndim'. Valid range: 1 <= ndim <= 2.");
   for (int a = 0; a < desc->ndim; ++a)
      if (!(desc->n[a] > 0)) return fail(HRWENO_EINVAL, "Invalid input 'ncells'. Valid range: ncells > 0."); // weno.f90:75
   if (!(desc->k >= 1 && desc->k <= 3)) return fail(HRWENO_EINVAL, "Invalid input 'k'. Valid range: 1 <= k <= 3."); // :84
   if (!(desc->eps > FLT_EPSILON)) return fail(HRWENO_EINVAL, "Invalid input 'eps'. Valid range: eps > epsilon.");   // :94, epsilon(1.0_rk)
   if (desc->flux_model != HRWENO_FLUX_BURGERS && desc->flux_model != HRWENO_FLUX_LINEAR) return fail(HRWENO_EINVAL, "hrweno_fv_f32_create: unknown flux_model");
   if (desc->flux_scheme != HRWENO_SCHEME_GODUNOV && desc->flux_scheme != HRWENO_SCHEME_LAX_FRIEDRICHS)
      return fail(HRWENO_EINVAL, "hrweno_fv_f32_create: unknown flux_scheme");
   if (desc->bc != HRWENO_BC_COPY_NEIGHBOUR && desc->bc != HRWENO_BC_ZERO_FLUX) return fail(HRWENO_EINVAL, "hrweno_fv_f32_create: unknown bc");
   if (desc->bc == HRWENO_BC_COPY_NEIGHBOUR)
      for (int a = 0; a < desc->ndim; ++a)
         if (desc->n[a] < 2) return fail(HRWENO_EINVAL, "hrweno_fv_f32_create: copy-neighbour boundary needs ncells >= 2");
   if (desc->nranks > 1) return fail(HRWENO_EINVAL, "hrweno_fv_f32_create: the REAL32 path runs on one GPU");
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(HRWENO_ECUDA, "no CUDA device (this library has no CPU fallback)");
   Fv32 *fv = new (std::nothrow) Fv32();
   if (!fv) return fail(HRWENO_ENOMEM, "out of host memory");
   fv->d = *desc;
   fv->n0 = desc->n[0];
   fv->n1 = desc->ndim == 2 ? desc->n[1] : 1;
   fv->rows = desc->ndim == 1 ? (desc->rows < 1 ? 1 : desc->rows) : 1;
   fv->neq = desc->ndim == 1 ? fv->n0 * fv->rows : fv->n0 * fv->n1;
   int st = HRWENO_OK;
   if (desc->grid_kind == HRWENO_GRID_LINEAR) {
      if (desc->ndim != 1 || !(desc->xmax > desc->xmin)) {
         st = fail(HRWENO_EINVAL, "hrweno_fv_f32_create: GRID_LINEAR is 1D with xmax > xmin");
      } else { // grid1%linear in real32: edges(i) = xmin + rx*i, width = edges(i) - edges(i-1)  (grids.f90:76-79,247)
         std::vector<float> w((size_t)fv->n0);
         volatile float rx = (desc->xmax - desc->xmin) / (float)fv->n0;
         volatile float el = desc->xmin;
         for (int64_t i = 1; i <= fv->n0; ++i) {
            volatile float p = rx * (float)i;
            volatile float er = desc->xmin + p;
            volatile float wd = er - el;
            w[(size_t)(i - 1)] = wd;
            el = er;
         }
         st = upload32(w.data(), fv->n0, &fv->d_w[0]);
      }
   } else if (desc->grid_kind == HRWENO_GRID_WIDTH_ARRAY) {
      for (int a = 0; a < desc->ndim && st == HRWENO_OK; ++a) {
         if (!desc->width[a]) st = fail(HRWENO_EINVAL, "hrweno_fv_f32_create: width array missing");
         else st = upload32(desc->width[a], desc->n[a], &fv->d_w[a]);
      }
   } else {
      st = fail(HRWENO_EINVAL, "hrweno_fv_f32_create: unknown grid_kind");
   }
   fv->d.width[0] = fv->d.width[1] = nullptr;
   if (st == HRWENO_OK) {
      cudaError_t e = cudaStreamCreateWithFlags(&fv->stream, cudaStreamNonBlocking);
      if (e != cudaSuccess) st = cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__);
   }
   if (st != HRWENO_OK) {
      delete fv;
      return st;
   }
   *out = fv;
   return HRWENO_OK;
}

template <int K>
static void launch32(bool two_d, unsigned blocks, const GenGeomT<float> &g, const StageArgsT<float> &a, int combine, cudaStream_t st) {
   if (two_d)
      fvgen_stage_kernel<K, true, float, StageArgsT<float>><<<blocks, GEN_NT, 0, st>>>(g, a, combine);
   else
      fvgen_stage_kernel<K, false, float, StageArgsT<float>><<<blocks, GEN_NT, 0, st>>>(g, a, combine);
}

// one fused stage on dense float vectors; t = the time of this rhs evaluation (read by the time factor only)
static int fv32_stage(Fv32 *fv, int combine, const StageArgsT<float> &args, float t, cudaStream_t st) {
   const hrweno_fv_desc_f32 &d = fv->d;
   const bool two_d = d.ndim == 2;
   GenGeomT<float> g{};
   g.n0 = fv->n0;
   g.n1 = two_d ? fv->n1 : fv->rows;
   g.ld = fv->n0; // dense rows
   g.bc = d.bc;
   g.cnu0 = fv->d_cnu[0];
   g.cnu1 = fv->d_cnu[1];
   g.w0 = fv->d_w[0];
   g.w1 = fv->d_w[1];
   g.fc0 = fv->d_fc[0];
   g.fc1 = fv->d_fc[1];
   g.cc0 = fv->d_cc[0];
   g.cc1 = fv->d_cc[1];
   g.eps = d.eps;
   g.fx0 = FluxCfgT<float>{d.flux_model, d.flux_scheme, d.flux_coef[0], d.alpha};
   g.fx1 = FluxCfgT<float>{d.flux_model, d.flux_scheme, d.flux_coef[1], d.alpha};
   g.phys_l = g.phys_r = 1;
   g.has_ts = fv->tfn != nullptr;
   g.ts = fv->tfn ? fv->tfn(fv->tfn_ctx, t) : 1.0f;
   const int64_t tx = two_d ? GenTile<true>::TX : GenTile<false>::TX, ty = two_d ? GenTile<true>::TY : GenTile<false>::TY;
   const int64_t tiles = ((g.n0 + tx - 1) / tx) * ((g.n1 + ty - 1) / ty);
   const int64_t cap = 148 * (two_d ? gen_minb<true>() : gen_minb<false>());
   const unsigned blocks = (unsigned)(tiles < cap ? tiles : cap);
   if (d.k == 1) launch32<1>(two_d, blocks, g, args, combine, st);
   else if (d.k == 2) launch32<2>(two_d, blocks, g, args, combine, st);
   else launch32<3>(two_d, blocks, g, args, combine, st);
   fv->launches++;
   HRW_CUDA(cudaGetLastError());
   return HRWENO_OK;
}

static StageArgsT<float> sargs(const float *vin, const float *a, const float *b, float *out, float *out2, int64_t ld, float c0, float c1) {
   StageArgsT<float> s{};
   s.vin = vin, s.a = a, s.b = b, s.out = out, s.out2 = out2, s.ld_out = ld, s.out_dense = 1, s.c0 = c0, s.c1 = c1;
   return s;
}

// ---- integrators (tvdode.f90) in real32: t, dt, tout and the stage coefficients are floats ------------------------------
struct Ode32 {
   bool is_ms = false;
   Fv32 *fv = nullptr;
   int order = 3;
   int64_t neq = 0, fevals = 0;
   int istate = 0, ring = 0;
   std::vector<float *> bufs; // rk: T1, T2 ; ms: u ring (5), L ring (4), T1, T2 ; last: dense staging for host u
   ~Ode32() {
      for (float *p : bufs) cudaFree(p);
   }
};

static int ode32_create(Ode32 **out, bool is_ms, Fv32 *fv, int order) {
   if (!out || !fv) return fail(HRWENO_EINVAL, "ode create: null argument");
   if (!(order >= 1 && order <= 3)) return fail(HRWENO_EINVAL, "Invalid input 'order' in 'rktvd'. Valid range: 1 <= k <= 3."); // tvdode.f90:89
   Ode32 *o = new (std::nothrow) Ode32();
   if (!o) return fail(HRWENO_ENOMEM, "out of host memory");
   o->is_ms = is_ms, o->fv = fv, o->order = is_ms ? 3 : order, o->neq = fv->neq;
   const size_t nb = (is_ms ? 5 + 4 + 2 : 2) + 1;
   for (size_t i = 0; i < nb; ++i) {
      float *p = nullptr;
      cudaError_t e = cudaMalloc(&p, (size_t)o->neq * sizeof(float));
      if (e != cudaSuccess) {
         delete o;
         return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
      }
      o->bufs.push_back(p);
   }
   o->istate = 1; // tvdode.f90:93,199
   *out = o;
   return HRWENO_OK;
}

// one RK step src -> dst (dst may equal src); stage times t, t+dt, t+dt/2 (tvdode.f90:138-167)
static int rk32_step(Ode32 *o, int order, float t, float *src, float *dst, float *t1, float *t2, float dt, cudaStream_t st) {
   Fv32 *fv = o->fv;
   const int64_t ld = fv->n0;
   HRW_TRY(fv32_stage(fv, C_EULER, sargs(src, nullptr, nullptr, t1, nullptr, ld, dt, 0.0f), t, st)); // ui = u + dt*udot
   if (order == 1) {
      HRW_CUDA(cudaMemcpyAsync(dst, t1, (size_t)o->neq * sizeof(float), cudaMemcpyDeviceToDevice, st));
      return HRWENO_OK;
   }
   if (order == 2) return fv32_stage(fv, C_RK2_FINAL, sargs(t1, src, nullptr, dst, nullptr, ld, dt, 0.0f), t + dt, st); // u = (u + ui + dt*udot)/2
   HRW_TRY(fv32_stage(fv, C_RK3_S2, sargs(t1, src, nullptr, t2, nullptr, ld, dt, 0.0f), t + dt, st));              // ui = (3u + ui + dt*udot)/4
   return fv32_stage(fv, C_RK3_S3, sargs(t2, src, nullptr, dst, nullptr, ld, 2 * dt, 0.0f), t + dt / 2, st);       // u = (u + 2ui + 2dt*udot)/3
}

static int ode32_integrate_dev(Ode32 *o, float *u, float *t, float tout, float dt, int itask, cudaStream_t st) {
   if (!o || !u || !t) return fail(HRWENO_EINVAL, "integrate: null argument");
   if (o->istate < 1) return HRWENO_OK;            // tvdode.f90:126,228
   if (is_done32(*t, tout, dt)) return HRWENO_OK;  // :127,229
   Fv32 *fv = o->fv;
   const int64_t ld = fv->n0;
   const size_t bytes = (size_t)o->neq * sizeof(float);
   if (!o->is_ms) { // rktvd_integrate :97-178
      float *T1 = o->bufs[0], *T2 = o->bufs[1];
      for (;;) {
         HRW_TRY(rk32_step(o, o->order, *t, u, u, T1, T2, dt, st));
         *t = *t + dt;
         o->fevals += o->order;
         if (is_done32(*t, tout, dt) || itask == 2) break;
      }
      if (o->istate == 1) o->istate = 2;
      return HRWENO_OK;
   }
   // mstvd_integrate :203-271; rings addressed by the step number as in ode.cu
   float **uring = &o->bufs[0], **lring = &o->bufs[5];
   float *T1 = o->bufs[9], *T2 = o->bufs[10];
   int &step = o->ring;
   HRW_CUDA(cudaMemcpyAsync(uring[step % 5], u, bytes, cudaMemcpyDeviceToDevice, st));
   if (o->istate == 1) {
      for (int i = 0; i < 4; ++i) {
         float *U = uring[step % 5], *Un = uring[(step + 1) % 5];
         HRW_TRY(fv32_stage(fv, C_RHS, sargs(U, nullptr, nullptr, lring[step % 4], nullptr, ld, 0.0f, 0.0f), *t, st)); // udotold(:,i) = fu(t,u)
         HRW_TRY(rk32_step(o, 3, *t, U, Un, T1, T2, dt, st));
         *t = *t + dt;
         step++;
      }
      o->fevals = 12; // :246
      o->istate = 2;
   }
   const float c50 = 50 * dt, c10 = 10 * dt;
   for (;;) {
      if (is_done32(*t, tout, dt)) break;
      float *U = uring[step % 5], *Uo4 = uring[(step + 1) % 5], *Lo4 = lring[step % 4];
      HRW_TRY(fv32_stage(fv, C_MS, sargs(U, Uo4, Lo4, Uo4, Lo4, ld, c50, c10), *t, st)); // :257
      *t = *t + dt;
      o->fevals += 1;
      step++;
   }
   HRW_CUDA(cudaMemcpyAsync(u, uring[step % 5], bytes, cudaMemcpyDeviceToDevice, st));
   step %= 20;
   return HRWENO_OK;
}

// ---- weno(ncells, k, eps, xedges) / reconstruct in real32 -------------------------------------------------------------
struct Weno32 {
   int64_t ncells = 0;
   int k = 3;
   float eps = 1e-6f;
   float *d_cnu = nullptr;
   ~Weno32() { cudaFree(d_cnu); }
};

template <int K>
__global__ void __launch_bounds__(256) recon32_kernel(const float *__restrict__ v, int64_t ldv, int64_t incv, float *__restrict__ vl,
                                                      float *__restrict__ vr, int64_t ldo, int64_t n, int64_t rows, float eps,
                                                      const float *__restrict__ cnu) {
   const int64_t total = rows * n;
   for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      const int64_t row = idx / n, i = idx - row * n;
      const float *vrow = v + row * ldv;
      float w[2 * K - 1];
#pragma unroll
      for (int o = -(K - 1); o <= K - 1; ++o) { // edge replicas (weno.f90:171-173)
         int64_t ii = i + o;
         ii = ii < 0 ? 0 : (ii > n - 1 ? n - 1 : ii);
         w[o + K - 1] = __ldg(vrow + ii * incv);
      }
      float ci[K * (K + 1)];
      if (cnu) {
#pragma unroll
         for (int q = 0; q < K * (K + 1); ++q) ci[q] = __ldg(cnu + (size_t)i * (K * (K + 1)) + q);
      } else {
         uniform_table<K, float>(ci);
      }
      float l, r;
      weno_cell_reference<K, float>(ci, w + (K - 1), eps, l, r);
      vl[row * ldo + i] = l;
      vr[row * ldo + i] = r;
   }
}

static int weno32_launch(const Weno32 *w, int64_t rows, const float *v, int64_t ldv, int64_t incv, float *vl, float *vr, int64_t ldo, cudaStream_t st) {
   int64_t blocks = (rows * w->ncells + 255) / 256;
   blocks = blocks < 1 ? 1 : (blocks > 148 * 16 ? 148 * 16 : blocks);
   if (w->k == 1) recon32_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(v, ldv, incv, vl, vr, ldo, w->ncells, rows, w->eps, w->d_cnu);
   else if (w->k == 2) recon32_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(v, ldv, incv, vl, vr, ldo, w->ncells, rows, w->eps, w->d_cnu);
   else recon32_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(v, ldv, incv, vl, vr, ldo, w->ncells, rows, w->eps, w->d_cnu);
   HRW_CUDA(cudaGetLastError());
   return HRWENO_OK;
}

} // namespace hrw

using namespace hrw;

extern "C" {

int hrweno_weno_f32_create(hrweno_weno_f32 **out, int64_t ncells, int k, float eps, const float *xedges) {
   if (!out) return fail(HRWENO_EINVAL, "hrweno_weno_f32_create: null argument");
   if (!(ncells > 0)) return fail(HRWENO_EINVAL, "Invalid input 'ncells'. Valid range: ncells > 0.");    // weno.f90:75
   if (!(k >= 1 && k <= 3)) return fail(HRWENO_EINVAL, "Invalid input 'k'. Valid range: 1 <= k <= 3."); // :84
   if (!(eps > FLT_EPSILON)) return fail(HRWENO_EINVAL, "Invalid input 'eps'. Valid range: eps > epsilon."); // :94
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(HRWENO_ECUDA, "no CUDA device (this library has no CPU fallback)");
   Weno32 *w = new (std::nothrow) Weno32();
   if (!w) return fail(HRWENO_ENOMEM, "out of host memory");
   w->ncells = ncells, w->k = k, w->eps = eps;
   if (xedges) {
      std::vector<float> cnu;
      calc_cnu_t<float>(ncells, k, xedges, cnu);
      const int st = upload32(cnu.data(), (int64_t)cnu.size(), &w->d_cnu);
      if (st != HRWENO_OK) {
         delete w;
         return st;
      }
   }
   *out = reinterpret_cast<hrweno_weno_f32 *>(w);
   return HRWENO_OK;
}
void hrweno_weno_f32_destroy(hrweno_weno_f32 *w) { delete reinterpret_cast<Weno32 *>(w); }
int hrweno_weno_f32_reconstruct_dev(const hrweno_weno_f32 *h, int64_t rows, const float *v_dev, int64_t ldv, int64_t incv, float *vl_dev,
                                    float *vr_dev, int64_t ldo, void *stream) {
   const Weno32 *w = reinterpret_cast<const Weno32 *>(h);
   if (!w || !v_dev || !vl_dev || !vr_dev) return fail(HRWENO_EINVAL, "null argument");
   if (rows < 1 || incv < 1 || ldo < w->ncells) return fail(HRWENO_EINVAL, "invalid rows/incv/ldo");
   return weno32_launch(w, rows, v_dev, ldv, incv, vl_dev, vr_dev, ldo, (cudaStream_t)stream);
}
int hrweno_weno_f32_reconstruct(const hrweno_weno_f32 *h, const float *v, float *vl, float *vr) {
   const Weno32 *w = reinterpret_cast<const Weno32 *>(h);
   if (!w || !v || !vl || !vr) return fail(HRWENO_EINVAL, "null argument");
   const size_t bytes = (size_t)w->ncells * sizeof(float);
   float *d = nullptr;
   HRW_CUDA(cudaMalloc(&d, 3 * bytes));
   int st = HRWENO_OK;
   cudaError_t e = cudaMemcpy(d, v, bytes, cudaMemcpyHostToDevice);
   if (e == cudaSuccess) st = weno32_launch(w, 1, d, w->ncells, 1, d + w->ncells, d + 2 * w->ncells, w->ncells, nullptr);
   if (e == cudaSuccess && st == HRWENO_OK) e = cudaMemcpy(vl, d + w->ncells, bytes, cudaMemcpyDeviceToHost);
   if (e == cudaSuccess && st == HRWENO_OK) e = cudaMemcpy(vr, d + 2 * w->ncells, bytes, cudaMemcpyDeviceToHost);
   if (e != cudaSuccess) st = cuda_fail(e, "hrweno_weno_f32_reconstruct", __FILE__, __LINE__);
   cudaFree(d);
   return st;
}

void hrweno_weno_f32_reconstruct_s(const hrweno_weno_f32 *h, const float *v, float *vl, float *vr, int *status) {
   const int st = hrweno_weno_f32_reconstruct(h, v, vl, vr);
   if (status) *status = st;
}
int hrweno_weno_f32_get_cnu(const hrweno_weno_f32 *h, float *cnu_host) {
   const Weno32 *w = reinterpret_cast<const Weno32 *>(h);
   if (!w || !cnu_host) return fail(HRWENO_EINVAL, "null argument");
   if (!w->d_cnu) return fail(HRWENO_EINVAL, "hrweno_weno_f32_get_cnu: uniform-grid object has no cnu (weno.f90:110-112)");
   HRW_CUDA(cudaMemcpy(cnu_host, w->d_cnu, (size_t)w->ncells * (size_t)(w->k * (w->k + 1)) * sizeof(float), cudaMemcpyDeviceToHost));
   return HRWENO_OK;
}

int hrweno_fv_f32_create(hrweno_fv_f32 **out, const hrweno_fv_desc_f32 *desc) { return fv32_create(reinterpret_cast<Fv32 **>(out), desc); }
void hrweno_fv_f32_destroy(hrweno_fv_f32 *fv) { delete reinterpret_cast<Fv32 *>(fv); }
int64_t hrweno_fv_f32_neq(const hrweno_fv_f32 *fv) { return fv ? reinterpret_cast<const Fv32 *>(fv)->neq : 0; }
int hrweno_fv_f32_rhs_dev(hrweno_fv_f32 *h, float t, const float *v_dev, float *vdot_dev, void *stream) {
   Fv32 *fv = reinterpret_cast<Fv32 *>(h);
   if (!fv || !v_dev || !vdot_dev) return fail(HRWENO_EINVAL, "null argument");
   return fv32_stage(fv, C_RHS, sargs(v_dev, nullptr, nullptr, vdot_dev, nullptr, fv->n0, 0.0f, 0.0f), t, (cudaStream_t)stream);
}
int hrweno_fv_f32_rhs(hrweno_fv_f32 *h, float t, const float *v, float *vdot) {
   Fv32 *fv = reinterpret_cast<Fv32 *>(h);
   if (!fv || !v || !vdot) return fail(HRWENO_EINVAL, "null argument");
   const size_t bytes = (size_t)fv->neq * sizeof(float);
   if (!fv->d_scratch) HRW_CUDA(cudaMalloc(&fv->d_scratch, 2 * bytes));
   float *din = fv->d_scratch, *dout = din + fv->neq;
   HRW_CUDA(cudaMemcpyAsync(din, v, bytes, cudaMemcpyHostToDevice, fv->stream));
   HRW_TRY(hrweno_fv_f32_rhs_dev(h, t, din, dout, fv->stream));
   HRW_CUDA(cudaMemcpyAsync(vdot, dout, bytes, cudaMemcpyDeviceToHost, fv->stream));
   HRW_CUDA(cudaStreamSynchronize(fv->stream));
   return HRWENO_OK;
}
int hrweno_fv_f32_set_xedges(hrweno_fv_f32 *h, int axis, const float *xedges) {
   Fv32 *fv = reinterpret_cast<Fv32 *>(h);
   if (!fv || !xedges) return fail(HRWENO_EINVAL, "hrweno_fv_f32_set_xedges: null argument");
   if (axis < 0 || axis >= fv->d.ndim) return fail(HRWENO_EINVAL, "hrweno_fv_f32_set_xedges: invalid axis");
   std::vector<float> cnu;
   calc_cnu_t<float>(axis == 0 ? fv->n0 : fv->n1, fv->d.k, xedges, cnu);
   cudaFree(fv->d_cnu[axis]);
   fv->d_cnu[axis] = nullptr;
   return upload32(cnu.data(), (int64_t)cnu.size(), &fv->d_cnu[axis]);
}
int hrweno_fv_f32_set_flux_coef(hrweno_fv_f32 *h, int axis, const float *face, const float *cross) {
   Fv32 *fv = reinterpret_cast<Fv32 *>(h);
   if (!fv) return fail(HRWENO_EINVAL, "null fv handle");
   if (axis < 0 || axis >= fv->d.ndim) return fail(HRWENO_EINVAL, "hrweno_fv_f32_set_flux_coef: invalid axis");
   if (cross && fv->d.ndim != 2) return fail(HRWENO_EINVAL, "hrweno_fv_f32_set_flux_coef: a cross coefficient needs ndim == 2");
   const int64_t n = axis == 0 ? fv->n0 : fv->n1, nother = axis == 0 ? fv->n1 : fv->n0;
   cudaFree(fv->d_fc[axis]);
   cudaFree(fv->d_cc[axis]);
   fv->d_fc[axis] = fv->d_cc[axis] = nullptr;
   if (face) HRW_TRY(upload32(face, n + 1, &fv->d_fc[axis]));
   if (cross) HRW_TRY(upload32(cross, nother, &fv->d_cc[axis]));
   return HRWENO_OK;
}
int hrweno_fv_f32_set_flux_time_fn(hrweno_fv_f32 *h, hrweno_time_fn_f32 g, void *ctx) {
   Fv32 *fv = reinterpret_cast<Fv32 *>(h);
   if (!fv) return fail(HRWENO_EINVAL, "null fv handle");
   fv->tfn = g, fv->tfn_ctx = ctx;
   return HRWENO_OK;
}

int hrweno_rktvd_f32_create_fused(hrweno_ode_f32 **out, hrweno_fv_f32 *fv, int order) {
   return ode32_create(reinterpret_cast<Ode32 **>(out), false, reinterpret_cast<Fv32 *>(fv), order);
}
int hrweno_mstvd_f32_create_fused(hrweno_ode_f32 **out, hrweno_fv_f32 *fv) {
   return ode32_create(reinterpret_cast<Ode32 **>(out), true, reinterpret_cast<Fv32 *>(fv), 3);
}
void hrweno_ode_f32_destroy(hrweno_ode_f32 *ode) { delete reinterpret_cast<Ode32 *>(ode); }
int hrweno_ode_f32_integrate_dev(hrweno_ode_f32 *ode, float *u_dev, float *t, float tout, float dt, int itask, void *stream) {
   return ode32_integrate_dev(reinterpret_cast<Ode32 *>(ode), u_dev, t, tout, dt, itask, (cudaStream_t)stream);
}
int hrweno_ode_f32_integrate(hrweno_ode_f32 *h, float *u, float *t, float tout, float dt, int itask) {
   Ode32 *o = reinterpret_cast<Ode32 *>(h);
   if (!o || !u || !t) return fail(HRWENO_EINVAL, "integrate: null argument");
   if (o->istate < 1 || is_done32(*t, tout, dt)) return HRWENO_OK;
   float *d = o->bufs.back();
   const size_t bytes = (size_t)o->neq * sizeof(float);
   cudaStream_t st = o->fv->stream;
   HRW_CUDA(cudaMemcpyAsync(d, u, bytes, cudaMemcpyHostToDevice, st));
   HRW_TRY(ode32_integrate_dev(o, d, t, tout, dt, itask, st));
   HRW_CUDA(cudaMemcpyAsync(u, d, bytes, cudaMemcpyDeviceToHost, st));
   HRW_CUDA(cudaStreamSynchronize(st));
   return HRWENO_OK;
}
int64_t hrweno_ode_f32_fevals(const hrweno_ode_f32 *ode) { return ode ? reinterpret_cast<const Ode32 *>(ode)->fevals : 0; }
int hrweno_ode_f32_istate(const hrweno_ode_f32 *ode) { return ode ? reinterpret_cast<const Ode32 *>(ode)->istate : -1; }
int64_t hrweno_ode_f32_launches(const hrweno_ode_f32 *ode) { return ode ? reinterpret_cast<const Ode32 *>(ode)->fv->launches : 0; }

} // extern "C"
