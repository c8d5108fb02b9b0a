// capi.cu -- the extern "C" boundary declared in include/hrweno_b200.h.
#include <cstring>
#include <limits>
#include <new>
#include <string>

#include "fv2d.cuh"
#include "internal.hpp"

namespace hrw {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }
const char *last_error_cstr() { return g_last_error.c_str(); }

int fail(int status, const std::string &msg) {
   g_last_error = msg;
   return status;
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
   g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" + std::to_string(line) + ")";
   return e == cudaErrorMemoryAllocation ? HRWENO_ENOMEM : HRWENO_ECUDA;
}

int weno_create(Weno **out, int64_t ncells, int k, double eps, const double *xedges);
int ode_create(Ode **out, bool is_ms, Fv *fv, hrweno_rhs_fn fu, void *ctx, int64_t neq, int order);
int ode_create_host(Ode **out, bool is_ms, hrweno_rhs_host_fn fu, void *ctx, int64_t neq, int order);
int ode_integrate_dev(Ode *o, double *u_dev, double *t, double tout, double dt, int itask, cudaStream_t st);
int ode_integrate_host(Ode *o, double *u, double *t, double tout, double dt, int itask);
int ode_attach(Ode *o, const double *u_dev, cudaStream_t st);
int ode_fetch(Ode *o, double *u_dev, cudaStream_t st);

// closed-set numerical flux per face (fluxes.f90:43,67-74)
__global__ void flux_faces_kernel(FluxCfg c, int64_t n, const double *__restrict__ vm, const double *__restrict__ vp,
                                  double *__restrict__ h) {
   for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      h[i] = face_flux<Strict>(c, vm[i], vp[i]);
}

static int ensure_buf(Weno *w, size_t doubles) {
   if (w->buf_doubles >= doubles) return HRWENO_OK;
   cudaFree(w->d_buf);
   w->d_buf = nullptr;
   w->buf_doubles = 0;
   HRW_CUDA(cudaMalloc(&w->d_buf, doubles * sizeof(double)));
   w->buf_doubles = doubles;
   return HRWENO_OK;
}

static int fv_rhs_any(Fv *fv, double t, const double *v_dev, double *vdot_dev, cudaStream_t st) {
   if (!fv->d_scratch_in) HRW_TRY(fv->alloc_state(&fv->d_scratch_in));
   HRW_TRY(fv_pack(fv, v_dev, fv->cell0(fv->d_scratch_in), st));
   HRW_TRY(fv_exchange(fv, fv->cell0(fv->d_scratch_in), st));
   StageArgs a{};
   a.vin = fv->cell0(fv->d_scratch_in);
   a.out = vdot_dev;
   a.ld_out = fv->n0;
   a.out_dense = 1;
   a.t = t;
   return fv_stage_halo(fv, C_RHS, a, false, st);
}

} // namespace hrw

using namespace hrw;

extern "C" {

int hrweno_abi_version(void) { return HRWENO_ABI_VERSION; }

const char *hrweno_last_error(void) { return g_last_error.c_str(); }

int hrweno_device_count(void) {
   int n = 0;
   if (cudaGetDeviceCount(&n) != cudaSuccess) {
      cudaGetLastError();
      return 0;
   }
   return n;
}

// ---- weno ------------------------------------------------------------------------------------------
int hrweno_weno_create(hrweno_weno **out, int64_t ncells, int k, double eps, const double *xedges) {
   return weno_create(reinterpret_cast<Weno **>(out), ncells, k, eps, xedges);
}

void hrweno_weno_destroy(hrweno_weno *w) { delete reinterpret_cast<Weno *>(w); }

int hrweno_weno_info(const hrweno_weno *h, int64_t *ncells, int *k, double *eps, int *uniform_grid) {
   const Weno *w = reinterpret_cast<const Weno *>(h);
   if (!w) return fail(HRWENO_EINVAL, "null weno handle");
   if (ncells) *ncells = w->ncells;
   if (k) *k = w->k;
   if (eps) *eps = w->eps;
   if (uniform_grid) *uniform_grid = w->uniform ? 1 : 0;
   return HRWENO_OK;
}

int hrweno_weno_set_mode(hrweno_weno *h, int mode) {
   Weno *w = reinterpret_cast<Weno *>(h);
   if (!w) return fail(HRWENO_EINVAL, "null weno handle");
   if (mode != HRWENO_MODE_STRICT && mode != HRWENO_MODE_FAST) return fail(HRWENO_EINVAL, "invalid mode");
   w->mode = mode;
   return HRWENO_OK;
}

int hrweno_weno_get_cnu(const hrweno_weno *h, double *cnu_host) {
   const Weno *w = reinterpret_cast<const Weno *>(h);
   if (!w || !cnu_host) return fail(HRWENO_EINVAL, "null argument");
   if (w->uniform) return fail(HRWENO_EINVAL, "cnu is not allocated for a uniform grid (weno.f90:100-112)");
   std::memcpy(cnu_host, w->cnu_host.data(), w->cnu_host.size() * sizeof(double));
   return HRWENO_OK;
}

int hrweno_weno_reconstruct_dev(const hrweno_weno *h, int64_t rows, const double *v, int64_t ldv, int64_t incv,
                                double *vl, double *vr, int64_t ldo, void *stream) {
   const Weno *w = reinterpret_cast<const Weno *>(h);
   if (!w || !v || !vl || !vr) return fail(HRWENO_EINVAL, "null argument");
   if (rows < 1 || incv < 1 || ldo < w->ncells) return fail(HRWENO_EINVAL, "invalid rows/incv/ldo");
   return weno_reconstruct_launch(w, rows, v, ldv, incv, vl, vr, ldo, (cudaStream_t)stream);
}

int hrweno_weno_reconstruct_batch(const hrweno_weno *h, int64_t rows, const double *v, int64_t ldv, int64_t incv,
                                  double *vl, double *vr, int64_t ldo) {
   Weno *w = const_cast<Weno *>(reinterpret_cast<const Weno *>(h));
   if (!w || !v || !vl || !vr) return fail(HRWENO_EINVAL, "null argument");
   if (rows < 1 || incv < 1 || ldv < 0 || ldo < w->ncells) return fail(HRWENO_EINVAL, "invalid rows/ldv/incv/ldo");
   const int64_t n = w->ncells;
   // extent of the strided input in doubles
   const size_t in_doubles = (size_t)((rows - 1) * ldv + (n - 1) * incv + 1);
   const size_t out_doubles = (size_t)(rows * n);
   std::lock_guard<std::mutex> lock(w->mtx); // the handle's staging buffers are shared: serialise host-pointer calls
   HRW_TRY(ensure_buf(w, in_doubles + 2 * out_doubles));
   double *dv = w->d_buf, *dl = dv + in_doubles, *dr = dl + out_doubles;
   HRW_CUDA(cudaMemcpy(dv, v, in_doubles * sizeof(double), cudaMemcpyHostToDevice));
   HRW_TRY(weno_reconstruct_launch(w, rows, dv, ldv, incv, dl, dr, n, nullptr));
   HRW_CUDA(cudaMemcpy2D(vl, (size_t)ldo * sizeof(double), dl, (size_t)n * sizeof(double), (size_t)n * sizeof(double),
                         (size_t)rows, cudaMemcpyDeviceToHost));
   HRW_CUDA(cudaMemcpy2D(vr, (size_t)ldo * sizeof(double), dr, (size_t)n * sizeof(double), (size_t)n * sizeof(double),
                         (size_t)rows, cudaMemcpyDeviceToHost));
   return HRWENO_OK;
}

int hrweno_weno_reconstruct(const hrweno_weno *h, const double *v, double *vl, double *vr) {
   const Weno *w = reinterpret_cast<const Weno *>(h);
   if (!w) return fail(HRWENO_EINVAL, "null weno handle");
   return hrweno_weno_reconstruct_batch(h, 1, v, w->ncells, 1, vl, vr, w->ncells);
}

void hrweno_weno_reconstruct_s(const hrweno_weno *h, const double *v, double *vl, double *vr, int *status) {
   const int st = hrweno_weno_reconstruct(h, v, vl, vr);
   if (status) *status = st;
}

// ---- fluxes ----------------------------------------------------------------------------------------
double hrweno_lax_friedrichs(hrweno_flux_fn f, void *ctx, double vm, double vp, const double *x, int nx, double t,
                             double alpha) {
   if (!f) { // the reference's dummy procedure cannot be absent; a C caller's pointer can
      fail(HRWENO_EINVAL, "hrweno_lax_friedrichs: null flux callback");
      return std::numeric_limits<double>::quiet_NaN();
   }
   return (f(ctx, vm, x, nx, t) + f(ctx, vp, x, nx, t) - alpha * (vp - vm)) / 2; // fluxes.f90:43
}

double hrweno_godunov(hrweno_flux_fn f, void *ctx, double vm, double vp, const double *x, int nx, double t) {
   if (!f) {
      fail(HRWENO_EINVAL, "hrweno_godunov: null flux callback");
      return std::numeric_limits<double>::quiet_NaN();
   }
   const double fm = f(ctx, vm, x, nx, t), fp = f(ctx, vp, x, nx, t); // fluxes.f90:67-68
   if (vm <= vp) return fm < fp ? fm : fp;                             // :70-74
   return fm > fp ? fm : fp;
}

int hrweno_flux_faces(int flux_scheme, int flux_model, double flux_coef, double alpha, int64_t n, const double *vm,
                      const double *vp, double *h) {
   if (n < 0 || !vm || !vp || !h) return fail(HRWENO_EINVAL, "invalid argument");
   if (n == 0) return HRWENO_OK;
   double *d = nullptr;
   HRW_CUDA(cudaMalloc(&d, 3 * (size_t)n * sizeof(double)));
   int st = HRWENO_OK;
   cudaError_t e = cudaMemcpy(d, vm, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
   if (e == cudaSuccess) e = cudaMemcpy(d + n, vp, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
   if (e == cudaSuccess) {
      int64_t blocks = (n + 255) / 256;
      if (blocks > 148 * 16) blocks = 148 * 16;
      flux_faces_kernel<<<(unsigned)blocks, 256>>>(FluxCfg{flux_model, flux_scheme, flux_coef, alpha}, n, d, d + n, d + 2 * n);
      e = cudaGetLastError();
   }
   if (e == cudaSuccess) e = cudaMemcpy(h, d + 2 * n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
   if (e != cudaSuccess) st = cuda_fail(e, "hrweno_flux_faces", __FILE__, __LINE__);
   cudaFree(d);
   return st;
}

// ---- fused finite-volume operator ------------------------------------------------------------------
int hrweno_fv_create(hrweno_fv **out, const hrweno_fv_desc *desc) { return fv_create(reinterpret_cast<Fv **>(out), desc); }

void hrweno_fv_destroy(hrweno_fv *fv) { delete reinterpret_cast<Fv *>(fv); }

int64_t hrweno_fv_neq(const hrweno_fv *fv) { return fv ? reinterpret_cast<const Fv *>(fv)->neq : 0; }

int hrweno_fv_rhs_dev(hrweno_fv *h, double t, const double *v_dev, double *vdot_dev, void *stream) {
   Fv *fv = reinterpret_cast<Fv *>(h); // t reaches the flux only through hrweno_fv_set_flux_time_fn (fluxes.f90:12-18)
   if (!fv || !v_dev || !vdot_dev) return fail(HRWENO_EINVAL, "null argument");
   std::lock_guard<std::mutex> lock(fv->mtx);
   return fv_rhs_any(fv, t, v_dev, vdot_dev, (cudaStream_t)stream);
}

int hrweno_fv_rhs(hrweno_fv *h, double t, const double *v, double *vdot) {
   Fv *fv = reinterpret_cast<Fv *>(h);
   if (!fv || !v || !vdot) return fail(HRWENO_EINVAL, "null argument");
   std::lock_guard<std::mutex> lock(fv->mtx);
   const size_t bytes = (size_t)fv->neq * sizeof(double);
   if (!fv->d_scratch_out) HRW_CUDA(cudaMalloc(&fv->d_scratch_out, 2 * bytes));
   double *din = fv->d_scratch_out, *dout = din + fv->neq;
   HRW_CUDA(cudaMemcpyAsync(din, v, bytes, cudaMemcpyHostToDevice, fv->stream));
   HRW_TRY(fv_rhs_any(fv, t, din, dout, fv->stream));
   HRW_CUDA(cudaMemcpyAsync(vdot, dout, bytes, cudaMemcpyDeviceToHost, fv->stream));
   HRW_CUDA(cudaStreamSynchronize(fv->stream));
   return fv_halo_status(fv);
}

int hrweno_fv_max_wavespeed_dev(hrweno_fv *h, const double *v_dev, double *out_dev, void *stream) {
   Fv *fv = reinterpret_cast<Fv *>(h);
   if (!fv || !v_dev || !out_dev) return fail(HRWENO_EINVAL, "null argument");
   std::lock_guard<std::mutex> lock(fv->mtx);
   return fv_max_wavespeed(fv, v_dev, out_dev, (cudaStream_t)stream);
}

int hrweno_fv_set_alpha(hrweno_fv *h, double alpha) {
   Fv *fv = reinterpret_cast<Fv *>(h);
   if (!fv) return fail(HRWENO_EINVAL, "null fv handle");
   if (!(alpha >= 0.0)) return fail(HRWENO_EINVAL, "Invalid input 'alpha'. Valid range: alpha >= 0.");
   std::lock_guard<std::mutex> lock(fv->mtx);
   fv->d.alpha = alpha; // read at every stage launch
   return HRWENO_OK;
}

int hrweno_fv_set_xedges(hrweno_fv *h, int axis, const double *xedges) {
   if (!h) return fail(HRWENO_EINVAL, "null fv handle");
   return fv_set_xedges(reinterpret_cast<Fv *>(h), axis, xedges);
}

int hrweno_fv_set_flux_coef(hrweno_fv *h, int axis, const double *face_coef, const double *cross_coef) {
   if (!h) return fail(HRWENO_EINVAL, "null fv handle");
   return fv_set_flux_coef(reinterpret_cast<Fv *>(h), axis, face_coef, cross_coef);
}

int hrweno_fv_set_flux_time_fn(hrweno_fv *h, hrweno_time_fn g, void *ctx) {
   if (!h) return fail(HRWENO_EINVAL, "null fv handle");
   return fv_set_flux_time_fn(reinterpret_cast<Fv *>(h), g, ctx);
}

int hrweno_fv_export_halo(hrweno_fv *h, void *handle_out) {
   if (!h) return fail(HRWENO_EINVAL, "null fv handle");
   return fv_halo_export(reinterpret_cast<Fv *>(h), handle_out);
}
int hrweno_fv_import_halo(hrweno_fv *h, const void *left_handle, const void *right_handle) {
   if (!h) return fail(HRWENO_EINVAL, "null fv handle");
   return fv_halo_import(reinterpret_cast<Fv *>(h), left_handle, right_handle);
}
int hrweno_fv_halo_status(hrweno_fv *h) {
   if (!h) return fail(HRWENO_EINVAL, "null fv handle");
   return fv_halo_status(reinterpret_cast<Fv *>(h));
}

// ---- integrators -----------------------------------------------------------------------------------
int hrweno_rktvd_create(hrweno_ode **out, hrweno_rhs_fn fu, void *ctx, int64_t neq, int order) {
   return ode_create(reinterpret_cast<Ode **>(out), false, nullptr, fu, ctx, neq, order);
}
int hrweno_mstvd_create(hrweno_ode **out, hrweno_rhs_fn fu, void *ctx, int64_t neq) {
   return ode_create(reinterpret_cast<Ode **>(out), true, nullptr, fu, ctx, neq, 3);
}
int hrweno_rktvd_create_host(hrweno_ode **out, hrweno_rhs_host_fn fu, void *ctx, int64_t neq, int order) {
   return ode_create_host(reinterpret_cast<Ode **>(out), false, fu, ctx, neq, order);
}
int hrweno_mstvd_create_host(hrweno_ode **out, hrweno_rhs_host_fn fu, void *ctx, int64_t neq) {
   return ode_create_host(reinterpret_cast<Ode **>(out), true, fu, ctx, neq, 3);
}
int hrweno_rktvd_create_fused(hrweno_ode **out, hrweno_fv *fv, int order) {
   if (!fv) return fail(HRWENO_EINVAL, "null fv handle");
   return ode_create(reinterpret_cast<Ode **>(out), false, reinterpret_cast<Fv *>(fv), nullptr, nullptr, 0, order);
}
int hrweno_mstvd_create_fused(hrweno_ode **out, hrweno_fv *fv) {
   if (!fv) return fail(HRWENO_EINVAL, "null fv handle");
   return ode_create(reinterpret_cast<Ode **>(out), true, reinterpret_cast<Fv *>(fv), nullptr, nullptr, 0, 3);
}
void hrweno_ode_destroy(hrweno_ode *ode) { delete reinterpret_cast<Ode *>(ode); }

int hrweno_ode_integrate(hrweno_ode *ode, double *u, double *t, double tout, double dt, int itask) {
   return ode_integrate_host(reinterpret_cast<Ode *>(ode), u, t, tout, dt, itask);
}
int hrweno_ode_integrate_dev(hrweno_ode *ode, double *u_dev, double *t, double tout, double dt, int itask, void *stream) {
   return ode_integrate_dev(reinterpret_cast<Ode *>(ode), u_dev, t, tout, dt, itask, (cudaStream_t)stream);
}
int hrweno_ode_attach(hrweno_ode *ode, const double *u_dev, void *stream) {
   return ode_attach(reinterpret_cast<Ode *>(ode), u_dev, (cudaStream_t)stream);
}
int hrweno_ode_integrate_attached(hrweno_ode *ode, double *t, double tout, double dt, int itask, void *stream) {
   return ode_integrate_dev(reinterpret_cast<Ode *>(ode), nullptr, t, tout, dt, itask, (cudaStream_t)stream);
}
int hrweno_ode_fetch(hrweno_ode *ode, double *u_dev, void *stream) {
   return ode_fetch(reinterpret_cast<Ode *>(ode), u_dev, (cudaStream_t)stream);
}
int64_t hrweno_ode_fevals(const hrweno_ode *ode) { return ode ? reinterpret_cast<const Ode *>(ode)->fevals : 0; }
int hrweno_ode_istate(const hrweno_ode *ode) { return ode ? reinterpret_cast<const Ode *>(ode)->istate : -1; }
int hrweno_ode_order(const hrweno_ode *ode) { return ode ? reinterpret_cast<const Ode *>(ode)->order : 0; }
int64_t hrweno_ode_neq(const hrweno_ode *ode) { return ode ? reinterpret_cast<const Ode *>(ode)->neq : 0; }
int64_t hrweno_ode_launches(const hrweno_ode *ode) {
   if (!ode) return 0;
   const Ode *o = reinterpret_cast<const Ode *>(ode);
   // fused path: every kernel (pack, stages, halo send/signal/recv, unpack) is counted by the fv operator
   return o->fused ? o->fv->launches : o->launches;
}

// ---- pinned host memory ----------------------------------------------------------------------------
int hrweno_host_alloc(void **out, int64_t bytes) {
   if (!out || bytes <= 0) return fail(HRWENO_EINVAL, "invalid argument");
   HRW_CUDA(cudaMallocHost(out, (size_t)bytes));
   return HRWENO_OK;
}
void hrweno_host_free(void *p) {
   if (p) cudaFreeHost(p);
}

} // extern "C"
