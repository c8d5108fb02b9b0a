// mgpu.cu -- one process, N GPUs: the slab decomposition of SURVEY 8e driven through the C ABI alone (hrweno_mgpu_*).
//
// The reference is a single-process program; a Fortran or C++ host that wants the 8 GPUs of a box should not
// need one process per GPU, a launcher and an out-of-band handle exchange.  hrweno_mgpu_create takes the GLOBAL
// operator description, cuts the slowest axis into contiguous slabs (independent rows are simply dealt out), creates
// one hrweno_fv per device, enables peer access and wires the halo mailboxes of neighbouring slabs directly (same address
// space: no IPC).  hrweno_mgpu_integrate runs the reference's `integrate` on every slab from one host thread per
// device; the per-stage halo traffic is the same in-kernel peer stores + sequence flags as in the one-process-per-GPU
// mode (fv1d.cuh / halo.cu), so the result is bit-identical to the single-domain run.  The Lax-Friedrichs alpha
// (extension, BASELINE north star) is reduced inside the library: per-slab max kernels, then one kernel on device 0
// that reads the partial maxima of its peers over NVLink.
#include <algorithm>
#include <cstring>
#include <memory>
#include <thread>

#include "internal.hpp"

namespace hrw {

int ode_create(Ode **out, bool is_ms, Fv *fv, hrweno_rhs_fn fu, void *ctx, int64_t neq, int order);
int ode_integrate_dev(Ode *o, double *u_dev, double *t, double tout, double dt, int itask, cudaStream_t st);
int ode_integrate_host(Ode *o, double *u, double *t, double tout, double dt, int itask);
const char *last_error_cstr();

struct Mgpu {
   int n = 0;
   bool split_rows = false;             // independent rows dealt out (no halos) instead of slabs of one grid
   std::vector<int> dev;
   std::vector<Fv *> fv;
   std::vector<Ode *> ode;
   std::vector<double *> u;             // resident dense state of each slab
   std::vector<double *> part;          // per-device scalar (partial maximum)
   std::vector<int64_t> off, len;       // offset / length of each slab in the caller's global vector (unknowns)
   std::vector<int64_t> soff, slen;     // offset / length of each slab along the decomposed axis (cells or rows)
   int ndim = 1;
   std::vector<cudaEvent_t> ev;
   double **d_ptrs = nullptr;           // device 0: table of the partial-maximum pointers
   double *d_max = nullptr;             // device 0: the reduced value
   bool resident = false;
   ~Mgpu();
};

Mgpu::~Mgpu() {
   for (int r = 0; r < (int)dev.size(); ++r) {
      cudaSetDevice(dev[r]);
      if (r < (int)ode.size()) delete ode[r];
      if (r < (int)fv.size()) delete fv[r];
      if (r < (int)u.size()) cudaFree(u[r]);
      if (r < (int)part.size()) cudaFree(part[r]);
      if (r < (int)ev.size() && ev[r]) cudaEventDestroy(ev[r]);
   }
   if (!dev.empty()) {
      cudaSetDevice(dev[0]);
      cudaFree(d_ptrs);
      cudaFree(d_max);
   }
}

// runs fn(rank) on one host thread per device with that device current; the first failure wins
template <class F>
static int for_each_slab(Mgpu *m, F fn) {
   std::vector<int> st((size_t)m->n, HRWENO_OK);
   std::vector<std::string> msg((size_t)m->n);
   auto body = [&](int r) {
      cudaError_t e = cudaSetDevice(m->dev[r]);
      if (e != cudaSuccess) {
         st[r] = HRWENO_ECUDA;
         msg[r] = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
         return;
      }
      st[r] = fn(r);
      if (st[r] != HRWENO_OK) msg[r] = last_error_cstr();
   };
   if (m->n == 1) {
      body(0);
   } else {
      std::vector<std::thread> th;
      for (int r = 0; r < m->n; ++r) th.emplace_back(body, r);
      for (auto &t : th) t.join();
   }
   for (int r = 0; r < m->n; ++r)
      if (st[r] != HRWENO_OK) return fail(st[r], "slab " + std::to_string(r) + " (device " + std::to_string(m->dev[r]) + "): " + msg[r]);
   return HRWENO_OK;
}

__global__ void mgpu_combine_max_kernel(double *const *parts, int n, double *out) {
   double mx = 0.0;
   for (int i = 0; i < n; ++i) mx = fmax(mx, *reinterpret_cast<const volatile double *>(parts[i])); // peer loads over NVLink
   *out = mx;
}

static int mgpu_create(Mgpu **out, const hrweno_fv_desc *desc, int ngpus, const int *devices) {
   if (!out || !desc) return fail(HRWENO_EINVAL, "hrweno_mgpu_create: null argument");
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(HRWENO_ECUDA, "no CUDA device (this library has no CPU fallback)");
   if (ngpus <= 0) ngpus = ndev;
   if (ngpus > ndev && !devices) return fail(HRWENO_EINVAL, "hrweno_mgpu_create: more GPUs requested than visible");
   if (desc->ndim != 1 && desc->ndim != 2) return fail(HRWENO_EINVAL, "hrweno_mgpu_create: ndim must be 1 or 2");
   std::unique_ptr<Mgpu> m(new Mgpu());
   const bool rows_mode = desc->ndim == 1 && desc->rows > 1;
   const int64_t nsplit = rows_mode ? desc->rows : desc->n[desc->ndim - 1];
   if (nsplit < ngpus) return fail(HRWENO_EINVAL, "hrweno_mgpu_create: fewer cells (rows) along the decomposed axis than GPUs");
   m->n = ngpus;
   m->ndim = desc->ndim;
   m->split_rows = rows_mode;
   for (int r = 0; r < ngpus; ++r) m->dev.push_back(devices ? devices[r] : r);
   // peer access between neighbouring slabs (and from device 0 to everyone for the reduction)
   for (int r = 0; r < ngpus; ++r) {
      HRW_CUDA(cudaSetDevice(m->dev[r]));
      for (int q = 0; q < ngpus; ++q) {
         if (q == r || m->dev[q] == m->dev[r]) continue;
         if (!(q == r - 1 || q == r + 1 || r == 0)) continue;
         int can = 0;
         HRW_CUDA(cudaDeviceCanAccessPeer(&can, m->dev[r], m->dev[q]));
         if (!can) return fail(HRWENO_ECOMM, "hrweno_mgpu_create: no peer access between devices " + std::to_string(m->dev[r]) + " and " + std::to_string(m->dev[q]));
         {  // peer access is a per-process, per-pair switch: enable each pair once (a second group reuses it)
            static std::mutex pm;
            static std::vector<std::pair<int, int>> enabled;
            std::lock_guard<std::mutex> lock(pm);
            const std::pair<int, int> pr(m->dev[r], m->dev[q]);
            if (std::find(enabled.begin(), enabled.end(), pr) == enabled.end()) {
               cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[q], 0);
               if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); // enabled by the caller's own code
               else if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
               enabled.push_back(pr);
            }
         }
      }
   }
   const int64_t base = nsplit / ngpus, rem = nsplit % ngpus;
   const int64_t unit = rows_mode ? desc->n[0] : (desc->ndim == 2 ? desc->n[0] : 1); // unknowns per split index
   m->fv.assign((size_t)ngpus, nullptr);
   m->ode.assign((size_t)ngpus, nullptr);
   m->u.assign((size_t)ngpus, nullptr);
   m->part.assign((size_t)ngpus, nullptr);
   m->ev.assign((size_t)ngpus, nullptr);
   for (int r = 0; r < ngpus; ++r) {
      const int64_t nloc = base + (r < rem ? 1 : 0), o = (int64_t)r * base + std::min<int64_t>(r, rem);
      m->off.push_back(o * unit);
      m->len.push_back(nloc * unit);
      m->soff.push_back(o);
      m->slen.push_back(nloc);
      hrweno_fv_desc d = *desc;
      if (rows_mode) {
         d.rows = nloc;
         d.rank = 0;
         d.nranks = 1;
      } else {
         const int ax = desc->ndim - 1;
         d.n[ax] = nloc;
         d.rank = r;
         d.nranks = ngpus;
         d.global_n = nsplit;
         d.global_offset = o;
         if (d.grid_kind == HRWENO_GRID_WIDTH_ARRAY && desc->width[ax]) d.width[ax] = desc->width[ax] + o;
      }
      HRW_CUDA(cudaSetDevice(m->dev[r]));
      HRW_TRY(fv_create(&m->fv[r], &d));
      HRW_CUDA(cudaMalloc(&m->u[r], (size_t)m->len[r] * sizeof(double)));
      HRW_CUDA(cudaMalloc(&m->part[r], sizeof(double)));
      HRW_CUDA(cudaMemset(m->part[r], 0, sizeof(double)));
      HRW_CUDA(cudaEventCreateWithFlags(&m->ev[r], cudaEventDisableTiming));
      if (!rows_mode && ngpus > 1) HRW_TRY(fv_halo_prepare_local(m->fv[r]));
   }
   if (!rows_mode && ngpus > 1)
      for (int r = 0; r < ngpus; ++r) {
         HRW_CUDA(cudaSetDevice(m->dev[r]));
         HRW_TRY(fv_halo_connect_local(m->fv[r], r > 0 ? m->fv[r - 1] : nullptr, r < ngpus - 1 ? m->fv[r + 1] : nullptr));
      }
   HRW_CUDA(cudaSetDevice(m->dev[0]));
   HRW_CUDA(cudaMalloc(&m->d_ptrs, (size_t)ngpus * sizeof(double *)));
   HRW_CUDA(cudaMemcpy(m->d_ptrs, m->part.data(), (size_t)ngpus * sizeof(double *), cudaMemcpyHostToDevice));
   HRW_CUDA(cudaMalloc(&m->d_max, sizeof(double)));
   *out = m.release();
   return HRWENO_OK;
}

// general operators on slabs: the global arrays are cut like the grid.  Along the decomposed axis a slab takes the GLOBAL
// edge array (fv_set_xedges derives the tables of its cells and of the two neighbour cells from it) and its own part of the
// face / cross coefficients; arrays along the other axis are shared.
static int mgpu_set_xedges(Mgpu *m, int axis, const double *xedges) {
   if (!m || !xedges) return fail(HRWENO_EINVAL, "hrweno_mgpu_set_xedges: null argument");
   for (int r = 0; r < m->n; ++r) {
      HRW_CUDA(cudaSetDevice(m->dev[r]));
      const bool dec = !m->split_rows && axis == m->ndim - 1;
      // nranks > 1: global edges on the decomposed axis; a single slab (or rows mode) has local == global
      (void)dec;
      HRW_TRY(fv_set_xedges(m->fv[r], axis, xedges));
   }
   return HRWENO_OK;
}

static int mgpu_set_flux_coef(Mgpu *m, int axis, const double *face, const double *cross) {
   if (!m) return fail(HRWENO_EINVAL, "null mgpu handle");
   for (int r = 0; r < m->n; ++r) {
      HRW_CUDA(cudaSetDevice(m->dev[r]));
      const bool dec_axis = !m->split_rows && axis == m->ndim - 1;        // faces of the decomposed axis: this slab's n+1 faces
      const bool dec_cross = !m->split_rows && m->ndim == 2 && axis == 0; // cross index runs along the decomposed axis (x2 cells)
      HRW_TRY(fv_set_flux_coef(m->fv[r], axis, face ? face + (dec_axis ? m->soff[r] : 0) : nullptr,
                               cross ? cross + (dec_cross ? m->soff[r] : 0) : nullptr));
   }
   return HRWENO_OK;
}

static int mgpu_make_ode(Mgpu *m, bool is_ms, int order) {
   if (!m) return fail(HRWENO_EINVAL, "null mgpu handle");
   for (int r = 0; r < m->n; ++r) {
      HRW_CUDA(cudaSetDevice(m->dev[r]));
      delete m->ode[r];
      m->ode[r] = nullptr;
      HRW_TRY(ode_create(&m->ode[r], is_ms, m->fv[r], nullptr, nullptr, 0, order));
   }
   return HRWENO_OK;
}

static int mgpu_copy(Mgpu *m, double *u, bool to_device) {
   return for_each_slab(m, [&](int r) -> int {
      cudaStream_t st = m->fv[r]->stream;
      if (to_device)
         HRW_CUDA(cudaMemcpyAsync(m->u[r], u + m->off[r], (size_t)m->len[r] * sizeof(double), cudaMemcpyHostToDevice, st));
      else
         HRW_CUDA(cudaMemcpyAsync(u + m->off[r], m->u[r], (size_t)m->len[r] * sizeof(double), cudaMemcpyDeviceToHost, st));
      HRW_CUDA(cudaStreamSynchronize(st));
      return HRWENO_OK;
   });
}

// the reference's integrate on every slab at once; host = true: u is the caller's GLOBAL host vector (each slab runs the
// host-pointer entry point on its part, including its chunk pipeline where that applies)
static int mgpu_integrate(Mgpu *m, double *u, double *t, double tout, double dt, int itask, bool host) {
   if (!m || !t) return fail(HRWENO_EINVAL, "hrweno_mgpu_integrate: null argument");
   for (int r = 0; r < m->n; ++r)
      if (!m->ode[r]) return fail(HRWENO_ESTATE, "hrweno_mgpu_integrate: create the integrator first (hrweno_mgpu_rktvd / _mstvd)");
   if (host && !u) return fail(HRWENO_EINVAL, "hrweno_mgpu_integrate: null u");
   if (!host && !m->resident) return fail(HRWENO_ESTATE, "hrweno_mgpu_integrate_resident: no resident state (hrweno_mgpu_upload first)");
   std::vector<double> tr((size_t)m->n, *t);
   HRW_TRY(for_each_slab(m, [&](int r) -> int {
      if (host) return ode_integrate_host(m->ode[r], u + m->off[r], &tr[r], tout, dt, itask);
      cudaStream_t st = m->fv[r]->stream;
      HRW_TRY(ode_integrate_dev(m->ode[r], m->u[r], &tr[r], tout, dt, itask, st));
      HRW_CUDA(cudaStreamSynchronize(st));
      return fv_halo_status(m->fv[r]);
   }));
   *t = tr[0]; // every slab advances the same scalar sequence t = t + dt (tvdode.f90:168)
   return HRWENO_OK;
}

static int mgpu_max_wavespeed(Mgpu *m, double *alpha_out, bool install) {
   if (!m) return fail(HRWENO_EINVAL, "null mgpu handle");
   if (!m->resident) return fail(HRWENO_ESTATE, "hrweno_mgpu_max_wavespeed: no resident state (hrweno_mgpu_upload first)");
   for (int r = 0; r < m->n; ++r) {
      HRW_CUDA(cudaSetDevice(m->dev[r]));
      cudaStream_t st = m->fv[r]->stream;
      HRW_TRY(fv_max_wavespeed(m->fv[r], m->u[r], m->part[r], st));
      HRW_CUDA(cudaEventRecord(m->ev[r], st));
   }
   HRW_CUDA(cudaSetDevice(m->dev[0]));
   cudaStream_t s0 = m->fv[0]->stream;
   for (int r = 1; r < m->n; ++r) HRW_CUDA(cudaStreamWaitEvent(s0, m->ev[r], 0));
   mgpu_combine_max_kernel<<<1, 1, 0, s0>>>(m->d_ptrs, m->n, m->d_max);
   HRW_CUDA(cudaGetLastError());
   double a = 0.0;
   HRW_CUDA(cudaMemcpyAsync(&a, m->d_max, sizeof(double), cudaMemcpyDeviceToHost, s0));
   HRW_CUDA(cudaStreamSynchronize(s0));
   if (alpha_out) *alpha_out = a;
   if (install)
      for (int r = 0; r < m->n; ++r) m->fv[r]->d.alpha = a;
   return HRWENO_OK;
}

} // namespace hrw

using namespace hrw;

extern "C" {

int hrweno_mgpu_create(hrweno_mgpu **out, const hrweno_fv_desc *desc, int ngpus, const int *devices) {
   int prev = 0;
   cudaGetDevice(&prev);
   const int st = mgpu_create(reinterpret_cast<Mgpu **>(out), desc, ngpus, devices);
   cudaSetDevice(prev);
   return st;
}
void hrweno_mgpu_destroy(hrweno_mgpu *h) {
   int prev = 0;
   cudaGetDevice(&prev);
   delete reinterpret_cast<Mgpu *>(h);
   cudaSetDevice(prev);
}
int hrweno_mgpu_ngpus(const hrweno_mgpu *h) { return h ? reinterpret_cast<const Mgpu *>(h)->n : 0; }
int hrweno_mgpu_slab(const hrweno_mgpu *h, int rank, int *device, int64_t *offset, int64_t *count) {
   const Mgpu *m = reinterpret_cast<const Mgpu *>(h);
   if (!m || rank < 0 || rank >= m->n) return fail(HRWENO_EINVAL, "hrweno_mgpu_slab: invalid handle or rank");
   if (device) *device = m->dev[rank];
   if (offset) *offset = m->off[rank];
   if (count) *count = m->len[rank];
   return HRWENO_OK;
}
int hrweno_mgpu_set_xedges(hrweno_mgpu *h, int axis, const double *xedges) {
   int prev = 0;
   cudaGetDevice(&prev);
   const int st = mgpu_set_xedges(reinterpret_cast<Mgpu *>(h), axis, xedges);
   cudaSetDevice(prev);
   return st;
}
int hrweno_mgpu_set_flux_coef(hrweno_mgpu *h, int axis, const double *face_coef, const double *cross_coef) {
   int prev = 0;
   cudaGetDevice(&prev);
   const int st = mgpu_set_flux_coef(reinterpret_cast<Mgpu *>(h), axis, face_coef, cross_coef);
   cudaSetDevice(prev);
   return st;
}
int hrweno_mgpu_set_flux_time_fn(hrweno_mgpu *h, hrweno_time_fn g, void *ctx) {
   Mgpu *m = reinterpret_cast<Mgpu *>(h);
   if (!m) return fail(HRWENO_EINVAL, "null mgpu handle");
   for (int r = 0; r < m->n; ++r) {
      cudaSetDevice(m->dev[r]);
      HRW_TRY(fv_set_flux_time_fn(m->fv[r], g, ctx));
   }
   return HRWENO_OK;
}
int hrweno_mgpu_rktvd(hrweno_mgpu *h, int order) {
   int prev = 0;
   cudaGetDevice(&prev);
   const int st = mgpu_make_ode(reinterpret_cast<Mgpu *>(h), false, order);
   cudaSetDevice(prev);
   return st;
}
int hrweno_mgpu_mstvd(hrweno_mgpu *h) {
   int prev = 0;
   cudaGetDevice(&prev);
   const int st = mgpu_make_ode(reinterpret_cast<Mgpu *>(h), true, 3);
   cudaSetDevice(prev);
   return st;
}
int hrweno_mgpu_integrate(hrweno_mgpu *h, double *u, double *t, double tout, double dt, int itask) {
   return mgpu_integrate(reinterpret_cast<Mgpu *>(h), u, t, tout, dt, itask, true);
}
int hrweno_mgpu_upload(hrweno_mgpu *h, const double *u) {
   Mgpu *m = reinterpret_cast<Mgpu *>(h);
   if (!m || !u) return fail(HRWENO_EINVAL, "hrweno_mgpu_upload: null argument");
   HRW_TRY(mgpu_copy(m, const_cast<double *>(u), true));
   m->resident = true;
   return HRWENO_OK;
}
int hrweno_mgpu_download(hrweno_mgpu *h, double *u) {
   Mgpu *m = reinterpret_cast<Mgpu *>(h);
   if (!m || !u) return fail(HRWENO_EINVAL, "hrweno_mgpu_download: null argument");
   if (!m->resident) return fail(HRWENO_ESTATE, "hrweno_mgpu_download: no resident state");
   return mgpu_copy(m, u, false);
}
int hrweno_mgpu_integrate_resident(hrweno_mgpu *h, double *t, double tout, double dt, int itask) {
   return mgpu_integrate(reinterpret_cast<Mgpu *>(h), nullptr, t, tout, dt, itask, false);
}
int hrweno_mgpu_max_wavespeed(hrweno_mgpu *h, double *alpha_out, int install) {
   int prev = 0;
   cudaGetDevice(&prev);
   const int st = mgpu_max_wavespeed(reinterpret_cast<Mgpu *>(h), alpha_out, install != 0);
   cudaSetDevice(prev);
   return st;
}
int hrweno_mgpu_set_alpha(hrweno_mgpu *h, double alpha) {
   Mgpu *m = reinterpret_cast<Mgpu *>(h);
   if (!m) return fail(HRWENO_EINVAL, "null mgpu handle");
   if (!(alpha >= 0.0)) return fail(HRWENO_EINVAL, "Invalid input 'alpha'. Valid range: alpha >= 0.");
   for (int r = 0; r < m->n; ++r) m->fv[r]->d.alpha = alpha;
   return HRWENO_OK;
}
int64_t hrweno_mgpu_fevals(const hrweno_mgpu *h) {
   const Mgpu *m = reinterpret_cast<const Mgpu *>(h);
   return (m && !m->ode.empty() && m->ode[0]) ? m->ode[0]->fevals : 0;
}
int64_t hrweno_mgpu_launches(const hrweno_mgpu *h) {
   const Mgpu *m = reinterpret_cast<const Mgpu *>(h);
   int64_t s = 0;
   if (m)
      for (Fv *f : m->fv) s += f ? f->launches : 0;
   return s;
}

} // extern "C"
