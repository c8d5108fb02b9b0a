// hrweno.hpp -- C++ host-side mirror of HR-WENO's Fortran modules over the C ABI (include/hrweno_b200.h).
//
// The reference's host language is Fortran; no Fortran compiler exists in this image, so the host side that
// sits above the C ABI is written in C++ with the reference's module / type / procedure names and argument
// meaning (the ISO_C_BINDING modules a Fortran program would use instead are in fortran/).
//   namespace hrweno::hrweno_grids   { class grid1 }                    src/hrweno_grids.f90   (host-side set-up)
//   namespace hrweno::hrweno_weno    { class weno }                     src/hrweno_weno.f90:23-50
//   namespace hrweno::hrweno_fluxes  { godunov, lax_friedrichs }        src/hrweno_fluxes.f90:22,47
//   namespace hrweno::hrweno_tvdode  { class rktvd, class mstvd }       src/hrweno_tvdode.f90:36-48
//   namespace hrweno::hrweno_fv      { class fv }                       the example `rhs` as one fused device operator
//   namespace hrweno::hrweno_multi   { class mgpu }                     the same operator + integrators on all GPUs of the box
//   namespace hrweno::real32         { weno, fv, rktvd, mstvd }         the REAL32 build (src/hrweno_kinds.F90:9-10)
// Errors: where the reference executes `error stop msg`, these throw hrweno::error carrying the same message.
#pragma once

#include <cmath>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "hrweno_b200.h"

namespace hrweno {
struct error : std::runtime_error {
   int status;
   error(int st, const std::string &m) : std::runtime_error(m), status(st) {}
};
inline void check(int st) {
   if (st != HRWENO_OK) throw error(st, hrweno_last_error());
}
} // namespace hrweno

namespace hrweno {
namespace hrweno_grids {
// type(grid1), grids.f90:10-37 -- only what the finite-volume path consumes
class grid1 {
 public:
   int64_t ncells = 0;
   std::string scale, name;
   std::vector<double> edges, center, width; // edges(0:ncells); left = edges[0..n-1], right = edges[1..n]
   const double *left() const { return edges.data(); }
   const double *right() const { return edges.data() + 1; }
   void linear(double xmin, double xmax, int64_t n, const std::string &nm = "") { // grids.f90:41-84
      if (xmax <= xmin) throw hrweno::error(HRWENO_EINVAL, "Invalid input 'xmin', 'xmax'. Valid range: xmax > xmin.");
      if (n < 1) throw hrweno::error(HRWENO_EINVAL, "Invalid input 'ncells'. Valid range: ncells > 1.");
      std::vector<double> xe((size_t)n + 1);
      const double rx = (xmax - xmin) / (double)n;
      for (int64_t i = 0; i <= n; ++i) {
         volatile double p = rx * (double)i; // keep mul and add separately rounded (no contraction)
         xe[(size_t)i] = xmin + p;
      }
      scale = "linear";
      compute(xe, nm);
   }
   void bilinear(double xmin, double xcross, double xmax, const int64_t (&n)[2], const std::string &nm = "") { // grids.f90:86-136
      if (xcross <= xmin) throw hrweno::error(HRWENO_EINVAL, "Invalid input 'xmin', 'xcross'. Valid range: xcross > xmin.");
      if (xmax <= xcross) throw hrweno::error(HRWENO_EINVAL, "Invalid input 'xcross', 'xmax'. Valid range: xmax > xcross.");
      if (n[0] < 1 || n[1] < 1) throw hrweno::error(HRWENO_EINVAL, "Invalid input 'ncells'. Valid range: ncells(i) >= 1.");
      std::vector<double> xe((size_t)(n[0] + n[1]) + 1);
      double rx = (xcross - xmin) / (double)n[0];
      for (int64_t i = 0; i <= n[0]; ++i) {
         volatile double p = rx * (double)i;
         xe[(size_t)i] = xmin + p;
      }
      rx = (xmax - xcross) / (double)n[1];
      for (int64_t i = 1; i <= n[1]; ++i) {
         volatile double p = rx * (double)i;
         xe[(size_t)(n[0] + i)] = xcross + p;
      }
      scale = "bilinear";
      compute(xe, nm);
   }
   void log(double xmin, double xmax, int64_t n, const std::string &nm = "") { // grids.f90:138-180
      if (xmin <= 0.0) throw hrweno::error(HRWENO_EINVAL, "Invalid input 'xmin'. Valid range: xmin > 0.");
      if (!(xmax > xmin)) throw hrweno::error(HRWENO_EINVAL, "Invalid input 'xmin', 'xmax'. Valid range: xmax > xmin.");
      if (n < 1) throw hrweno::error(HRWENO_EINVAL, "Invalid input 'ncells'. Valid range: ncells > 1.");
      std::vector<double> xe((size_t)n + 1);
      const double rx = std::log(xmax / xmin) / (double)n, l0 = std::log(xmin);
      for (int64_t i = 0; i <= n; ++i) {
         volatile double p = rx * (double)i;
         xe[(size_t)i] = std::exp(l0 + p);
      }
      scale = "log";
      compute(xe, nm);
   }
   void geometric(double xmin, double xmax, double ratio, int64_t n, const std::string &nm = "") { // grids.f90:182-230
      if (xmax <= xmin) throw hrweno::error(HRWENO_EINVAL, "Invalid input 'xmin', 'xmax'. Valid range: xmax > xmin");
      if (ratio <= 0.0) throw hrweno::error(HRWENO_EINVAL, "Invalid input 'ratio'. Valid range: ratio > 0");
      if (n < 1) throw hrweno::error(HRWENO_EINVAL, "Invalid input 'ncells'. Valid range: ncells > 1");
      std::vector<double> xe((size_t)n + 1);
      const double a = (xmax - xmin) / (powi(ratio, n) - 1.0);
      for (int64_t i = 0; i <= n; ++i) {
         volatile double p = a * (powi(ratio, i) - 1.0);
         xe[(size_t)i] = xmin + p;
      }
      scale = "geometric";
      compute(xe, nm);
   }
   // real**integer as gfortran evaluates it (binary exponentiation, low bit first: libgcc __powidf2 / libgfortran pow_r8_i4)
   static double powi(double x, int64_t m) {
      uint64_t nn = (uint64_t)(m < 0 ? -m : m);
      volatile double y = (nn & 1) ? x : 1.0;
      volatile double xx = x;
      while (nn >>= 1) {
         xx = xx * xx;
         if (nn & 1) y = y * xx;
      }
      return m < 0 ? 1.0 / y : y;
   }
 private:
   void compute(const std::vector<double> &xe, const std::string &nm) { // grids.f90:232-250
      ncells = (int64_t)xe.size() - 1;
      edges = xe;
      center.resize((size_t)ncells);
      width.resize((size_t)ncells);
      for (int64_t i = 0; i < ncells; ++i) {
         center[(size_t)i] = (edges[(size_t)i] + edges[(size_t)i + 1]) / 2;
         width[(size_t)i] = edges[(size_t)i + 1] - edges[(size_t)i];
      }
      name = nm;
   }
};
} // namespace hrweno_grids

namespace hrweno_weno {
class weno { // weno.f90:23-50
 public:
   std::string msg;
   int ierr = 0;
   int64_t ncells = 0;
   int k = 3;
   double eps = 1e-6;
   weno() = default;
   weno(int64_t ncells_, int k_ = 3, double eps_ = 1e-6, const std::vector<double> *xedges = nullptr) { // weno_init, :54-127
      if (xedges && (int64_t)xedges->size() != ncells_ + 1) fail("Invalid input 'xedges': size(xedges) /= ncells + 1.");
      const int st = hrweno_weno_create(&h_, ncells_, k_, eps_, xedges ? xedges->data() : nullptr);
      if (st != HRWENO_OK) fail(hrweno_last_error(), st);
      ncells = ncells_;
      k = k_;
      eps = eps_;
   }
   weno(weno &&o) noexcept { *this = std::move(o); }
   weno &operator=(weno &&o) noexcept { // `myweno = weno(...)` (example1:44) -- move of the function result
      if (this != &o) {
         hrweno_weno_destroy(h_);
         h_ = o.h_;
         o.h_ = nullptr;
         msg = o.msg; ierr = o.ierr; ncells = o.ncells; k = o.k; eps = o.eps;
      }
      return *this;
   }
   weno(const weno &) = delete;
   weno &operator=(const weno &) = delete;
   ~weno() { hrweno_weno_destroy(h_); }
   // call w%reconstruct(v, vl, vr)  (weno.f90:129-219)
   void reconstruct(const double *v, double *vl, double *vr) const { hrweno::check(hrweno_weno_reconstruct(h_, v, vl, vr)); }
   // strided section v(i::inc) (example2:107)
   void reconstruct(const double *v, int64_t inc, double *vl, double *vr) const {
      hrweno::check(hrweno_weno_reconstruct_batch(h_, 1, v, 0, inc, vl, vr, ncells));
   }
   std::vector<double> cnu() const { // cnu(0:k-1,-1:k-1,1:ncells), weno.f90:41
      std::vector<double> c((size_t)(k * (k + 1) * ncells));
      hrweno::check(hrweno_weno_get_cnu(h_, c.data()));
      return c;
   }
   const ::hrweno_weno *handle() const { return h_; }
 private:
   ::hrweno_weno *h_ = nullptr;
   [[noreturn]] void fail(const std::string &m, int st = HRWENO_EINVAL) {
      msg = m;
      ierr = 1;
      throw hrweno::error(st, m);
   }
};
} // namespace hrweno_weno

namespace hrweno_fluxes {
using flux = std::function<double(double u, const std::vector<double> &x, double t)>; // fluxes.f90:12-18
namespace detail {
struct Ctx { const flux *f; };
inline double tramp(void *ctx, double u, const double *x, int nx, double t) {
   return (*static_cast<Ctx *>(ctx)->f)(u, std::vector<double>(x, x + nx), t);
}
} // namespace detail
inline double lax_friedrichs(const flux &f, double vm, double vp, const std::vector<double> &x, double t, double alpha) {
   detail::Ctx c{&f};
   return hrweno_lax_friedrichs(detail::tramp, &c, vm, vp, x.data(), (int)x.size(), t, alpha); // fluxes.f90:22-45
}
inline double godunov(const flux &f, double vm, double vp, const std::vector<double> &x, double t) {
   detail::Ctx c{&f};
   return hrweno_godunov(detail::tramp, &c, vm, vp, x.data(), (int)x.size(), t); // fluxes.f90:47-76
}
} // namespace hrweno_fluxes

namespace hrweno_fv {
// the example `rhs` as one fused device operator (example1:72-109, example2:73-129)
class fv {
 public:
   explicit fv(const hrweno_fv_desc &d) { hrweno::check(hrweno_fv_create(&h_, &d)); }
   fv(const fv &) = delete;
   fv &operator=(const fv &) = delete;
   ~fv() { hrweno_fv_destroy(h_); }
   int64_t neq() const { return hrweno_fv_neq(h_); }
   void rhs(double t, const double *v, double *vdot) { hrweno::check(hrweno_fv_rhs(h_, t, v, vdot)); }
   // weno(ncells, k, eps, xedges) for the sweep along `axis` (weno.f90:100-112,177): xedges(0:n[axis]) = grid1%edges
   void set_xedges(int axis, const double *xedges) { hrweno::check(hrweno_fv_set_xedges(h_, axis, xedges)); }
   // x-dependent flux f = (model(v)*cross[c])*face[f] along `axis` (example2:100-101,109-110,140,153); nullptr = absent
   void set_flux_coef(int axis, const double *face, const double *cross = nullptr) {
      hrweno::check(hrweno_fv_set_flux_coef(h_, axis, face, cross));
   }
   // t-dependent flux f(v, x, t) = f(v, x)*g(t) (fluxes.f90:12-18 passes t to the flux); an empty function removes it
   void set_flux_time_fn(std::function<double(double)> g) {
      tfn_ = std::move(g);
      hrweno::check(hrweno_fv_set_flux_time_fn(h_, tfn_ ? &fv::tramp_time : nullptr, this));
   }
   // device-resident callers
   void rhs_dev(double t, const double *v_dev, double *vdot_dev, void *stream = nullptr) {
      hrweno::check(hrweno_fv_rhs_dev(h_, t, v_dev, vdot_dev, stream));
   }
   void max_wavespeed_dev(const double *v_dev, double *out_dev, void *stream = nullptr) {
      hrweno::check(hrweno_fv_max_wavespeed_dev(h_, v_dev, out_dev, stream));
   }
   void set_alpha(double alpha) { hrweno::check(hrweno_fv_set_alpha(h_, alpha)); } // lax_friedrichs' alpha, fluxes.f90:22-45
   ::hrweno_fv *handle() const { return h_; }
   static hrweno_fv_desc desc1d(int64_t nc, int k, double eps, const double *width) {
      hrweno_fv_desc d{};
      d.abi_version = HRWENO_ABI_VERSION;
      d.ndim = 1; d.n[0] = nc; d.n[1] = 1; d.rows = 1; d.k = k; d.eps = eps;
      d.flux_model = HRWENO_FLUX_BURGERS; d.flux_scheme = HRWENO_SCHEME_GODUNOV; d.bc = HRWENO_BC_COPY_NEIGHBOUR;
      d.grid_kind = HRWENO_GRID_WIDTH_ARRAY; d.mode = HRWENO_MODE_STRICT; d.flux_coef[0] = d.flux_coef[1] = 1.0; d.alpha = 1.0;
      d.width[0] = width; d.nranks = 1;
      return d;
   }
   static hrweno_fv_desc desc2d(int64_t nc1, int64_t nc2, int k, double eps, const double *w1, const double *w2) {
      hrweno_fv_desc d = desc1d(nc1, k, eps, w1);
      d.ndim = 2; d.n[1] = nc2; d.width[1] = w2;
      d.flux_model = HRWENO_FLUX_LINEAR; d.bc = HRWENO_BC_ZERO_FLUX;
      return d;
   }
 private:
   ::hrweno_fv *h_ = nullptr;
   std::function<double(double)> tfn_;
   static double tramp_time(void *ctx, double t) { return static_cast<fv *>(ctx)->tfn_(t); }
};
} // namespace hrweno_fv

namespace hrweno_tvdode {
// integrand(t, u(:), udot(:)) with device-resident u/udot (tvdode.f90:50-57)
using integrand = std::function<void(double t, int64_t neq, const double *u_dev, double *udot_dev, void *stream)>;
// the reference's integrand as it is: host arrays (the library stages u and udot in pinned memory around every call)
using host_integrand = std::function<void(double t, int64_t neq, const double *u, double *udot)>;

class tvdode { // tvdode.f90:14-34
 public:
   std::string msg;
   int64_t neq() const { return hrweno_ode_neq(h_); }
   int order() const { return hrweno_ode_order(h_); }
   int64_t fevals() const { return hrweno_ode_fevals(h_); }
   int istate() const { return h_ ? hrweno_ode_istate(h_) : -1; }
   int64_t launches() const { return hrweno_ode_launches(h_); }
   // device-resident state: u_dev is a dense device vector, integrated in place
   void integrate_dev(double *u_dev, double &t, double tout, double dt, int itask = 1, void *stream = nullptr) {
      hrweno::check(hrweno_ode_integrate_dev(h_, u_dev, &t, tout, dt, itask, stream));
   }
   // fused integrators: hand the state over once, integrate from output time to output time without the dense <-> padded
   // copies of every call, fetch it when output is due
   void attach(const double *u_dev, void *stream = nullptr) { hrweno::check(hrweno_ode_attach(h_, u_dev, stream)); }
   void integrate_attached(double &t, double tout, double dt, int itask = 1, void *stream = nullptr) {
      hrweno::check(hrweno_ode_integrate_attached(h_, &t, tout, dt, itask, stream));
   }
   void fetch(double *u_dev, void *stream = nullptr) { hrweno::check(hrweno_ode_fetch(h_, u_dev, stream)); }
   tvdode(const tvdode &) = delete;
   tvdode &operator=(const tvdode &) = delete;
   virtual ~tvdode() { hrweno_ode_destroy(h_); }
 protected:
   tvdode() = default;
   hrweno_ode *h_ = nullptr;
   integrand fu_;
   host_integrand fu_host_;
   static void tramp_host(void *ctx, double t, int64_t n, const double *u, double *udot) {
      static_cast<tvdode *>(ctx)->fu_host_(t, n, u, udot);
   }
   static void tramp(void *ctx, double t, int64_t n, const double *u, double *udot, void *stream) {
      static_cast<tvdode *>(ctx)->fu_(t, n, u, udot, stream);
   }
   void created(int st) {
      if (st != HRWENO_OK) {
         msg = hrweno_last_error(); // error_msg: istate = -1 then error stop (tvdode.f90:286-297)
         throw hrweno::error(st, msg);
      }
   }
};

class rktvd : public tvdode { // rktvd(fu, neq, order), tvdode.f90:69-95
 public:
   rktvd(integrand fu, int64_t neq, int order) {
      fu_ = std::move(fu);
      created(hrweno_rktvd_create(&h_, &tvdode::tramp, this, neq, order));
   }
   rktvd(host_integrand fu, int64_t neq, int order) { // the reference's own call: a host integrand
      fu_host_ = std::move(fu);
      created(hrweno_rktvd_create_host(&h_, &tvdode::tramp_host, this, neq, order));
   }
   rktvd(hrweno_fv::fv &rhs, int64_t neq, int order) { // fused path: rhs + stage combination in one kernel
      if (neq != rhs.neq()) throw hrweno::error(HRWENO_EINVAL, "neq does not match the finite-volume operator");
      created(hrweno_rktvd_create_fused(&h_, rhs.handle(), order));
   }
   // call ode%integrate(u, t, tout, dt [, itask])  (tvdode.f90:97-178)
   void integrate(double *u, double &t, double tout, double dt, int itask = 1) {
      hrweno::check(hrweno_ode_integrate(h_, u, &t, tout, dt, itask));
   }
};

class mstvd : public tvdode { // mstvd(fu, neq), tvdode.f90:180-201
 public:
   mstvd(host_integrand fu, int64_t neq) { // the reference's own call: a host integrand
      fu_host_ = std::move(fu);
      created(hrweno_mstvd_create_host(&h_, &tvdode::tramp_host, this, neq));
   }
   mstvd(integrand fu, int64_t neq) {
      fu_ = std::move(fu);
      created(hrweno_mstvd_create(&h_, &tvdode::tramp, this, neq));
   }
   mstvd(hrweno_fv::fv &rhs, int64_t neq) {
      if (neq != rhs.neq()) throw hrweno::error(HRWENO_EINVAL, "neq does not match the finite-volume operator");
      created(hrweno_mstvd_create_fused(&h_, rhs.handle()));
   }
   void integrate(double *u, double &t, double tout, double dt) { // tvdode.f90:203-271
      hrweno::check(hrweno_ode_integrate(h_, u, &t, tout, dt, 1));
   }
};
} // namespace hrweno_tvdode

namespace hrweno_multi {
// The fused operator and its integrator on every GPU of the box from ONE process (hrweno_mgpu_*): the grid is cut into
// slabs (1D: along x; 2D: along x2; ensembles: rows), u is the caller's global vector in the reference's storage order.
class mgpu {
 public:
   // ngpus <= 0: all visible devices; devices == nullptr: 0 .. ngpus-1
   explicit mgpu(const hrweno_fv_desc &d, int ngpus = 0, const int *devices = nullptr) {
      hrweno::check(hrweno_mgpu_create(&h_, &d, ngpus, devices));
   }
   mgpu(const mgpu &) = delete;
   mgpu &operator=(const mgpu &) = delete;
   ~mgpu() { hrweno_mgpu_destroy(h_); }
   int ngpus() const { return hrweno_mgpu_ngpus(h_); }
   struct slab_info {
      int device;
      int64_t offset, count; // unknowns [offset, offset + count) of the global vector
   };
   slab_info slab(int rank) const {
      slab_info s{};
      hrweno::check(hrweno_mgpu_slab(h_, rank, &s.device, &s.offset, &s.count));
      return s;
   }
   // general operators, GLOBAL arrays (weno(..., xedges); example2:140,153; fluxes.f90:12-18)
   void set_xedges(int axis, const double *xedges) { hrweno::check(hrweno_mgpu_set_xedges(h_, axis, xedges)); }
   void set_flux_coef(int axis, const double *face, const double *cross = nullptr) {
      hrweno::check(hrweno_mgpu_set_flux_coef(h_, axis, face, cross));
   }
   void set_flux_time_fn(std::function<double(double)> g) {
      tfn_ = std::move(g);
      hrweno::check(hrweno_mgpu_set_flux_time_fn(h_, tfn_ ? &mgpu::tramp_time : nullptr, this));
   }
   void rktvd(int order) { hrweno::check(hrweno_mgpu_rktvd(h_, order)); } // ode = rktvd(rhs, neq, order), tvdode.f90:69-95
   void mstvd() { hrweno::check(hrweno_mgpu_mstvd(h_)); }                 // ode = mstvd(rhs, neq), tvdode.f90:180-201
   // call ode%integrate(u, t, tout, dt [, itask]) with the global host vector
   void integrate(double *u, double &t, double tout, double dt, int itask = 1) {
      hrweno::check(hrweno_mgpu_integrate(h_, u, &t, tout, dt, itask));
   }
   // state resident on the GPUs between output times
   void upload(const double *u) { hrweno::check(hrweno_mgpu_upload(h_, u)); }
   void integrate_resident(double &t, double tout, double dt, int itask = 1) {
      hrweno::check(hrweno_mgpu_integrate_resident(h_, &t, tout, dt, itask));
   }
   void download(double *u) { hrweno::check(hrweno_mgpu_download(h_, u)); }
   // alpha = max |f'(u)| over all slabs of the resident state; install: also the Lax-Friedrichs alpha from the next stage on
   double max_wavespeed(bool install = false) {
      double a = 0.0;
      hrweno::check(hrweno_mgpu_max_wavespeed(h_, &a, install ? 1 : 0));
      return a;
   }
   void set_alpha(double alpha) { hrweno::check(hrweno_mgpu_set_alpha(h_, alpha)); }
   int64_t fevals() const { return hrweno_mgpu_fevals(h_); }
   int64_t launches() const { return hrweno_mgpu_launches(h_); }
 private:
   ::hrweno_mgpu *h_ = nullptr;
   std::function<double(double)> tfn_;
   static double tramp_time(void *ctx, double t) { return static_cast<mgpu *>(ctx)->tfn_(t); }
};
} // namespace hrweno_multi

// The REAL32 build of the reference (src/hrweno_kinds.F90:9-10: -DREAL32 makes rk = real32 for the whole library): the
// same types with float in every position, over the hrweno_*_f32 entry points.  One GPU, the reference's operation order.
namespace real32 {
class weno { // weno.f90:23-50 with rk = real32
 public:
   int64_t ncells = 0;
   int k = 3;
   float eps = 1e-6f;
   weno(int64_t ncells_, int k_ = 3, float eps_ = 1e-6f, const std::vector<float> *xedges = nullptr) : ncells(ncells_), k(k_), eps(eps_) {
      if (xedges && (int64_t)xedges->size() != ncells_ + 1) throw hrweno::error(HRWENO_EINVAL, "Invalid input 'xedges': size(xedges) /= ncells + 1.");
      hrweno::check(hrweno_weno_f32_create(&h_, ncells_, k_, eps_, xedges ? xedges->data() : nullptr));
   }
   weno(const weno &) = delete;
   weno &operator=(const weno &) = delete;
   ~weno() { hrweno_weno_f32_destroy(h_); }
   void reconstruct(const float *v, float *vl, float *vr) const { hrweno::check(hrweno_weno_f32_reconstruct(h_, v, vl, vr)); } // weno.f90:129-219
   std::vector<float> cnu() const { // cnu(0:k-1,-1:k-1,1:ncells), weno.f90:41 (objects built with xedges)
      std::vector<float> c((size_t)(k * (k + 1) * ncells));
      hrweno::check(hrweno_weno_f32_get_cnu(h_, c.data()));
      return c;
   }
 private:
   ::hrweno_weno_f32 *h_ = nullptr;
};

class fv { // the example `rhs` (example1:72-109, example2:73-129) with rk = real32
 public:
   explicit fv(const hrweno_fv_desc_f32 &d) { hrweno::check(hrweno_fv_f32_create(&h_, &d)); }
   fv(const fv &) = delete;
   fv &operator=(const fv &) = delete;
   ~fv() { hrweno_fv_f32_destroy(h_); }
   int64_t neq() const { return hrweno_fv_f32_neq(h_); }
   void rhs(float t, const float *v, float *vdot) { hrweno::check(hrweno_fv_f32_rhs(h_, t, v, vdot)); }
   void rhs_dev(float t, const float *v_dev, float *vdot_dev, void *stream = nullptr) {
      hrweno::check(hrweno_fv_f32_rhs_dev(h_, t, v_dev, vdot_dev, stream));
   }
   void set_xedges(int axis, const float *xedges) { hrweno::check(hrweno_fv_f32_set_xedges(h_, axis, xedges)); }
   void set_flux_coef(int axis, const float *face, const float *cross = nullptr) {
      hrweno::check(hrweno_fv_f32_set_flux_coef(h_, axis, face, cross));
   }
   void set_flux_time_fn(std::function<float(float)> g) {
      tfn_ = std::move(g);
      hrweno::check(hrweno_fv_f32_set_flux_time_fn(h_, tfn_ ? &fv::tramp_time : nullptr, this));
   }
   ::hrweno_fv_f32 *handle() const { return h_; }
   static hrweno_fv_desc_f32 desc1d(int64_t nc, int k, float eps, const float *width) {
      hrweno_fv_desc_f32 d{};
      d.abi_version = HRWENO_ABI_VERSION;
      d.ndim = 1; d.n[0] = nc; d.n[1] = 1; d.rows = 1; d.k = k; d.eps = eps;
      d.flux_model = HRWENO_FLUX_BURGERS; d.flux_scheme = HRWENO_SCHEME_GODUNOV; d.bc = HRWENO_BC_COPY_NEIGHBOUR;
      d.grid_kind = HRWENO_GRID_WIDTH_ARRAY; d.mode = HRWENO_MODE_STRICT; d.flux_coef[0] = d.flux_coef[1] = 1.0f; d.alpha = 1.0f;
      d.width[0] = width; d.nranks = 1;
      return d;
   }
   static hrweno_fv_desc_f32 desc2d(int64_t nc1, int64_t nc2, int k, float eps, const float *w1, const float *w2) {
      hrweno_fv_desc_f32 d = desc1d(nc1, k, eps, w1);
      d.ndim = 2; d.n[1] = nc2; d.width[1] = w2;
      d.flux_model = HRWENO_FLUX_LINEAR; d.bc = HRWENO_BC_ZERO_FLUX;
      return d;
   }
 private:
   ::hrweno_fv_f32 *h_ = nullptr;
   std::function<float(float)> tfn_;
   static float tramp_time(void *ctx, float t) { return static_cast<fv *>(ctx)->tfn_(t); }
};

class tvdode { // tvdode.f90:14-34 with rk = real32
 public:
   int64_t fevals() const { return hrweno_ode_f32_fevals(h_); }
   int istate() const { return h_ ? hrweno_ode_f32_istate(h_) : -1; }
   int64_t launches() const { return hrweno_ode_f32_launches(h_); }
   // call ode%integrate(u, t, tout, dt [, itask])  (tvdode.f90:97-178, 203-271): t = t + dt accumulates in real32
   void integrate(float *u, float &t, float tout, float dt, int itask = 1) { hrweno::check(hrweno_ode_f32_integrate(h_, u, &t, tout, dt, itask)); }
   void integrate_dev(float *u_dev, float &t, float tout, float dt, int itask = 1, void *stream = nullptr) {
      hrweno::check(hrweno_ode_f32_integrate_dev(h_, u_dev, &t, tout, dt, itask, stream));
   }
   tvdode(const tvdode &) = delete;
   tvdode &operator=(const tvdode &) = delete;
   virtual ~tvdode() { hrweno_ode_f32_destroy(h_); }
 protected:
   tvdode() = default;
   ::hrweno_ode_f32 *h_ = nullptr;
};
class rktvd : public tvdode { // rktvd(rhs, neq, order), tvdode.f90:69-95
 public:
   rktvd(fv &rhs, int64_t neq, int order) {
      if (neq != rhs.neq()) throw hrweno::error(HRWENO_EINVAL, "neq does not match the finite-volume operator");
      hrweno::check(hrweno_rktvd_f32_create_fused(&h_, rhs.handle(), order));
   }
};
class mstvd : public tvdode { // mstvd(rhs, neq), tvdode.f90:180-201
 public:
   mstvd(fv &rhs, int64_t neq) {
      if (neq != rhs.neq()) throw hrweno::error(HRWENO_EINVAL, "neq does not match the finite-volume operator");
      hrweno::check(hrweno_mstvd_f32_create_fused(&h_, rhs.handle()));
   }
};
} // namespace real32
} // namespace hrweno
