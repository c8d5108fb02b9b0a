/* abi_on_oracle.c -- TEST INFRASTRUCTURE.  The entry points of include/hrweno_b200.h that the Fortran shim
 * (fortran/hrweno_b200_shim.f90) binds, implemented on the CPU oracle (oracle/hrweno_oracle.c).
 *
 * Why: the container that runs `pytest -m "not gpu"` has no GPU, so the product library cannot execute there.  To run the
 * shim's SOURCE (through tools/f90exec) under the reference's own programs on that machine, its bind(c) calls need a
 * library with the product's C ABI.  This file is that library and nothing else: same symbol names, same prototypes (it
 * includes the product header, so a signature that drifts is a compile error here), every call forwarded to the oracle
 * function that restates the same reference lines.  It is built by tests/test_fortran_shim_exec.py into a temporary
 * directory, never shipped, never loaded by hr-weno_b200/.  On the GPU box the same tests bind libhrweno_b200.so instead. */
#include <stdlib.h>
#include <string.h>

#include "hrweno_b200.h"
#include "hrweno_oracle.h"

static _Thread_local char g_err[256] = "";

static int fail(int st, const char *msg) {
   strncpy(g_err, msg, sizeof g_err - 1);
   return st;
}

const char *hrweno_last_error(void) { return g_err; }

/* ---- weno (weno.f90:54-127, 129-219) ---- */
struct hrweno_weno {
   int64_t nc;
   int k;
   double eps;
   double *cnu; /* NULL: uniform tables */
};

int hrweno_weno_create(hrweno_weno **out, int64_t ncells, int k, double eps, const double *xedges) {
   if (!out) return fail(HRWENO_EINVAL, "null output pointer");
   *out = NULL;
   if (!(ncells > 0)) return fail(HRWENO_EINVAL, "Invalid input 'ncells'. Valid range: ncells > 0.");
   if (!(k >= 1 && k <= 3)) return fail(HRWENO_EINVAL, "Invalid input 'k'. Valid range: 1 <= k <= 3.");
   if (hrweno_ref_weno_check(ncells, k, eps)) return fail(HRWENO_EINVAL, "Invalid input 'eps'. Valid range: eps > epsilon.");
   hrweno_weno *w = (hrweno_weno *)calloc(1, sizeof *w);
   if (!w) return fail(HRWENO_ENOMEM, "out of memory");
   w->nc = ncells, w->k = k, w->eps = eps;
   if (xedges) {
      w->cnu = (double *)malloc(sizeof(double) * (size_t)(k * (k + 1)) * (size_t)ncells);
      if (!w->cnu || hrweno_ref_weno_calc_cnu(ncells, k, xedges, w->cnu)) {
         free(w->cnu), free(w);
         return fail(HRWENO_EINVAL, "weno_calc_cnu failed");
      }
   }
   *out = w;
   return HRWENO_OK;
}

void hrweno_weno_destroy(hrweno_weno *w) {
   if (w) free(w->cnu), free(w);
}

int hrweno_weno_get_cnu(const hrweno_weno *w, double *cnu) {
   if (!w || !w->cnu || !cnu) return fail(HRWENO_EINVAL, "no cnu");
   memcpy(cnu, w->cnu, sizeof(double) * (size_t)(w->k * (w->k + 1)) * (size_t)w->nc);
   return HRWENO_OK;
}

static int reconstruct(const hrweno_weno *w, const double *v, double *vl, double *vr) {
   if (!w) return fail(HRWENO_EINVAL, "null weno handle");
   return hrweno_ref_weno_reconstruct(w->nc, w->k, w->eps, w->cnu, v, 1, vl, vr);
}

int hrweno_weno_reconstruct(const hrweno_weno *w, const double *v, double *vl, double *vr) { return reconstruct(w, v, vl, vr); }

void hrweno_weno_reconstruct_s(const hrweno_weno *w, const double *v, double *vl, double *vr, int *status) {
   const int st = reconstruct(w, v, vl, vr);
   if (status) *status = st;
}

/* ---- the example rhs as one operator (example1:72-109, example2:73-129) ---- */
int hrweno_fv_create(hrweno_fv **out, const hrweno_fv_desc *desc) {
   if (!out || !desc) return fail(HRWENO_EINVAL, "null argument");
   const int st = hrweno_ref_fv_create((hrweno_ref_fv **)out, desc);
   return st ? fail(st, "hrweno_fv_create: invalid descriptor") : st;
}
void hrweno_fv_destroy(hrweno_fv *fv) { hrweno_ref_fv_destroy((hrweno_ref_fv *)fv); }
int64_t hrweno_fv_neq(const hrweno_fv *fv) { return hrweno_ref_fv_neq((const hrweno_ref_fv *)fv); }
int hrweno_fv_rhs(hrweno_fv *fv, double t, const double *v, double *vdot) {
   return hrweno_ref_fv_rhs((hrweno_ref_fv *)fv, t, v, vdot);
}
/* "device" pointers are host pointers here, the stream is ignored */
int hrweno_fv_rhs_dev(hrweno_fv *fv, double t, const double *v_dev, double *vdot_dev, void *stream) {
   (void)stream;
   return hrweno_ref_fv_rhs((hrweno_ref_fv *)fv, t, v_dev, vdot_dev);
}
int hrweno_fv_set_xedges(hrweno_fv *fv, int axis, const double *xedges) {
   return hrweno_ref_fv_set_xedges((hrweno_ref_fv *)fv, axis, xedges);
}
int hrweno_fv_set_flux_coef(hrweno_fv *fv, int axis, const double *face, const double *cross) {
   return hrweno_ref_fv_set_flux_coef((hrweno_ref_fv *)fv, axis, face, cross);
}
int hrweno_fv_set_flux_time_fn(hrweno_fv *fv, hrweno_time_fn g, void *ctx) {
   return hrweno_ref_fv_set_flux_time_fn((hrweno_ref_fv *)fv, g, ctx);
}

/* ---- integrators (tvdode.f90:69-271) ---- */
int hrweno_rktvd_create_host(hrweno_ode **out, hrweno_rhs_host_fn fu, void *ctx, int64_t neq, int order) {
   if (!(neq > 0)) return fail(HRWENO_EINVAL, "Invalid input 'neq'. Valid range: neq >= 1.");
   if (!(order >= 1 && order <= 3)) return fail(HRWENO_EINVAL, "Invalid input 'order' in 'rktvd'. Valid range: 1 <= k <= 3.");
   return hrweno_ref_rktvd_create((hrweno_ref_ode **)out, (hrweno_ref_rhs_fn)fu, ctx, neq, order);
}
int hrweno_mstvd_create_host(hrweno_ode **out, hrweno_rhs_host_fn fu, void *ctx, int64_t neq) {
   if (!(neq > 0)) return fail(HRWENO_EINVAL, "Invalid input 'neq'. Valid range: neq >= 1.");
   return hrweno_ref_mstvd_create((hrweno_ref_ode **)out, (hrweno_ref_rhs_fn)fu, ctx, neq);
}
/* device-integrand constructors: hrweno_rhs_fn has a stream argument the oracle's callback type lacks */
struct dev_adapter {
   hrweno_rhs_fn fu;
   void *ctx;
};
static void dev_trampoline(void *p, double t, int64_t neq, const double *u, double *udot) {
   struct dev_adapter *a = (struct dev_adapter *)p;
   a->fu(a->ctx, t, neq, u, udot, NULL);
}
int hrweno_rktvd_create(hrweno_ode **out, hrweno_rhs_fn fu, void *ctx, int64_t neq, int order) {
   if (!(order >= 1 && order <= 3)) return fail(HRWENO_EINVAL, "Invalid input 'order' in 'rktvd'. Valid range: 1 <= k <= 3.");
   struct dev_adapter *a = (struct dev_adapter *)malloc(sizeof *a); /* lives as long as the process: test infrastructure */
   a->fu = fu, a->ctx = ctx;
   return hrweno_ref_rktvd_create((hrweno_ref_ode **)out, dev_trampoline, a, neq, order);
}
int hrweno_mstvd_create(hrweno_ode **out, hrweno_rhs_fn fu, void *ctx, int64_t neq) {
   struct dev_adapter *a = (struct dev_adapter *)malloc(sizeof *a);
   a->fu = fu, a->ctx = ctx;
   return hrweno_ref_mstvd_create((hrweno_ref_ode **)out, dev_trampoline, a, neq);
}
int hrweno_rktvd_create_fused(hrweno_ode **out, hrweno_fv *fv, int order) {
   if (!(order >= 1 && order <= 3)) return fail(HRWENO_EINVAL, "Invalid input 'order' in 'rktvd'. Valid range: 1 <= k <= 3.");
   return hrweno_ref_rktvd_create_fv((hrweno_ref_ode **)out, (hrweno_ref_fv *)fv, order);
}
int hrweno_mstvd_create_fused(hrweno_ode **out, hrweno_fv *fv) {
   return hrweno_ref_mstvd_create_fv((hrweno_ref_ode **)out, (hrweno_ref_fv *)fv);
}
void hrweno_ode_destroy(hrweno_ode *ode) { hrweno_ref_ode_destroy((hrweno_ref_ode *)ode); }
int hrweno_ode_integrate(hrweno_ode *ode, double *u, double *t, double tout, double dt, int itask) {
   return hrweno_ref_ode_integrate((hrweno_ref_ode *)ode, u, t, tout, dt, itask);
}
int64_t hrweno_ode_fevals(const hrweno_ode *ode) { return hrweno_ref_ode_fevals((const hrweno_ref_ode *)ode); }
int hrweno_ode_istate(const hrweno_ode *ode) { return hrweno_ref_ode_istate((const hrweno_ref_ode *)ode); }

/* ---- one process, N GPUs (hrweno_mgpu_*): on the oracle there is one domain and no slabs ---- */
struct hrweno_mgpu {
   hrweno_ref_fv *fv;
   hrweno_ref_ode *ode;
   int64_t neq;
   double *state; /* resident state between upload and download */
};

int hrweno_mgpu_create(hrweno_mgpu **out, const hrweno_fv_desc *desc, int ngpus, const int *devices) {
   (void)ngpus, (void)devices;
   if (!out || !desc) return fail(HRWENO_EINVAL, "null argument");
   hrweno_mgpu *m = (hrweno_mgpu *)calloc(1, sizeof *m);
   if (!m) return fail(HRWENO_ENOMEM, "out of memory");
   const int st = hrweno_ref_fv_create(&m->fv, desc);
   if (st) {
      free(m);
      return fail(st, "hrweno_mgpu_create: invalid descriptor");
   }
   m->neq = hrweno_ref_fv_neq(m->fv);
   *out = m;
   return HRWENO_OK;
}
void hrweno_mgpu_destroy(hrweno_mgpu *m) {
   if (!m) return;
   hrweno_ref_ode_destroy(m->ode), hrweno_ref_fv_destroy(m->fv), free(m->state), free(m);
}
int hrweno_mgpu_ngpus(const hrweno_mgpu *m) { return m ? 1 : 0; }
int hrweno_mgpu_rktvd(hrweno_mgpu *m, int order) {
   hrweno_ref_ode_destroy(m->ode), m->ode = NULL;
   return hrweno_ref_rktvd_create_fv(&m->ode, m->fv, order);
}
int hrweno_mgpu_mstvd(hrweno_mgpu *m) {
   hrweno_ref_ode_destroy(m->ode), m->ode = NULL;
   return hrweno_ref_mstvd_create_fv(&m->ode, m->fv);
}
int hrweno_mgpu_integrate(hrweno_mgpu *m, double *u, double *t, double tout, double dt, int itask) {
   if (!m->ode) return fail(HRWENO_EINVAL, "hrweno_mgpu_integrate: no integrator (call hrweno_mgpu_rktvd / _mstvd first)");
   return hrweno_ref_ode_integrate(m->ode, u, t, tout, dt, itask);
}
int hrweno_mgpu_upload(hrweno_mgpu *m, const double *u) {
   if (!m->state) m->state = (double *)malloc(sizeof(double) * (size_t)m->neq);
   if (!m->state) return fail(HRWENO_ENOMEM, "out of memory");
   memcpy(m->state, u, sizeof(double) * (size_t)m->neq);
   return HRWENO_OK;
}
int hrweno_mgpu_integrate_resident(hrweno_mgpu *m, double *t, double tout, double dt, int itask) {
   if (!m->ode || !m->state) return fail(HRWENO_EINVAL, "hrweno_mgpu_integrate_resident: no integrator or no resident state");
   return hrweno_ref_ode_integrate(m->ode, m->state, t, tout, dt, itask);
}
int hrweno_mgpu_download(hrweno_mgpu *m, double *u) {
   if (!m->state) return fail(HRWENO_EINVAL, "hrweno_mgpu_download: no resident state");
   memcpy(u, m->state, sizeof(double) * (size_t)m->neq);
   return HRWENO_OK;
}
int64_t hrweno_mgpu_fevals(const hrweno_mgpu *m) { return m->ode ? hrweno_ref_ode_fevals(m->ode) : 0; }
