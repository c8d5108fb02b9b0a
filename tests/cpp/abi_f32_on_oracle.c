/* abi_f32_on_oracle.c -- TEST INFRASTRUCTURE.  The REAL32 entry points (hrweno_*_f32 of include/hrweno_b200.h) that the
 * Fortran shim binds when it is compiled with -DREAL32, implemented on the REAL32 build of the CPU oracle
 * (oracle/libhrweno_oracle_f32.so = hrweno_oracle.c with every `double` read as `float`).  Same purpose and rules as
 * abi_on_oracle.c: built by the tests into a temporary directory, never shipped, never loaded by hr-weno_b200/. */
#include <stdlib.h>
#include <string.h>

#include "hrweno_b200.h" /* the product prototypes, float signatures: a drift is a compile error here */

/* the oracle's header as its REAL32 build reads it (hrweno_oracle.c:23-31) */
#define double float
#define hrweno_fv_desc hrweno_fv_desc_f32
#define hrweno_time_fn hrweno_time_fn_f32
#include "hrweno_oracle.h"
#undef double
#undef hrweno_fv_desc
#undef hrweno_time_fn

static _Thread_local char g_err[256] = "";
static int fail(int st, const char *msg) {
   strncpy(g_err, msg, sizeof g_err - 1);
   return st;
}
const char *hrweno_last_error(void) { return g_err; }

struct hrweno_weno_f32 {
   int64_t nc;
   int k;
   float eps;
   float *cnu;
};

int hrweno_weno_f32_create(hrweno_weno_f32 **out, int64_t ncells, int k, float eps, const float *xedges) {
   if (!out) return fail(HRWENO_EINVAL, "null output pointer");
   *out = NULL;
   if (!(ncells > 0)) return fail(HRWENO_EINVAL, "Invalid input 'ncells'. Valid range: ncells > 0.");
   if (!(k >= 1 && k <= 3)) return fail(HRWENO_EINVAL, "Invalid input 'k'. Valid range: 1 <= k <= 3.");
   if (hrweno_ref_weno_check(ncells, k, eps)) return fail(HRWENO_EINVAL, "Invalid input 'eps'. Valid range: eps > epsilon.");
   hrweno_weno_f32 *w = (hrweno_weno_f32 *)calloc(1, sizeof *w);
   if (!w) return fail(HRWENO_ENOMEM, "out of memory");
   w->nc = ncells, w->k = k, w->eps = eps;
   if (xedges) {
      w->cnu = (float *)malloc(sizeof(float) * (size_t)(k * (k + 1)) * (size_t)ncells);
      if (!w->cnu || hrweno_ref_weno_calc_cnu(ncells, k, xedges, w->cnu)) {
         free(w->cnu), free(w);
         return fail(HRWENO_EINVAL, "weno_calc_cnu failed");
      }
   }
   *out = w;
   return HRWENO_OK;
}
void hrweno_weno_f32_destroy(hrweno_weno_f32 *w) {
   if (w) free(w->cnu), free(w);
}
int hrweno_weno_f32_get_cnu(const hrweno_weno_f32 *w, float *cnu) {
   if (!w || !w->cnu || !cnu) return fail(HRWENO_EINVAL, "no cnu");
   memcpy(cnu, w->cnu, sizeof(float) * (size_t)(w->k * (w->k + 1)) * (size_t)w->nc);
   return HRWENO_OK;
}
static int reconstruct(const hrweno_weno_f32 *w, const float *v, float *vl, float *vr) {
   if (!w) return fail(HRWENO_EINVAL, "null weno handle");
   return hrweno_ref_weno_reconstruct(w->nc, w->k, w->eps, w->cnu, v, 1, vl, vr);
}
int hrweno_weno_f32_reconstruct(const hrweno_weno_f32 *w, const float *v, float *vl, float *vr) { return reconstruct(w, v, vl, vr); }
void hrweno_weno_f32_reconstruct_s(const hrweno_weno_f32 *w, const float *v, float *vl, float *vr, int *status) {
   const int st = reconstruct(w, v, vl, vr);
   if (status) *status = st;
}

int hrweno_fv_f32_create(hrweno_fv_f32 **out, const hrweno_fv_desc_f32 *desc) {
   if (!out || !desc) return fail(HRWENO_EINVAL, "null argument");
   const int st = hrweno_ref_fv_create((hrweno_ref_fv **)out, desc);
   return st ? fail(st, "hrweno_fv_f32_create: invalid descriptor") : st;
}
void hrweno_fv_f32_destroy(hrweno_fv_f32 *fv) { hrweno_ref_fv_destroy((hrweno_ref_fv *)fv); }
int64_t hrweno_fv_f32_neq(const hrweno_fv_f32 *fv) { return hrweno_ref_fv_neq((const hrweno_ref_fv *)fv); }
int hrweno_fv_f32_rhs(hrweno_fv_f32 *fv, float t, const float *v, float *vdot) { return hrweno_ref_fv_rhs((hrweno_ref_fv *)fv, t, v, vdot); }
int hrweno_fv_f32_rhs_dev(hrweno_fv_f32 *fv, float t, const float *v_dev, float *vdot_dev, void *stream) {
   (void)stream;
   return hrweno_ref_fv_rhs((hrweno_ref_fv *)fv, t, v_dev, vdot_dev);
}
int hrweno_fv_f32_set_xedges(hrweno_fv_f32 *fv, int axis, const float *xedges) {
   return hrweno_ref_fv_set_xedges((hrweno_ref_fv *)fv, axis, xedges);
}
int hrweno_fv_f32_set_flux_coef(hrweno_fv_f32 *fv, int axis, const float *face, const float *cross) {
   return hrweno_ref_fv_set_flux_coef((hrweno_ref_fv *)fv, axis, face, cross);
}
int hrweno_fv_f32_set_flux_time_fn(hrweno_fv_f32 *fv, hrweno_time_fn_f32 g, void *ctx) {
   return hrweno_ref_fv_set_flux_time_fn((hrweno_ref_fv *)fv, g, ctx);
}

int hrweno_rktvd_f32_create_fused(hrweno_ode_f32 **out, hrweno_fv_f32 *fv, int order) {
   if (!(order >= 1 && order <= 3)) return fail(HRWENO_EINVAL, "Invalid input 'order' in 'rktvd'. Valid range: 1 <= k <= 3.");
   return hrweno_ref_rktvd_create_fv((hrweno_ref_ode **)out, (hrweno_ref_fv *)fv, order);
}
int hrweno_mstvd_f32_create_fused(hrweno_ode_f32 **out, hrweno_fv_f32 *fv) {
   return hrweno_ref_mstvd_create_fv((hrweno_ref_ode **)out, (hrweno_ref_fv *)fv);
}
void hrweno_ode_f32_destroy(hrweno_ode_f32 *ode) { hrweno_ref_ode_destroy((hrweno_ref_ode *)ode); }
int hrweno_ode_f32_integrate(hrweno_ode_f32 *ode, float *u, float *t, float tout, float dt, int itask) {
   return hrweno_ref_ode_integrate((hrweno_ref_ode *)ode, u, t, tout, dt, itask);
}
int64_t hrweno_ode_f32_fevals(const hrweno_ode_f32 *ode) { return hrweno_ref_ode_fevals((const hrweno_ref_ode *)ode); }
int hrweno_ode_f32_istate(const hrweno_ode_f32 *ode) { return hrweno_ref_ode_istate((const hrweno_ref_ode *)ode); }
