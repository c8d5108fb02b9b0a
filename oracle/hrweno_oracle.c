/*
 * hrweno_oracle.c -- CPU restatement of the reference's finite-volume update path.
 * TEST INFRASTRUCTURE (see hrweno_oracle.h for the parity status: pinned to the
 * reference's own test tolerances and to an independent NumPy restatement;
 * unpinned beyond that because no Fortran compiler exists in this image).
 *
 * Build: gcc -O3 -funroll-loops -ffp-contract=off -fno-fast-math -fopenmp  (oracle/Makefile)
 *   -O3 -funroll-loops mirror fpm's gfortran release profile; -ffp-contract=off because the
 *   reference's x86-64 build has no FMA contraction; OpenMP is only used by the baseline legs
 *   (the shipped reference is serial: its !$omp lines are inert, example2:22,34,95-114).
 *
 * Every routine cites the reference lines it follows (paths relative to /root/reference).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* REAL32 build (src/hrweno_kinds.F90:9-17: `rk` is a compile-time switch of the WHOLE reference).  The same source with
 * every `double` of this file AND of the two headers below read as `float`, built with -fsingle-precision-constant so that
 * a literal like 11.0/6 is the float quotient `11.0_rk/6` is for rk = real32 (x86-64 SSE evaluates float expressions in
 * float: FLT_EVAL_METHOD == 0).  The system headers above keep their prototypes; copysign's detour through double only
 * carries a sign.  -> oracle/libhrweno_oracle_f32.so, same symbol names, float in every signature and in hrweno_fv_desc. */
#ifdef HRW_REAL32
#define double float
#undef DBL_EPSILON
#define DBL_EPSILON FLT_EPSILON
#endif
#include "hrweno_oracle.h"

static int g_threads = 1;

void hrweno_ref_set_threads(int nthreads) {
   g_threads = nthreads < 1 ? 1 : nthreads;
#ifdef _OPENMP
   omp_set_num_threads(g_threads);
#endif
}

int hrweno_ref_max_threads(void) {
#ifdef _OPENMP
   return omp_get_num_procs();
#else
   return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Constant tables -- src/hrweno_weno.f90:12-21.  The literals are written exactly like the
 * source (11.0_rk/6 etc.) so each entry is the correctly rounded fp64 quotient.  reshape(...,
 * order=[1,2]) is the plain column-major fill, so the source list is c(0,-1),c(1,-1),...
 * ---------------------------------------------------------------------------------------- */
static const double D1[1] = {1.0};
static const double D2[2] = {2.0 / 3, 1.0 / 3};
static const double D3[3] = {0.3, 0.6, 0.1};
static const double C1[2] = {1.0, 1.0};
static const double C2[6] = {3.0 / 2, -1.0 / 2, 1.0 / 2, 1.0 / 2, -1.0 / 2, 3.0 / 2};
static const double C3[12] = {11.0 / 6, -7.0 / 6, 1.0 / 3, 1.0 / 3, 5.0 / 6, -1.0 / 6,
                              -1.0 / 6, 5.0 / 6, 1.0 / 3, 1.0 / 3, -7.0 / 6, 11.0 / 6};

int hrweno_ref_tables(int k, double *d, double *c) {
   const double *ds, *cs;
   switch (k) {
   case 1: ds = D1; cs = C1; break;
   case 2: ds = D2; cs = C2; break;
   case 3: ds = D3; cs = C3; break;
   default: return HRWENO_EINVAL;
   }
   if (d) memcpy(d, ds, sizeof(double) * (size_t)k);
   if (c) memcpy(c, cs, sizeof(double) * (size_t)(k * (k + 1)));
   return HRWENO_OK;
}

/* weno_init validation -- src/hrweno_weno.f90:71-98 */
int hrweno_ref_weno_check(int64_t nc, int k, double eps) {
   if (!(nc > 0)) return HRWENO_EINVAL;        /* :72-78 */
   if (!(k >= 1 && k <= 3)) return HRWENO_EINVAL; /* :80-88 */
   if (!(eps > DBL_EPSILON)) return HRWENO_EINVAL; /* :90-98, epsilon(1.0_rk) */
   return HRWENO_OK;
}

/* ------------------------------------------------------------------------------------------
 * weno_calc_cnu -- src/hrweno_weno.f90:221-297 (Shu eq. 2.20).
 * xext(-(k+1) : nc+(k+1)); xl(m) = xext(m-1), xr(m) = xext(m)  (:258-259).
 * ---------------------------------------------------------------------------------------- */
int hrweno_ref_weno_calc_cnu(int64_t nc, int k, const double *xedges, double *cnu) {
   if (hrweno_ref_weno_check(nc, k, 1.0) || !xedges || !cnu) return HRWENO_EINVAL;
   const int ng = k + 1;
   double *buf = (double *)malloc(sizeof(double) * (size_t)(nc + 2 * ng + 1));
   if (!buf) return HRWENO_ENOMEM;
   double *xext = buf + ng; /* xext[-ng .. nc+ng] */
   for (int64_t i = 0; i <= nc; ++i) xext[i] = xedges[i]; /* :243 */
   double dx = xext[1] - xext[0];                           /* :246 */
   for (int64_t i = -1; i >= -ng; --i) xext[i] = xext[i + 1] - dx; /* :247-249 */
   dx = xext[nc] - xext[nc - 1];                            /* :252 */
   for (int64_t i = nc + 1; i <= nc + ng; ++i) xext[i] = xext[i - 1] + dx; /* :253-255 */
#define XL(m) xext[(m)-1]
#define XR(m) xext[(m)]
   for (int64_t i = 1; i <= nc; ++i) {
      for (int r = -1; r <= k - 1; ++r) {
         for (int j = 0; j <= k - 1; ++j) {
            double sum2 = 0.0; /* :265 */
            for (int m = j + 1; m <= k; ++m) {
               double prod2 = 1.0; /* :268-272 */
               for (int l = 0; l <= k; ++l) {
                  if (l == m) continue;
                  prod2 = prod2 * (XL(i - r + m) - XL(i - r + l));
               }
               double sum1 = 0.0; /* :274-287 */
               for (int l = 0; l <= k; ++l) {
                  if (l == m) continue;
                  double prod1 = 1.0;
                  for (int q = 0; q <= k; ++q) {
                     if (q == m || q == l) continue;
                     prod1 = prod1 * (XR(i) - XL(i - r + q));
                  }
                  sum1 = sum1 + prod1;
               }
               sum2 = sum2 + sum1 / prod2; /* :289 */
            }
            cnu[j + k * ((r + 1) + (k + 1) * (i - 1))] = sum2 * (XR(i - r + j) - XL(i - r + j)); /* :293 */
         }
      }
   }
#undef XL
#undef XR
   free(buf);
   return HRWENO_OK;
}

/* ------------------------------------------------------------------------------------------
 * weno_reconstruct -- src/hrweno_weno.f90:129-219.  One cell, literal order:
 *   vrr(r) = sum(ci(:,r)  *vext(i-r:i-r+k-1))   (:179)   sum() accumulates from 0 left to right
 *   vlr(r) = sum(ci(:,r-1)*vext(i-r:i-r+k-1))   (:180)
 *   beta                                          (:183-204)
 *   alfa = d/(eps+beta)**2 ; alfatilde = d(k-1:0:-1)/(eps+beta)**2   (:207-208)
 *   w = alfa/sum(alfa) ; wtilde = alfatilde/sum(alfatilde)           (:209-210)
 *   vr(i) = sum(w*vrr) ; vl(i) = sum(wtilde*vlr)                     (:213-214)
 * `ve` points at vext(i); ci is c(j,r) at ci[j + k*(r+1)].
 * ---------------------------------------------------------------------------------------- */
static inline void recon_cell(const int k, const double eps, const double *d, const double *ci,
                              const double *ve, double *vl_i, double *vr_i) {
   double vlr[3], vrr[3], beta[3], alfa[3], alfatilde[3];
   for (int r = 0; r < k; ++r) {
      double sr = 0.0, sl = 0.0;
      for (int j = 0; j < k; ++j) {
         sr = sr + ci[j + k * (r + 1)] * ve[-r + j];
         sl = sl + ci[j + k * r] * ve[-r + j];
      }
      vrr[r] = sr;
      vlr[r] = sl;
   }
   switch (k) {
   case 1:
      beta[0] = 0.0; /* :186 */
      break;
   case 2: /* :190-191 */
      beta[0] = (ve[1] - ve[0]) * (ve[1] - ve[0]);
      beta[1] = (ve[0] - ve[-1]) * (ve[0] - ve[-1]);
      break;
   default: { /* :195-202 */
      double a, b;
      a = (ve[0] - 2 * ve[1]) + ve[2];
      b = (3 * ve[0] - 4 * ve[1]) + ve[2];
      beta[0] = 13.0 / 12 * (a * a) + 1.0 / 4 * (b * b);
      a = (ve[-1] - 2 * ve[0]) + ve[1];
      b = ve[-1] - ve[1];
      beta[1] = 13.0 / 12 * (a * a) + 1.0 / 4 * (b * b);
      a = (ve[-2] - 2 * ve[-1]) + ve[0];
      b = (ve[-2] - 4 * ve[-1]) + 3 * ve[0];
      beta[2] = 13.0 / 12 * (a * a) + 1.0 / 4 * (b * b);
   }
   }
   double sa = 0.0, sat = 0.0;
   for (int r = 0; r < k; ++r) {
      const double den = (eps + beta[r]) * (eps + beta[r]);
      alfa[r] = d[r] / den;
      alfatilde[r] = d[k - 1 - r] / den;
   }
   for (int r = 0; r < k; ++r) {
      sa = sa + alfa[r];
      sat = sat + alfatilde[r];
   }
   double svr = 0.0, svl = 0.0;
   for (int r = 0; r < k; ++r) {
      svr = svr + (alfa[r] / sa) * vrr[r];
      svl = svl + (alfatilde[r] / sat) * vlr[r];
   }
   *vr_i = svr;
   *vl_i = svl;
}

static void recon_row(const int k, const int64_t nc, const double eps, const double *cnu, const double *v,
                      const int64_t incv, double *vl, double *vr, double *vext_buf, const int par) {
   double d[3], c[12];
   hrweno_ref_tables(k, d, c);
   double *vext = vext_buf + (k - 1); /* vext[-(k-1) .. nc-1+(k-1)], 0-based cell index */
   /* :171-173 */
   for (int64_t i = 0; i < nc; ++i) vext[i] = v[i * incv];
   for (int g = 1; g <= k - 1; ++g) {
      vext[-g] = v[0];
      vext[nc - 1 + g] = v[(nc - 1) * incv];
   }
   const int kk = k * (k + 1);
   if (par) {
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < nc; ++i) {
         const double *ci = cnu ? cnu + (size_t)kk * (size_t)i : c; /* :177 */
         recon_cell(k, eps, d, ci, vext + i, vl + i, vr + i);
      }
   } else {
      for (int64_t i = 0; i < nc; ++i) {
         const double *ci = cnu ? cnu + (size_t)kk * (size_t)i : c;
         recon_cell(k, eps, d, ci, vext + i, vl + i, vr + i);
      }
   }
}

int hrweno_ref_weno_reconstruct(int64_t nc, int k, double eps, const double *cnu, const double *v,
                                int64_t incv, double *vl, double *vr) {
   if (hrweno_ref_weno_check(nc, k, eps) || !v || !vl || !vr || incv < 1) return HRWENO_EINVAL;
   /* the automatic array vext of weno.f90:163 -- allocated per call like the reference */
   double *vext = (double *)malloc(sizeof(double) * (size_t)(nc + 2 * (k - 1)));
   if (!vext) return HRWENO_ENOMEM;
   recon_row(k, nc, eps, cnu, v, incv, vl, vr, vext, g_threads > 1 && nc >= 65536);
   free(vext);
   return HRWENO_OK;
}

/* ------------------------------------------------------------------------------------------
 * Fluxes -- src/hrweno_fluxes.f90
 * ---------------------------------------------------------------------------------------- */
double hrweno_ref_lax_friedrichs(hrweno_flux_fn f, void *ctx, double vm, double vp, const double *x, int nx,
                                 double t, double alpha) {
   return (f(ctx, vm, x, nx, t) + f(ctx, vp, x, nx, t) - alpha * (vp - vm)) / 2; /* :43 */
}

double hrweno_ref_godunov(hrweno_flux_fn f, void *ctx, double vm, double vp, const double *x, int nx, double t) {
   const double fm = f(ctx, vm, x, nx, t); /* :67 */
   const double fp = f(ctx, vp, x, nx, t); /* :68 */
   if (vm <= vp)                           /* :70-74 */
      return fm < fp ? fm : fp;
   else
      return fm > fp ? fm : fp;
}

double hrweno_ref_flux_model(int model, double coef, double v) {
   if (model == HRWENO_FLUX_BURGERS) return (v * v) / 2; /* example1:120  (v**2)/2 */
   return coef * v;                                       /* example2:140,153  v (coef = 1) */
}

double hrweno_ref_face_flux(int scheme, int model, double coef, double alpha, double vm, double vp) {
   const double fm = hrweno_ref_flux_model(model, coef, vm);
   const double fp = hrweno_ref_flux_model(model, coef, vp);
   if (scheme == HRWENO_SCHEME_LAX_FRIEDRICHS) return (fm + fp - alpha * (vp - vm)) / 2;
   if (vm <= vp) return fm < fp ? fm : fp;
   return fm > fp ? fm : fp;
}

/* x-dependent physical flux of the commented-out growth terms of example2:140,153 (`flux1 = v*x(1)**2`,
 * `flux2 = v*x(1)*x(2)`): f = (model(v)*cross)*face, left to right like the Fortran expression; a factor that is
 * absent is not multiplied in. */
static inline double flux_model_x(int model, double coef, double v, const double *cross, const double *face, const double *tfac) {
   double f = hrweno_ref_flux_model(model, coef, v);
   if (cross) f = f * *cross;
   if (face) f = f * *face;
   if (tfac) f = f * *tfac; /* separable time factor g(t): f(u, x, t), fluxes.f90:12-18 */
   return f;
}

static inline double face_flux_x(int scheme, int model, double coef, double alpha, double vm, double vp,
                                 const double *cross, const double *face, const double *tfac) {
   const double fm = flux_model_x(model, coef, vm, cross, face, tfac);
   const double fp = flux_model_x(model, coef, vp, cross, face, tfac);
   if (scheme == HRWENO_SCHEME_LAX_FRIEDRICHS) return (fm + fp - alpha * (vp - vm)) / 2; /* fluxes.f90:43 */
   if (vm <= vp) return fm < fp ? fm : fp;                                              /* fluxes.f90:70-74 */
   return fm > fp ? fm : fp;
}

/* ------------------------------------------------------------------------------------------
 * grid1%linear -- src/hrweno_grids.f90:76-79 (edges), :246-247 (center, width)
 * ---------------------------------------------------------------------------------------- */
void hrweno_ref_grid_linear(double xmin, double xmax, int64_t n, double *edges, double *center, double *width) {
   const double rx = (xmax - xmin) / (double)n; /* :76 */
   double el = xmin + rx * 0.0;
   if (edges) edges[0] = el;
   for (int64_t i = 1; i <= n; ++i) {
      const double er = xmin + rx * (double)i; /* :78 */
      if (edges) edges[i] = er;
      if (center) center[i - 1] = (el + er) / 2; /* :246 */
      if (width) width[i - 1] = er - el;         /* :247 */
      el = er;
   }
}

/* ------------------------------------------------------------------------------------------
 * The example right-hand sides
 * ---------------------------------------------------------------------------------------- */
struct hrweno_ref_fv {
   hrweno_fv_desc d;
   double *w[2]; /* width arrays */
   int64_t neq;
   double *cnu[2];   /* per-axis cnu(0:k-1,-1:k-1,1:n) of weno(ncells,k,eps,xedges) (weno.f90:100-112), or NULL */
   double *fcoef[2]; /* per-axis face coefficient, index 0..n like edges(0:n), or NULL */
   double *ccoef[2]; /* per-axis cross coefficient, index = cell along the other axis, or NULL */
   hrweno_time_fn tfn; /* time factor g(t) of the flux or NULL */
   void *tfn_ctx;
   double tfac;        /* g(t) of the evaluation in progress */
};

int hrweno_ref_fv_create(hrweno_ref_fv **out, const hrweno_fv_desc *desc) {
   if (!out || !desc) return HRWENO_EINVAL;
   if (desc->ndim != 1 && desc->ndim != 2) return HRWENO_EINVAL;
   if (desc->nranks > 1) return HRWENO_EINVAL; /* the oracle is always the global problem */
   for (int a = 0; a < desc->ndim; ++a)
      if (hrweno_ref_weno_check(desc->n[a], desc->k, desc->eps)) return HRWENO_EINVAL;
   hrweno_ref_fv *fv = (hrweno_ref_fv *)calloc(1, sizeof(*fv));
   if (!fv) return HRWENO_ENOMEM;
   fv->d = *desc;
   if (fv->d.ndim == 1 && fv->d.rows < 1) fv->d.rows = 1;
   for (int a = 0; a < desc->ndim; ++a) {
      fv->w[a] = (double *)malloc(sizeof(double) * (size_t)desc->n[a]);
      if (desc->grid_kind == HRWENO_GRID_LINEAR)
         hrweno_ref_grid_linear(desc->xmin, desc->xmax, desc->n[a], NULL, NULL, fv->w[a]);
      else
         memcpy(fv->w[a], desc->width[a], sizeof(double) * (size_t)desc->n[a]);
   }
   fv->neq = desc->ndim == 1 ? desc->n[0] * fv->d.rows : desc->n[0] * desc->n[1];
   *out = fv;
   return HRWENO_OK;
}

void hrweno_ref_fv_destroy(hrweno_ref_fv *fv) {
   if (!fv) return;
   for (int a = 0; a < 2; ++a) {
      free(fv->w[a]);
      free(fv->cnu[a]);
      free(fv->fcoef[a]);
      free(fv->ccoef[a]);
   }
   free(fv);
}

/* weno(ncells, k, eps, xedges) for the reconstruction along `axis` (weno.f90:100-112,177) */
int hrweno_ref_fv_set_xedges(hrweno_ref_fv *fv, int axis, const double *xedges) {
   if (!fv || !xedges || axis < 0 || axis >= fv->d.ndim) return HRWENO_EINVAL;
   const int k = fv->d.k;
   const int64_t n = fv->d.n[axis];
   free(fv->cnu[axis]);
   fv->cnu[axis] = (double *)malloc(sizeof(double) * (size_t)(n * k * (k + 1)));
   if (!fv->cnu[axis]) return HRWENO_ENOMEM;
   return hrweno_ref_weno_calc_cnu(n, k, xedges, fv->cnu[axis]);
}

static double *dup_doubles(const double *src, int64_t n) {
   double *p = (double *)malloc(sizeof(double) * (size_t)n);
   if (p) memcpy(p, src, sizeof(double) * (size_t)n);
   return p;
}

/* f(v, x) = (model(v)*cross[i_other])*face[i_face] along `axis`; face[0..n[axis]], cross[0..n[other]-1] */
int hrweno_ref_fv_set_flux_coef(hrweno_ref_fv *fv, int axis, const double *face, const double *cross) {
   if (!fv || axis < 0 || axis >= fv->d.ndim) return HRWENO_EINVAL;
   if (cross && fv->d.ndim != 2) return HRWENO_EINVAL;
   free(fv->fcoef[axis]);
   free(fv->ccoef[axis]);
   fv->fcoef[axis] = face ? dup_doubles(face, fv->d.n[axis] + 1) : NULL;
   fv->ccoef[axis] = cross ? dup_doubles(cross, fv->d.n[1 - axis]) : NULL;
   if ((face && !fv->fcoef[axis]) || (cross && !fv->ccoef[axis])) return HRWENO_ENOMEM;
   return HRWENO_OK;
}

/* f(v, x, t) = ((model(v)*cross)*face)*g(t): g is called once per rhs evaluation with that evaluation's time */
int hrweno_ref_fv_set_flux_time_fn(hrweno_ref_fv *fv, hrweno_time_fn g, void *ctx) {
   if (!fv) return HRWENO_EINVAL;
   fv->tfn = g;
   fv->tfn_ctx = ctx;
   return HRWENO_OK;
}

int64_t hrweno_ref_fv_neq(const hrweno_ref_fv *fv) { return fv->neq; }

/* faces 1..nc-1 then the boundary rule; fe[0..nc] */
static void faces_row(const hrweno_fv_desc *d, int axis, int64_t nc, const double *vl, const double *vr,
                      double *fe, int par, const double *fcoef, const double *cross, const double *tfac) {
   const double coef = d->flux_coef[axis];
   if (fcoef || cross || tfac) { /* x/t-dependent flux: x = [right(i), center_other] (example2:100-101,109-110) */
      for (int64_t i = 1; i <= nc - 1; ++i)
         fe[i] = face_flux_x(d->flux_scheme, d->flux_model, coef, d->alpha, vr[i - 1], vl[i], cross, fcoef ? fcoef + i : NULL, tfac);
   } else if (par) {
#pragma omp parallel for schedule(static)
      for (int64_t i = 1; i <= nc - 1; ++i) /* example1:97-100 ; example2:99-102,108-111 */
         fe[i] = hrweno_ref_face_flux(d->flux_scheme, d->flux_model, coef, d->alpha, vr[i - 1], vl[i]);
   } else {
      for (int64_t i = 1; i <= nc - 1; ++i)
         fe[i] = hrweno_ref_face_flux(d->flux_scheme, d->flux_model, coef, d->alpha, vr[i - 1], vl[i]);
   }
   if (d->bc == HRWENO_BC_COPY_NEIGHBOUR) { /* example1:103-104 */
      fe[0] = fe[1];
      fe[nc] = fe[nc - 1];
   } else { /* example2:117-120 */
      fe[0] = 0;
      fe[nc] = 0;
   }
}

static int rhs_1d(hrweno_ref_fv *fv, const double *v, double *vdot) {
   const hrweno_fv_desc *d = &fv->d;
   const int64_t nc = d->n[0];
   const int par_rows = g_threads > 1 && d->rows > 1;
   const int par_cells = g_threads > 1 && d->rows == 1 && nc >= 65536;
   int err = 0;
#pragma omp parallel if (par_rows)
   {
      double *vl = (double *)malloc(sizeof(double) * (size_t)nc);
      double *vr = (double *)malloc(sizeof(double) * (size_t)nc);
      double *fe = (double *)malloc(sizeof(double) * (size_t)(nc + 1));
      double *vext = (double *)malloc(sizeof(double) * (size_t)(nc + 2 * (d->k - 1)));
      if (!vl || !vr || !fe || !vext) {
#pragma omp atomic write
         err = 1;
      } else {
#pragma omp for schedule(static)
         for (int64_t row = 0; row < d->rows; ++row) {
            const double *vrow = v + row * nc;
            double *orow = vdot + row * nc;
            recon_row(d->k, nc, d->eps, fv->cnu[0], vrow, 1, vl, vr, vext, par_cells); /* example1:93 */
            faces_row(d, 0, nc, vl, vr, fe, par_cells, fv->fcoef[0], NULL, fv->tfn ? &fv->tfac : NULL);
            const double *w = fv->w[0];
            if (par_cells) {
#pragma omp parallel for schedule(static)
               for (int64_t i = 0; i < nc; ++i) orow[i] = -(fe[i + 1] - fe[i]) / w[i]; /* example1:107 */
            } else {
               for (int64_t i = 0; i < nc; ++i) orow[i] = -(fe[i + 1] - fe[i]) / w[i];
            }
         }
      }
      free(vl);
      free(vr);
      free(fe);
      free(vext);
   }
   return err ? HRWENO_ENOMEM : HRWENO_OK;
}

static int rhs_2d(hrweno_ref_fv *fv, const double *v, double *vdot) {
   const hrweno_fv_desc *d = &fv->d;
   const int64_t n1 = d->n[0], n2 = d->n[1];
   /* fedges1(0:nc1, 1:nc2) and fedges2(0:nc2, 1:nc1)  (example2:91) */
   double *f1 = (double *)malloc(sizeof(double) * (size_t)((n1 + 1) * n2));
   double *f2 = (double *)malloc(sizeof(double) * (size_t)((n2 + 1) * n1));
   if (!f1 || !f2) {
      free(f1);
      free(f2);
      return HRWENO_ENOMEM;
   }
   int err = 0;
#pragma omp parallel if (g_threads > 1)
   {
      const int64_t nmax = n1 > n2 ? n1 : n2;
      double *vl = (double *)malloc(sizeof(double) * (size_t)nmax);
      double *vr = (double *)malloc(sizeof(double) * (size_t)nmax);
      double *vext = (double *)malloc(sizeof(double) * (size_t)(nmax + 2 * (d->k - 1)));
      if (!vl || !vr || !vext) {
#pragma omp atomic write
         err = 1;
      } else {
#pragma omp for schedule(static) nowait
         for (int64_t j = 0; j < n2; ++j) { /* example2:97-103: contiguous rows */
            recon_row(d->k, n1, d->eps, fv->cnu[0], v + j * n1, 1, vl, vr, vext, 0);
            faces_row(d, 0, n1, vl, vr, f1 + j * (n1 + 1), 0, fv->fcoef[0], fv->ccoef[0] ? fv->ccoef[0] + j : NULL, fv->tfn ? &fv->tfac : NULL);
         }
#pragma omp for schedule(static)
         for (int64_t i = 0; i < n1; ++i) { /* example2:106-112: stride-nc1 columns */
            recon_row(d->k, n2, d->eps, fv->cnu[1], v + i, n1, vl, vr, vext, 0);
            faces_row(d, 1, n2, vl, vr, f2 + i * (n2 + 1), 0, fv->fcoef[1], fv->ccoef[1] ? fv->ccoef[1] + i : NULL, fv->tfn ? &fv->tfac : NULL);
         }
         const double *w1 = fv->w[0], *w2 = fv->w[1];
#pragma omp for schedule(static)
         for (int64_t j = 0; j < n2; ++j) /* example2:123-127 */
            for (int64_t i = 0; i < n1; ++i)
               vdot[j * n1 + i] = -(f1[j * (n1 + 1) + i + 1] - f1[j * (n1 + 1) + i]) / w1[i] -
                                  (f2[i * (n2 + 1) + j + 1] - f2[i * (n2 + 1) + j]) / w2[j];
      }
      free(vl);
      free(vr);
      free(vext);
   }
   free(f1);
   free(f2);
   return err ? HRWENO_ENOMEM : HRWENO_OK;
}

int hrweno_ref_fv_rhs(hrweno_ref_fv *fv, double t, const double *v, double *vdot) {
   if (!fv || !v || !vdot) return HRWENO_EINVAL;
   if (fv->tfn) fv->tfac = fv->tfn(fv->tfn_ctx, t); /* t enters through the separable factor g(t) only; x through the coefficient arrays */
   return fv->d.ndim == 1 ? rhs_1d(fv, v, vdot) : rhs_2d(fv, v, vdot);
}

/* ------------------------------------------------------------------------------------------
 * TVD integrators -- src/hrweno_tvdode.f90
 * ---------------------------------------------------------------------------------------- */
struct hrweno_ref_ode {
   hrweno_ref_rhs_fn fu;
   void *ctx;
   int64_t neq;
   int order;
   int64_t fevals;
   int64_t rhs_calls;
   int istate;
   int is_ms;
   double *ui, *udot;
   double *uold, *udotold; /* (neq,4), column c (1-based) at +(c-1)*neq */
};

/* is_done -- tvdode.f90:273-284: (t - tout)*sign(1,dt) > 0 */
int hrweno_ref_is_done(double t, double tout, double dt) { return (t - tout) * copysign(1.0, dt) > 0.0; }

static void call_fu(hrweno_ref_ode *o, double t, const double *u, double *udot) {
   o->fu(o->ctx, t, o->neq, u, udot);
   o->rhs_calls++;
}

static int ode_create(hrweno_ref_ode **out, hrweno_ref_rhs_fn fu, void *ctx, int64_t neq, int order, int ms) {
   if (!out || !fu) return HRWENO_EINVAL;
   if (!(neq > 0)) return HRWENO_EINVAL;                    /* :80-84, :189-193 */
   if (!(order >= 1 && order <= 3)) return HRWENO_EINVAL; /* :86-90 */
   hrweno_ref_ode *o = (hrweno_ref_ode *)calloc(1, sizeof(*o));
   if (!o) return HRWENO_ENOMEM;
   o->fu = fu;
   o->ctx = ctx;
   o->neq = neq;
   o->order = order;
   o->is_ms = ms;
   o->ui = (double *)malloc(sizeof(double) * (size_t)neq);
   o->udot = (double *)malloc(sizeof(double) * (size_t)neq);
   if (ms) {
      o->uold = (double *)malloc(sizeof(double) * (size_t)neq * 4);
      o->udotold = (double *)malloc(sizeof(double) * (size_t)neq * 4);
   }
   o->istate = 1; /* :93, :199 */
   *out = o;
   return HRWENO_OK;
}

int hrweno_ref_rktvd_create(hrweno_ref_ode **out, hrweno_ref_rhs_fn fu, void *ctx, int64_t neq, int order) {
   return ode_create(out, fu, ctx, neq, order, 0);
}

int hrweno_ref_mstvd_create(hrweno_ref_ode **out, hrweno_ref_rhs_fn fu, void *ctx, int64_t neq) {
   return ode_create(out, fu, ctx, neq, 3, 1); /* order = 3, :195 */
}

static void fv_trampoline(void *ctx, double t, int64_t neq, const double *u, double *udot) {
   (void)neq;
   hrweno_ref_fv_rhs((hrweno_ref_fv *)ctx, t, u, udot);
}

int hrweno_ref_rktvd_create_fv(hrweno_ref_ode **out, hrweno_ref_fv *fv, int order) {
   if (!fv) return HRWENO_EINVAL;
   return ode_create(out, fv_trampoline, fv, fv->neq, order, 0);
}

int hrweno_ref_mstvd_create_fv(hrweno_ref_ode **out, hrweno_ref_fv *fv) {
   if (!fv) return HRWENO_EINVAL;
   return ode_create(out, fv_trampoline, fv, fv->neq, 3, 1);
}

void hrweno_ref_ode_destroy(hrweno_ref_ode *o) {
   if (!o) return;
   free(o->ui);
   free(o->udot);
   free(o->uold);
   free(o->udotold);
   free(o);
}

int64_t hrweno_ref_ode_fevals(const hrweno_ref_ode *o) { return o->fevals; }
int64_t hrweno_ref_ode_rhs_calls(const hrweno_ref_ode *o) { return o->rhs_calls; }
int hrweno_ref_ode_istate(const hrweno_ref_ode *o) { return o->istate; }

#define PAR_VEC _Pragma("omp parallel for schedule(static) if (g_threads > 1 && n >= 65536)")

/* rktvd_integrate -- tvdode.f90:97-178 */
static void rk_integrate(hrweno_ref_ode *o, double *u, double *t, double tout, double dt, int itask) {
   if (o->istate < 1) return;                   /* :126 */
   if (hrweno_ref_is_done(*t, tout, dt)) return; /* :127 */
   const int64_t n = o->neq;
   double *ui = o->ui, *udot = o->udot;
   switch (o->order) {
   case 1: /* :136-143 */
      for (;;) {
         call_fu(o, *t, u, udot);
         PAR_VEC for (int64_t i = 0; i < n; ++i) u[i] = u[i] + dt * udot[i];
         *t = *t + dt;
         o->fevals += 1;
         if (hrweno_ref_is_done(*t, tout, dt) || itask == 2) break;
      }
      break;
   case 2: /* :147-156 */
      for (;;) {
         call_fu(o, *t, u, udot);
         PAR_VEC for (int64_t i = 0; i < n; ++i) ui[i] = u[i] + dt * udot[i];
         call_fu(o, *t + dt, ui, udot);
         PAR_VEC for (int64_t i = 0; i < n; ++i) u[i] = (u[i] + ui[i] + dt * udot[i]) / 2;
         *t = *t + dt;
         o->fevals += 2;
         if (hrweno_ref_is_done(*t, tout, dt) || itask == 2) break;
      }
      break;
   default: /* :160-171 */
      for (;;) {
         call_fu(o, *t, u, udot);
         PAR_VEC for (int64_t i = 0; i < n; ++i) ui[i] = u[i] + dt * udot[i];
         call_fu(o, *t + dt, ui, udot);
         PAR_VEC for (int64_t i = 0; i < n; ++i) ui[i] = (3 * u[i] + ui[i] + dt * udot[i]) / 4;
         call_fu(o, *t + dt / 2, ui, udot);
         {
            const double dt2 = 2 * dt; /* 2*dt*udot parses as (2*dt)*udot */
            PAR_VEC for (int64_t i = 0; i < n; ++i) u[i] = (u[i] + 2 * ui[i] + dt2 * udot[i]) / 3;
         }
         *t = *t + dt;
         o->fevals += 3;
         if (hrweno_ref_is_done(*t, tout, dt) || itask == 2) break;
      }
   }
   if (o->istate == 1) o->istate = 2; /* :176 */
}

/* mstvd_integrate -- tvdode.f90:203-271 */
static void ms_integrate(hrweno_ref_ode *o, double *u, double *t, double tout, double dt) {
   if (o->istate < 1) return;                   /* :228 */
   if (hrweno_ref_is_done(*t, tout, dt)) return; /* :229 */
   const int64_t n = o->neq;
   double *ui = o->ui, *udot = o->udot, *uold = o->uold, *udotold = o->udotold;
   if (o->istate == 1) { /* :236-249 */
      hrweno_ref_ode *start = NULL;
      ode_create(&start, o->fu, o->ctx, n, o->order, 0);
      for (int c = o->order + 1; c >= 1; --c) {
         memcpy(uold + (size_t)(c - 1) * n, u, sizeof(double) * (size_t)n);
         call_fu(o, *t, u, udotold + (size_t)(c - 1) * n);
         rk_integrate(start, u, t, *t + 2 * dt, dt, 2);
      }
      o->fevals = start->fevals; /* :246 -- the 4 extra fu calls above are not counted */
      o->rhs_calls += start->rhs_calls;
      o->istate = 2;
      hrweno_ref_ode_destroy(start);
   }
   const double c50 = 50 * dt, c10 = 10 * dt;
   for (;;) { /* :252-268 */
      if (hrweno_ref_is_done(*t, tout, dt)) break;
      call_fu(o, *t, u, udot);
      {
         const double *uo4 = uold + (size_t)3 * n, *udo4 = udotold + (size_t)3 * n;
         PAR_VEC for (int64_t i = 0; i < n; ++i)
            ui[i] = (25 * u[i] + c50 * udot[i] + 7 * uo4[i] + c10 * udo4[i]) / 32; /* :257 */
      }
      *t = *t + dt;
      o->fevals += 1;
      /* eoshift(.., shift=-1, dim=2): column c -> c+1, column 4 dropped (:262-263) */
      memmove(udotold + n, udotold, sizeof(double) * (size_t)n * 3);
      memmove(uold + n, uold, sizeof(double) * (size_t)n * 3);
      memcpy(udotold, udot, sizeof(double) * (size_t)n); /* :264 */
      memcpy(uold, u, sizeof(double) * (size_t)n);       /* :265 */
      memcpy(u, ui, sizeof(double) * (size_t)n);         /* :266 */
   }
}

int hrweno_ref_ode_integrate(hrweno_ref_ode *o, double *u, double *t, double tout, double dt, int itask) {
   if (!o || !u || !t) return HRWENO_EINVAL;
   if (o->is_ms)
      ms_integrate(o, u, t, tout, dt);
   else
      rk_integrate(o, u, t, tout, dt, itask);
   return HRWENO_OK;
}

/* test/test_tvdode.f90:15 and :107-111 */
void hrweno_ref_test_ode_rhs(void *ctx, double t, int64_t neq, const double *u, double *udot) {
   (void)ctx;
   (void)t;
   for (int64_t ii = 1; ii <= neq; ++ii) {
      const double a = -1.0 + (double)(ii - 1) * 4 / (double)(neq - 1);
      udot[ii - 1] = a * u[ii - 1];
   }
}
