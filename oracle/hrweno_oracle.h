/*
 * hrweno_oracle.h -- CPU oracle for the HR-WENO finite-volume update path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the
 * shipped path (hr-weno_b200/) never links or calls it.
 *
 * Parity status: the reference is Fortran and no Fortran compiler exists in this
 * image, so a reference BINARY cannot be run here (oracle/_ref is empty).  The oracle
 * is pinned three ways:
 *   1. to the reference's own SOURCE TEXT executed mechanically: tools/f90exec/f90py.py
 *      translates the procedures of /root/reference/src and of both example programs
 *      statement by statement into Python and runs them on IEEE binary64 (nothing
 *      restated by hand); this file reproduces those outputs BIT FOR BIT --
 *      reconstruct, calc_cnu, both fluxes, rktvd 1-3 / mstvd, example1 as shipped over
 *      all 101 outputs (+ k x order sweep, Lax-Friedrichs variant), example2 (40x40 over
 *      all outputs, 250x250 first outputs) and
 *      example2 with geometric grids + growth terms (fixtures tests/golden/ref_exec_*.npz,
 *      generator tests/golden/make_ref_exec_golden.py, tests/test_reference_source_exec.py);
 *   2. to every assertion of the reference's own test-drive suites (test/test_hrweno.f90,
 *      test_tvdode.f90, test_fluxes.f90, test_grid.f90 -- tolerances 1e-3..1e-8; the
 *      reference holds no golden vectors);
 *   3. to an independent NumPy restatement (oracle/np_oracle.py) bit for bit.
 * What remains unpinned: the output of a gfortran BINARY.  (1) is the source's operation
 * order under IEEE evaluation without contraction or re-association -- what gfortran
 * does on x86-64 without -ffast-math / -march=native -- not a compiler run.
 */
#ifndef HRWENO_ORACLE_H
#define HRWENO_ORACLE_H

#include <stdint.h>
#include "hrweno_b200.h" /* hrweno_fv_desc and the enums only */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hrweno_ref_fv hrweno_ref_fv;
typedef struct hrweno_ref_ode hrweno_ref_ode;
typedef void (*hrweno_ref_rhs_fn)(void *ctx, double t, int64_t neq, const double *u, double *udot);

void hrweno_ref_set_threads(int nthreads); /* OpenMP threads for the baseline legs (1 = as shipped) */
int hrweno_ref_max_threads(void);

/* weno.f90:12-21 -- d(0:k-1) and c(j,r) at c[j + k*(r+1)] */
int hrweno_ref_tables(int k, double *d, double *c);
/* weno.f90:221-297 -- cnu(j,r,i) at cnu[j + k*((r+1) + (k+1)*(i-1))], i = 1..nc */
int hrweno_ref_weno_calc_cnu(int64_t nc, int k, const double *xedges, double *cnu);
/* weno.f90:129-219 -- v[i*incv], i = 0..nc-1 */
int hrweno_ref_weno_reconstruct(int64_t nc, int k, double eps, const double *cnu, const double *v,
                                int64_t incv, double *vl, double *vr);
/* weno.f90:71-98 validation: returns HRWENO_EINVAL where the reference would error stop */
int hrweno_ref_weno_check(int64_t nc, int k, double eps);

/* fluxes.f90:22-45, 47-76 */
double hrweno_ref_lax_friedrichs(hrweno_flux_fn f, void *ctx, double vm, double vp, const double *x, int nx,
                                 double t, double alpha);
double hrweno_ref_godunov(hrweno_flux_fn f, void *ctx, double vm, double vp, const double *x, int nx, double t);
/* closed-set flux models used by the fused path */
double hrweno_ref_flux_model(int model, double coef, double v);
double hrweno_ref_face_flux(int scheme, int model, double coef, double alpha, double vm, double vp);

/* grids.f90:41-84, 232-250 -- edges(0:n), center(n), width(n) of grid1%linear */
void hrweno_ref_grid_linear(double xmin, double xmax, int64_t n, double *edges, double *center, double *width);

/* example1:72-109 / example2:73-129 */
int hrweno_ref_fv_create(hrweno_ref_fv **out, const hrweno_fv_desc *desc);
void hrweno_ref_fv_destroy(hrweno_ref_fv *fv);
int64_t hrweno_ref_fv_neq(const hrweno_ref_fv *fv);
/* weno(ncells,k,eps,xedges): per-cell tables cnu for the sweep along `axis` (weno.f90:100-112,177,221-297) */
int hrweno_ref_fv_set_xedges(hrweno_ref_fv *fv, int axis, const double *xedges);
/* x-dependent flux f = (model(v)*cross[i_other])*face[i_face] (the growth terms hinted at example2:140,153) */
int hrweno_ref_fv_set_flux_coef(hrweno_ref_fv *fv, int axis, const double *face, const double *cross);
int hrweno_ref_fv_set_flux_time_fn(hrweno_ref_fv *fv, hrweno_time_fn g, void *ctx);
int hrweno_ref_fv_rhs(hrweno_ref_fv *fv, double t, const double *v, double *vdot);

/* tvdode.f90:69-95, 97-178, 180-201, 203-271, 273-284 */
int hrweno_ref_rktvd_create(hrweno_ref_ode **out, hrweno_ref_rhs_fn fu, void *ctx, int64_t neq, int order);
int hrweno_ref_mstvd_create(hrweno_ref_ode **out, hrweno_ref_rhs_fn fu, void *ctx, int64_t neq);
int hrweno_ref_rktvd_create_fv(hrweno_ref_ode **out, hrweno_ref_fv *fv, int order);
int hrweno_ref_mstvd_create_fv(hrweno_ref_ode **out, hrweno_ref_fv *fv);
void hrweno_ref_ode_destroy(hrweno_ref_ode *ode);
int hrweno_ref_ode_integrate(hrweno_ref_ode *ode, double *u, double *t, double tout, double dt, int itask);
int64_t hrweno_ref_ode_fevals(const hrweno_ref_ode *ode);
int64_t hrweno_ref_ode_rhs_calls(const hrweno_ref_ode *ode); /* real number of fu calls (mstvd under-counts) */
int hrweno_ref_ode_istate(const hrweno_ref_ode *ode);
int hrweno_ref_is_done(double t, double tout, double dt);

/* test ODE of test/test_tvdode.f90:107-111: udot = a*u with a = linspace(-1,3,10) */
void hrweno_ref_test_ode_rhs(void *ctx, double t, int64_t neq, const double *u, double *udot);

#ifdef __cplusplus
}
#endif
#endif
