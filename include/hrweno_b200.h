/*
 * hrweno_b200.h -- C ABI of the B200-native HR-WENO finite-volume update path.
 *
 * The reference (HugoMVale/HR-WENO) is pure Fortran and has no FFI boundary of
 * its own; its surface for this path is the set of module procedures cited
 * below (paths relative to the reference tree).  This header is the boundary a
 * thin ISO_C_BINDING shim (fortran/) binds so that the
 * Fortran-side names keep working:
 *
 *   weno(ncells,k,eps,xedges)         src/hrweno_weno.f90:54-127   -> hrweno_weno_create
 *   weno%reconstruct(v,vl,vr)         src/hrweno_weno.f90:129-219  -> hrweno_weno_reconstruct[_batch|_dev]
 *   weno%cnu                          src/hrweno_weno.f90:41,221-297 -> hrweno_weno_get_cnu
 *   godunov(f,vm,vp,x,t)              src/hrweno_fluxes.f90:47-76  -> hrweno_godunov / hrweno_flux_faces
 *   lax_friedrichs(f,vm,vp,x,t,alpha) src/hrweno_fluxes.f90:22-45  -> hrweno_lax_friedrichs / hrweno_flux_faces
 *   rktvd(fu,neq,order)               src/hrweno_tvdode.f90:69-95  -> hrweno_rktvd_create[_fused]
 *   rktvd%integrate(u,t,tout,dt,itask)src/hrweno_tvdode.f90:97-178 -> hrweno_ode_integrate[_dev]
 *   mstvd(fu,neq)                     src/hrweno_tvdode.f90:180-201-> hrweno_mstvd_create[_fused]
 *   mstvd%integrate(u,t,tout,dt)      src/hrweno_tvdode.f90:203-271-> hrweno_ode_integrate[_dev]
 *   example rhs (1D)                  example/example1_burgers_1d_fv.f90:72-109 -> hrweno_fv_* (ndim=1)
 *   example rhs (2D, split)           example/example2_pbe_2d_fv.f90:73-129     -> hrweno_fv_* (ndim=2)
 *
 * Conventions: plain pointers and sizes only; every entry point returns an
 * int status (0 = ok) unless it is a pure value function; the message of the
 * last failure on the calling thread is available from hrweno_last_error().
 * "host" pointers are ordinary CPU memory; "dev" pointers are CUDA device
 * memory on the current device, and `stream` is a cudaStream_t passed as
 * void* (NULL = legacy default stream).  All arithmetic is IEEE fp64
 * (rk = real64, src/hrweno_kinds.F90:16).
 *
 * There is no CPU fallback: every compute entry point runs CUDA kernels and
 * fails with HRWENO_ECUDA when no device is usable.
 */
#ifndef HRWENO_B200_H
#define HRWENO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HRWENO_ABI_VERSION 1

/* ---- status codes --------------------------------------------------------- */
enum hrweno_status {
   HRWENO_OK = 0,
   HRWENO_EINVAL = 1, /* invalid argument: the reference would `error stop` (weno.f90:75-108, tvdode.f90:83,89) */
   HRWENO_ECUDA = 2,  /* CUDA runtime failure or no device */
   HRWENO_ENOMEM = 3,
   HRWENO_ESTATE = 4, /* object in error state (istate < 1, tvdode.f90:126,228) */
   HRWENO_ECOMM = 5   /* multi-GPU halo exchange failure / timeout */
};

/* ---- enums for the fused finite-volume path ------------------------------- */
enum hrweno_flux_model {
   HRWENO_FLUX_BURGERS = 0, /* f(v) = (v**2)/2          example1:120 */
   HRWENO_FLUX_LINEAR = 1   /* f(v) = a*v (a=1: f = v)  example2:140,153 */
};
enum hrweno_flux_scheme {
   HRWENO_SCHEME_GODUNOV = 0,       /* fluxes.f90:67-74 */
   HRWENO_SCHEME_LAX_FRIEDRICHS = 1 /* fluxes.f90:43, alpha supplied by the caller */
};
enum hrweno_bc {
   HRWENO_BC_COPY_NEIGHBOUR = 0, /* fedges(0)=fedges(1), fedges(nc)=fedges(nc-1)   example1:103-104 */
   HRWENO_BC_ZERO_FLUX = 1       /* fedges(0)=fedges(nc)=0                          example2:117-120 */
};
enum hrweno_grid_kind {
   HRWENO_GRID_WIDTH_ARRAY = 0, /* width(:) arrays streamed from memory (any grid1 kind) */
   HRWENO_GRID_LINEAR = 1       /* 1D only: width recomputed on device bit-identically to grid1%linear
                                   (grids.f90:76-79,247): edges(i)=xmin+rx*i, width=edges(i)-edges(i-1) */
};
enum hrweno_mode {
   HRWENO_MODE_STRICT = 0, /* reference operation order, no FMA contraction: bit-identical to the oracle */
   HRWENO_MODE_FAST = 1    /* same formulas, division-light weights and FMA contraction; within the
                              north-star tolerances (1e-12 normwise per output time, few ULP reconstruct).
                              Magnitude range: the division-light weights form products of smoothness terms
                              that grow like v^9 where the reference forms v^4, so runs of cells are
                              reconstructed directly while every |v| of their stencil window is < 2^100
                              (1.3e30); windows holding larger values are rescaled by an exact power of two
                              first (the scheme is homogeneous), so fast mode stays finite wherever the
                              reference does (|v| up to ~1e77) at the same few-ULP accuracy.  eps must be
                              > epsilon(1d0) as in the reference (weno.f90:92-96). */
};

/* ---- opaque handles ------------------------------------------------------- */
typedef struct hrweno_weno hrweno_weno; /* type(weno)          weno.f90:23-46 */
typedef struct hrweno_fv hrweno_fv;     /* the example `rhs`: reconstruct + flux + divergence */
typedef struct hrweno_ode hrweno_ode;   /* type(rktvd)/type(mstvd)  tvdode.f90:14-48 */

/* ---- library -------------------------------------------------------------- */
int hrweno_abi_version(void);
const char *hrweno_last_error(void); /* thread-local, never NULL */
int hrweno_device_count(void);       /* number of usable CUDA devices (0 if none) */

/* ---- WENO reconstruction (src/hrweno_weno.f90) ---------------------------- */

/* weno_init (weno.f90:54-127).  Validation mirrors the reference: ncells > 0,
 * 1 <= k <= 3, eps > epsilon(1d0).  xedges == NULL selects the uniform-grid
 * tables c1/c2/c3 (weno.f90:12-21); otherwise xedges[0..ncells] are the cell
 * edges and cnu(0:k-1,-1:k-1,1:ncells) is computed per weno_calc_cnu
 * (weno.f90:221-297, Shu eq. 2.20). */
int hrweno_weno_create(hrweno_weno **out, int64_t ncells, int k, double eps, const double *xedges);
void hrweno_weno_destroy(hrweno_weno *w);
int hrweno_weno_info(const hrweno_weno *w, int64_t *ncells, int *k, double *eps, int *uniform_grid);
/* Extension (no counterpart in the reference): arithmetic of reconstruct on uniform tables, HRWENO_MODE_STRICT
 * (default; bit-identical to the reference's operation order) or HRWENO_MODE_FAST (same scheme in difference form,
 * a few ULP of max|v|).  Non-uniform (cnu) and strided reconstructions always run in the reference order. */
int hrweno_weno_set_mode(hrweno_weno *w, int mode);
/* copy of cnu in the reference's column-major order cnu(j,r,i): index j + k*((r+1) + (k+1)*(i-1)) */
int hrweno_weno_get_cnu(const hrweno_weno *w, double *cnu_host);

/* weno_reconstruct (weno.f90:129-219): v(ncells) -> vl(ncells), vr(ncells), host memory.
 * Re-entrant for one handle (the reference procedure is pure with intent(in) self). */
int hrweno_weno_reconstruct(const hrweno_weno *w, const double *v, double *vl, double *vr);
/* The same call in subroutine form (status through the last argument).  The reference's `reconstruct` is `pure`
 * (weno.f90:129) and is called from `pure subroutine rhs` (example1:72,93): Fortran lets an interface body promise
 * `pure` for a C procedure, but a pure FUNCTION may not have intent(out) arguments, so the Fortran shim binds this one
 * as `pure subroutine` and the reference's programs keep their `pure` prefixes. */
void hrweno_weno_reconstruct_s(const hrweno_weno *w, const double *v, double *vl, double *vr, int *status);
/* `rows` independent rows of ncells cells; element (row, i) of v at v[row*ldv + i*incv]
 * (incv > 1 expresses the strided sections of example2:107); outputs at [row*ldo + i]. */
int hrweno_weno_reconstruct_batch(const hrweno_weno *w, int64_t rows, const double *v, int64_t ldv,
                                  int64_t incv, double *vl, double *vr, int64_t ldo);
/* same, device pointers, asynchronous on `stream` */
int hrweno_weno_reconstruct_dev(const hrweno_weno *w, int64_t rows, const double *v_dev, int64_t ldv,
                                int64_t incv, double *vl_dev, double *vr_dev, int64_t ldo, void *stream);

/* ---- numerical fluxes (src/hrweno_fluxes.f90) ------------------------------ */

/* flux callback: abstract interface `flux(u, x(:), t)` (fluxes.f90:12-18) */
typedef double (*hrweno_flux_fn)(void *ctx, double u, const double *x, int nx, double t);
/* pointwise host helpers with a user callback, exactly the reference signatures */
double hrweno_lax_friedrichs(hrweno_flux_fn f, void *ctx, double vm, double vp, const double *x, int nx,
                             double t, double alpha);
double hrweno_godunov(hrweno_flux_fn f, void *ctx, double vm, double vp, const double *x, int nx, double t);
/* closed-set device evaluation over n faces: h[i] = scheme(model; vm[i], vp[i]) (host pointers) */
int hrweno_flux_faces(int flux_scheme, int flux_model, double flux_coef, double alpha, int64_t n,
                      const double *vm, const double *vp, double *h);

/* ---- fused finite-volume right-hand side ---------------------------------- */
typedef struct hrweno_fv_desc {
   int32_t abi_version; /* HRWENO_ABI_VERSION */
   int32_t ndim;        /* 1 or 2 */
   int64_t n[2];        /* cells per dimension; n[0] is the contiguous axis; storage u[(j)*n[0]+i] (example2:50) */
   int64_t rows;        /* ndim==1: number of independent 1D problems stored row after row (>=1) */
   int32_t k;           /* 1..3, order 2k-1 */
   int32_t flux_model;  /* enum hrweno_flux_model */
   int32_t flux_scheme; /* enum hrweno_flux_scheme */
   int32_t bc;          /* enum hrweno_bc */
   int32_t grid_kind;   /* enum hrweno_grid_kind */
   int32_t mode;        /* enum hrweno_mode */
   double eps;          /* WENO smoothing factor (> epsilon) */
   double flux_coef[2]; /* LINEAR: a per dimension */
   double alpha;        /* LAX_FRIEDRICHS: max|f'(v)|, caller supplied (fluxes.f90:40) */
   double xmin, xmax;   /* GRID_LINEAR: domain of the *global* grid */
   const double *width[2]; /* GRID_WIDTH_ARRAY: host arrays width[d][0..n[d]-1] (copied) */
   /* slab decomposition across GPUs (one process per GPU).  The decomposed axis is the
      slowest one (axis 0 for ndim==1, axis 1 for ndim==2).  nranks <= 1: single GPU. */
   int32_t rank, nranks;
   int64_t global_n;      /* global number of cells along the decomposed axis */
   int64_t global_offset; /* global index of this slab's first cell along that axis */
} hrweno_fv_desc;

int hrweno_fv_create(hrweno_fv **out, const hrweno_fv_desc *desc);
void hrweno_fv_destroy(hrweno_fv *fv);
int64_t hrweno_fv_neq(const hrweno_fv *fv); /* local number of unknowns */
/* one evaluation of the example `rhs` (example1:72-109 / example2:73-129): vdot = L(v), host memory */
int hrweno_fv_rhs(hrweno_fv *fv, double t, const double *v, double *vdot);
int hrweno_fv_rhs_dev(hrweno_fv *fv, double t, const double *v_dev, double *vdot_dev, void *stream);

/* Non-uniform grids in the fused path: the sweep along `axis` (0 = x1, 1 = x2) reconstructs with the per-cell tables
 * cnu(:,:,i) of `weno(ncells, k, eps, xedges)` (weno.f90:100-112, 177, 221-297) instead of c1/c2/c3.  xedges[0..n[axis]]
 * (host) are the cell edges of that axis (grid1%edges, grids.f90:232-250; any grid1 kind).  The cell widths stay the
 * ones of the descriptor.  Call after hrweno_fv_create and before the first rhs / integrate call; from then on this
 * operator runs the general stage kernels: the reference operation order in MODE_STRICT (bit-identical to the oracle);
 * MODE_FAST keeps the reference order in 1D and uses division-light weights with FMA contraction in 2D (1e-12 normwise per
 * output time, like the uniform fast path).  Slabs (nranks > 1): along the
 * DECOMPOSED axis pass the GLOBAL edge array xedges[0..global_n] -- the tables of a slab's first and last cells, and of
 * the neighbour cells its interface faces need, depend on edges beyond the slab; other axes: the local (= global) edges. */
int hrweno_fv_set_xedges(hrweno_fv *fv, int axis, const double *xedges);
/* x-dependent fluxes in the fused path.  The reference hands the face coordinates to the flux callback
 * (`f(u, x(:), t)`, fluxes.f90:12-18; example2:100-101 passes [right1(i), center2(j)], :109-110 [center1(i), right2(j)])
 * and hints at the growth terms `v*x(1)**2` and `v*x(1)*x(2)` (example2:140,153).  Closed-set form of those:
 *     f(v, x) = (model(v) * cross_coef[c]) * face_coef[f]        evaluated left to right, an absent (NULL) factor skipped
 * for the faces along `axis`: face_coef[0..n[axis]] is indexed like edges(0:n) (face f lies between cells f-1 and f;
 * entries 0 and n belong to the boundary rule and are not read), cross_coef[0..n[other axis]-1] by the cell index
 * along the other axis (ndim == 2 only).  `flux1 = v*x(1)**2`: face_coef = edges1**2; `flux2 = v*x(1)*x(2)`:
 * cross_coef = center1, face_coef = edges2.  Same calling rules as hrweno_fv_set_xedges. */
int hrweno_fv_set_flux_coef(hrweno_fv *fv, int axis, const double *face_coef, const double *cross_coef);
/* t-dependent fluxes in the fused path.  The reference's flux receives the time of the evaluation (`f(u, x(:), t)`,
 * fluxes.f90:12-18; the integrators call the rhs at t, t+dt and t+dt/2, tvdode.f90:162-166, and at t, :256).  Closed-set
 * form: a separable time factor,
 *     f(v, x, t) = ((model(v) * cross_coef[c]) * face_coef[f]) * g(t)
 * with g a HOST function evaluated once per right-hand-side evaluation at that evaluation's time and handed to the stage
 * kernel as a scalar, so the stage stays one fused launch (no PCIe round trip per stage, unlike the host-integrand
 * constructors).  g == NULL removes the factor.  Same calling rules as hrweno_fv_set_xedges (general stage kernel). */
typedef double (*hrweno_time_fn)(void *ctx, double t);
int hrweno_fv_set_flux_time_fn(hrweno_fv *fv, hrweno_time_fn g, void *ctx);

/* multi-GPU plumbing: each rank exports a 64-byte CUDA IPC handle of its halo mailbox and
 * imports the handles of its left/right neighbours (NULL at a physical boundary). */
#define HRWENO_IPC_HANDLE_BYTES 64
/* Extension (the reference takes the Lax-Friedrichs alpha from the caller, fluxes.f90:40, and has no reduction):
 * max |f'(v)| over this rank's cells of the dense device vector v_dev, written to the device scalar out_dev,
 * asynchronous on `stream` (Burgers: max|v|; linear flux: max|a|).  NaN cells are ignored.  Ranks combine their values
 * with one max all-reduce (hr-weno_b200/slab.py: global_max_wavespeed) and pass the result to hrweno_fv_set_alpha,
 * which takes effect from the next stage launch.  Off in every parity run: results then depend on alpha's history. */
int hrweno_fv_max_wavespeed_dev(hrweno_fv *fv, const double *v_dev, double *out_dev, void *stream);
int hrweno_fv_set_alpha(hrweno_fv *fv, double alpha);
int hrweno_fv_export_halo(hrweno_fv *fv, void *handle_out);
int hrweno_fv_import_halo(hrweno_fv *fv, const void *left_handle, const void *right_handle);
/* synchronises the device and reports HRWENO_ECOMM if a halo wait timed out (a neighbour rank died) */
int hrweno_fv_halo_status(hrweno_fv *fv);

/* ---- one process, N GPUs (SURVEY 8b "ngpus", 8e) ---------------------------------------------------------------
 * The reference is a single-process program; these entry points let a Fortran / C++ host keep that shape and still use
 * every GPU of the box.  `desc` describes the GLOBAL problem (rank / nranks / global_* are ignored): the slowest axis
 * (x for ndim == 1 with rows == 1, x2 for ndim == 2) is cut into `ngpus` contiguous slabs, independent rows
 * (ndim == 1, rows > 1) are dealt out without halos.  The library opens the devices (`devices` == NULL: 0..ngpus-1;
 * ngpus <= 0: all visible), enables peer access, wires the halo mailboxes of neighbouring slabs and runs one host
 * thread + one stream per device inside each call; the k halo cells (rows) per stage travel as peer stores over
 * NVLink from inside the stage kernels.  Results are bit-identical to the single-GPU run of the same mode.
 * The integrators keep the semantics of rktvd%integrate / mstvd%integrate (tvdode.f90:97-178, 203-271); u is the
 * caller's global vector in the reference's storage order (example2:50). */
typedef struct hrweno_mgpu hrweno_mgpu;
int hrweno_mgpu_create(hrweno_mgpu **out, const hrweno_fv_desc *desc, int ngpus, const int *devices);
void hrweno_mgpu_destroy(hrweno_mgpu *m);
int hrweno_mgpu_ngpus(const hrweno_mgpu *m);
/* slab `rank`: its device, and offset / count of its unknowns inside the global vector */
int hrweno_mgpu_slab(const hrweno_mgpu *m, int rank, int *device, int64_t *offset, int64_t *count);
/* general operators on the slabs, GLOBAL arrays (the library cuts them like the grid): hrweno_fv_set_xedges /
 * _set_flux_coef / _set_flux_time_fn for the whole problem */
int hrweno_mgpu_set_xedges(hrweno_mgpu *m, int axis, const double *xedges);
int hrweno_mgpu_set_flux_coef(hrweno_mgpu *m, int axis, const double *face_coef, const double *cross_coef);
int hrweno_mgpu_set_flux_time_fn(hrweno_mgpu *m, hrweno_time_fn g, void *ctx);
int hrweno_mgpu_rktvd(hrweno_mgpu *m, int order); /* rktvd(fu, neq, order) with the fused rhs, tvdode.f90:69-95 */
int hrweno_mgpu_mstvd(hrweno_mgpu *m);            /* mstvd(fu, neq), tvdode.f90:180-201 */
/* integrate with the global HOST vector u (copied to the slabs and back inside the call) */
int hrweno_mgpu_integrate(hrweno_mgpu *m, double *u, double *t, double tout, double dt, int itask);
/* device-resident state between calls: upload once, integrate any number of times, download when output is due */
int hrweno_mgpu_upload(hrweno_mgpu *m, const double *u);
int hrweno_mgpu_integrate_resident(hrweno_mgpu *m, double *t, double tout, double dt, int itask);
int hrweno_mgpu_download(hrweno_mgpu *m, double *u);
/* Extension (north star: "CFL / max-wavespeed reduction"): alpha = max |f'(u)| over ALL slabs of the resident state,
 * reduced on the GPUs (per-slab kernels, then one kernel on the first device reading its peers' partial maxima over
 * NVLink).  install != 0 also makes it the Lax-Friedrichs alpha of every slab from the next stage on. */
int hrweno_mgpu_max_wavespeed(hrweno_mgpu *m, double *alpha_out, int install);
int hrweno_mgpu_set_alpha(hrweno_mgpu *m, double alpha);
int64_t hrweno_mgpu_fevals(const hrweno_mgpu *m);
int64_t hrweno_mgpu_launches(const hrweno_mgpu *m);

/* ---- TVD time integrators (src/hrweno_tvdode.f90) -------------------------- */

/* integrand callback `fu(t, u(:), udot(:))` (tvdode.f90:50-57).  u and udot are DEVICE pointers
 * (the state is device resident); the callback must enqueue its work on `stream`. */
typedef void (*hrweno_rhs_fn)(void *ctx, double t, int64_t neq, const double *u_dev, double *udot_dev,
                              void *stream);

/* the reference's integrand exactly as it is (tvdode.f90:50-57): a HOST procedure on HOST arrays.  The state and the
 * stage combinations stay on the device; u is copied to pinned host memory before every call and udot back after it,
 * so a reference program runs unchanged (its rhs calling weno%reconstruct and godunov), only slower than the fused path. */
typedef void (*hrweno_rhs_host_fn)(void *ctx, double t, int64_t neq, const double *u, double *udot);
int hrweno_rktvd_create_host(hrweno_ode **out, hrweno_rhs_host_fn fu, void *ctx, int64_t neq, int order);
int hrweno_mstvd_create_host(hrweno_ode **out, hrweno_rhs_host_fn fu, void *ctx, int64_t neq);

/* rktvd_init (tvdode.f90:69-95): neq >= 1, 1 <= order <= 3 */
int hrweno_rktvd_create(hrweno_ode **out, hrweno_rhs_fn fu, void *ctx, int64_t neq, int order);
/* mstvd_init (tvdode.f90:180-201) */
int hrweno_mstvd_create(hrweno_ode **out, hrweno_rhs_fn fu, void *ctx, int64_t neq);
/* same integrators with the rhs, flux and stage combination fused into one kernel per stage */
int hrweno_rktvd_create_fused(hrweno_ode **out, hrweno_fv *fv, int order);
int hrweno_mstvd_create_fused(hrweno_ode **out, hrweno_fv *fv);
void hrweno_ode_destroy(hrweno_ode *ode);

/* rktvd_integrate / mstvd_integrate (tvdode.f90:97-178, 203-271).  Semantics reproduced
 * exactly: returns immediately when istate < 1 or is_done(t,tout,dt) (strict '>' test,
 * tvdode.f90:282); t is advanced by repeated t = t + dt; itask = 1 integrate past tout,
 * itask = 2 one single step (ignored by mstvd, which has no itask).  u is host memory. */
int hrweno_ode_integrate(hrweno_ode *ode, double *u, double *t, double tout, double dt, int itask);
/* device-resident u; asynchronous w.r.t. the host except for the scalar bookkeeping */
int hrweno_ode_integrate_dev(hrweno_ode *ode, double *u_dev, double *t, double tout, double dt, int itask,
                             void *stream);
/* Device callers that keep the state inside the integrator between output times (fused integrators only).  The
 * reference's `u` is inout on every call (tvdode.f90:97,203); a caller that does not touch u between calls can hand it over
 * once: attach copies the dense device vector into the library's padded state, integrate_attached advances that state
 * with the semantics of integrate (no dense <-> padded copies, which are 2 x 16 B per cell and call), fetch writes the
 * current state to a dense device vector.  A later hrweno_ode_integrate[_dev] with an explicit u replaces the state. */
int hrweno_ode_attach(hrweno_ode *ode, const double *u_dev, void *stream);
int hrweno_ode_integrate_attached(hrweno_ode *ode, double *t, double tout, double dt, int itask, void *stream);
int hrweno_ode_fetch(hrweno_ode *ode, double *u_dev, void *stream);
/* public fields of type(tvdode) (tvdode.f90:18-29) */
int64_t hrweno_ode_fevals(const hrweno_ode *ode); /* keeps the reference's count, incl. mstvd's 12 for the start-up */
int hrweno_ode_istate(const hrweno_ode *ode);
int hrweno_ode_order(const hrweno_ode *ode);
int64_t hrweno_ode_neq(const hrweno_ode *ode);
/* number of kernels launched by this object so far (bench bookkeeping) */
int64_t hrweno_ode_launches(const hrweno_ode *ode);

/* ---- REAL32 (src/hrweno_kinds.F90:9-17) -------------------------------------------------------------------------
 * `rk` is a compile-time switch of the whole reference: built with -DREAL32 every real of the path (state, widths, tables,
 * eps, t, dt) is single precision.  These entry points are that build of the path: same semantics as their fp64
 * counterparts above with float in every signature.  One GPU; the reference's operation order in separately rounded float
 * operations (the general kernels instantiated for float), bit-identical to the REAL32 build of the oracle; the
 * integrators work in place on the caller's dense vector.  (REAL128 has no GPU type.) */
typedef struct hrweno_fv_desc_f32 {
   int32_t abi_version, ndim;
   int64_t n[2];
   int64_t rows;
   int32_t k, flux_model, flux_scheme, bc, grid_kind, mode; /* mode is ignored: always the reference order */
   float eps;
   float flux_coef[2];
   float alpha;
   float xmin, xmax;
   const float *width[2];
   int32_t rank, nranks; /* must describe one GPU (nranks <= 1) */
   int64_t global_n, global_offset;
} hrweno_fv_desc_f32;
typedef struct hrweno_weno_f32 hrweno_weno_f32;
typedef struct hrweno_fv_f32 hrweno_fv_f32;
typedef struct hrweno_ode_f32 hrweno_ode_f32;
typedef float (*hrweno_time_fn_f32)(void *ctx, float t);
int hrweno_weno_f32_create(hrweno_weno_f32 **out, int64_t ncells, int k, float eps, const float *xedges); /* weno.f90:54-127 */
void hrweno_weno_f32_destroy(hrweno_weno_f32 *w);
int hrweno_weno_f32_reconstruct(const hrweno_weno_f32 *w, const float *v, float *vl, float *vr);          /* weno.f90:129-219, host */
void hrweno_weno_f32_reconstruct_s(const hrweno_weno_f32 *w, const float *v, float *vl, float *vr, int *status); /* subroutine form (a `pure` Fortran binding), like hrweno_weno_reconstruct_s */
int hrweno_weno_f32_get_cnu(const hrweno_weno_f32 *w, float *cnu_host); /* weno%cnu (weno.f90:41): cnu(j,r,i) at j + k*((r+1) + (k+1)*(i-1)) */
int hrweno_weno_f32_reconstruct_dev(const hrweno_weno_f32 *w, int64_t rows, const float *v_dev, int64_t ldv, int64_t incv,
                                    float *vl_dev, float *vr_dev, int64_t ldo, void *stream);
int hrweno_fv_f32_create(hrweno_fv_f32 **out, const hrweno_fv_desc_f32 *desc);
void hrweno_fv_f32_destroy(hrweno_fv_f32 *fv);
int64_t hrweno_fv_f32_neq(const hrweno_fv_f32 *fv);
int hrweno_fv_f32_rhs(hrweno_fv_f32 *fv, float t, const float *v, float *vdot); /* example1:72-109 / example2:73-129, host */
int hrweno_fv_f32_rhs_dev(hrweno_fv_f32 *fv, float t, const float *v_dev, float *vdot_dev, void *stream);
int hrweno_fv_f32_set_xedges(hrweno_fv_f32 *fv, int axis, const float *xedges);
int hrweno_fv_f32_set_flux_coef(hrweno_fv_f32 *fv, int axis, const float *face_coef, const float *cross_coef);
int hrweno_fv_f32_set_flux_time_fn(hrweno_fv_f32 *fv, hrweno_time_fn_f32 g, void *ctx);
int hrweno_rktvd_f32_create_fused(hrweno_ode_f32 **out, hrweno_fv_f32 *fv, int order); /* tvdode.f90:69-95 */
int hrweno_mstvd_f32_create_fused(hrweno_ode_f32 **out, hrweno_fv_f32 *fv);            /* tvdode.f90:180-201 */
void hrweno_ode_f32_destroy(hrweno_ode_f32 *ode);
int hrweno_ode_f32_integrate(hrweno_ode_f32 *ode, float *u, float *t, float tout, float dt, int itask); /* tvdode.f90:97-178, 203-271 */
int hrweno_ode_f32_integrate_dev(hrweno_ode_f32 *ode, float *u_dev, float *t, float tout, float dt, int itask, void *stream);
int64_t hrweno_ode_f32_fevals(const hrweno_ode_f32 *ode);
int hrweno_ode_f32_istate(const hrweno_ode_f32 *ode);
int64_t hrweno_ode_f32_launches(const hrweno_ode_f32 *ode);

/* ---- pinned host memory helpers (so host-buffer calls reach PCIe speed) ---- */
int hrweno_host_alloc(void **out, int64_t bytes);
void hrweno_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* HRWENO_B200_H */
