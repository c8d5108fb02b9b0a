!> ISO_C_BINDING shim: the reference's module / type / procedure names on top of libhrweno_b200.so.
!!
!! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran compiler (see DESIGN.md).  What the CI does
!! instead: tools/f90exec (f90py.py + f90c.py) EXECUTES this file statement by statement, with every `bind(c)` interface
!! body below marshalled from its own dummy declarations (F2018 18.3.6) into a real call of the shared library, under the
!! reference's unmodified example programs and test suites (tests/test_fortran_shim_exec.py) -- and, on the GPU box, against
!! libhrweno_b200.so itself (tests/test_zzzz_gpu_fortran_shim_exec.py; profiles/r2af_fortran_shim_on_b200.txt).  That is an interpreter, not a compiler: syntax a
!! compiler would reject and the interpreter accepts remains possible.
!! It is the binding a maintainer adds to HR-WENO so that example1/example2 re-link against the B200 library:
!!     gfortran -cpp -DREAL64 -c hrweno_b200_shim.f90 && gfortran example1.f90 hrweno_b200_shim.o -lhrweno_b200
!! (preprocessed like the reference's hrweno_kinds.F90: -DREAL32 selects the hrweno_*_f32 entry points, see `crk` below)
!! `hrweno_kinds` and `hrweno_grids` are used unchanged from the reference (grids, set-up and I/O stay
!! on the host); this file replaces hrweno_weno, hrweno_fluxes (re-exported unchanged: pointwise host
!! helpers) and hrweno_tvdode.
module hrweno_b200_c
   use, intrinsic :: iso_c_binding
   implicit none
   public

   integer(c_int), parameter :: HRWENO_ABI_VERSION = 1
   integer(c_int), parameter :: FLUX_BURGERS = 0, FLUX_LINEAR = 1
   integer(c_int), parameter :: SCHEME_GODUNOV = 0, SCHEME_LAX_FRIEDRICHS = 1
   integer(c_int), parameter :: BC_COPY_NEIGHBOUR = 0, BC_ZERO_FLUX = 1
   integer(c_int), parameter :: GRID_WIDTH_ARRAY = 0, GRID_LINEAR = 1
   integer(c_int), parameter :: MODE_STRICT = 0, MODE_FAST = 1

   ! The kind of the reference is a compile-time switch of the whole library (src/hrweno_kinds.F90:9-17: -DREAL32 / -DREAL64).
   ! Compile this file with the same switch: `crk` is the C kind of `rk`, and the bind(c) names below select the fp64 entry
   ! points or their REAL32 counterparts (hrweno_*_f32: float in every position, include/hrweno_b200.h).  The REAL32 build
   ! has the fused operators and integrators and type(weno); the host-integrand constructors rktvd(fu, ...) / mstvd(fu, ...),
   ! the device-integrand ones and hrweno_mgpu_* exist for REAL64 only.
#ifdef REAL32
   integer, parameter :: crk = c_float
#define N_WENO_CREATE "hrweno_weno_f32_create"
#define N_WENO_DESTROY "hrweno_weno_f32_destroy"
#define N_WENO_GET_CNU "hrweno_weno_f32_get_cnu"
#define N_WENO_RECONSTRUCT "hrweno_weno_f32_reconstruct"
#define N_WENO_RECONSTRUCT_S "hrweno_weno_f32_reconstruct_s"
#define N_FV_CREATE "hrweno_fv_f32_create"
#define N_FV_DESTROY "hrweno_fv_f32_destroy"
#define N_FV_RHS "hrweno_fv_f32_rhs"
#define N_FV_RHS_DEV "hrweno_fv_f32_rhs_dev"
#define N_FV_SET_XEDGES "hrweno_fv_f32_set_xedges"
#define N_FV_SET_FLUX_COEF "hrweno_fv_f32_set_flux_coef"
#define N_FV_SET_FLUX_TIME_FN "hrweno_fv_f32_set_flux_time_fn"
#define N_RKTVD_CREATE_FUSED "hrweno_rktvd_f32_create_fused"
#define N_MSTVD_CREATE_FUSED "hrweno_mstvd_f32_create_fused"
#define N_ODE_DESTROY "hrweno_ode_f32_destroy"
#define N_ODE_INTEGRATE "hrweno_ode_f32_integrate"
#define N_ODE_FEVALS "hrweno_ode_f32_fevals"
#define N_ODE_ISTATE "hrweno_ode_f32_istate"
#else
   integer, parameter :: crk = c_double
#define N_WENO_CREATE "hrweno_weno_create"
#define N_WENO_DESTROY "hrweno_weno_destroy"
#define N_WENO_GET_CNU "hrweno_weno_get_cnu"
#define N_WENO_RECONSTRUCT "hrweno_weno_reconstruct"
#define N_WENO_RECONSTRUCT_S "hrweno_weno_reconstruct_s"
#define N_FV_CREATE "hrweno_fv_create"
#define N_FV_DESTROY "hrweno_fv_destroy"
#define N_FV_RHS "hrweno_fv_rhs"
#define N_FV_RHS_DEV "hrweno_fv_rhs_dev"
#define N_FV_SET_XEDGES "hrweno_fv_set_xedges"
#define N_FV_SET_FLUX_COEF "hrweno_fv_set_flux_coef"
#define N_FV_SET_FLUX_TIME_FN "hrweno_fv_set_flux_time_fn"
#define N_RKTVD_CREATE_FUSED "hrweno_rktvd_create_fused"
#define N_MSTVD_CREATE_FUSED "hrweno_mstvd_create_fused"
#define N_ODE_DESTROY "hrweno_ode_destroy"
#define N_ODE_INTEGRATE "hrweno_ode_integrate"
#define N_ODE_FEVALS "hrweno_ode_fevals"
#define N_ODE_ISTATE "hrweno_ode_istate"
#endif

   type, bind(c) :: hrweno_fv_desc
      integer(c_int32_t) :: abi_version = HRWENO_ABI_VERSION
      integer(c_int32_t) :: ndim = 1
      integer(c_int64_t) :: n(2) = [1, 1]
      integer(c_int64_t) :: rows = 1
      integer(c_int32_t) :: k = 3
      integer(c_int32_t) :: flux_model = FLUX_BURGERS
      integer(c_int32_t) :: flux_scheme = SCHEME_GODUNOV
      integer(c_int32_t) :: bc = BC_COPY_NEIGHBOUR
      integer(c_int32_t) :: grid_kind = GRID_WIDTH_ARRAY
      integer(c_int32_t) :: mode = MODE_STRICT
      real(crk) :: eps = 1e-6_crk
      real(crk) :: flux_coef(2) = [1.0_crk, 1.0_crk]
      real(crk) :: alpha = 1.0_crk
      real(crk) :: xmin = 0.0_crk, xmax = 1.0_crk
      type(c_ptr) :: width(2) = c_null_ptr
      integer(c_int32_t) :: rank = 0, nranks = 1
      integer(c_int64_t) :: global_n = 0, global_offset = 0
   end type

   interface
      function hrweno_last_error() bind(c, name="hrweno_last_error") result(p)
         import :: c_ptr
         type(c_ptr) :: p
      end function
      function hrweno_weno_create(out, ncells, k, eps, xedges) bind(c, name=N_WENO_CREATE) result(st)
         import :: c_ptr, c_int, c_int64_t, crk
         type(c_ptr), intent(out) :: out
         integer(c_int64_t), value :: ncells
         integer(c_int), value :: k
         real(crk), value :: eps
         type(c_ptr), value :: xedges
         integer(c_int) :: st
      end function
      subroutine hrweno_weno_destroy(w) bind(c, name=N_WENO_DESTROY)
         import :: c_ptr
         type(c_ptr), value :: w
      end subroutine
      function hrweno_weno_get_cnu(w, cnu) bind(c, name=N_WENO_GET_CNU) result(st)
         import :: c_ptr, c_int, crk
         type(c_ptr), value :: w
         real(crk), intent(out) :: cnu(*)
         integer(c_int) :: st
      end function
      function hrweno_weno_reconstruct(w, v, vl, vr) bind(c, name=N_WENO_RECONSTRUCT) result(st)
         import :: c_ptr, c_int, crk
         type(c_ptr), value :: w
         real(crk), intent(in) :: v(*)
         real(crk), intent(out) :: vl(*), vr(*)
         integer(c_int) :: st
      end function
      ! the same call in subroutine form: a pure FUNCTION may not have intent(out) dummies, a pure subroutine may, and
      ! `reconstruct` has to stay pure for the reference's `pure subroutine rhs` (example1:72,93) to compile
      pure subroutine hrweno_weno_reconstruct_s(w, v, vl, vr, st) bind(c, name=N_WENO_RECONSTRUCT_S)
         import :: c_ptr, c_int, crk
         type(c_ptr), value :: w
         real(crk), intent(in) :: v(*)
         real(crk), intent(out) :: vl(*), vr(*)
         integer(c_int), intent(out) :: st
      end subroutine
      function hrweno_fv_create(out, desc) bind(c, name=N_FV_CREATE) result(st)
         import :: c_ptr, c_int, hrweno_fv_desc
         type(c_ptr), intent(out) :: out
         type(hrweno_fv_desc), intent(in) :: desc
         integer(c_int) :: st
      end function
      subroutine hrweno_fv_destroy(fv) bind(c, name=N_FV_DESTROY)
         import :: c_ptr
         type(c_ptr), value :: fv
      end subroutine
      function hrweno_fv_rhs(fv, t, v, vdot) bind(c, name=N_FV_RHS) result(st)
         import :: c_ptr, c_int, crk
         type(c_ptr), value :: fv
         real(crk), value :: t
         real(crk), intent(in) :: v(*)
         real(crk), intent(out) :: vdot(*)
         integer(c_int) :: st
      end function
      ! the same on device pointers, asynchronous on `stream` (what a device integrand of rktvd_dev / mstvd_dev calls)
      function hrweno_fv_rhs_dev(fv, t, v_dev, vdot_dev, stream) bind(c, name=N_FV_RHS_DEV) result(st)
         import :: c_ptr, c_int, crk
         type(c_ptr), value :: fv
         real(crk), value :: t
         type(c_ptr), value :: v_dev, vdot_dev, stream
         integer(c_int) :: st
      end function
      ! weno(ncells,k,eps,xedges) inside the fused operator: per-cell tables for the sweep along `axis` (0-based)
      function hrweno_fv_set_xedges(fv, axis, xedges) bind(c, name=N_FV_SET_XEDGES) result(st)
         import :: c_ptr, c_int, crk
         type(c_ptr), value :: fv
         integer(c_int), value :: axis
         real(crk), intent(in) :: xedges(*)
         integer(c_int) :: st
      end function
      ! x-dependent flux f = (model(v)*cross(c))*face(f) along `axis`; pass c_null_ptr for an absent factor
      function hrweno_fv_set_flux_coef(fv, axis, face_coef, cross_coef) bind(c, name=N_FV_SET_FLUX_COEF) result(st)
         import :: c_ptr, c_int
         type(c_ptr), value :: fv
         integer(c_int), value :: axis
         type(c_ptr), value :: face_coef, cross_coef
         integer(c_int) :: st
      end function
      ! t-dependent flux f = ((model(v)*cross)*face)*g(t); g is a bind(C) function `real(c_double) function g(ctx, t)` passed
      ! with c_funloc (c_null_funptr removes the factor)
      function hrweno_fv_set_flux_time_fn(fv, g, ctx) bind(c, name=N_FV_SET_FLUX_TIME_FN) result(st)
         import :: c_ptr, c_funptr, c_int
         type(c_ptr), value :: fv
         type(c_funptr), value :: g
         type(c_ptr), value :: ctx
         integer(c_int) :: st
      end function
#ifndef REAL32
      ! ---- one process, N GPUs: the GLOBAL descriptor, slabs / halos / alpha reduction inside the library ----
      function hrweno_mgpu_create(out, desc, ngpus, devices) bind(c, name="hrweno_mgpu_create") result(st)
         import :: c_ptr, c_int, hrweno_fv_desc
         type(c_ptr), intent(out) :: out
         type(hrweno_fv_desc), intent(in) :: desc
         integer(c_int), value :: ngpus          ! <= 0: all visible devices
         type(c_ptr), value :: devices           ! c_null_ptr: devices 0 .. ngpus-1
         integer(c_int) :: st
      end function
      subroutine hrweno_mgpu_destroy(m) bind(c, name="hrweno_mgpu_destroy")
         import :: c_ptr
         type(c_ptr), value :: m
      end subroutine
      function hrweno_mgpu_ngpus(m) bind(c, name="hrweno_mgpu_ngpus") result(n)
         import :: c_ptr, c_int
         type(c_ptr), value :: m
         integer(c_int) :: n
      end function
      function hrweno_mgpu_rktvd(m, order) bind(c, name="hrweno_mgpu_rktvd") result(st)
         import :: c_ptr, c_int
         type(c_ptr), value :: m
         integer(c_int), value :: order
         integer(c_int) :: st
      end function
      function hrweno_mgpu_mstvd(m) bind(c, name="hrweno_mgpu_mstvd") result(st)
         import :: c_ptr, c_int
         type(c_ptr), value :: m
         integer(c_int) :: st
      end function
      function hrweno_mgpu_integrate(m, u, t, tout, dt, itask) bind(c, name="hrweno_mgpu_integrate") result(st)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: m
         real(c_double), intent(inout) :: u(*)   ! the global vector, storage order of example2:50
         real(c_double), intent(inout) :: t
         real(c_double), value :: tout, dt
         integer(c_int), value :: itask
         integer(c_int) :: st
      end function
      function hrweno_mgpu_upload(m, u) bind(c, name="hrweno_mgpu_upload") result(st)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: m
         real(c_double), intent(in) :: u(*)
         integer(c_int) :: st
      end function
      function hrweno_mgpu_integrate_resident(m, t, tout, dt, itask) bind(c, name="hrweno_mgpu_integrate_resident") result(st)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: m
         real(c_double), intent(inout) :: t
         real(c_double), value :: tout, dt
         integer(c_int), value :: itask
         integer(c_int) :: st
      end function
      function hrweno_mgpu_download(m, u) bind(c, name="hrweno_mgpu_download") result(st)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: m
         real(c_double), intent(inout) :: u(*)
         integer(c_int) :: st
      end function
      function hrweno_mgpu_max_wavespeed(m, alpha, install) bind(c, name="hrweno_mgpu_max_wavespeed") result(st)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: m
         real(c_double), intent(out) :: alpha
         integer(c_int), value :: install
         integer(c_int) :: st
      end function
      function hrweno_mgpu_set_xedges(m, axis, xedges) bind(c, name="hrweno_mgpu_set_xedges") result(st)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: m
         integer(c_int), value :: axis
         real(c_double), intent(in) :: xedges(*) ! global edges(0:n(axis))
         integer(c_int) :: st
      end function
      function hrweno_mgpu_set_flux_coef(m, axis, face_coef, cross_coef) bind(c, name="hrweno_mgpu_set_flux_coef") result(st)
         import :: c_ptr, c_int
         type(c_ptr), value :: m
         integer(c_int), value :: axis
         type(c_ptr), value :: face_coef, cross_coef
         integer(c_int) :: st
      end function
      function hrweno_mgpu_set_flux_time_fn(m, g, ctx) bind(c, name="hrweno_mgpu_set_flux_time_fn") result(st)
         import :: c_ptr, c_funptr, c_int
         type(c_ptr), value :: m
         type(c_funptr), value :: g
         type(c_ptr), value :: ctx
         integer(c_int) :: st
      end function
      function hrweno_mgpu_set_alpha(m, alpha) bind(c, name="hrweno_mgpu_set_alpha") result(st)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: m
         real(c_double), value :: alpha
         integer(c_int) :: st
      end function
      ! slab `rank` (0-based) holds unknowns [offset, offset + count) of the global vector on CUDA device `device`
      function hrweno_mgpu_slab(m, rank, device, offset, count) bind(c, name="hrweno_mgpu_slab") result(st)
         import :: c_ptr, c_int, c_int64_t
         type(c_ptr), value :: m
         integer(c_int), value :: rank
         integer(c_int), intent(out) :: device
         integer(c_int64_t), intent(out) :: offset, count
         integer(c_int) :: st
      end function
      function hrweno_mgpu_fevals(m) bind(c, name="hrweno_mgpu_fevals") result(n)
         import :: c_ptr, c_int64_t
         type(c_ptr), value :: m
         integer(c_int64_t) :: n
      end function
#endif
      function hrweno_rktvd_create_fused(out, fv, order) bind(c, name=N_RKTVD_CREATE_FUSED) result(st)
         import :: c_ptr, c_int
         type(c_ptr), intent(out) :: out
         type(c_ptr), value :: fv
         integer(c_int), value :: order
         integer(c_int) :: st
      end function
      function hrweno_mstvd_create_fused(out, fv) bind(c, name=N_MSTVD_CREATE_FUSED) result(st)
         import :: c_ptr, c_int
         type(c_ptr), intent(out) :: out
         type(c_ptr), value :: fv
         integer(c_int) :: st
      end function
#ifndef REAL32
      function hrweno_rktvd_create(out, fu, ctx, neq, order) bind(c, name="hrweno_rktvd_create") result(st)
         import :: c_ptr, c_funptr, c_int, c_int64_t
         type(c_ptr), intent(out) :: out
         type(c_funptr), value :: fu
         type(c_ptr), value :: ctx
         integer(c_int64_t), value :: neq
         integer(c_int), value :: order
         integer(c_int) :: st
      end function
      function hrweno_rktvd_create_host(out, fu, ctx, neq, order) bind(c, name="hrweno_rktvd_create_host") result(st)
         import :: c_ptr, c_funptr, c_int, c_int64_t
         type(c_ptr), intent(out) :: out
         type(c_funptr), value :: fu
         type(c_ptr), value :: ctx
         integer(c_int64_t), value :: neq
         integer(c_int), value :: order
         integer(c_int) :: st
      end function
      function hrweno_mstvd_create_host(out, fu, ctx, neq) bind(c, name="hrweno_mstvd_create_host") result(st)
         import :: c_ptr, c_funptr, c_int, c_int64_t
         type(c_ptr), intent(out) :: out
         type(c_funptr), value :: fu
         type(c_ptr), value :: ctx
         integer(c_int64_t), value :: neq
         integer(c_int) :: st
      end function
      function hrweno_mstvd_create(out, fu, ctx, neq) bind(c, name="hrweno_mstvd_create") result(st)
         import :: c_ptr, c_funptr, c_int, c_int64_t
         type(c_ptr), intent(out) :: out
         type(c_funptr), value :: fu
         type(c_ptr), value :: ctx
         integer(c_int64_t), value :: neq
         integer(c_int) :: st
      end function
#endif
      subroutine hrweno_ode_destroy(ode) bind(c, name=N_ODE_DESTROY)
         import :: c_ptr
         type(c_ptr), value :: ode
      end subroutine
      function hrweno_ode_integrate(ode, u, t, tout, dt, itask) bind(c, name=N_ODE_INTEGRATE) result(st)
         import :: c_ptr, c_int, crk
         type(c_ptr), value :: ode
         real(crk), intent(inout) :: u(*)
         real(crk), intent(inout) :: t
         real(crk), value :: tout, dt
         integer(c_int), value :: itask
         integer(c_int) :: st
      end function
      function hrweno_ode_fevals(ode) bind(c, name=N_ODE_FEVALS) result(n)
         import :: c_ptr, c_int64_t
         type(c_ptr), value :: ode
         integer(c_int64_t) :: n
      end function
      function hrweno_ode_istate(ode) bind(c, name=N_ODE_ISTATE) result(n)
         import :: c_ptr, c_int
         type(c_ptr), value :: ode
         integer(c_int) :: n
      end function
   end interface

contains

   function last_error_string() result(msg)
      character(:), allocatable :: msg
      character(kind=c_char), pointer :: p(:)
      integer :: n
      call c_f_pointer(hrweno_last_error(), p, [4096])
      n = 0
      do while (p(n + 1) /= c_null_char)
         n = n + 1
      end do
      allocate (character(n) :: msg)
      msg = transfer(p(1:n), msg)
   end function

end module hrweno_b200_c

module hrweno_weno
!! Drop-in for src/hrweno_weno.f90: same public names (weno, c1, c2, c3), same constructor and
!! reconstruct signatures (weno.f90:44,48-50,54,129).  `reconstruct` stays `pure` (the C entry point is declared pure
!! in its interface body: it defines nothing but vl, vr and its status); `error stop` in a pure procedure is F2018.
   use, intrinsic :: iso_c_binding
   use hrweno_kinds, only: rk
   use hrweno_b200_c
   implicit none
   private
   public :: weno, c1, c2, c3

   real(rk), parameter :: &
      c1(0:0, -1:0) = reshape([1.0_rk, 1.0_rk], [1, 2]), &
      c2(0:1, -1:1) = reshape([3.0_rk/2, -1.0_rk/2, 1.0_rk/2, 1.0_rk/2, -1.0_rk/2, 3.0_rk/2], [2, 3]), &
      c3(0:2, -1:2) = reshape([11.0_rk/6, -7.0_rk/6, 1.0_rk/3, 1.0_rk/3, 5.0_rk/6, -1.0_rk/6, &
                               -1.0_rk/6, 5.0_rk/6, 1.0_rk/3, 1.0_rk/3, -7.0_rk/6, 11.0_rk/6], [3, 4])

   type :: weno
      character(:), allocatable :: msg
      integer :: ierr = 0
      integer :: ncells
      integer :: k = 3
      real(rk) :: eps = 1e-6_rk
      real(rk), allocatable :: cnu(:, :, :)
      type(c_ptr), private :: handle = c_null_ptr  ! owned by the C library; released by destroy()
   contains
      procedure, pass(self) :: reconstruct => weno_reconstruct
      procedure, pass(self) :: destroy => weno_destroy  ! explicit: a finaliser would free the handle of the
                                                        ! function-result temporary in `w = weno(...)` (example1:44)
   end type weno

   interface weno
      module procedure :: weno_init
   end interface weno

contains

   type(weno) function weno_init(ncells, k, eps, xedges) result(self)
      integer, intent(in) :: ncells
      integer, intent(in), optional :: k
      real(rk), intent(in), optional :: eps
      real(rk), intent(in), optional, contiguous, target :: xedges(0:)   ! contiguous: its address goes to C
      integer(c_int) :: st
      type(c_ptr) :: xe
      if (rk /= crk) error stop "hrweno_b200_shim: compile the shim with the same -DREAL32 / -DREAL64 as hrweno_kinds"
      self%ncells = ncells
      if (present(k)) self%k = k
      if (present(eps)) self%eps = eps
      xe = c_null_ptr
      if (present(xedges)) then
         if (size(xedges) /= ncells + 1) then
            self%msg = "Invalid input 'xedges': size(xedges) /= ncells + 1."
            self%ierr = 1
            error stop self%msg
         end if
         xe = c_loc(xedges)
      end if
      st = hrweno_weno_create(self%handle, int(ncells, c_int64_t), int(self%k, c_int), real(self%eps, crk), xe)
      if (st /= 0) then
         self%msg = last_error_string()   ! same texts as weno.f90:75,84,94
         self%ierr = 1
         error stop self%msg
      end if
      if (present(xedges)) then
         allocate (self%cnu(0:self%k - 1, -1:self%k - 1, 1:ncells))
         st = hrweno_weno_get_cnu(self%handle, self%cnu)
      end if
   end function weno_init

   pure subroutine weno_reconstruct(self, v, vl, vr)
      class(weno), intent(in) :: self
      real(rk), intent(in), contiguous :: v(:)   ! strided actuals (example2:107) are copied in by the compiler
      real(rk), intent(out), contiguous :: vl(:), vr(:)
      integer(c_int) :: st
      call hrweno_weno_reconstruct_s(self%handle, v, vl, vr, st)
      if (st /= 0) error stop "hrweno_weno_reconstruct failed (hrweno_last_error() has the text)"
   end subroutine weno_reconstruct

   subroutine weno_destroy(self)
      class(weno), intent(inout) :: self
      call hrweno_weno_destroy(self%handle)
      self%handle = c_null_ptr
   end subroutine weno_destroy

end module hrweno_weno

module hrweno_tvdode
!! Drop-in for src/hrweno_tvdode.f90 (rktvd, mstvd; tvdode.f90:36-48,59-65).  Constructors:
!!  * the reference's own `rktvd(fu, neq, order)` / `mstvd(fu, neq)` with the reference's HOST integrand
!!    `fu(t, u(:), udot(:))` (tvdode.f90:50-57): an unmodified program (its rhs calling w%reconstruct and godunov)
!!    links and runs; the state and the stage combinations live on the GPU, u/udot are staged around every call;
!!  * the fused `rktvd(fv, neq, order)` / `mstvd(fv, neq)` where `fv` is a hrweno_fv handle describing the example rhs
!!    (one kernel per stage, the fast path);
!!  * `rktvd_dev(fu_dev, neq, order)` / `mstvd_dev(fu_dev, neq)` for an integrand that works on device pointers.
   use, intrinsic :: iso_c_binding
   use hrweno_kinds, only: rk
   use hrweno_b200_c
   implicit none
   private
   public :: rktvd, mstvd, integrand, integrand_dev
#ifndef REAL32
   public :: rktvd_dev, mstvd_dev
#endif

   abstract interface
      subroutine integrand(t, u, udot)
         !! the reference's interface, unchanged (tvdode.f90:50-57)
         import :: rk
         real(rk), intent(in) :: t, u(:)
         real(rk), intent(out) :: udot(:)
      end subroutine
   end interface

   type :: integrand_holder
      !! keeps the user's procedure alive behind the C callback's context pointer
      procedure(integrand), pointer, nopass :: fu => null()
   end type

   abstract interface
      subroutine integrand_dev(ctx, t, neq, u_dev, udot_dev, stream) bind(c)
         !! device-resident integrand: u_dev/udot_dev are CUDA device pointers (hrweno_rhs_fn)
         import :: c_ptr, c_double, c_int64_t
         type(c_ptr), value :: ctx
         real(c_double), value :: t
         integer(c_int64_t), value :: neq
         type(c_ptr), value :: u_dev, udot_dev, stream
      end subroutine
   end interface

   type, abstract :: tvdode
      integer :: neq
      integer :: order
      integer :: fevals = 0   ! tvdode.f90:22-23: public components; mirrored from the C object after every call,
      integer :: istate = 0   ! so `ode%fevals` / `ode%istate` of the reference's programs and tests compile unchanged
      character(:), allocatable :: msg
      type(c_ptr) :: handle = c_null_ptr
      type(integrand_holder), pointer :: holder => null()
   contains
      procedure, pass(self) :: integrate => tvdode_integrate
      procedure, pass(self) :: destroy => tvdode_destroy
   end type tvdode

   type, extends(tvdode) :: rktvd
   end type
   type, extends(tvdode) :: mstvd
   end type

#ifdef REAL32
   ! the REAL32 entry points have the fused constructors only (no host-integrand / device-integrand integrators in float)
   interface rktvd
      module procedure :: rktvd_init_fused
   end interface
   interface mstvd
      module procedure :: mstvd_init_fused
   end interface
#else
   interface rktvd
      module procedure :: rktvd_init, rktvd_init_fused
   end interface
   interface mstvd
      module procedure :: mstvd_init, mstvd_init_fused
   end interface
#endif

contains

#ifndef REAL32
   subroutine host_trampoline(ctx, t, neq, u, udot) bind(c)
      !! hrweno_rhs_host_fn: hands the library's pinned staging arrays to the user's assumed-shape integrand
      type(c_ptr), value :: ctx
      real(c_double), value :: t
      integer(c_int64_t), value :: neq
      real(c_double), intent(in) :: u(*)
      real(c_double), intent(out) :: udot(*)
      type(integrand_holder), pointer :: h
      call c_f_pointer(ctx, h)
      call h%fu(real(t, rk), u(1:neq), udot(1:neq))
   end subroutine

   type(rktvd) function rktvd_init(fu, neq, order) result(self)
      !! tvdode.f90:69-95, same argument list
      procedure(integrand) :: fu
      integer, intent(in) :: neq, order
      integer(c_int) :: st
      self%neq = neq; self%order = order
      allocate (self%holder)
      self%holder%fu => fu
      st = hrweno_rktvd_create_host(self%handle, c_funloc(host_trampoline), c_loc(self%holder), int(neq, c_int64_t), int(order, c_int))
      call created(self, st)
   end function

   type(rktvd) function rktvd_dev(fu, neq, order) result(self)
      procedure(integrand_dev) :: fu
      integer, intent(in) :: neq, order
      integer(c_int) :: st
      self%neq = neq; self%order = order
      st = hrweno_rktvd_create(self%handle, c_funloc(fu), c_null_ptr, int(neq, c_int64_t), int(order, c_int))
      call created(self, st)
   end function

#endif

   type(rktvd) function rktvd_init_fused(fv, neq, order) result(self)
      type(c_ptr), intent(in) :: fv
      integer, intent(in) :: neq, order
      integer(c_int) :: st
      self%neq = neq; self%order = order
      st = hrweno_rktvd_create_fused(self%handle, fv, int(order, c_int))
      call created(self, st)
   end function

#ifndef REAL32
   type(mstvd) function mstvd_init(fu, neq) result(self)
      !! tvdode.f90:180-201, same argument list
      procedure(integrand) :: fu
      integer, intent(in) :: neq
      integer(c_int) :: st
      self%neq = neq; self%order = 3
      allocate (self%holder)
      self%holder%fu => fu
      st = hrweno_mstvd_create_host(self%handle, c_funloc(host_trampoline), c_loc(self%holder), int(neq, c_int64_t))
      call created(self, st)
   end function

   type(mstvd) function mstvd_dev(fu, neq) result(self)
      procedure(integrand_dev) :: fu
      integer, intent(in) :: neq
      integer(c_int) :: st
      self%neq = neq; self%order = 3
      st = hrweno_mstvd_create(self%handle, c_funloc(fu), c_null_ptr, int(neq, c_int64_t))
      call created(self, st)
   end function

#endif

   type(mstvd) function mstvd_init_fused(fv, neq) result(self)
      type(c_ptr), intent(in) :: fv
      integer, intent(in) :: neq
      integer(c_int) :: st
      self%neq = neq; self%order = 3
      st = hrweno_mstvd_create_fused(self%handle, fv)
      call created(self, st)
   end function

   subroutine created(self, st)
      class(tvdode), intent(inout) :: self
      integer(c_int), intent(in) :: st
      if (rk /= crk) error stop "hrweno_b200_shim: compile the shim with the same -DREAL32 / -DREAL64 as hrweno_kinds"
      if (st /= 0) then
         self%msg = last_error_string()   ! tvdode.f90:83,89 texts
         self%istate = -1
         error stop self%msg
      end if
      call mirror(self)
   end subroutine

   subroutine mirror(self)
      class(tvdode), intent(inout) :: self
      self%fevals = int(hrweno_ode_fevals(self%handle))
      self%istate = int(hrweno_ode_istate(self%handle))
   end subroutine

   subroutine tvdode_integrate(self, u, t, tout, dt, itask)
      !! tvdode.f90:97 / :203 -- same argument list; mstvd ignores itask like the reference (it has none)
      class(tvdode), intent(inout) :: self
      real(rk), intent(inout), contiguous :: u(:)
      real(rk), intent(inout) :: t
      real(rk), intent(in) :: tout, dt
      integer, intent(in), optional :: itask
      integer(c_int) :: st, itask_
      itask_ = 1
      if (present(itask)) itask_ = int(itask, c_int)
      st = hrweno_ode_integrate(self%handle, u, t, tout, dt, itask_)
      call mirror(self)
      if (st /= 0) then
         self%msg = last_error_string()
         error stop self%msg
      end if
   end subroutine

   subroutine tvdode_destroy(self)
      class(tvdode), intent(inout) :: self
      call hrweno_ode_destroy(self%handle)
      self%handle = c_null_ptr
      if (associated(self%holder)) deallocate (self%holder)
   end subroutine

end module hrweno_tvdode
