!> Growth rates that depend on time as well as on size: f1 = v*x1**2*g(t), f2 = v*x1*x2*g(t) with g(t) = 1 + t/4, on the
!! FUSED general operator.  The reference hands the evaluation time to the flux function (fluxes.f90:12-18) and its
!! integrators evaluate the right-hand side at t, t + dt and t + dt/2 (tvdode.f90:162-166); on the device the separable
!! factor g(t) is a host function the library calls once per right-hand-side evaluation, at those very times, so a stage
!! stays one kernel launch.  g is a `bind(c)` function handed over with `c_funloc`.
!! Geometric x1 grid (ratio 1.01, 130 cells), uniform x2 grid (97 cells), WENO5 + Godunov + SSP-RK3, dt = 1e-4, outputs
!! after 1, 5 and 11 steps.  Self-contained: executed by tests/test_fortran_shim_exec.py and, on the GPU box, by
!! tests/test_zzzz_gpu_fortran_shim_exec.py; both compare with the oracle on this program's own arrays.
module growth_clock
   use, intrinsic :: iso_c_binding
   implicit none
contains
   function growth_factor(ctx, t) bind(c) result(g)
      type(c_ptr), value :: ctx
      real(c_double), value :: t
      real(c_double) :: g
      g = 1.0_c_double + 0.25_c_double*t
   end function growth_factor
end module growth_clock

program pbe2d_growth_time_factor
   use, intrinsic :: iso_c_binding
   use hrweno_kinds, only: rk
   use hrweno_b200_c
   use hrweno_tvdode, only: rktvd
   use growth_clock, only: growth_factor
   implicit none

   integer, parameter :: n1 = 130, n2 = 97, nout = 2
   real(rk), target :: e1(0:n1), e2(0:n2), dx1(n1), dx2(n2), c1x(n1), c2x(n2), g1face(0:n1)
   real(rk) :: q(n1*n2), q0(n1*n2), history(n1*n2, 0:nout), tgrid(0:nout)
   real(rk) :: t, step, scale1
   integer :: io, i, j, nfev
   logical :: inside
   type(hrweno_fv_desc) :: desc
   type(c_ptr) :: op
   type(rktvd) :: solver
   integer(c_int) :: st

   scale1 = (10.0_rk - 0.0_rk)/(1.01_rk**n1 - 1.0_rk)
   do i = 0, n1
      e1(i) = 0.0_rk + scale1*(1.01_rk**i - 1)
   end do
   do j = 0, n2
      e2(j) = 0.0_rk + (10.0_rk/n2)*j
   end do
   do i = 1, n1
      c1x(i) = (e1(i - 1) + e1(i))/2
      dx1(i) = e1(i) - e1(i - 1)
   end do
   do j = 1, n2
      c2x(j) = (e2(j - 1) + e2(j))/2
      dx2(j) = e2(j) - e2(j - 1)
   end do
   do i = 0, n1
      g1face(i) = e1(i)**2
   end do

   do j = 1, n2
      do i = 1, n1
         inside = c1x(i) >= 1.0_rk .and. c1x(i) <= 3.0_rk .and. c2x(j) >= 1.0_rk .and. c2x(j) <= 3.0_rk
         q((j - 1)*n1 + i) = 0.0_rk
         if (inside) q((j - 1)*n1 + i) = 1.0_rk
      end do
   end do
   q0 = q

   desc%ndim = 2
   desc%n = [int(n1, c_int64_t), int(n2, c_int64_t)]
   desc%flux_model = FLUX_LINEAR
   desc%bc = BC_ZERO_FLUX
   desc%width(1) = c_loc(dx1)
   desc%width(2) = c_loc(dx2)
   st = hrweno_fv_create(op, desc)
   if (st /= 0) error stop last_error_string()
   st = hrweno_fv_set_xedges(op, 0_c_int, e1)
   if (st /= 0) error stop last_error_string()
   st = hrweno_fv_set_flux_coef(op, 0_c_int, c_loc(g1face), c_null_ptr)
   if (st /= 0) error stop last_error_string()
   st = hrweno_fv_set_flux_coef(op, 1_c_int, c_loc(e2), c_loc(c1x))
   if (st /= 0) error stop last_error_string()
   st = hrweno_fv_set_flux_time_fn(op, c_funloc(growth_factor), c_null_ptr)
   if (st /= 0) error stop last_error_string()

   solver = rktvd(op, n1*n2, order=3)

   t = 0.0_rk
   step = 1e-4_rk
   tgrid(0) = 0.0_rk
   tgrid(1) = 4.5_rk*step
   tgrid(2) = 10.5_rk*step
   do io = 0, nout
      call solver%integrate(q, t, tgrid(io), step)
      history(:, io) = q
      tgrid(io) = t
   end do
   nfev = solver%fevals

   call solver%destroy()
   call hrweno_fv_destroy(op)
end program pbe2d_growth_time_factor
