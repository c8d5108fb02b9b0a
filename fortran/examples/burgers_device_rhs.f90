!> Burgers' equation with the generic integrators and a DEVICE integrand: `rktvd_dev(fu, neq, order)` / `mstvd_dev(fu, neq)`.
!!
!! For a right-hand side that is more than one finite-volume operator -- source terms, several coupled operators -- a
!! program keeps the reference's integrators (state and stage combinations resident on the GPU, tvdode.f90:97-271) and
!! supplies the integrand as a `bind(c)` procedure that receives DEVICE pointers and a stream and enqueues its work there.
!! Here the integrand is the library's own operator on device pointers (`hrweno_fv_rhs_dev`), so the result must equal the
!! fused integrators': 200 cells, WENO5 + Godunov, SSP-RK3 and then the multi-step integrator, dt = 5e-3, outputs at
!! t = 0 (one step), 0.1 and 0.3.  Also recorded: the times the first three evaluations were asked for (t, t + dt, t + dt/2).
!! Self-contained: executed by tests/test_fortran_shim_exec.py and tests/test_zzzz_gpu_fortran_shim_exec.py.
module device_rhs
   use, intrinsic :: iso_c_binding
   use hrweno_b200_c
   implicit none
   type(c_ptr) :: op = c_null_ptr        ! the finite-volume operator the integrand applies
   integer :: ncalls = 0
   real(c_double) :: asked(3) = 0.0_c_double
contains
   subroutine apply_operator(ctx, t, neq, u_dev, udot_dev, stream) bind(c)
      type(c_ptr), value :: ctx
      real(c_double), value :: t
      integer(c_int64_t), value :: neq
      type(c_ptr), value :: u_dev, udot_dev, stream
      integer(c_int) :: st
      ncalls = ncalls + 1
      if (ncalls <= 3) asked(ncalls) = t
      st = hrweno_fv_rhs_dev(op, t, u_dev, udot_dev, stream)
      if (st /= 0) error stop "hrweno_fv_rhs_dev failed"
   end subroutine apply_operator
end module device_rhs

program burgers_device_rhs
   use, intrinsic :: iso_c_binding
   use hrweno_kinds, only: rk
   use hrweno_b200_c
   use hrweno_tvdode, only: rktvd, mstvd, rktvd_dev, mstvd_dev
   use device_rhs
   implicit none

   integer, parameter :: ncell = 200, nout = 2
   real(rk), target :: dx(ncell)
   real(rk) :: edge(0:ncell), xc(ncell), q0(ncell), q(ncell), history(ncell, 0:nout, 2), tgrid(0:nout, 2), marks(0:nout)
   real(rk) :: t, step, slope
   integer :: io, j, which, nfev(2)
   type(hrweno_fv_desc) :: desc
   type(rktvd) :: rk3
   type(mstvd) :: ms
   integer(c_int) :: st

   do j = 0, ncell
      edge(j) = -5.0_rk + (10.0_rk/ncell)*j
   end do
   do j = 1, ncell
      xc(j) = (edge(j - 1) + edge(j))/2
      dx(j) = edge(j) - edge(j - 1)
   end do
   slope = -1.5_rk/6.0_rk
   do j = 1, ncell
      q0(j) = 1.0_rk + slope*(xc(j) + 4.0_rk)
      q0(j) = max(min(q0(j), 1.0_rk), -0.5_rk)
   end do

   desc%ndim = 1
   desc%n = [int(ncell, c_int64_t), 1_c_int64_t]
   desc%width(1) = c_loc(dx)
   st = hrweno_fv_create(op, desc)
   if (st /= 0) error stop last_error_string()

   step = 5e-3_rk
   marks(0) = 0.0_rk
   marks(1) = 0.1_rk
   marks(2) = 0.3_rk
   rk3 = rktvd_dev(apply_operator, ncell, 3)
   ms = mstvd_dev(apply_operator, ncell)
   do which = 1, 2
      q = q0
      t = 0.0_rk
      do io = 0, nout
         if (which == 1) call rk3%integrate(q, t, marks(io), step)
         if (which == 2) call ms%integrate(q, t, marks(io), step)
         history(:, io, which) = q
         tgrid(io, which) = t
      end do
   end do
   nfev(1) = rk3%fevals
   nfev(2) = ms%fevals

   call rk3%destroy()
   call ms%destroy()
   call hrweno_fv_destroy(op)
end program burgers_device_rhs
