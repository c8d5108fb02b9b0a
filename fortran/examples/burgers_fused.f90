!> Burgers' equation, WENO5 + Godunov + SSP-RK3, through the Fortran shim with the FUSED device operator.
!!
!! A caller's Fortran `rhs` cannot run on the GPU, so the finite-volume right-hand side is described by a
!! `hrweno_fv_desc` (include/hrweno_b200.h) and the integrator is built from the operator handle:
!! `rktvd(op, neq, order)` instead of `rktvd(rhs, neq, order)`.  Same problem as BASELINE.json configs[0]: 100 cells on
!! [-5, 5], clipped linear profile, dt = 1e-2, 101 output times up to t = 12.
!!
!! Self-contained on purpose (its own grid and initial profile, no reference module besides the kind parameter), because
!! tests/test_zzzz_gpu_fortran_shim_exec.py executes it on the GPU box, where the reference tree does not exist.
program burgers_fused
   use, intrinsic :: iso_c_binding
   use hrweno_kinds, only: rk
   use hrweno_b200_c
   use hrweno_tvdode, only: rktvd
   implicit none

   integer, parameter :: ncell = 100, nout = 100
   real(rk), target :: dx(ncell)
   real(rk) :: edge(0:ncell), xc(ncell), q(ncell), history(ncell, 0:nout), tgrid(0:nout)
   real(rk) :: t, t_stop, step, span, slope
   integer :: io, j, nfev
   type(hrweno_fv_desc) :: desc
   type(c_ptr) :: op
   type(rktvd) :: solver
   integer(c_int) :: st

   ! uniform edges, centres and widths
   span = 10.0_rk
   do j = 0, ncell
      edge(j) = -5.0_rk + (span/ncell)*j
   end do
   do j = 1, ncell
      xc(j) = (edge(j - 1) + edge(j))/2
      dx(j) = edge(j) - edge(j - 1)
   end do

   ! ramp from 1 at x = -4 down to -1/2 at x = 2, constant outside
   slope = -1.5_rk/6.0_rk
   do j = 1, ncell
      q(j) = 1.0_rk + slope*(xc(j) + 4.0_rk)
      q(j) = max(min(q(j), 1.0_rk), -0.5_rk)
   end do

   ! the right-hand side as a descriptor: f(v) = v**2/2, Godunov faces, outermost face fluxes copied from their neighbours
   desc%ndim = 1
   desc%n = [int(ncell, c_int64_t), 1_c_int64_t]
   desc%k = 3
   desc%eps = 1e-6_rk
   desc%flux_model = FLUX_BURGERS
   desc%flux_scheme = SCHEME_GODUNOV
   desc%bc = BC_COPY_NEIGHBOUR
   desc%grid_kind = GRID_WIDTH_ARRAY
   desc%width(1) = c_loc(dx)
   desc%mode = MODE_STRICT
   st = hrweno_fv_create(op, desc)
   if (st /= 0) error stop last_error_string()

   solver = rktvd(op, ncell, order=3)

   t = 0.0_rk
   t_stop = 12.0_rk
   step = 1e-2_rk
   do io = 0, nout
      tgrid(io) = t_stop*io/nout
      call solver%integrate(q, t, tgrid(io), step)
      history(:, io) = q
      tgrid(io) = t
   end do
   nfev = solver%fevals

   call solver%destroy()
   call hrweno_fv_destroy(op)
end program burgers_fused
