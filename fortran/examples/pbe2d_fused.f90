!> 2D population balance (pure advection in both internal coordinates), WENO5 + Godunov + the third-order multi-step
!! integrator, through the Fortran shim with the FUSED device operator: `mstvd(op, neq)` instead of `mstvd(rhs, neq)`.
!!
!! The problem of BASELINE.json configs[1] on a 40 x 40 grid (the size of tests/golden/ref_exec_example2_40.npz): unit
!! square pulse on [1, 3]^2 inside [0, 10]^2, f1 = f2 = v, zero-flux walls, dt = 5e-3, 101 output times up to t = 5.
!! The state vector holds x1 fastest.  Self-contained (no reference module besides the kind parameter): executed on the
!! GPU box by tests/test_zzzz_gpu_fortran_shim_exec.py.
program pbe2d_fused
   use, intrinsic :: iso_c_binding
   use hrweno_kinds, only: rk
   use hrweno_b200_c
   use hrweno_tvdode, only: mstvd
   implicit none

   integer, parameter :: n1 = 40, n2 = 40, nout = 100
   real(rk), target :: dx1(n1), dx2(n2)
   real(rk) :: e1(0:n1), e2(0:n2), c1x(n1), c2x(n2), q(n1*n2), history(n1*n2, 0:nout), tgrid(0:nout)
   real(rk) :: t, t_stop, step
   integer :: io, i, j, nfev
   logical :: inside
   type(hrweno_fv_desc) :: desc
   type(c_ptr) :: op
   type(mstvd) :: solver
   integer(c_int) :: st

   do i = 0, n1
      e1(i) = 0.0_rk + (10.0_rk/n1)*i
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

   do j = 1, n2
      do i = 1, n1
         inside = c1x(i) >= 1.0_rk .and. c1x(i) <= 3.0_rk .and. c2x(j) >= 1.0_rk .and. c2x(j) <= 3.0_rk
         q((j - 1)*n1 + i) = 0.0_rk
         if (inside) q((j - 1)*n1 + i) = 1.0_rk
      end do
   end do

   desc%ndim = 2
   desc%n = [int(n1, c_int64_t), int(n2, c_int64_t)]
   desc%k = 3
   desc%eps = 1e-6_rk
   desc%flux_model = FLUX_LINEAR
   desc%flux_coef = [1.0_rk, 1.0_rk]
   desc%flux_scheme = SCHEME_GODUNOV
   desc%bc = BC_ZERO_FLUX
   desc%grid_kind = GRID_WIDTH_ARRAY
   desc%width(1) = c_loc(dx1)
   desc%width(2) = c_loc(dx2)
   st = hrweno_fv_create(op, desc)
   if (st /= 0) error stop last_error_string()

   solver = mstvd(op, n1*n2)

   t = 0.0_rk
   t_stop = 5.0_rk
   step = 5e-3_rk
   do io = 0, nout
      tgrid(io) = t_stop*io/nout
      call solver%integrate(q, t, tgrid(io), step)
      history(:, io) = q
      tgrid(io) = t
   end do
   nfev = solver%fevals

   call solver%destroy()
   call hrweno_fv_destroy(op)
end program pbe2d_fused
