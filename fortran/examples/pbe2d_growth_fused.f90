!> 2D population balance with size-dependent growth on geometric grids, through the Fortran shim with the FUSED general
!! operator: growth rates G1 = x1**2 and G2 = x1*x2 (fluxes f1 = v*x1**2 at the x1-faces, f2 = v*x1*x2 at the x2-faces --
!! the terms example2's comments point at), WENO5 on non-uniform grids (per-cell reconstruction tables from the edges),
!! Godunov faces, zero-flux walls, the third-order multi-step integrator.  24 x 18 cells on [0, 10]^2 with ratios 1.02 and
!! 1.03, dt = 2.5e-4, 21 output times: the problem of tests/golden/ref_exec_example2_growth.npz.
!!
!! On the device the x-dependence of the flux is a per-face coefficient and a per-cross-cell coefficient,
!! f = (v*cross(c))*face(f): along x1 face = (x1 edge)**2; along x2 face = x2 edge, cross = x1 centre.
!! Self-contained: executed by tests/test_fortran_shim_exec.py and tests/test_zzzz_gpu_fortran_shim_exec.py.
program pbe2d_growth_fused
   use, intrinsic :: iso_c_binding
   use hrweno_kinds, only: rk
   use hrweno_b200_c
   use hrweno_tvdode, only: mstvd
   implicit none

   integer, parameter :: n1 = 24, n2 = 18, nout = 20
   real(rk), target :: e1(0:n1), e2(0:n2), dx1(n1), dx2(n2), c1x(n1), c2x(n2), g1face(0:n1)
   real(rk) :: q(n1*n2), history(n1*n2, 0:nout), tgrid(0:nout)
   real(rk) :: t, t_stop, step, scale1, scale2
   integer :: io, i, j, nfev
   logical :: inside
   type(hrweno_fv_desc) :: desc
   type(c_ptr) :: op
   type(mstvd) :: solver
   integer(c_int) :: st

   ! geometric edges: every cell is `ratio` times as wide as the one before it
   scale1 = (10.0_rk - 0.0_rk)/(1.02_rk**n1 - 1.0_rk)
   do i = 0, n1
      e1(i) = 0.0_rk + scale1*(1.02_rk**i - 1)
   end do
   scale2 = (10.0_rk - 0.0_rk)/(1.03_rk**n2 - 1.0_rk)
   do j = 0, n2
      e2(j) = 0.0_rk + scale2*(1.03_rk**j - 1)
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
   desc%flux_scheme = SCHEME_GODUNOV
   desc%bc = BC_ZERO_FLUX
   desc%grid_kind = GRID_WIDTH_ARRAY
   desc%width(1) = c_loc(dx1)
   desc%width(2) = c_loc(dx2)
   st = hrweno_fv_create(op, desc)
   if (st /= 0) error stop last_error_string()

   ! was: myweno(1) = weno(nc(1), k, eps, xedges=gx(1)%edges) and the same for axis 2
   st = hrweno_fv_set_xedges(op, 0_c_int, e1)
   if (st /= 0) error stop last_error_string()
   st = hrweno_fv_set_xedges(op, 1_c_int, e2)
   if (st /= 0) error stop last_error_string()
   ! f1 = v*x1**2 at the x1-faces; f2 = (v*x1)*x2 with x1 at the cell centre, x2 at the x2-faces
   do i = 0, n1
      g1face(i) = e1(i)**2
   end do
   st = hrweno_fv_set_flux_coef(op, 0_c_int, c_loc(g1face), c_null_ptr)
   if (st /= 0) error stop last_error_string()
   st = hrweno_fv_set_flux_coef(op, 1_c_int, c_loc(e2), c_loc(c1x))
   if (st /= 0) error stop last_error_string()

   solver = mstvd(op, n1*n2)

   t = 0.0_rk
   t_stop = 0.5_rk
   step = 2.5e-4_rk
   do io = 0, nout
      tgrid(io) = t_stop*io/100
      call solver%integrate(q, t, tgrid(io), step)
      history(:, io) = q
      tgrid(io) = t
   end do
   nfev = solver%fevals

   call solver%destroy()
   call hrweno_fv_destroy(op)
end program pbe2d_growth_fused
