!> Burgers' equation with the caller's OWN Fortran right-hand side, through the Fortran shim.
!!
!! This is the shape of a program that links against the shim without being rewritten: the integrator is built from a
!! host procedure, `rktvd(rhs, neq, order)`, exactly as with the reference's module; the library keeps the state and the
!! stage combinations on the GPU and calls `rhs` back on host arrays; inside it, `w%reconstruct` is the GPU WENO
!! reconstruction behind the reference's call (and stays `pure`).  Lax-Friedrichs faces with alpha = 1 (BASELINE.json's
!! wording of configs[0]), written out in place so that the program needs no reference module besides the kind parameter:
!! tests/test_zzzz_gpu_fortran_shim_exec.py executes it on the GPU box, where the reference tree does not exist.
program burgers_host_rhs
   use hrweno_kinds, only: rk
   use hrweno_weno, only: weno
   use hrweno_tvdode, only: rktvd
   implicit none

   integer, parameter :: ncell = 100, nout = 100
   real(rk) :: edge(0:ncell), xc(ncell), dx(ncell), q(ncell), history(ncell, 0:nout), tgrid(0:nout)
   real(rk) :: t, t_stop, step, span, slope
   integer :: io, j, nfev
   type(weno) :: w
   type(rktvd) :: solver

   span = 10.0_rk
   do j = 0, ncell
      edge(j) = -5.0_rk + (span/ncell)*j
   end do
   do j = 1, ncell
      xc(j) = (edge(j - 1) + edge(j))/2
      dx(j) = edge(j) - edge(j - 1)
   end do
   slope = -1.5_rk/6.0_rk
   do j = 1, ncell
      q(j) = 1.0_rk + slope*(xc(j) + 4.0_rk)
      q(j) = max(min(q(j), 1.0_rk), -0.5_rk)
   end do

   w = weno(ncells=ncell, k=3, eps=1e-6_rk)
   solver = rktvd(rhs, ncell, order=3)

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
   call w%destroy()

contains

   pure subroutine rhs(t, v, vdot)
      real(rk), intent(in) :: t, v(:)
      real(rk), intent(out) :: vdot(:)
      real(rk) :: face(0:ncell), below(ncell), above(ncell)
      real(rk), parameter :: alpha = 1.0_rk
      integer :: i

      call w%reconstruct(v, below, above)
      ! face i lies between cells i and i+1: left state above(i), right state below(i+1)
      do i = 1, ncell - 1
         face(i) = (half_square(above(i)) + half_square(below(i + 1)) - alpha*(below(i + 1) - above(i)))/2
      end do
      face(0) = face(1)
      face(ncell) = face(ncell - 1)
      do i = 1, ncell
         vdot(i) = -(face(i) - face(i - 1))/dx(i)
      end do
   end subroutine rhs

   pure real(rk) function half_square(v)
      real(rk), intent(in) :: v
      half_square = (v**2)/2
   end function half_square

end program burgers_host_rhs
