!> Burgers' equation on every GPU of the box from ONE Fortran process (no launcher, no MPI): the `hrweno_mgpu_*` entry
!! points take the GLOBAL problem; the library cuts it into slabs, exchanges the k halo cells of every stage over NVLink
!! peer memory and keeps the state resident between output times.  WENO5 + Godunov + SSP-RK3 on 40000 cells, CFL 0.2,
!! outputs after 0 (one step, as the reference's driver does at t = tout), 10 and 25 steps' worth of time; the state is
!! handed over once, advanced on the GPUs and fetched at the outputs.
!! Self-contained: executed by tests/test_fortran_shim_exec.py (CPU stand-in of the ABI) and by
!! tests/test_zzzz_gpu_fortran_shim_exec.py on the GPU box with all visible devices.
program burgers_multi_gpu
   use, intrinsic :: iso_c_binding
   use hrweno_kinds, only: rk
   use hrweno_b200_c
   implicit none

   integer, parameter :: ncell = 40000, nout = 2
   real(rk), target :: dx(ncell)
   real(rk) :: edge(0:ncell), xc(ncell), q(ncell), q0(ncell), history(ncell, 0:nout), tgrid(0:nout)
   real(rk) :: t, step, slope
   integer :: io, j, ngpu, nfev
   type(hrweno_fv_desc) :: desc
   type(c_ptr) :: box
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
      q(j) = 1.0_rk + slope*(xc(j) + 4.0_rk)
      q(j) = max(min(q(j), 1.0_rk), -0.5_rk)
   end do
   q0 = q

   desc%ndim = 1
   desc%n = [int(ncell, c_int64_t), 1_c_int64_t]
   desc%k = 3
   desc%flux_model = FLUX_BURGERS
   desc%flux_scheme = SCHEME_GODUNOV
   desc%bc = BC_COPY_NEIGHBOUR
   desc%grid_kind = GRID_WIDTH_ARRAY
   desc%width(1) = c_loc(dx)

   st = hrweno_mgpu_create(box, desc, 0_c_int, c_null_ptr)   ! 0: all visible devices
   if (st /= 0) error stop last_error_string()
   ngpu = hrweno_mgpu_ngpus(box)
   st = hrweno_mgpu_rktvd(box, 3_c_int)
   if (st /= 0) error stop last_error_string()

   step = 0.2_rk*10.0_rk/ncell
   tgrid(0) = 0.0_rk
   tgrid(1) = 10*step
   tgrid(2) = 25*step
   t = 0.0_rk
   st = hrweno_mgpu_upload(box, q)
   if (st /= 0) error stop last_error_string()
   do io = 0, nout
      st = hrweno_mgpu_integrate_resident(box, t, tgrid(io), step, 1_c_int)
      if (st /= 0) error stop last_error_string()
      st = hrweno_mgpu_download(box, q)
      if (st /= 0) error stop last_error_string()
      history(:, io) = q
      tgrid(io) = t
   end do
   nfev = int(hrweno_mgpu_fevals(box))

   call hrweno_mgpu_destroy(box)
end program burgers_multi_gpu
