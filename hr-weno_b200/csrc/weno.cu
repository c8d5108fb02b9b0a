// weno.cu -- type(weno): constructor, cnu set-up and the reconstruct kernel (K1).
#include <float.h>

#include <cstring>
#include <new>

#include "internal.hpp"
#include "weno_core.cuh"

namespace hrw {

Weno::~Weno() {
   cudaFree(d_cnu);
   cudaFree(d_buf);
}

// weno_calc_cnu (weno.f90:221-297, Shu eq. 2.20).  Host side, once per object ("only called a very
// small number of times at the start of the simulation", weno.f90:225-228); fp64, reference order.
void weno_calc_cnu_host(int64_t nc, int k, const double *xedges, std::vector<double> &cnu) {
   const int ng = k + 1;
   std::vector<double> buf((size_t)(nc + 2 * ng + 1));
   double *xext = buf.data() + ng;
   for (int64_t i = 0; i <= nc; ++i) xext[i] = xedges[i];
   double dx = xext[1] - xext[0];
   for (int64_t i = -1; i >= -ng; --i) xext[i] = xext[i + 1] - dx;
   dx = xext[nc] - xext[nc - 1];
   for (int64_t i = nc + 1; i <= nc + ng; ++i) xext[i] = xext[i - 1] + dx;
   auto xl = [&](int64_t m) { return xext[m - 1]; };
   auto xr = [&](int64_t m) { return xext[m]; };
   cnu.assign((size_t)nc * k * (k + 1), 0.0);
   for (int64_t i = 1; i <= nc; ++i)
      for (int r = -1; r <= k - 1; ++r)
         for (int j = 0; j <= k - 1; ++j) {
            volatile double sum2 = 0.0; // volatile: forbid contraction of the accumulations by the host compiler
            for (int m = j + 1; m <= k; ++m) {
               volatile double prod2 = 1.0;
               for (int l = 0; l <= k; ++l) {
                  if (l == m) continue;
                  const double diff = xl(i - r + m) - xl(i - r + l);
                  prod2 = prod2 * diff;
               }
               volatile double sum1 = 0.0;
               for (int l = 0; l <= k; ++l) {
                  if (l == m) continue;
                  volatile double prod1 = 1.0;
                  for (int q = 0; q <= k; ++q) {
                     if (q == m || q == l) continue;
                     const double diff = xr(i) - xl(i - r + q);
                     prod1 = prod1 * diff;
                  }
                  sum1 = sum1 + prod1;
               }
               const double quot = sum1 / prod2;
               sum2 = sum2 + quot;
            }
            const double wdt = xr(i - r + j) - xl(i - r + j);
            cnu[(size_t)(j + k * ((r + 1) + (k + 1) * (i - 1)))] = sum2 * wdt;
         }
}

int weno_create(Weno **out, int64_t ncells, int k, double eps, const double *xedges) {
   if (!out) return fail(HRWENO_EINVAL, "hrweno_weno_create: null argument");
   if (!(ncells > 0)) return fail(HRWENO_EINVAL, "Invalid input 'ncells'. Valid range: ncells > 0.");    // weno.f90:75
   if (!(k >= 1 && k <= 3)) return fail(HRWENO_EINVAL, "Invalid input 'k'. Valid range: 1 <= k <= 3."); // :84
   if (!(eps > DBL_EPSILON)) return fail(HRWENO_EINVAL, "Invalid input 'eps'. Valid range: eps > epsilon."); // :94
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
      return fail(HRWENO_ECUDA, "no CUDA device (this library has no CPU fallback)");
   Weno *w = new (std::nothrow) Weno();
   if (!w) return fail(HRWENO_ENOMEM, "out of host memory");
   w->ncells = ncells;
   w->k = k;
   w->eps = eps;
   w->uniform = xedges == nullptr;
   if (xedges) {
      weno_calc_cnu_host(ncells, k, xedges, w->cnu_host);
      const size_t bytes = w->cnu_host.size() * sizeof(double);
      cudaError_t e = cudaMalloc(&w->d_cnu, bytes);
      if (e == cudaSuccess) e = cudaMemcpy(w->d_cnu, w->cnu_host.data(), bytes, cudaMemcpyHostToDevice);
      if (e != cudaSuccess) {
         delete w;
         return cuda_fail(e, "upload cnu", __FILE__, __LINE__);
      }
   }
   *out = w;
   return HRWENO_OK;
}

// ------------------------------------------------------------------------------------------------
// K1: reconstruct `rows` rows of n cells.  Element (row, i) of v at v[row*ldv + i*incv]; ghost cells
// are edge replicas (weno.f90:171-173) realised by clamping the cell index.
// ------------------------------------------------------------------------------------------------
template <int K, class M>
__global__ void __launch_bounds__(256)
   recon_kernel(const double *__restrict__ v, int64_t ldv, int64_t incv, double *__restrict__ vl, double *__restrict__ vr,
                int64_t ldo, int64_t n, int64_t rows, const WenoK kc, const double *__restrict__ cnu, int cell_fastest) {
   const double eps = kc.eps;
   const int64_t total = rows * n;
   for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      int64_t row, i;
      if (cell_fastest) {
         row = idx / n;
         i = idx - row * n;
      } else {
         i = idx / rows;
         row = idx - i * rows;
      }
      const double *vrow = v + row * ldv;
      double w[2 * K - 1];
#pragma unroll
      for (int o = -(K - 1); o <= K - 1; ++o) {
         int64_t ii = i + o;
         ii = ii < 0 ? 0 : (ii > n - 1 ? n - 1 : ii);
         w[o + K - 1] = __ldg(vrow + ii * incv);
      }
      double l, r;
      if (cnu) {
         double ci[K * (K + 1)];
#pragma unroll
         for (int q = 0; q < K * (K + 1); ++q) ci[q] = __ldg(cnu + (size_t)i * (K * (K + 1)) + q);
         weno_cell_nonuniform<K, M>(ci, w + (K - 1), eps, l, r);
      } else {
         weno_run<K, 1, M>(w, kc, &l, &r);
      }
      vl[row * ldo + i] = l;
      vr[row * ldo + i] = r;
   }
}

// ------------------------------------------------------------------------------------------------
// K1r: uniform tables, contiguous rows (incv == 1).  A thread reconstructs a run of R = 4 consecutive cells from a
// register window, so the products and differences neighbouring cells share are computed once (weno_core.cuh);
// 16-B loads and stores when the caller's pointers and pitches allow it, clamped scalar accesses at row ends.
// ------------------------------------------------------------------------------------------------
template <int K, class M, bool VEC>
__global__ void __launch_bounds__(256)
   recon_run_kernel(const double *__restrict__ v, int64_t ldv, double *__restrict__ vl, double *__restrict__ vr, int64_t ldo, int64_t n,
                    int64_t rows, const WenoK kc) {
   constexpr int R = 4;
   const int64_t runs_per_row = (n + R - 1) / R;
   const int64_t total = rows * runs_per_row;
   for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      const int64_t row = idx / runs_per_row;
      const int64_t i0 = (idx - row * runs_per_row) * R;
      const double *vrow = v + row * ldv;
      double w[R + 4]; // cells i0-2 .. i0+R+1; only [JLO, JHI) is needed for this K (even bounds: 16-B loads)
      constexpr int JLO = (2 - (K - 1)) & ~1, JHI = (R + 2 + (K - 1) + 1) & ~1;
#pragma unroll
      for (int j = 0; j < R + 4; ++j) w[j] = 0.0;
      const bool inner = i0 >= 2 && i0 + R + 2 <= n;
      if (VEC && inner) {
#pragma unroll
         for (int j = JLO; j < JHI; j += 2) {
            const double2 t = __ldg(reinterpret_cast<const double2 *>(vrow + i0 - 2 + j));
            w[j] = t.x;
            w[j + 1] = t.y;
         }
      } else {
#pragma unroll
         for (int j = JLO; j < JHI; ++j) {
            int64_t ii = i0 - 2 + j;
            ii = ii < 0 ? 0 : (ii > n - 1 ? n - 1 : ii); // edge replicas (weno.f90:171-173)
            w[j] = __ldg(vrow + ii);
         }
      }
      double l[R], r[R];
      weno_run<K, R, M>(w + (2 - (K - 1)), kc, l, r);
      double *pl = vl + row * ldo + i0, *pr = vr + row * ldo + i0;
      if (VEC && i0 + R <= n) {
#pragma unroll
         for (int j = 0; j < R; j += 2) {
            *reinterpret_cast<double2 *>(pl + j) = make_double2(l[j], l[j + 1]);
            *reinterpret_cast<double2 *>(pr + j) = make_double2(r[j], r[j + 1]);
         }
      } else {
#pragma unroll
         for (int j = 0; j < R; ++j)
            if (i0 + j < n) {
               pl[j] = l[j];
               pr[j] = r[j];
            }
      }
   }
}

template <int K, class M>
static void launch_run(const Weno *w, int64_t rows, const double *v, int64_t ldv, double *vl, double *vr, int64_t ldo, cudaStream_t st) {
   const int64_t total = rows * ((w->ncells + 3) / 4);
   int64_t blocks = (total + 255) / 256;
   if (blocks > 148 * 32) blocks = 148 * 32;
   const bool vec = ((reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(vl) | reinterpret_cast<uintptr_t>(vr)) & 15u) == 0 &&
                    (rows == 1 || ((ldv | ldo) & 1) == 0);
   if (vec)
      recon_run_kernel<K, M, true><<<(unsigned)blocks, 256, 0, st>>>(v, ldv, vl, vr, ldo, w->ncells, rows, make_wenok(w->eps));
   else
      recon_run_kernel<K, M, false><<<(unsigned)blocks, 256, 0, st>>>(v, ldv, vl, vr, ldo, w->ncells, rows, make_wenok(w->eps));
}

template <class M>
static void launch_run_k(const Weno *w, int64_t rows, const double *v, int64_t ldv, double *vl, double *vr, int64_t ldo, cudaStream_t st) {
   if (w->k == 1)
      launch_run<1, M>(w, rows, v, ldv, vl, vr, ldo, st);
   else if (w->k == 2)
      launch_run<2, M>(w, rows, v, ldv, vl, vr, ldo, st);
   else
      launch_run<3, M>(w, rows, v, ldv, vl, vr, ldo, st);
}

int weno_reconstruct_launch(const Weno *w, int64_t rows, const double *v, int64_t ldv, int64_t incv, double *vl,
                            double *vr, int64_t ldo, cudaStream_t st) {
   if (w->uniform && incv == 1) { // the common case: contiguous rows, uniform tables
      if (w->mode == HRWENO_MODE_FAST)
         launch_run_k<Fast>(w, rows, v, ldv, vl, vr, ldo, st);
      else
         launch_run_k<Strict>(w, rows, v, ldv, vl, vr, ldo, st);
      HRW_CUDA(cudaGetLastError());
      return HRWENO_OK;
   }
   // strided sections (example2:107) and per-cell coefficient tables: one cell per thread, reference order
   const int64_t total = rows * w->ncells;
   int64_t blocks = (total + 255) / 256;
   if (blocks > 148 * 32) blocks = 148 * 32;
   const int cf = incv == 1;
   switch (w->k) {
   case 1:
      recon_kernel<1, Strict><<<(unsigned)blocks, 256, 0, st>>>(v, ldv, incv, vl, vr, ldo, w->ncells, rows, make_wenok(w->eps), w->d_cnu, cf);
      break;
   case 2:
      recon_kernel<2, Strict><<<(unsigned)blocks, 256, 0, st>>>(v, ldv, incv, vl, vr, ldo, w->ncells, rows, make_wenok(w->eps), w->d_cnu, cf);
      break;
   default:
      recon_kernel<3, Strict><<<(unsigned)blocks, 256, 0, st>>>(v, ldv, incv, vl, vr, ldo, w->ncells, rows, make_wenok(w->eps), w->d_cnu, cf);
   }
   HRW_CUDA(cudaGetLastError());
   return HRWENO_OK;
}

} // namespace hrw
