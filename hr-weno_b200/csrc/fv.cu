// fv.cu -- the fused finite-volume operator object (hrweno_fv) and its stage dispatch.
#include "fv1d.cuh"
#include "fv2d.cuh"

#include <float.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <new>

namespace hrw {

// ------------------------------------------------------------------------------------------------
size_t Fv::state_doubles() const { return (size_t)nrows_alloc * (size_t)pitch; }

double *Fv::cell0(double *base) const { return base + (d.ndim == 2 ? (int64_t)PAD2 * pitch : 0) + PAD; }
const double *Fv::cell0(const double *base) const { return base + (d.ndim == 2 ? (int64_t)PAD2 * pitch : 0) + PAD; }

int Fv::alloc_state(double **out) const {
   double *p = nullptr;
   HRW_CUDA(cudaMalloc(&p, state_doubles() * sizeof(double)));
   cudaError_t e = cudaMemset(p, 0, state_doubles() * sizeof(double));
   if (e != cudaSuccess) {
      cudaFree(p);
      return cuda_fail(e, "cudaMemset", __FILE__, __LINE__);
   }
   *out = p;
   return HRWENO_OK;
}

Fv::~Fv() {
   fv_halo_free(this);
   cudaFree(d_width[0]);
   cudaFree(d_width[1]);
   cudaFree(d_rwidth[0]);
   cudaFree(d_rwidth[1]);
   cudaFree(d_wtab);
   cudaFree(d_widx);
   for (int a = 0; a < 2; ++a) {
      cudaFree(d_cnu_base[a]);
      cudaFree(d_fcoef[a]);
      cudaFree(d_ccoef[a]);
   }
   cudaFree(d_scratch_in);
   cudaFree(d_scratch_out);
   if (stream) cudaStreamDestroy(stream);
}

static int upload_width(const double *host, int64_t n, double **dev) {
   std::vector<double> tmp((size_t)n + PAD, 1.0);
   std::memcpy(tmp.data(), host, sizeof(double) * (size_t)n);
   HRW_CUDA(cudaMalloc(dev, tmp.size() * sizeof(double)));
   HRW_CUDA(cudaMemcpy(*dev, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice));
   return HRWENO_OK;
}

__global__ void recip_array_kernel(const double *w, double *rw, int64_t n) {
   const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
   if (i < n) rw[i] = exact_recip(w[i]);
}

// refined reciprocals of the dictionary widths, computed with the very sequence the kernels' divisions use
__global__ void wtab_recip_kernel(double2 *tab, int count) {
   const int i = threadIdx.x;
   if (i < 256) {
      double2 e = tab[i];
      if (i >= count) e.x = 1.0;
      e.y = exact_recip(e.x);
      tab[i] = e;
   }
}

// 1D widths.  The cell widths of the local slab -- grid1%linear's own arithmetic for GRID_LINEAR (grids.f90:76-79,247:
// edges(i) = xmin + rx*i, width = edges(i) - edges(i-1), separately rounded) or the caller's array -- are deduplicated
// into a dictionary when they take at most 256 distinct values (a linear grid takes a handful: the roundings of the
// edges), else streamed as an array.
static int setup_width_1d(Fv *fv, const hrweno_fv_desc *desc) {
   const int64_t n = fv->n0;
   std::vector<double> tab;
   std::vector<unsigned char> idx((size_t)n + 2048, 0); // the stage kernel stages a whole tile of indices (<= 1056 B from a 16-B boundary below n)
   bool dict = true;
   int last = 0;
   double el = 0.0;
   const bool linear = desc->grid_kind == HRWENO_GRID_LINEAR;
   if (linear) {
      fv->rx = (desc->xmax - desc->xmin) / (double)fv->d.global_n; // grids.f90:76
      volatile double p = fv->rx * (double)fv->d.global_offset;
      el = desc->xmin + p;
   }
   for (int64_t i = 0; i < n && dict; ++i) {
      double wv;
      if (linear) {
         volatile double p = fv->rx * (double)(fv->d.global_offset + i + 1); // grids.f90:78 (mul and add rounded separately)
         const double er = desc->xmin + p;
         wv = er - el; // grids.f90:247
         el = er;
      } else {
         wv = desc->width[0][i];
      }
      if (!(wv > 0x1p-200 && wv < 0x1p200)) { // outside the range the shared-reciprocal division is proven for
         dict = false;
         break;
      }
      int found = -1;
      if (!tab.empty() && tab[(size_t)last] == wv) {
         found = last;
      } else {
         for (size_t q = 0; q < tab.size(); ++q)
            if (tab[q] == wv) {
               found = (int)q;
               break;
            }
      }
      if (found < 0) {
         if (tab.size() == 256) {
            dict = false;
            break;
         }
         tab.push_back(wv);
         found = (int)tab.size() - 1;
      }
      last = found;
      idx[(size_t)i] = (unsigned char)found;
   }
   if (dict) {
      std::vector<double2> t2(256, make_double2(1.0, 1.0));
      for (size_t q = 0; q < tab.size(); ++q) t2[q].x = tab[q];
      HRW_CUDA(cudaMalloc(&fv->d_wtab, 256 * sizeof(double2)));
      HRW_CUDA(cudaMemcpy(fv->d_wtab, t2.data(), 256 * sizeof(double2), cudaMemcpyHostToDevice));
      wtab_recip_kernel<<<1, 256>>>(fv->d_wtab, (int)tab.size());
      HRW_CUDA(cudaGetLastError());
      HRW_CUDA(cudaMalloc(&fv->d_widx, idx.size()));
      HRW_CUDA(cudaMemcpy(fv->d_widx, idx.data(), idx.size(), cudaMemcpyHostToDevice));
      fv->width_dict = true;
      return HRWENO_OK;
   }
   // array mode
   if (linear) {
      std::vector<double> wv((size_t)n);
      volatile double p0 = fv->rx * (double)fv->d.global_offset;
      double e0 = desc->xmin + p0;
      for (int64_t i = 0; i < n; ++i) {
         volatile double p = fv->rx * (double)(fv->d.global_offset + i + 1);
         const double er = desc->xmin + p;
         wv[(size_t)i] = er - e0;
         e0 = er;
      }
      return upload_width(wv.data(), n, &fv->d_width[0]);
   }
   return upload_width(desc->width[0], n, &fv->d_width[0]);
}

int fv_create(Fv **out, const hrweno_fv_desc *desc) {
   if (!out || !desc) return fail(HRWENO_EINVAL, "hrweno_fv_create: null argument");
   if (desc->abi_version != HRWENO_ABI_VERSION) return fail(HRWENO_EINVAL, "hrweno_fv_create: abi_version mismatch");
   if (desc->ndim != 1 && desc->ndim != 2) return fail(HRWENO_EINVAL, "Invalid input 'ndim'. Valid range: 1 <= ndim <= 2.");
   for (int a = 0; a < desc->ndim; ++a)
      if (!(desc->n[a] > 0)) return fail(HRWENO_EINVAL, "Invalid input 'ncells'. Valid range: ncells > 0."); // weno.f90:75
   if (!(desc->k >= 1 && desc->k <= 3)) return fail(HRWENO_EINVAL, "Invalid input 'k'. Valid range: 1 <= k <= 3."); // :84
   if (!(desc->eps > DBL_EPSILON)) return fail(HRWENO_EINVAL, "Invalid input 'eps'. Valid range: eps > epsilon.");    // :94
   if (desc->flux_model != HRWENO_FLUX_BURGERS && desc->flux_model != HRWENO_FLUX_LINEAR)
      return fail(HRWENO_EINVAL, "hrweno_fv_create: unknown flux_model");
   if (desc->flux_scheme != HRWENO_SCHEME_GODUNOV && desc->flux_scheme != HRWENO_SCHEME_LAX_FRIEDRICHS)
      return fail(HRWENO_EINVAL, "hrweno_fv_create: unknown flux_scheme");
   if (desc->bc != HRWENO_BC_COPY_NEIGHBOUR && desc->bc != HRWENO_BC_ZERO_FLUX)
      return fail(HRWENO_EINVAL, "hrweno_fv_create: unknown bc");
   if (desc->mode != HRWENO_MODE_STRICT && desc->mode != HRWENO_MODE_FAST)
      return fail(HRWENO_EINVAL, "hrweno_fv_create: unknown mode");
   if (desc->bc == HRWENO_BC_COPY_NEIGHBOUR)
      for (int a = 0; a < desc->ndim; ++a)
         if (desc->n[a] < 2) return fail(HRWENO_EINVAL, "hrweno_fv_create: copy-neighbour boundary needs ncells >= 2");
   for (int a = 0; a < desc->ndim; ++a)
      if (desc->n[a] > 2000000000LL) return fail(HRWENO_EINVAL, "hrweno_fv_create: more than 2e9 cells along one axis of one slab");
   if (desc->grid_kind == HRWENO_GRID_LINEAR) {
      if (desc->ndim != 1) return fail(HRWENO_EINVAL, "hrweno_fv_create: GRID_LINEAR is 1D only (pass width arrays)");
      if (!(desc->xmax > desc->xmin)) return fail(HRWENO_EINVAL, "Invalid input 'xmin', 'xmax'. Valid range: xmax > xmin."); // grids.f90:66
   } else if (desc->grid_kind == HRWENO_GRID_WIDTH_ARRAY) {
      for (int a = 0; a < desc->ndim; ++a)
         if (!desc->width[a]) return fail(HRWENO_EINVAL, "hrweno_fv_create: width array missing");
   } else {
      return fail(HRWENO_EINVAL, "hrweno_fv_create: unknown grid_kind");
   }
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
      return fail(HRWENO_ECUDA, "no CUDA device (this library has no CPU fallback)");

   Fv *fv = new (std::nothrow) Fv();
   if (!fv) return fail(HRWENO_ENOMEM, "out of host memory");
   fv->d = *desc;
   fv->d.width[0] = fv->d.width[1] = nullptr;
   fv->n0 = desc->n[0];
   fv->n1 = desc->ndim == 2 ? desc->n[1] : 1;
   fv->rows = desc->ndim == 1 ? (desc->rows < 1 ? 1 : desc->rows) : 1;
   fv->neq = desc->ndim == 1 ? fv->n0 * fv->rows : fv->n0 * fv->n1;
   fv->pitch = padded_pitch(fv->n0);
   fv->nrows_alloc = desc->ndim == 1 ? fv->rows : fv->n1 + 2 * PAD2;
   if (fv->d.nranks < 1) fv->d.nranks = 1;
   if (fv->d.nranks == 1) {
      fv->d.rank = 0;
      fv->d.global_offset = 0;
      fv->d.global_n = desc->ndim == 1 ? fv->n0 : fv->n1;
   }
   int st = HRWENO_OK;
   if (desc->ndim == 1) {
      st = setup_width_1d(fv, desc);
   } else {
      for (int a = 0; a < desc->ndim && st == HRWENO_OK; ++a) {
         for (int64_t i = 0; i < desc->n[a]; ++i)
            if (!(desc->width[a][i] > 0x1p-200 && desc->width[a][i] < 0x1p200)) st = fail(HRWENO_EINVAL, "hrweno_fv_create: cell width outside (2^-200, 2^200)");
         if (st == HRWENO_OK) st = upload_width(desc->width[a], desc->n[a], &fv->d_width[a]);
         if (st == HRWENO_OK) {
            const int64_t np = desc->n[a] + PAD;
            cudaError_t e = cudaMalloc(&fv->d_rwidth[a], (size_t)np * sizeof(double));
            if (e != cudaSuccess) {
               st = cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
            } else {
               recip_array_kernel<<<(unsigned)((np + 255) / 256), 256>>>(fv->d_width[a], fv->d_rwidth[a], np);
               e = cudaGetLastError();
               if (e != cudaSuccess) st = cuda_fail(e, "recip_array_kernel", __FILE__, __LINE__);
            }
         }
      }
   }
   if (st == HRWENO_OK) {
      cudaError_t e = cudaStreamCreateWithFlags(&fv->stream, cudaStreamNonBlocking);
      if (e != cudaSuccess) st = cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__);
   }
   if (st != HRWENO_OK) {
      delete fv;
      return st;
   }
   *out = fv;
   return HRWENO_OK;
}

// ------------------------------------------------------------------------------------------------
// General path (fvgen.cu): non-uniform grids and x-dependent fluxes.  Both setters move the operator from the tuned
// stage kernels to the general one; they are called after hrweno_fv_create and before the first rhs / integrate call.
// ------------------------------------------------------------------------------------------------
__global__ void expand_width_kernel(const double2 *__restrict__ tab, const unsigned char *__restrict__ idx, double *__restrict__ w, int64_t n) {
   for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) w[i] = tab[idx[i]].x;
}

static int upload_doubles(const double *host, int64_t n, double **dev) {
   cudaFree(*dev);
   *dev = nullptr;
   HRW_CUDA(cudaMalloc(dev, (size_t)n * sizeof(double)));
   HRW_CUDA(cudaMemcpy(*dev, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
   return HRWENO_OK;
}

static int fv_enable_general(Fv *fv, const char *who) {
   (void)who;
   if (fv->d.ndim == 1 && !fv->d_width[0]) { // the 1D widths live in the dictionary: the general kernel reads a plain array
      double *w = nullptr;
      HRW_CUDA(cudaMalloc(&w, (size_t)(fv->n0 + PAD) * sizeof(double)));
      int64_t blocks = (fv->n0 + 255) / 256;
      if (blocks > 148 * 16) blocks = 148 * 16;
      expand_width_kernel<<<(unsigned)blocks, 256>>>(fv->d_wtab, fv->d_widx, w, fv->n0);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
         cudaFree(w);
         return cuda_fail(e, "expand_width_kernel", __FILE__, __LINE__);
      }
      fv->d_width[0] = w;
   }
   fv->general = true;
   return HRWENO_OK;
}

// weno(ncells, k, eps, xedges) for the sweep along `axis` (weno.f90:100-112): cnu on the host, once (weno.f90:225-228)
int fv_set_xedges(Fv *fv, int axis, const double *xedges) {
   if (!xedges) return fail(HRWENO_EINVAL, "hrweno_fv_set_xedges: null argument");
   if (axis < 0 || axis >= fv->d.ndim) return fail(HRWENO_EINVAL, "hrweno_fv_set_xedges: invalid axis");
   std::lock_guard<std::mutex> lock(fv->mtx);
   HRW_TRY(fv_enable_general(fv, "hrweno_fv_set_xedges"));
   const int64_t n = axis == 0 ? fv->n0 : fv->n1;
   const int k = fv->d.k, KK = k * (k + 1);
   // Tables of the local cells plus one ghost cell on either side (the faces on a slab interface need the neighbour's edge
   // cell reconstructed with ITS table).  On the decomposed axis of a slab, xedges is the GLOBAL edge array: the tables of
   // cells [lo, hi) are computed from the global edges around them (a table reaches k+1 edges beyond its cell, so a
   // window cut k+2 cells away from the cells we keep gives the very same arithmetic as the whole grid); beyond the
   // global ends weno_calc_cnu's own linear ghost edges apply (weno.f90:242-255).
   const bool decomposed = fv->d.nranks > 1 && axis == fv->d.ndim - 1;
   const int64_t gn = decomposed ? fv->d.global_n : n, off = decomposed ? fv->d.global_offset : 0;
   const int64_t lo = std::max<int64_t>(0, off - (k + 2)), hi = std::min<int64_t>(gn, off + n + (k + 2));
   std::vector<double> win;
   weno_calc_cnu_host(hi - lo, k, xedges + lo, win);
   std::vector<double> cnu((size_t)(n + 2) * KK);
   for (int64_t c = -1; c <= n; ++c) {
      int64_t gc = off + c; // global cell; ghost tables beyond the global ends replicate the end cell's (never used: boundary rule)
      gc = gc < 0 ? 0 : (gc > gn - 1 ? gn - 1 : gc);
      std::memcpy(&cnu[(size_t)(c + 1) * KK], &win[(size_t)(gc - lo) * KK], sizeof(double) * (size_t)KK);
   }
   cudaFree(fv->d_cnu_base[axis]);
   fv->d_cnu_base[axis] = fv->d_cnu[axis] = nullptr;
   HRW_TRY(upload_doubles(cnu.data(), (int64_t)cnu.size(), &fv->d_cnu_base[axis]));
   fv->d_cnu[axis] = fv->d_cnu_base[axis] + KK;
   return HRWENO_OK;
}

// f(v, x) = (model(v)*cross[i_other])*face[i_face] for the faces along `axis`; nullptr = factor absent
int fv_set_flux_coef(Fv *fv, int axis, const double *face, const double *cross) {
   if (axis < 0 || axis >= fv->d.ndim) return fail(HRWENO_EINVAL, "hrweno_fv_set_flux_coef: invalid axis");
   if (cross && fv->d.ndim != 2) return fail(HRWENO_EINVAL, "hrweno_fv_set_flux_coef: a cross coefficient needs ndim == 2");
   std::lock_guard<std::mutex> lock(fv->mtx);
   HRW_TRY(fv_enable_general(fv, "hrweno_fv_set_flux_coef"));
   const int64_t n = axis == 0 ? fv->n0 : fv->n1, nother = axis == 0 ? fv->n1 : fv->n0;
   cudaFree(fv->d_fcoef[axis]);
   cudaFree(fv->d_ccoef[axis]);
   fv->d_fcoef[axis] = fv->d_ccoef[axis] = nullptr;
   if (face) HRW_TRY(upload_doubles(face, n + 1, &fv->d_fcoef[axis]));
   if (cross) HRW_TRY(upload_doubles(cross, nother, &fv->d_ccoef[axis]));
   return HRWENO_OK;
}

int fv_set_flux_time_fn(Fv *fv, hrweno_time_fn g, void *ctx) {
   std::lock_guard<std::mutex> lock(fv->mtx);
   if (g) HRW_TRY(fv_enable_general(fv, "hrweno_fv_set_flux_time_fn"));
   fv->tfn = g;
   fv->tfn_ctx = ctx;
   return HRWENO_OK;
}

// ------------------------------------------------------------------------------------------------
// dense <-> padded copies.  pack also writes the ghost cells of physical boundaries.
// ------------------------------------------------------------------------------------------------
__global__ void pack_kernel(const double *__restrict__ dense, double *__restrict__ padded0, int64_t n, int64_t rows,
                            int64_t pitch, int k, int phys_left, int phys_right) {
   // one thread per (row, i) with i in [-k, n+k)
   const int64_t span = n + 2 * k;
   const int64_t total = rows * span;
   for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      const int64_t row = idx / span;
      const int64_t i = idx - row * span - k;
      if (i < 0 && !phys_left) continue;
      if (i >= n && !phys_right) continue;
      const int64_t ic = i < 0 ? 0 : (i >= n ? n - 1 : i);
      padded0[row * pitch + i] = dense[row * n + ic];
   }
}

__global__ void unpack_kernel(const double *__restrict__ padded0, double *__restrict__ dense, int64_t n, int64_t rows,
                              int64_t pitch) {
   const int64_t total = rows * n;
   for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      const int64_t row = idx / n;
      const int64_t i = idx - row * n;
      dense[idx] = padded0[row * pitch + i];
   }
}

// 2D: ghost rows/columns of physical boundaries are edge replicas along the respective axis
__global__ void pack2d_kernel(const double *__restrict__ dense, double *__restrict__ padded0, int64_t n0, int64_t n1,
                              int64_t pitch, int k, int phys_lo, int phys_hi) {
   const int64_t span0 = n0 + 2 * k, span1 = n1 + 2 * k;
   const int64_t total = span0 * span1;
   for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      const int64_t jj = idx / span0;
      const int64_t i = idx - jj * span0 - k;
      const int64_t j = jj - k;
      if (j < 0 && !phys_lo) continue;
      if (j >= n1 && !phys_hi) continue;
      const int64_t ic = i < 0 ? 0 : (i >= n0 ? n0 - 1 : i);
      const int64_t jc = j < 0 ? 0 : (j >= n1 ? n1 - 1 : j);
      padded0[j * pitch + i] = dense[jc * n0 + ic];
   }
}

static int grid_for(int64_t total, int threads) {
   int64_t b = (total + threads - 1) / threads;
   const int64_t cap = 148 * 16;
   return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

int fv_pack(Fv *fv, const double *dense, double *padded0, cudaStream_t st) {
   const int k = fv->d.k;
   const int phys_lo = fv->d.rank == 0, phys_hi = fv->d.rank == fv->d.nranks - 1;
   if (fv->d.ndim == 1) {
      const int64_t total = fv->rows * (fv->n0 + 2 * k);
      pack_kernel<<<grid_for(total, 256), 256, 0, st>>>(dense, padded0, fv->n0, fv->rows, fv->pitch, k, phys_lo, phys_hi);
   } else {
      const int64_t total = (fv->n0 + 2 * k) * (fv->n1 + 2 * k);
      pack2d_kernel<<<grid_for(total, 256), 256, 0, st>>>(dense, padded0, fv->n0, fv->n1, fv->pitch, k, phys_lo, phys_hi);
   }
   fv->launches++;
   HRW_CUDA(cudaGetLastError());
   return HRWENO_OK;
}

int fv_unpack(Fv *fv, const double *padded0, double *dense, cudaStream_t st) {
   const int64_t rows = fv->d.ndim == 1 ? fv->rows : fv->n1;
   const int64_t total = rows * fv->n0;
   unpack_kernel<<<grid_for(total, 256), 256, 0, st>>>(padded0, dense, fv->n0, rows, fv->pitch);
   fv->launches++;
   HRW_CUDA(cudaGetLastError());
   return HRWENO_OK;
}

// ------------------------------------------------------------------------------------------------
// K5: local max |f'(v)| (the Lax-Friedrichs alpha, fluxes.f90:40-43).  The reference takes alpha from the caller and has
// no reduction; this helper lets a caller compute it on the device -- Burgers (example1:120): f'(v) = v; linear flux:
// |a| -- and combine the per-GPU values with one max all-reduce (hr-weno_b200/slab.py: global_max_wavespeed, NCCL).
// Non-negative doubles order like their bit patterns, so the grid-wide maximum is one integer atomicMax per CTA.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) max_abs_kernel(const double *__restrict__ v, int64_t n, unsigned long long *out) {
   double m = 0.0;
   const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
   const bool vec = (reinterpret_cast<uintptr_t>(v) & 15u) == 0;
   if (vec) {
      const double2 *v2 = reinterpret_cast<const double2 *>(v);
      for (int64_t i = tid; i < n / 2; i += nth) {
         const double2 t = __ldg(v2 + i);
         m = fmax(m, fmax(fabs(t.x), fabs(t.y))); // fmax drops a NaN operand, as Fortran's maxval(abs(v)) does not: see the header
      }
      if (tid == 0 && (n & 1)) m = fmax(m, fabs(v[n - 1]));
   } else {
      for (int64_t i = tid; i < n; i += nth) m = fmax(m, fabs(__ldg(v + i)));
   }
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
   __shared__ double s_m[8];
   if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
   __syncthreads();
   if (threadIdx.x < 8) {
      m = s_m[threadIdx.x];
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffu, m, o));
      if (threadIdx.x == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
   }
}

int fv_max_wavespeed(Fv *fv, const double *v_dev, double *out_dev, cudaStream_t st) {
   const hrweno_fv_desc &d = fv->d;
   if (fv->d_fcoef[0] || fv->d_fcoef[1] || fv->d_ccoef[0] || fv->d_ccoef[1])
      return fail(HRWENO_EINVAL, "hrweno_fv_max_wavespeed_dev: not available with x-dependent flux coefficients");
   if (d.flux_model == HRWENO_FLUX_LINEAR) {
      const double a = d.ndim == 2 ? fmax(fabs(d.flux_coef[0]), fabs(d.flux_coef[1])) : fabs(d.flux_coef[0]);
      HRW_CUDA(cudaMemcpyAsync(out_dev, &a, sizeof(double), cudaMemcpyHostToDevice, st));
      HRW_CUDA(cudaStreamSynchronize(st)); // `a` lives on this stack frame
      return HRWENO_OK;
   }
   HRW_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double), st));
   int64_t blocks = (fv->neq / 2 + 255) / 256;
   blocks = blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks);
   max_abs_kernel<<<(unsigned)blocks, 256, 0, st>>>(v_dev, fv->neq, reinterpret_cast<unsigned long long *>(out_dev));
   fv->launches++;
   HRW_CUDA(cudaGetLastError());
   return HRWENO_OK;
}

// ------------------------------------------------------------------------------------------------
// stage dispatch
// ------------------------------------------------------------------------------------------------
int fv1d_launch_k1_m0(int, int, int, int, const Fv1dGeom &, const StageArgs &, cudaStream_t);
int fv1d_launch_k2_m0(int, int, int, int, const Fv1dGeom &, const StageArgs &, cudaStream_t);
int fv1d_launch_k3_m0(int, int, int, int, const Fv1dGeom &, const StageArgs &, cudaStream_t);
int fv1d_launch_k1_m1(int, int, int, int, const Fv1dGeom &, const StageArgs &, cudaStream_t);
int fv1d_launch_k2_m1(int, int, int, int, const Fv1dGeom &, const StageArgs &, cudaStream_t);
int fv1d_launch_k3_m1(int, int, int, int, const Fv1dGeom &, const StageArgs &, cudaStream_t);
int fv1d_tile_cells(int mode, int half_tile);
int fv1d_tile_slots(int mode, int half_tile);

int fv1d_launch(int k, int mode, int combine, int fk, int wk, int half_tile, const Fv1dGeom &g, const StageArgs &a, cudaStream_t st) {
   if (mode == HRWENO_MODE_STRICT) {
      if (k == 1) return fv1d_launch_k1_m0(combine, fk, wk, half_tile, g, a, st);
      if (k == 2) return fv1d_launch_k2_m0(combine, fk, wk, half_tile, g, a, st);
      return fv1d_launch_k3_m0(combine, fk, wk, half_tile, g, a, st);
   }
   if (k == 1) return fv1d_launch_k1_m1(combine, fk, wk, half_tile, g, a, st);
   if (k == 2) return fv1d_launch_k2_m1(combine, fk, wk, half_tile, g, a, st);
   return fv1d_launch_k3_m1(combine, fk, wk, half_tile, g, a, st);
}

static int fv1d_flux_kind(const hrweno_fv_desc &d) {
   return (d.flux_model == HRWENO_FLUX_BURGERS && d.flux_scheme == HRWENO_SCHEME_GODUNOV) ? FK_BURGERS_GODUNOV : FK_GENERIC;
}

// tile = (threads - 2) * R cells.  The half-size tile exists for Burgers/Godunov with a width dictionary; it is used
// when it covers the row with fewer (partly idle) thread runs, e.g. 4096-cell rows: 9 x 504 instead of 5 x 1016
static int fv1d_half_tile(const Fv *fv) {
   if (fv1d_flux_kind(fv->d) != FK_BURGERS_GODUNOV || !fv->width_dict) return 0;
   if (const char *e = std::getenv("HRWENO_FORCE_HALF_TILE")) return std::atoi(e) != 0; // tuning A/B only
   const int64_t t0 = fv1d_tile_cells(fv->d.mode, 0), t1 = fv1d_tile_cells(fv->d.mode, 1);
   const double slots0 = (double)((fv->n0 + t0 - 1) / t0) * (double)fv1d_tile_slots(fv->d.mode, 0);
   const double slots1 = (double)((fv->n0 + t1 - 1) / t1) * (double)fv1d_tile_slots(fv->d.mode, 1);
   return slots1 < 0.97 * slots0;
}

void fv_tiling_1d(const Fv *fv, int *tile_cells, int *tiles_per_row) {
   const int tile = fv1d_tile_cells(fv->d.mode, fv1d_half_tile(fv));
   *tile_cells = tile;
   *tiles_per_row = (int)((fv->n0 + tile - 1) / tile);
}

static int fv_stage_impl(Fv *fv, int combine, const StageArgs &args, const HaloIO *io, cudaStream_t st) {
   const hrweno_fv_desc &d = fv->d;
   if (fv->general && d.ndim == 1) { // non-uniform grid and/or x-/t-dependent flux in 1D: the general kernel, reference operation order
      HRW_TRY(fvgen_stage(fv, combine, args, st));
      fv->launches++;
      return HRWENO_OK;
   }
   if (d.ndim == 2) {
      HRW_TRY(fv2d_stage(fv, combine, args, st));
      fv->launches++;
      return HRWENO_OK;
   }
   Fv1dGeom g{};
   const int fk = fv1d_flux_kind(d);
   const int half_tile = fv1d_half_tile(fv);
   const int tile = fv1d_tile_cells(d.mode, half_tile);
   g.n = fv->n0;
   g.ld = fv->pitch;
   g.tiles_per_row = (fv->n0 + tile - 1) / tile;
   g.rows = fv->rows;
   g.tile_begin = args.tile_begin;
   g.tile_end = args.tile_end > 0 ? args.tile_end : (int)(g.tiles_per_row * g.rows);
   g.width = fv->d_width[0];
   g.wtab = fv->d_wtab;
   g.widx = fv->d_widx;
   g.kc = make_wenok(d.eps);
   g.flux = FluxCfg{d.flux_model, d.flux_scheme, d.flux_coef[0], d.alpha};
   g.bc = d.bc;
   g.phys_left = d.rank == 0;
   g.phys_right = d.rank == d.nranks - 1;
   if (io) g.halo = *io;
   HRW_TRY(fv1d_launch(d.k, d.mode, combine, fk, fv->width_dict ? WK_DICT : WK_ARRAY, half_tile, g, args, st));
   fv->launches++;
   return HRWENO_OK;
}

int fv_stage(Fv *fv, int combine, const StageArgs &args, cudaStream_t st) { return fv_stage_impl(fv, combine, args, nullptr, st); }

int fv_stage_halo(Fv *fv, int combine, const StageArgs &args, bool halo_out, cudaStream_t st) {
   if (fv->d.nranks <= 1) return fv_stage_impl(fv, combine, args, nullptr, st);
   if (fv->general) { // the general kernels take their ghost cells from the padded state: exchange after the stage
      HRW_TRY(fv_stage_impl(fv, combine, args, nullptr, st));
      if (halo_out && !args.out_dense) HRW_TRY(fv_exchange(fv, args.out, st));
      return HRWENO_OK;
   }
   HaloIO io;
   HRW_TRY(fv_halo_io(fv, args.vin, args.out, halo_out && !args.out_dense, &io));
   if (io.edge_first) return fv_stage_impl(fv, combine, args, &io, st); // halo traffic inside the stage kernel
   HRW_TRY(fv_stage_impl(fv, combine, args, nullptr, st));
   if (halo_out && !args.out_dense) HRW_TRY(fv_exchange(fv, args.out, st));
   return HRWENO_OK;
}

} // namespace hrw
