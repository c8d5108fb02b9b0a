// fv.cu -- the fused finite-volume operator object (hrweno_fv) and its stage dispatch.
#include "fv1d.cuh"
#include "fv2d.cuh"

#include <float.h>

#include <cstring>
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
   if (desc->grid_kind == HRWENO_GRID_LINEAR) {
      fv->rx = (desc->xmax - desc->xmin) / (double)fv->d.global_n; // grids.f90:76
   } else {
      for (int a = 0; a < desc->ndim && st == HRWENO_OK; ++a) st = upload_width(desc->width[a], desc->n[a], &fv->d_width[a]);
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
// stage dispatch
// ------------------------------------------------------------------------------------------------
constexpr int R1 = 4, NT1 = 256;

template <int K, int COMBINE, class M>
static void launch1d(const Fv1dGeom &g, const StageArgs &a, int64_t rows, cudaStream_t st) {
   // persistent grid: SMs x resident CTAs of this instantiation (queried once), never more than there are tiles
   static int resident = 0;
   if (resident == 0) {
      int dev = 0, sms = 0, per_sm = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fv1d_stage_kernel<K, COMBINE, M, R1, NT1>, NT1, 0);
      resident = (sms > 0 ? sms : 148) * (per_sm > 0 ? per_sm : 1);
   }
   int64_t blocks = rows * g.tiles_per_row;
   if (blocks > resident) blocks = resident;
   fv1d_stage_kernel<K, COMBINE, M, R1, NT1><<<(unsigned)blocks, NT1, 0, st>>>(g, a);
}

template <int K, class M>
static void launch1d_c(int combine, const Fv1dGeom &g, const StageArgs &a, int64_t rows, cudaStream_t st) {
   switch (combine) {
   case C_RHS: launch1d<K, C_RHS, M>(g, a, rows, st); break;
   case C_EULER: launch1d<K, C_EULER, M>(g, a, rows, st); break;
   case C_RK2_FINAL: launch1d<K, C_RK2_FINAL, M>(g, a, rows, st); break;
   case C_RK3_S2: launch1d<K, C_RK3_S2, M>(g, a, rows, st); break;
   case C_RK3_S3: launch1d<K, C_RK3_S3, M>(g, a, rows, st); break;
   default: launch1d<K, C_MS, M>(g, a, rows, st); break;
   }
}

template <class M>
static void launch1d_k(int k, int combine, const Fv1dGeom &g, const StageArgs &a, int64_t rows, cudaStream_t st) {
   if (k == 1)
      launch1d_c<1, M>(combine, g, a, rows, st);
   else if (k == 2)
      launch1d_c<2, M>(combine, g, a, rows, st);
   else
      launch1d_c<3, M>(combine, g, a, rows, st);
}

int fv_stage(Fv *fv, int combine, const StageArgs &args, cudaStream_t st) {
   const hrweno_fv_desc &d = fv->d;
   if (d.ndim == 2) {
      HRW_TRY(fv2d_stage(fv, combine, args, st));
      fv->launches++;
      return HRWENO_OK;
   }
   Fv1dGeom g{};
   g.n = fv->n0;
   g.ld = fv->pitch;
   g.tiles_per_row = (fv->n0 + (NT1 - 2) * R1 - 1) / ((NT1 - 2) * R1);
   g.rows = fv->rows;
   g.width = fv->d_width[0];
   g.xmin = d.xmin;
   g.rx = fv->rx;
   g.goff = d.global_offset;
   g.eps = d.eps;
   g.flux = FluxCfg{d.flux_model, d.flux_scheme, d.flux_coef[0], d.alpha};
   g.bc = d.bc;
   g.grid_kind = d.grid_kind;
   g.phys_left = d.rank == 0;
   g.phys_right = d.rank == d.nranks - 1;
   if (d.mode == HRWENO_MODE_STRICT)
      launch1d_k<Strict>(d.k, combine, g, args, fv->rows, st);
   else
      launch1d_k<Fast>(d.k, combine, g, args, fv->rows, st);
   fv->launches++;
   HRW_CUDA(cudaGetLastError());
   return HRWENO_OK;
}

} // namespace hrw
