// halo.cu -- slab-interface ghost cells over NVLink peer memory (halo = k cells per side per stage).
#include <cstring>

#include "fv2d.cuh"
#include "internal.hpp"

namespace hrw {

static __host__ __device__ inline size_t flag_off(int side, int slot) { return (size_t)(side * Halo::NSLOTS + slot) * 128; }
static constexpr size_t ERR_OFF = 2 * Halo::NSLOTS * 128;

static inline unsigned char *slot_ptr(unsigned char *mbox, size_t halo_doubles, int side, int slot) {
   return mbox + Halo::HDR_BYTES + ((size_t)side * Halo::NSLOTS + slot) * halo_doubles * sizeof(double);
}

// my first k cells (rows) -> left neighbour's "from right" slot; my last k -> right neighbour's "from left" slot
__global__ void halo_send_kernel(const double *__restrict__ cell0, int ndim, int64_t n0, int64_t n1, int64_t rows,
                                 int64_t pitch, int k, double *__restrict__ to_left, double *__restrict__ to_right) {
   const int64_t total = ndim == 1 ? rows * k : (int64_t)k * n0;
   for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      if (ndim == 1) {
         const int64_t row = idx / k, q = idx - row * k;
         if (to_left) to_left[idx] = cell0[row * pitch + q];
         if (to_right) to_right[idx] = cell0[row * pitch + n0 - k + q];
      } else {
         const int64_t q = idx / n0, x = idx - q * n0;
         if (to_left) to_left[idx] = cell0[q * pitch + x];
         if (to_right) to_right[idx] = cell0[(n1 - k + q) * pitch + x];
      }
   }
}

// runs after halo_send_kernel on the same stream: its peer stores are complete; publish the sequence number
__global__ void halo_signal_kernel(unsigned long long *flag_left, unsigned long long *flag_right, unsigned long long seq) {
   __threadfence_system();
   if (flag_left) *reinterpret_cast<volatile unsigned long long *>(flag_left) = seq;
   if (flag_right) *reinterpret_cast<volatile unsigned long long *>(flag_right) = seq;
   __threadfence_system();
}

__global__ void halo_recv_kernel(double *__restrict__ cell0, int ndim, int64_t n0, int64_t n1, int64_t rows, int64_t pitch,
                                 int k, const double *from_left, const double *from_right,
                                 const unsigned long long *flag_from_left, const unsigned long long *flag_from_right,
                                 unsigned long long seq, unsigned int *err, long long timeout_cycles) {
   __shared__ int ok;
   if (threadIdx.x == 0) {
      int good = 1;
      if (*reinterpret_cast<volatile unsigned int *>(err) != 0) {
         good = 0; // a previous exchange timed out: do not wait again
      } else {
         const long long t0 = clock64();
         for (int side = 0; side < 2 && good; ++side) {
            const unsigned long long *f = side == 0 ? flag_from_left : flag_from_right;
            if (!f) continue;
            while (*reinterpret_cast<const volatile unsigned long long *>(f) < seq) {
               if (clock64() - t0 > timeout_cycles) {
                  atomicExch(err, 1u);
                  good = 0;
                  break;
               }
               __nanosleep(100);
            }
         }
      }
      __threadfence_system();
      ok = good;
   }
   __syncthreads();
   if (!ok) return;
   const int64_t total = ndim == 1 ? rows * k : (int64_t)k * n0;
   for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      if (ndim == 1) {
         const int64_t row = idx / k, q = idx - row * k;
         if (from_left) cell0[row * pitch - k + q] = __ldcv(from_left + idx);
         if (from_right) cell0[row * pitch + n0 + q] = __ldcv(from_right + idx);
      } else {
         const int64_t q = idx / n0, x = idx - q * n0;
         if (from_left) cell0[(q - k) * pitch + x] = __ldcv(from_left + idx);
         if (from_right) cell0[(n1 + q) * pitch + x] = __ldcv(from_right + idx);
      }
   }
}

void fv_halo_free(Fv *fv) {
   Halo &h = fv->halo;
   for (int s = 0; s < 2; ++s)
      if (h.peer[s] && !h.local_peers) cudaIpcCloseMemHandle(h.peer[s]);
   cudaFree(h.mbox);
   h = Halo{};
}

// this rank's mailbox (allocated on the current device on first use)
static int halo_alloc_mbox(Fv *fv) {
   Halo &h = fv->halo;
   if (h.mbox) return HRWENO_OK;
   const int k = fv->d.k;
   h.halo_doubles = fv->d.ndim == 1 ? (size_t)fv->rows * k : (size_t)k * fv->n0;
   h.bytes = Halo::HDR_BYTES + 2 * Halo::NSLOTS * h.halo_doubles * sizeof(double);
   if (fv->d.ndim == 1 && fv->rows == 1) { // wide slots [side][2][WIDE_H] behind the per-stage ones
      h.bytes = (h.bytes + 255) & ~(size_t)255;
      h.wide_off = h.bytes;
      h.bytes += (size_t)2 * 2 * Halo::WIDE_H * sizeof(double);
   }
   HRW_CUDA(cudaMalloc(&h.mbox, h.bytes));
   HRW_CUDA(cudaMemset(h.mbox, 0, h.bytes));
   HRW_CUDA(cudaDeviceSynchronize());
   return HRWENO_OK;
}

// slabs of ONE process (mgpu.cu): the neighbours' mailboxes are ordinary device pointers of this address space; the
// caller has enabled peer access between the devices.  Call on every slab with its own device current.
int fv_halo_connect_local(Fv *fv, Fv *left, Fv *right) {
   HRW_TRY(halo_alloc_mbox(fv));
   Halo &h = fv->halo;
   if ((left && !left->halo.mbox) || (right && !right->halo.mbox)) return fail(HRWENO_ECOMM, "neighbour slab has no mailbox yet");
   h.peer[0] = left ? left->halo.mbox : nullptr;
   h.peer[1] = right ? right->halo.mbox : nullptr;
   h.local_peers = true;
   h.ready = true;
   return HRWENO_OK;
}
int fv_halo_prepare_local(Fv *fv) { return halo_alloc_mbox(fv); }

int fv_halo_export(Fv *fv, void *handle_out) {
   static_assert(sizeof(cudaIpcMemHandle_t) == HRWENO_IPC_HANDLE_BYTES, "IPC handle size");
   if (!handle_out) return fail(HRWENO_EINVAL, "null handle buffer");
   Halo &h = fv->halo;
   HRW_TRY(halo_alloc_mbox(fv));
   cudaIpcMemHandle_t ipc;
   HRW_CUDA(cudaIpcGetMemHandle(&ipc, h.mbox));
   std::memcpy(handle_out, &ipc, sizeof(ipc));
   return HRWENO_OK;
}

int fv_halo_import(Fv *fv, const void *left, const void *right) {
   Halo &h = fv->halo;
   if (!h.mbox) return fail(HRWENO_ECOMM, "hrweno_fv_import_halo: call hrweno_fv_export_halo first");
   const bool need_left = fv->d.rank > 0, need_right = fv->d.rank < fv->d.nranks - 1;
   if ((need_left && !left) || (need_right && !right))
      return fail(HRWENO_ECOMM, "hrweno_fv_import_halo: neighbour handle missing for an interior slab interface");
   const void *src[2] = {need_left ? left : nullptr, need_right ? right : nullptr};
   for (int s = 0; s < 2; ++s) {
      if (!src[s]) continue;
      cudaIpcMemHandle_t ipc;
      std::memcpy(&ipc, src[s], sizeof(ipc));
      void *p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
         cuda_fail(e, "cudaIpcOpenMemHandle", __FILE__, __LINE__);
         return HRWENO_ECOMM;
      }
      h.peer[s] = static_cast<unsigned char *>(p);
   }
   h.ready = true;
   return HRWENO_OK;
}

int fv_halo_status(Fv *fv) {
   Halo &h = fv->halo;
   if (fv->d.nranks <= 1 || !h.mbox) return HRWENO_OK;
   unsigned int err = 0;
   HRW_CUDA(cudaMemcpy(&err, h.mbox + ERR_OFF, sizeof(err), cudaMemcpyDeviceToHost));
   if (err) return fail(HRWENO_ECOMM, "halo exchange timed out waiting for a neighbour rank");
   return HRWENO_OK;
}

int fv_exchange(Fv *fv, double *cell0, cudaStream_t st) {
   if (fv->d.nranks <= 1) return HRWENO_OK;
   Halo &h = fv->halo;
   if (h.pending_buf == cell0) h.pending_buf = nullptr; // ghosts are about to be in place
   if (!h.ready) return fail(HRWENO_ECOMM, "slab decomposition needs hrweno_fv_export_halo / hrweno_fv_import_halo first");
   const unsigned long long seq = ++h.seq;
   const int slot = (int)(seq % Halo::NSLOTS);
   const int k = fv->d.k;
   const size_t hd = h.halo_doubles;
   // I am the right neighbour of my left neighbour: my first cells go to its side 1 ("from right"), and vice versa
   double *to_left = h.peer[0] ? reinterpret_cast<double *>(slot_ptr(h.peer[0], hd, 1, slot)) : nullptr;
   double *to_right = h.peer[1] ? reinterpret_cast<double *>(slot_ptr(h.peer[1], hd, 0, slot)) : nullptr;
   auto *fl = h.peer[0] ? reinterpret_cast<unsigned long long *>(h.peer[0] + flag_off(1, slot)) : nullptr;
   auto *fr = h.peer[1] ? reinterpret_cast<unsigned long long *>(h.peer[1] + flag_off(0, slot)) : nullptr;
   const int64_t total = (int64_t)hd;
   int blocks = (int)((total + 255) / 256);
   if (blocks > 296) blocks = 296;
   if (blocks < 1) blocks = 1;
   halo_send_kernel<<<blocks, 256, 0, st>>>(cell0, fv->d.ndim, fv->n0, fv->n1, fv->rows, fv->pitch, k, to_left, to_right);
   halo_signal_kernel<<<1, 1, 0, st>>>(fl, fr, seq);
   const double *from_left = h.peer[0] ? reinterpret_cast<const double *>(slot_ptr(h.mbox, hd, 0, slot)) : nullptr;
   const double *from_right = h.peer[1] ? reinterpret_cast<const double *>(slot_ptr(h.mbox, hd, 1, slot)) : nullptr;
   auto *ffl = h.peer[0] ? reinterpret_cast<const unsigned long long *>(h.mbox + flag_off(0, slot)) : nullptr;
   auto *ffr = h.peer[1] ? reinterpret_cast<const unsigned long long *>(h.mbox + flag_off(1, slot)) : nullptr;
   const long long timeout_cycles = 60LL * 1900000000LL; // ~60 s at 1.9 GHz: a dead neighbour, not a slow one
   halo_recv_kernel<<<blocks, 256, 0, st>>>(cell0, fv->d.ndim, fv->n0, fv->n1, fv->rows, fv->pitch, k, from_left, from_right, ffl,
                                            ffr, seq, reinterpret_cast<unsigned int *>(h.mbox + ERR_OFF), timeout_cycles);
   fv->launches += 3;
   HRW_CUDA(cudaGetLastError());
   return HRWENO_OK;
}

// ---- wide halos: once per host-pointer integrate call (ode.cu: chunk pipeline on slabs) -------------------------------
static constexpr size_t WIDE_FLAG_OFF = ERR_OFF + 128; // two flags on their own 128-B lines behind the error word
static inline double *wide_slot(unsigned char *mbox, size_t wide_off, int side, int slot) {
   return reinterpret_cast<double *>(mbox + wide_off) + ((size_t)side * 2 + slot) * Halo::WIDE_H;
}

__global__ void wide_send_kernel(const double *__restrict__ first, const double *__restrict__ last, double *__restrict__ to_left,
                                 double *__restrict__ to_right, int h) {
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < h; i += gridDim.x * blockDim.x) {
      if (to_left) to_left[i] = first[i];
      if (to_right) to_right[i] = last[i];
   }
}

__global__ void wide_recv_kernel(double *__restrict__ dst_left, double *__restrict__ dst_right, const double *from_left, const double *from_right,
                                 const unsigned long long *flag_from_left, const unsigned long long *flag_from_right, unsigned long long seq,
                                 unsigned int *err, long long timeout_cycles, int h) {
   __shared__ int ok;
   if (threadIdx.x == 0) {
      int good = *reinterpret_cast<volatile unsigned int *>(err) == 0;
      const long long t0 = clock64();
      for (int side = 0; side < 2 && good; ++side) {
         const unsigned long long *f = side == 0 ? flag_from_left : flag_from_right;
         if (!f) continue;
         while (*reinterpret_cast<const volatile unsigned long long *>(f) < seq) {
            if (clock64() - t0 > timeout_cycles) {
               atomicExch(err, 1u);
               good = 0;
               break;
            }
            __nanosleep(200);
         }
      }
      __threadfence_system();
      ok = good;
   }
   __syncthreads();
   if (!ok) return;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < h; i += gridDim.x * blockDim.x) {
      if (from_left) dst_left[i] = __ldcv(from_left + i);
      if (from_right) dst_right[i] = __ldcv(from_right + i);
   }
}

int fv_exchange_wide(Fv *fv, double *cell0, int64_t n, cudaStream_t st) {
   Halo &h = fv->halo;
   if (fv->d.nranks <= 1) return HRWENO_OK;
   if (!h.ready || !h.wide_off) return fail(HRWENO_ECOMM, "wide halo exchange: mailboxes not connected");
   constexpr int H = Halo::WIDE_H;
   const unsigned long long seq = ++h.wide_seq;
   const int slot = (int)(seq & 1); // two slots: a rank is at most one call ahead of its neighbours (its receive waits for their send)
   // I am the right neighbour of my left neighbour: my first cells go to its side 1 ("from right"), and vice versa
   double *to_left = h.peer[0] ? wide_slot(h.peer[0], h.wide_off, 1, slot) : nullptr;
   double *to_right = h.peer[1] ? wide_slot(h.peer[1], h.wide_off, 0, slot) : nullptr;
   auto *fl = h.peer[0] ? reinterpret_cast<unsigned long long *>(h.peer[0] + WIDE_FLAG_OFF + 128) : nullptr;
   auto *fr = h.peer[1] ? reinterpret_cast<unsigned long long *>(h.peer[1] + WIDE_FLAG_OFF) : nullptr;
   wide_send_kernel<<<32, 256, 0, st>>>(cell0, cell0 + n - H, to_left, to_right, H);
   halo_signal_kernel<<<1, 1, 0, st>>>(fl, fr, seq);
   const double *from_left = h.peer[0] ? wide_slot(h.mbox, h.wide_off, 0, slot) : nullptr;
   const double *from_right = h.peer[1] ? wide_slot(h.mbox, h.wide_off, 1, slot) : nullptr;
   auto *ffl = h.peer[0] ? reinterpret_cast<const unsigned long long *>(h.mbox + WIDE_FLAG_OFF) : nullptr;
   auto *ffr = h.peer[1] ? reinterpret_cast<const unsigned long long *>(h.mbox + WIDE_FLAG_OFF + 128) : nullptr;
   wide_recv_kernel<<<32, 256, 0, st>>>(cell0 - H, cell0 + n, from_left, from_right, ffl, ffr, seq, reinterpret_cast<unsigned int *>(h.mbox + ERR_OFF),
                                        60LL * 1900000000LL, H);
   fv->launches += 3;
   HRW_CUDA(cudaGetLastError());
   return HRWENO_OK;
}

// Build the in-kernel halo description for one fused stage (see HaloIO).  Returns with io->edge_first == 0 when the
// configuration is not eligible (then the caller uses fv_exchange after the stage).
int fv_halo_io(Fv *fv, const double *vin_cell0, const double *out_cell0, bool halo_out, HaloIO *io) {
   *io = HaloIO{};
   Halo &h = fv->halo;
   if (fv->d.nranks <= 1 || fv->d.ndim != 1 || fv->rows != 1) return HRWENO_OK;
   if (!h.ready) return fail(HRWENO_ECOMM, "slab decomposition needs hrweno_fv_export_halo / hrweno_fv_import_halo first");
   // the k boundary cells must lie inside one tile on either side (the CTA that sends them must have written them)
   int tile = 0, tpr = 0;
   fv_tiling_1d(fv, &tile, &tpr);
   if (fv->n0 - (int64_t)(tpr - 1) * tile < fv->d.k || fv->n0 < fv->d.k) return HRWENO_OK;
   const size_t hd = h.halo_doubles;
   io->edge_first = 1;
   io->err = reinterpret_cast<unsigned int *>(h.mbox + ERR_OFF);
   io->timeout_cycles = 60LL * 1900000000LL;
   if (h.pending_buf == vin_cell0 && h.pending_seq != 0) { // the input's ghost cells sit in my mailbox
      const int slot = (int)(h.pending_seq % Halo::NSLOTS);
      io->seq_in = h.pending_seq;
      if (h.peer[0]) {
         io->recv_left = reinterpret_cast<const double *>(slot_ptr(h.mbox, hd, 0, slot));
         io->rflag_left = reinterpret_cast<const unsigned long long *>(h.mbox + flag_off(0, slot));
      }
      if (h.peer[1]) {
         io->recv_right = reinterpret_cast<const double *>(slot_ptr(h.mbox, hd, 1, slot));
         io->rflag_right = reinterpret_cast<const unsigned long long *>(h.mbox + flag_off(1, slot));
      }
   }
   if (halo_out) {
      const unsigned long long seq = ++h.seq;
      const int slot = (int)(seq % Halo::NSLOTS);
      io->seq_out = seq;
      if (h.peer[0]) { // I am my left neighbour's right neighbour: its side 1 ("from right")
         io->send_left = reinterpret_cast<double *>(slot_ptr(h.peer[0], hd, 1, slot));
         io->sflag_left = reinterpret_cast<unsigned long long *>(h.peer[0] + flag_off(1, slot));
      }
      if (h.peer[1]) {
         io->send_right = reinterpret_cast<double *>(slot_ptr(h.peer[1], hd, 0, slot));
         io->sflag_right = reinterpret_cast<unsigned long long *>(h.peer[1] + flag_off(0, slot));
      }
      h.pending_buf = out_cell0;
      h.pending_seq = seq;
   } else if (h.pending_buf == out_cell0) {
      h.pending_buf = nullptr; // overwritten without a send (its ghost cells are not needed)
   }
   return HRWENO_OK;
}

} // namespace hrw
