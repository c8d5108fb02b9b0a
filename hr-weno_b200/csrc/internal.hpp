// internal.hpp -- C++ objects behind the opaque C handles.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace hrw {

// stage combinations fused behind the right-hand side L(v)  (tvdode.f90:141,149-167,257)
enum Combine {
   C_RHS = 0,       // out = L(v)                                          example rhs
   C_EULER = 1,     // out = v + dt*L(v)                                   RK1, first stage of RK2/RK3
   C_RK2_FINAL = 2, // out = ((a + v) + dt*L(v))/2          a = u^n        tvdode.f90:152
   C_RK3_S2 = 3,    // out = ((3a + v) + dt*L(v))/4         a = u^n        tvdode.f90:165
   C_RK3_S3 = 4,    // out = ((a + 2v) + (2dt)*L(v))/3      a = u^n        tvdode.f90:167
   C_MS = 5         // out = ((((25v) + (50dt)L(v)) + 7a) + (10dt)b)/32, out2 = L(v);  a = u^{n-4}, b = L^{n-4}   :257
};

struct StageArgs {
   const double *vin;  // stage input  (padded layout, pointer at cell 0 of row 0)
   const double *a;    // pointwise operand (padded layout) or nullptr
   const double *b;    // second pointwise operand or nullptr
   double *out;        // result, pointer at cell 0 of row 0
   double *out2;       // C_MS: L(v) (padded layout)
   int64_t ld_out;     // row pitch of out (padded pitch, or n for a caller's dense array)
   int out_dense;      // 1: out is a caller's dense array: no ghost writes, no alignment assumptions
   double c0, c1;      // dt-like coefficients: (dt) | (2dt) | (50dt, 10dt)
   int tile_begin = 0, tile_end = 0; // 1D only: restrict the launch to a range of tiles (0,0 = the whole state)
   double t = 0.0;     // time of this right-hand-side evaluation (tvdode.f90:162,164,166,256): read by t-dependent fluxes only
};

// Halo traffic fused into the 1D stage kernel (one slab per GPU, rows == 1).  The CTA that computed the first / last
// k cells of the slab stores them into the neighbour's mailbox over NVLink and publishes seq_out; the CTA that owns an
// edge tile of the NEXT stage waits for seq_in and patches the ghost cells of its shared-memory tile from its own
// mailbox.  Edge tiles are walked first, so the transfer overlaps the bulk of the stage.  nullptr = side not used.
struct HaloIO {
   double *send_left = nullptr, *send_right = nullptr;                      // peer mailbox slots for this stage's output
   unsigned long long *sflag_left = nullptr, *sflag_right = nullptr;        // peer flags to publish
   const double *recv_left = nullptr, *recv_right = nullptr;                // my mailbox slots with the input's ghost cells
   const unsigned long long *rflag_left = nullptr, *rflag_right = nullptr;  // my flags to wait on
   unsigned long long seq_out = 0, seq_in = 0;
   unsigned int *err = nullptr;
   long long timeout_cycles = 0;
   int edge_first = 0; // walk tile 0 and the last tile first
};

// Halo exchange between slab neighbours over NVLink peer memory (CUDA IPC), one process per GPU.
// Every rank owns a mailbox; its neighbours store their boundary cells straight into it (peer stores)
// and then publish a sequence number; the owner's receive kernel spins on that number and moves the
// cells into the ghost region of the state vector.  No host synchronisation, no NCCL call per stage.
struct Halo {
   static constexpr int NSLOTS = 4;
   static constexpr size_t HDR_BYTES = 2048; // flags[2][NSLOTS] on separate 128-B lines + error word + 2 wide-halo flags
   // wide halos (1D, one row): the WIDE_H cells next to each slab interface, exchanged ONCE per host-pointer integrate
   // call so that every slab can run its own time-skewed chunk pipeline without per-stage traffic (ode.cu)
   static constexpr int WIDE_H = 32768;
   size_t wide_off = 0;                      // byte offset of the wide slots [side][2][WIDE_H] inside the mailbox (0: none)
   unsigned long long wide_seq = 0;
   size_t halo_doubles = 0;                  // doubles per side per slot
   size_t bytes = 0;
   unsigned char *mbox = nullptr;            // my mailbox (device memory, cudaMalloc'ed so it can be IPC-exported)
   unsigned char *peer[2] = {nullptr, nullptr}; // left / right neighbour's mailbox mapped into this process
   unsigned long long seq = 0;               // number of exchanges done (identical on all ranks: SPMD)
   bool ready = false;
   bool local_peers = false;                 // peer[] are mailboxes of slabs in this process (mgpu.cu), not IPC mappings
   // which padded state has its slab-interface ghost cells in a mailbox slot (sequence number) instead of in place
   const double *pending_buf = nullptr;
   unsigned long long pending_seq = 0;
};

struct Fv {
   hrweno_fv_desc d{};
   Halo halo;
   int64_t n0 = 0, n1 = 1;      // cells along x1 (contiguous) and x2
   int64_t rows = 1;             // 1D: independent rows
   int64_t neq = 0;
   int64_t pitch = 0;            // padded row pitch (doubles)
   int64_t nrows_alloc = 0;      // rows (1D) or n1 + 2*PAD2 (2D)
   double *d_width[2] = {nullptr, nullptr};
   double *d_rwidth[2] = {nullptr, nullptr}; // 2D: refined reciprocals of the widths (exact_recip), same padding
   // 1D width dictionary: <= 256 distinct widths {w, refined reciprocal} + one byte per cell (fv1d.cuh)
   double2 *d_wtab = nullptr;
   unsigned char *d_widx = nullptr;
   bool width_dict = false;
   double rx = 0.0;              // GRID_LINEAR: (xmax-xmin)/global_n
   // general path (fvgen.cu, K7), switched on by hrweno_fv_set_xedges / hrweno_fv_set_flux_coef
   bool general = false;
   double *d_cnu[2] = {nullptr, nullptr};   // cnu(0:k-1,-1:k-1,1:n[a]) of weno(ncells,k,eps,xedges)  (weno.f90:41,100-112): pointer at the
                                            // table of cell 0; one ghost table on either side (the neighbour slab's edge cell, or a replica)
   double *d_cnu_base[2] = {nullptr, nullptr}; // their allocations
   double *d_fcoef[2] = {nullptr, nullptr}; // face coefficient along axis a, index 0..n[a] like edges(0:n)
   double *d_ccoef[2] = {nullptr, nullptr}; // cross coefficient for the faces of axis a, index = cell along the other axis
   hrweno_time_fn tfn = nullptr;            // time factor g(t) of the flux (hrweno_fv_set_flux_time_fn), evaluated on the host per stage
   void *tfn_ctx = nullptr;
   // scratch for hrweno_fv_rhs[_dev]
   double *d_scratch_in = nullptr, *d_scratch_out = nullptr;
   cudaStream_t stream = nullptr;
   std::mutex mtx;
   int64_t launches = 0;
   // TMA tensor maps of the padded 2D states, per (buffer, box shape) (tmap.cu)
   struct TmapEntry {
      const double *ptr;
      uint32_t box0, box1;
      alignas(64) CUtensorMap map;
   };
   std::vector<TmapEntry> tmaps;
   // geometry helpers
   size_t state_doubles() const; // doubles of one padded state vector
   double *cell0(double *base) const; // pointer at cell 0 of row 0 inside a padded allocation
   const double *cell0(const double *base) const;
   int alloc_state(double **out) const;
   ~Fv();
};

int fv_create(Fv **out, const hrweno_fv_desc *desc);
// dense caller array <-> padded state (fills ghost cells of physical boundaries)
int fv_pack(Fv *fv, const double *dense_dev, double *padded_cell0, cudaStream_t st);
int fv_unpack(Fv *fv, const double *padded_cell0, double *dense_dev, cudaStream_t st);
// one fused stage: reconstruct + face fluxes + divergence + combination
int fv_stage(Fv *fv, int combine, const StageArgs &args, cudaStream_t st);
// 1D tiling of the current configuration: cells per tile and tiles per row
void fv_tiling_1d(const Fv *fv, int *tile_cells, int *tiles_per_row);
int fv_max_wavespeed(Fv *fv, const double *v_dev, double *out_dev, cudaStream_t st);
// fvgen.cu: general fused stage (per-cell reconstruction tables and/or x-dependent flux coefficients), reference order
int fvgen_stage(Fv *fv, int combine, const StageArgs &args, cudaStream_t st);
int fv_set_xedges(Fv *fv, int axis, const double *xedges);
int fv_set_flux_coef(Fv *fv, int axis, const double *face, const double *cross);
int fv_set_flux_time_fn(Fv *fv, hrweno_time_fn g, void *ctx);
// weno.cu: weno_calc_cnu (weno.f90:221-297) on the host
void weno_calc_cnu_host(int64_t nc, int k, const double *xedges, std::vector<double> &cnu);
// fv1d_small.cu: whole integrate call of a small 1D problem (<= 1024 cells) in one single-CTA launch
bool fv_small_eligible(const Fv *fv);
int fv_small_integrate(Fv *fv, double *u_dev, int order, long long nsteps, double dt, cudaStream_t st);
struct Fv1dGeom;
// fv1d_inst.cu, compiled once per (k, mode): launches the 1D stage kernel specialised for (combine, flux kind, width kind)
int fv1d_launch(int k, int mode, int combine, int flux_kind, int width_kind, int half_tile, const Fv1dGeom &g, const StageArgs &a,
                cudaStream_t st);
// fill the slab-interface ghost cells of a padded state from the neighbouring ranks (no-op on one GPU)
int fv_exchange(Fv *fv, double *padded_cell0, cudaStream_t st);
// stage + halo: fused into the kernel where supported (1D, one row, slabs), else fv_stage followed by fv_exchange.
// `halo_out`: the result will be a stencil input (its ghost cells must reach the neighbours).
int fv_stage_halo(Fv *fv, int combine, const StageArgs &args, bool halo_out, cudaStream_t st);
int fv_halo_io(Fv *fv, const double *vin_cell0, const double *out_cell0, bool halo_out, HaloIO *io);
int fv_halo_export(Fv *fv, void *handle_out);
int fv_halo_import(Fv *fv, const void *left, const void *right);
int fv_halo_status(Fv *fv);
void fv_halo_free(Fv *fv);
// wide halo exchange of the padded state at cell0 (interior cells [0, n)): my first / last WIDE_H cells go to the
// neighbours, theirs arrive in cells [-WIDE_H, 0) and [n, n + WIDE_H) (the caller's buffer has room for them)
int fv_exchange_wide(Fv *fv, double *cell0, int64_t n, cudaStream_t st);
int fv_halo_prepare_local(Fv *fv);                    // allocate this slab's mailbox on the current device
int fv_halo_connect_local(Fv *fv, Fv *left, Fv *right); // same-process neighbours (peer access enabled by the caller)

struct Weno {
   int64_t ncells = 0;
   int k = 3;
   double eps = 1e-6;
   bool uniform = true;
   int mode = HRWENO_MODE_STRICT; // arithmetic of the uniform-table reconstruct kernel (hrweno_weno_set_mode)
   std::vector<double> cnu_host;
   double *d_cnu = nullptr;
   // cached device buffers for the host-pointer entry points
   double *d_buf = nullptr;
   size_t buf_doubles = 0;
   std::mutex mtx;
   ~Weno();
};

int weno_reconstruct_launch(const Weno *w, int64_t rows, const double *v, int64_t ldv, int64_t incv, double *vl,
                            double *vr, int64_t ldo, cudaStream_t st);

struct Ode {
   bool is_ms = false;
   bool fused = false;
   Fv *fv = nullptr;
   hrweno_rhs_fn fu = nullptr;
   void *ctx = nullptr;
   // the reference's integrand as it is: a HOST procedure on host arrays (tvdode.f90:50-57), reached through an
   // adapter that stages u and udot in pinned memory around every call
   hrweno_rhs_host_fn fu_host = nullptr;
   void *ctx_host = nullptr;
   double *h_u = nullptr, *h_udot = nullptr;
   int cb_status = HRWENO_OK;
   int64_t neq = 0;
   int order = 3;
   int64_t fevals = 0;
   int istate = 0;
   int64_t launches = 0;
   // device state (padded layout for the fused path, dense for the callback path)
   std::vector<double *> bufs;
   int ring = 0; // mstvd: index of the slot holding u^n inside the u ring
   bool attached = false; // the current state lives inside the integrator (hrweno_ode_attach): null u in integrate
   cudaStream_t stream = nullptr;
   // chunk pipeline of the host-pointer entry point (ode.cu: rk_integrate_pipelined)
   cudaStream_t s_in = nullptr, s_out = nullptr;
   std::vector<cudaEvent_t> ev_in, ev_fin;
   // slabs: the pipeline runs on an operator over the slab extended by the wide halos, with its own three states
   Fv *fv_ext = nullptr;
   std::vector<double *> ext_bufs;
   cudaEvent_t ev_edge = nullptr;
   ~Ode();
};

} // namespace hrw
