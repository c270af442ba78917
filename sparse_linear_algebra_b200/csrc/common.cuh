// common.cuh — internal structures and helpers of libsla_b200.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sla_b200.h"

#define SLA_NUM_SMS 148          // B200: 148 SMs (2 dies x 74)
#define SLA_SCAL_SLOTS 1024      // device scalar slots per ctx
#define SLA_MAX_KRYLOV 384       // max Arnoldi / GMRES basis size (H column lives in scal[S_HCOL..])
#define SLA_MAX_PARTIALS (1 << 20)
#define SLA_MAX_PANELS 64        // column panels per matrix (spmv.cu)
#define SLA_MAX_WORLD 16         // ranks of one NVSwitch domain the peer-memory collectives address (p2p.cu)
#define SLA_HALO_MAX_SEGS 8      // receive segments of an LL halo plan (p2p.cu mode 3)
#define SLA_MAX_PARKED 256       // retired peer-memory windows kept until sla_finalize (p2p.cu)

// device scalar slots (doubles living in ctx->scal)
enum {
  S_RHO = 0, S_D1, S_ALPHA, S_D3, S_D4, S_OMEGA, S_D5, S_BETA, S_RES2, S_TMP0, S_TMP1, S_TMP2, S_NRM,
  S_RR, S_PP, S_RHO_NEW, S_INVN, S_RAW = 48 /* .. + 8 : per-rank raw sums awaiting the all-reduce */, S_HCOL = 64 /* .. + SLA_MAX_KRYLOV + 1 : one Hessenberg column */,
  S_HCOL2 = 512 /* .. + SLA_MAX_KRYLOV + 1 : re-orthogonalisation correction (GMRES) */
};

struct sla_ctx {
  int device;
  int rank, world;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  double* scal;              // SLA_SCAL_SLOTS device doubles (alpha, omega, beta, dot results ...)
  double* partials;          // per-CTA partial sums for grid reductions (SLA_MAX_PARTIALS * 4 doubles)
  unsigned int* counter;     // "last block" tickets, one per reduction site
  double* h_scal;            // pinned host mirror for scalar read-back
  int64_t launches;
  uint64_t stamp;            // source of the vectors' write stamps: strictly increasing, so a stamp never repeats (krylov.cu rho cache)
  void* nccl;                // ncclComm_t when world > 1
  void* nccl_x;              // second communicator (ncclCommSplit) for the x exchange on comm_stream
  cudaStream_t comm_stream;  // the x exchange runs here so that it overlaps the column-panel kernels
  cudaEvent_t ev_x0;         // "x is ready / xfull may be overwritten" (compute -> comm)
  cudaEvent_t ev_panel[SLA_MAX_PANELS];   // "panel p of xfull has arrived" (comm -> compute)
  struct sla_vec *scratch_x, *scratch_y;   // device staging for the host-pointer entry points
  struct sla_vec* scratch_r;               // partial row sums of the panelised residual-norm SpMV
  int spmv_hints;            // bit0: matrix stream L2 evict_first, bit1: x gathers L2 evict_last (env SLA_SPMV_HINTS)
  cudaStream_t copy_stream;  // PCIe copies of the pipelined host-buffer (#>) (spmv.cu)
  cudaEvent_t ev_copy[SLA_MAX_PANELS + 16];
  int bsell_variant;         // launch shape of the sliced-ELL band kernel (spmv_bandsell.cuh; env SLA_BSELL_VARIANT / option "bsell_variant")
  int spmv_bulk;             // 1: the tile kernel stages its (col, val) tile with bulk copies instead of LDG (env SLA_SPMV_BULK / option "spmv_bulk")
  int spmv_tma;              // 0: LDG tile kernel; k > 0: TMA-staged persistent kernel with k CTAs per SM (env SLA_SPMV_TMA)
  const void* scal_owner;    // Krylov state whose recurrence scalars currently live in scal[]
  void* bfull; size_t bfull_bytes;   // gathered dense right operand of a row-partitioned (##) (spmm.cu / dist.cu)
  struct sla_p2p* p2p;       // peer-memory all-reduce window (p2p.cu); null / disabled: NCCL
  void* parked[SLA_MAX_PARKED]; int n_parked;   // exchange windows of freed matrices (peers may still map them)
  void* dense_cache; size_t dense_cache_bytes;   // the last freed dense block (sla_pool_alloc / sla_pool_free)
  int skip_exchange;         // diagnostic (sla_set_option "skip_exchange"): row-partitioned (#>) runs its kernels WITHOUT the x exchange (results invalid)
  char err[512];
};

struct sla_vec {
  sla_ctx* ctx;
  int64_t n;
  double* d;
  uint64_t version;          // stamp of the last write through the API, drawn from ctx->stamp (invalidates cached dots)
  bool owns;
};

// one column panel of a matrix: a CSR over all rows holding only the columns of the panel (spmv.cu)
struct sla_panel {
  int32_t* row_ptr; int32_t* col; double* val; int32_t* tile_row;
  int ntiles; int64_t nnz; int skew_a;
};

// row-partitioned (multi-GPU) matrices: which pieces of x travel before each (#>)   (dist.cu)
struct sla_xseg { int dir; int peer; int64_t goff; int64_t count; };   // dir 0 = receive, 1 = send
struct sla_dist_info {
  int64_t row0;              // global index of the first local row (= first entry of the local x slice)
  int nseg; sla_xseg* seg;
  double* xfull;             // n doubles; only the remote entries this rank references are kept current
  int allgather;             // the plan is a plain all-gather of equal slices (collective decision)
  int dense_equal;           // the host's collective decision as installed (allgather may be switched off by a transport choice)
  // dense plans are pipelined: the segments clipped to each column panel, exchanged panel by panel on
  // comm_stream while the kernels of the earlier panels run
  int pipelined;
  int pan_first[SLA_MAX_PANELS + 1];   // pseg[pan_first[p] .. pan_first[p+1]) belong to panel p
  sla_xseg* pseg;
  struct sla_xwin* xwin;     // peer-memory exchange window (p2p.cu); when enabled xfull points into it
  // LL halo plan (p2p.cu mode 3), installed by sla_csr_set_halo: compact index of every segment — for a receive segment the
  // offset of its first entry in THIS rank's halo buffer, for a send segment its offset in the DESTINATION's
  int64_t* seg_base;         // nseg entries, null: no halo plan
  int64_t halo_total;        // entries this rank receives per exchange
};

struct sla_csr {
  sla_ctx* ctx;
  int64_t m, n, nnz;
  int32_t* row_ptr;          // m + 1
  int32_t* col;              // nnz, padded to a multiple of the tile size
  double* val;               // nnz, padded
  int32_t* tile_row;         // ntiles + 1 : first row whose start offset lies in the tile
  int ntiles;
  sla_csr* T;                // cached transpose for (<#) / CGNE
  int is_diag;               // -1 unknown, 0 / 1
  int skew_a;                // shared-memory skew shift of the SpMV product buffer (spmv.cu)
  int hints;                 // cache-hint bits of the SpMV loads chosen by the plan (spmv.cu)
  int npanels;               // >= 2 when the column-panel copy exists
  int panel_width;           // columns per panel
  sla_panel* panels;         // host array of device pointers
  sla_dist_info* dist;       // non-null: this is the local row block of a distributed matrix (n = GLOBAL columns)
  void* band;                // sla_band_plan (spmv_band.cuh): the matrix re-sorted for the x-in-shared-memory kernel, null when not banded
  void* val_bf16;            // bf16 copy of val for the bf16 (##) path, built on first use (spmm.cu)
  int bsr_ready, bsr_nbr, bsr_nblk; int *bsr_row_ptr, *bsr_col; void* bsr_val;   // 16 x 16 bf16 block copy for the tcgen05 (##) path
  int chunk_ready, chunk_tile[8], chunk_row[8];   // row chunks of the last pass for the pipelined host (#>)
  // triangular solves (trisolve.cu): level schedules [0] forward / [1] backward, position of the diagonal in every row,
  // the polled partial solution and the chunk ticket
  void* tri[2]; int32_t* tri_diag; double* tri_w; unsigned int* tri_ticket;
};

// dimension a vector must have to be multiplied by A / to receive A's product, on this rank
static inline int64_t csr_xdim(const sla_csr* A) { return A->dist ? A->m : A->n; }

struct sla_dense {
  sla_ctx* ctx;
  size_t bytes;              // size of the allocation behind d
  int64_t rows, cols, ld;    // column-major (Krylov basis): ld = rows rounded up to 16 doubles; row-major (## operands): ld = cols
  int dtype, rowmajor;       // SLA_F64 / SLA_BF16 ; 1 = row-major block created by sla_dense_create
  double* d;
};

struct sla_krylov {
  sla_ctx* ctx;
  int kind;                  // SLA_BICGSTAB_ / SLA_CGS_ / SLA_CGNE_
  int64_t n;
  sla_vec *x, *r, *p, *u;    // state record
  sla_vec *t0, *t1, *t2;     // work vectors (aap, s, aas ...)
  // cached rho = r <.> r0hat from the previous step, valid while nobody touched r or r0hat
  bool rho_valid;
  const sla_vec* rho_r0hat;
  uint64_t rho_r0hat_version, rho_r_version;
};

// every write through the API re-stamps the vector with a value no vector of this context ever carried before
static inline void sla_touch(sla_vec* v) { if (v) v->version = ++v->ctx->stamp; }

static inline sla_status sla_fail(sla_ctx* c, sla_status s, const char* msg) {
  if (c) snprintf(c->err, sizeof(c->err), "%s", msg);
  return s;
}

#define SLA_CUDA(ctx, call)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), "CUDA error %s at %s:%d (%s)",       \
                        cudaGetErrorString(_e), __FILE__, __LINE__, #call);                  \
      return SLA_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

// The current device is per host thread and other code in the process (torch, another context) may change it: entry points that
// allocate or launch re-select the context's device first (a thread-local no-op when it is already current).
#define SLA_GUARD(ctx) do { if (ctx) cudaSetDevice((ctx)->device); } while (0)

#define SLA_TRY(call)                       \
  do {                                      \
    sla_status _s = (call);                 \
    if (_s != SLA_OK) return _s;            \
  } while (0)

#define SLA_LAUNCH_CHECK(ctx)               \
  do {                                      \
    (ctx)->launches++;                      \
    SLA_CUDA(ctx, cudaGetLastError());      \
  } while (0)

// ---- device helpers -----------------------------------------------------------------------

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of NV values per thread; result valid in thread 0.  smem: NV * 32 doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) smem[k * 32 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double t = lane < nwarp ? smem[k * 32 + lane] : 0.0;
      v[k] = warp_sum(t);
    }
  }
}

// Scalar post-processing run by the last CTA of a grid reduction, after the sums are final.
enum {
  FIN_STORE = 0,       // scal[dst + k] = sum_k
  FIN_BICG_ALPHA,      // d1 = sum0 ; alpha = rho / d1
  FIN_BICG_OMEGA,      // d3 = sum0, d4 = sum1 ; omega = d3 / d4
  FIN_BICG_BETA,       // d5 = sum0 ; beta = d5 / rho * alpha / omega ; rho = d5
  FIN_CGS_BETA,        // d = sum0 ; beta = d / rho ; rho = d
  FIN_CGNE_ALPHA,      // rr = sum0, pp = sum1 ; alpha = rr / pp
  FIN_CGNE_BETA,       // rr1 = sum0 ; beta = rr1 / rr ; rr = rr1
  FIN_NORM_INV,        // nrm = sqrt(sum0) ; invn = 1 / nrm
  FIN_DEFER = 0x100    // flag (multi-GPU): only store the raw sums; the host all-reduces, then finalize_kernel runs
};

__device__ __forceinline__ void finalize_scalars(int fin, int dst, double* scal, const double* sum, int nv) {
  switch (fin) {
    case FIN_STORE:
      for (int k = 0; k < nv; ++k) scal[dst + k] = sum[k];
      break;
    case FIN_BICG_ALPHA:                       // alphaj = (r <.> r0hat) / (aap <.> r0hat)   Sparse.hs:974
      scal[S_D1] = sum[0];
      scal[S_ALPHA] = scal[S_RHO] / sum[0];
      break;
    case FIN_BICG_OMEGA:                       // omegaj = (aasj <.> sj) / (aasj <.> aasj)   Sparse.hs:977
      scal[S_D3] = sum[0]; scal[S_D4] = sum[1];
      scal[S_OMEGA] = sum[0] / sum[1];
      break;
    case FIN_BICG_BETA:                        // betaj = (rj1 <.> r0hat)/(r <.> r0hat) * alphaj / omegaj   Sparse.hs:980
      scal[S_D5] = sum[0];
      scal[S_BETA] = sum[0] / scal[S_RHO] * scal[S_ALPHA] / scal[S_OMEGA];
      scal[S_RHO] = sum[0];
      break;
    case FIN_CGS_BETA:                         // betaj = (rj1 `dot` rhat) / (r `dot` rhat)   Sparse.hs:937
      scal[S_BETA] = sum[0] / scal[S_RHO];
      scal[S_RHO] = sum[0];
      break;
    case FIN_CGNE_ALPHA:                       // alphai = (r `dot` r) / (p `dot` p)   Sparse.hs:874
      scal[S_RR] = sum[0]; scal[S_PP] = sum[1];
      scal[S_ALPHA] = sum[0] / sum[1];
      break;
    case FIN_CGNE_BETA:                        // beta = (r1 `dot` r1) / (r `dot` r)   Sparse.hs:877
      scal[S_BETA] = sum[0] / scal[S_RR];
      scal[S_RR] = sum[0];
      break;
    case FIN_NORM_INV:                         // norm2 = sqrt (norm2Sq) ; recip   SpVector.hs:125-128, Class.hs:94-95
      scal[S_NRM] = sqrt(sum[0]);
      scal[S_INVN] = 1.0 / scal[S_NRM];
      break;
  }
}

// ---- peer-memory all-reduce, inlined into the kernel that ends a grid reduction (p2p.cu owns the windows) -------------
// Window of a context: flags[b][r] (u64) at byte 8 * (b * SLA_MAX_WORLD + r), values[b][r][k] at byte
// P2P_FLAG_BYTES + 8 * ((b * SLA_MAX_WORLD + r) * P2P_MAX_NV + k); b = sequence parity (double buffer).
#define P2P_FLAG_BYTES 256
#define P2P_MAX_NV 32                            // doubles per all-reduce (one chunk of Arnoldi dots)
#define P2P_TIMEOUT_CYCLES 60000000000LL         // ~30 s at 1.9 GHz

struct sla_p2p_args {          // world <= 1: the reduction is local (or completed by the host through NCCL)
  char* const* peer;           // device array: every rank's window as mapped here
  int* err;                    // device flag: a wait timed out
  unsigned long long seq;
  int rank, world;
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// spins until *p >= seq; false (and *err = 1) when the peer never shows up
__device__ __forceinline__ bool wait_flag(const unsigned long long* p, unsigned long long seq, int* err) {
  const long long t0 = clock64();
  while (ld_acquire_sys_u64(p) < seq) {
    if (clock64() - t0 > P2P_TIMEOUT_CYCLES) { atomicExch(err, 1); return false; }
  }
  return true;
}

// All threads of ONE CTA (>= world threads): vals[0..nv) (shared memory) holds this rank's sums on entry and the sums over
// all ranks — added in RANK ORDER, hence bit-identical on every rank — on exit.
__device__ __forceinline__ void p2p_allreduce_block(const sla_p2p_args& a, double* vals, int nv) {
  const int t = threadIdx.x, nt = blockDim.x;
  const int b = (int)(a.seq & 1ull);
  const int slot = b * SLA_MAX_WORLD + a.rank;
  for (int q = t; q < a.world * nv; q += nt) {                      // (peer, value) pairs
    const int p = q / nv, k = q - p * nv;
    st_relaxed_sys_f64(reinterpret_cast<double*>(a.peer[p] + P2P_FLAG_BYTES) + (size_t)slot * P2P_MAX_NV + k, vals[k]);
  }
  __threadfence_system();
  __syncthreads();
  char* mine = a.peer[a.rank];
  if (t < a.world) {
    __threadfence_system();                                         // cumulative over the CTA's stores observed through the barrier
    st_release_sys_u64(reinterpret_cast<unsigned long long*>(a.peer[t]) + slot, a.seq);
    wait_flag(reinterpret_cast<const unsigned long long*>(mine) + b * SLA_MAX_WORLD + t, a.seq, a.err);
  }
  __syncthreads();
  double s = 0.0;
  if (t < nv) {
    const double* in = reinterpret_cast<const double*>(mine + P2P_FLAG_BYTES) + (size_t)b * SLA_MAX_WORLD * P2P_MAX_NV;
    for (int r = 0; r < a.world; ++r) s += ld_relaxed_sys_f64(in + (size_t)r * P2P_MAX_NV + t);
  }
  __syncthreads();
  if (t < nv) vals[t] = s;
  __syncthreads();
}

// Deterministic grid reduction: every CTA writes its NV partials, takes a ticket; the last CTA sums all
// partials in a fixed order (thread-strided sequential, then the block tree) and post-processes scalars.
// Must be called by all threads of every CTA.  `mine` holds this CTA's sums in thread 0.
template <int NV>
__device__ __forceinline__ void grid_reduce_finish(double (&mine)[NV], double* partials, unsigned int* counter,
                                                   double* scal, int fin, int dst, double* smem, const sla_p2p_args& pa) {
  __shared__ bool is_last;
  const unsigned int nblk = gridDim.x;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) partials[(size_t)k * nblk + blockIdx.x] = mine[k];
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == nblk - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    acc[k] = 0.0;
    const volatile double* p = partials + (size_t)k * nblk;
    for (unsigned int i = threadIdx.x; i < nblk; i += blockDim.x) acc[k] += p[i];
  }
  __syncthreads();
  block_sum<NV>(acc, smem);
  if (pa.world > 1) {
    // multi-GPU: the last CTA completes the reduction itself over NVLink peer memory — no separate all-reduce launch
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < NV; ++k) smem[k] = acc[k];
    }
    __syncthreads();
    p2p_allreduce_block(pa, smem, NV);
    if (threadIdx.x == 0) {
      finalize_scalars(fin & 0xff, dst, scal, smem, NV);
      *counter = 0u;
    }
    return;
  }
  if (threadIdx.x == 0) {
    if (fin & FIN_DEFER) {
      const int base = (fin & 0xff) == FIN_STORE ? dst : S_RAW;
#pragma unroll
      for (int k = 0; k < NV; ++k) scal[base + k] = acc[k];
    } else {
      finalize_scalars(fin, dst, scal, acc, NV);
    }
    *counter = 0u;
  }
}

// where the (#>) kernel of a row block finds the x entries of remote columns (spmv.cu)
struct SpmvDist {
  const double* xr;          // gathered-x buffer, indexed by GLOBAL column
  int col0, ncl;             // local column range [col0, col0 + ncl): read from the local slice instead
};

// Dense blocks (the Arnoldi basis is ~1 GB at cfg 4): the context keeps the last freed block and hands it to the next
// request it fits, so that back-to-back arnoldi / gmres calls do not make the driver unmap and re-map a gigabyte each time
// (measured: 3-5 ms of a 41 ms arnoldi(A, b, 30) went there).  All users are ordered on the context stream, so no
// synchronisation is needed for the hand-over.  (cudaMallocAsync was tried first: 38-372 ms per call, erratic.)
static inline cudaError_t sla_pool_alloc(sla_ctx* c, void** p, size_t bytes) {
  if (bytes < 16) bytes = 16;
  if (c->dense_cache && c->dense_cache_bytes >= bytes && c->dense_cache_bytes <= 2 * bytes + (1u << 20)) {
    *p = c->dense_cache; c->dense_cache = nullptr; c->dense_cache_bytes = 0;
    return cudaSuccess;
  }
  const cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaSuccess) return e;
  cudaGetLastError();
  if (c->dense_cache) { cudaStreamSynchronize(c->stream); cudaFree(c->dense_cache); c->dense_cache = nullptr; c->dense_cache_bytes = 0; }
  return cudaMalloc(p, bytes);
}
static inline void sla_pool_free(sla_ctx* c, void* p, size_t bytes) {
  if (!p) return;
  if (bytes >= (8u << 20) && bytes > c->dense_cache_bytes) {        // keep the larger block
    void* old = c->dense_cache;
    c->dense_cache = p; c->dense_cache_bytes = bytes;
    p = old;
    if (!p) return;
  }
  cudaStreamSynchronize(c->stream);
  cudaFree(p);
}

// internal entry points shared between translation units
sla_status sla_spmv_launch(sla_ctx* c, const sla_csr* A, const double* x, double* y, int epi,
                           const double* u0, const double* u1, int fin, int dst);
sla_status sla_csr_build_plan(sla_ctx* c, sla_csr* A);
sla_status sla_spmv_host_pipelined(sla_ctx* c, const sla_csr* A, const double* x_host, double* y_host, double* dx, double* dy);
void sla_csr_free_panels(sla_csr* A);
void sla_csr_free_band(sla_csr* A);
sla_status sla_csr_alloc(sla_ctx* c, int64_t m, int64_t n, int64_t nnz, sla_csr** out);
sla_status sla_vec_alloc(sla_ctx* c, int64_t n, sla_vec** out);
sla_status sla_read_scalars(sla_ctx* c, int first, int count, double* host_out);
// multi-GPU (dist.cu)
sla_status sla_dist_finish_reduction(sla_ctx* c, int nv, int fin, int dst);
sla_status sla_dist_exchange_x(sla_ctx* c, const sla_csr* A, const double* x_local);
sla_status sla_dist_exchange_panel(sla_ctx* c, const sla_csr* A, const double* x_local, int p);   // on comm_stream
sla_status sla_csr_force_panels(sla_ctx* c, sla_csr* A, int P);                                   // spmv.cu
sla_status sla_dist_allreduce_int(sla_ctx* c, int* d_val, int count);
sla_status sla_dist_allgather_i32(sla_ctx* c, const int* d_send, int* d_recv, int count);
sla_status sla_dist_group_begin(sla_ctx* c);
sla_status sla_dist_group_end(sla_ctx* c);
sla_status sla_dist_send(sla_ctx* c, const void* p, size_t count, int bytes8, int peer);
sla_status sla_dist_recv(sla_ctx* c, void* p, size_t count, int bytes8, int peer);
sla_status sla_dist_gather_rows(sla_ctx* c, const sla_csr* A, const void* local, void* full, int64_t k, int dtype);
void sla_csr_free_dist(sla_csr* A);
// peer-memory collectives (p2p.cu)
bool sla_p2p_active(const sla_ctx* c);
sla_status sla_p2p_allreduce(sla_ctx* c, int nv, int src, int fin, int dst);
sla_status sla_p2p_check(sla_ctx* c);
void sla_p2p_free(sla_ctx* c);
void* sla_p2p_window(sla_ctx* c);                                                                 // multi.cu: one process, several GPUs
sla_status sla_p2p_attach_direct(sla_ctx* c, void* const* wins);
bool sla_xwin_active(const sla_csr* A);
int sla_xwin_mode(const sla_csr* A);                                                             // 0 off, 1 push kernel, 2 arrival order, 3 LL halo, 4 copy-engine all-gather, 5 two-phase push
sla_status sla_p2p_arrival_begin(sla_ctx* c, const sla_csr* A, const double* x_local);
sla_status sla_p2p_arrival_wait(sla_ctx* c, const sla_csr* A, int src);
sla_status sla_p2p_arrival_end(sla_ctx* c);
sla_status sla_p2p_twophase_begin(sla_ctx* c, const sla_csr* A, const double* x_local);          // mode 5 (phased push; named after its first, two-phase form)
sla_status sla_p2p_twophase_wait(sla_ctx* c, const sla_csr* A, int phase);
// Rotated column panels of a row-partitioned matrix with equal blocks (spmv.cu): column j belongs to the block of predecessor
// k = ((own_end - 1 - j) mod n) / m of this rank (k = 0: own block); panel p holds the predecessors kb[p] <= k < kb[p + 1].
#define SLA_ROT_MAX 8
struct sla_rot_spec { long long n, own_end, m; int P; int kb[SLA_ROT_MAX + 1]; };
sla_status sla_csr_force_rot_panels(sla_ctx* c, sla_csr* A, const sla_rot_spec* spec);
extern "C" int sla_p2p_phase_schedule(int world, const char* spec, int* sizes);                  // p2p.cu
sla_status sla_p2p_exchange_x(sla_ctx* c, const sla_csr* A, const double* x_local);
void sla_xwin_free(sla_csr* A);
void sla_csr_free_bsr(sla_csr* A);
void sla_csr_free_tri(sla_csr* A);                                                                // trisolve.cu
// How the kernel that ends a grid reduction completes it across ranks: inline over peer memory (pa.world > 1), or raw sums
// + an all-reduce issued by the host (NCCL, or the stand-alone peer kernel) when `host` is set; single GPU: neither.
struct sla_red_plan { int fin; sla_p2p_args pa; bool host; };
sla_p2p_args sla_p2p_next(sla_ctx* c);                                                           // p2p.cu: args of the next inline all-reduce (bumps the sequence)
bool sla_p2p_inline(const sla_ctx* c);
static inline sla_red_plan sla_red_begin(sla_ctx* c, int fin, int nv) {
  sla_red_plan r;
  r.fin = fin; r.host = false;
  r.pa.peer = nullptr; r.pa.err = nullptr; r.pa.seq = 0; r.pa.rank = 0; r.pa.world = 1;
  if (c->world > 1) {
    if (nv <= P2P_MAX_NV && sla_p2p_inline(c)) r.pa = sla_p2p_next(c);
    else { r.fin = fin | FIN_DEFER; r.host = true; }
  }
  return r;
}
static inline sla_status sla_red_end(sla_ctx* c, const sla_red_plan& r, int nv, int fin, int dst) {
  return r.host ? sla_dist_finish_reduction(c, nv, fin, dst) : SLA_OK;
}

// SpMV epilogues
enum {
  EPI_NONE = 0,
  EPI_DOT1,      // sum0 = y . u0
  EPI_DOT2_YY,   // sum0 = y . u0 ; sum1 = y . y
  EPI_RESNORM    // sum0 = sum (y - u0)^2 ; y is NOT stored (true-residual check, Sparse.hs:1041)
};

#ifndef SLA_SPMV_TILE
#define SLA_SPMV_TILE 1024     // measured best with 128 threads (8 entries per thread, 16 CTAs/SM): profiles/r01_tile_sweep.txt
#endif
