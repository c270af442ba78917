// p2p.cu — the two collectives of the Krylov hot path written over NVLink / NVSwitch PEER MEMORY instead of NCCL:
//
//   * the all-reduce behind every dot / norm (1..8 doubles): one single-CTA kernel stores this rank's raw sums
//     into every peer's window, raises a flag, waits for the peers' flags, adds the contributions in RANK ORDER
//     (bit-identical on every rank) and derives alpha / omega / beta — all-reduce + scalar post-processing in one
//     launch instead of ncclAllReduce + finalize_kernel.
//   * the x exchange before a row-partitioned (#>): one kernel pushes the pieces of the local x slice the plan
//     (dist.py: plan_exchange) says the peers need straight into THEIR gathered-x buffers with 16-byte NVLink
//     stores, the last CTA raises this rank's flag on every peer and waits for theirs.  No staging copy, no
//     rendezvous protocol, works for halos (Laplacian: 32 KB per neighbour) and dense plans (cfg 2: every block).
//
// Windows are plain cudaMalloc allocations exported with cudaIpcGetMemHandle; the 64-byte handles travel through
// the host plumbing (torch.distributed all_gather_object in dist.py) and are opened with cudaIpcOpenMemHandle.
// Both protocols are double-buffered by sequence parity.  Why that is enough: rank A can only issue collective
// s+2 after it completed s+1, which needed every peer's contribution to s+1, which a peer issues (stream order)
// only after its own kernel for s has finished reading buffer s&1 — so nobody overwrites a buffer still in use.
// Every wait has a cycle-count timeout that raises a device error flag (reported as SLA_ERR_COMM at the next
// synchronisation) instead of hanging the GPU.
//
// Mode 2 of the x exchange (SLA_P2P_X=2; WRITTEN BUT NOT YET RUN ON HARDWARE — round-2 experiment): for dense plans
// with equal blocks the blocks travel on the COPY ENGINES in a staggered order (rank r sends to r+1, r+2, ... on the
// comm stream, one flag per destination after each copy) and the (#>) runs one column panel per source rank in ARRIVAL
// order (own block, then r-1, r-2, ...), each panel kernel preceded by a one-warp wait on that source's flag — so the
// transfer of block k+1 hides behind the kernel of panel k.  Row sums are then folded in rotated column order:
// within the fp64 bound of SURVEY.md §8(d), no longer bit-identical to the single-GPU result.
//
// Enabling is a COLLECTIVE decision taken by the host (every rank exported and attached successfully and
// SLA_P2P != 0); otherwise the NCCL path of dist.cu runs unchanged.
#include "common.cuh"

#include <new>

#define P2P_FLAG_BYTES 256                       // flags live in the first 256 bytes of a window
#define P2P_MAX_NV 32                            // doubles per all-reduce (one Hessenberg column chunk of the Arnoldi dots)
#define P2P_AR_THREADS 128
#define P2P_AR_WINDOW_BYTES (P2P_FLAG_BYTES + 2 * SLA_MAX_WORLD * P2P_MAX_NV * 8)   // [2][W] u64 flags | [2][W][P2P_MAX_NV] doubles
#define P2P_ITEM_LEN 4096                        // doubles per push work item (one CTA)
#define P2P_PUSH_THREADS 256
#define P2P_TIMEOUT_CYCLES 60000000000LL         // ~30 s at 1.9 GHz

struct p2p_item { int peer; int len; int64_t goff; int64_t src; };

struct sla_p2p {                                 // per context: the all-reduce window
  int enabled;
  char* win;
  char* peer[SLA_MAX_WORLD];
  char** d_peer;
  unsigned long long seq;
  int* d_err;                                    // device flag: a wait timed out
  int* h_err;                                    // pinned mirror
};

struct sla_xwin {                                // per distributed matrix: [256 B flags][x buffer 0][x buffer 1]
  int enabled;
  int mode;                                      // 1: push kernel + wait for every peer; 2: copy engines, arrival-order panels
  char* win;
  size_t buf_bytes;
  char* peer[SLA_MAX_WORLD];
  char** d_peer;
  p2p_item* d_items;
  int nitems;
  unsigned int* d_ticket;
  unsigned long long seq;
};

namespace {

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

// ---- all-reduce + scalar post-processing -------------------------------------------------------------------
// window: flags[b][r] at byte 8 * (b * SLA_MAX_WORLD + r), values[b][r][k] at byte 256 + 8 * ((b * SLA_MAX_WORLD + r) * P2P_MAX_NV + k)
__global__ void __launch_bounds__(P2P_AR_THREADS)
p2p_allreduce_kernel(char* const* __restrict__ peer, int rank, int world, unsigned long long seq, int nv, int src,
                     int fin, int dst, double* scal, int* err) {
  __shared__ double sum[P2P_MAX_NV];
  const int t = threadIdx.x;
  const int b = (int)(seq & 1ull);
  const int slot = b * SLA_MAX_WORLD + rank;
  for (int q = t; q < world * nv; q += P2P_AR_THREADS) {          // (peer, value) pairs
    const int p = q / nv, k = q - p * nv;
    st_relaxed_sys_f64(reinterpret_cast<double*>(peer[p] + P2P_FLAG_BYTES) + (size_t)slot * P2P_MAX_NV + k, scal[src + k]);
  }
  __threadfence_system();
  __syncthreads();
  char* mine = peer[rank];
  if (t < world) {
    __threadfence_system();                                       // cumulative over the CTA's stores observed through the barrier
    st_release_sys_u64(reinterpret_cast<unsigned long long*>(peer[t]) + slot, seq);
    wait_flag(reinterpret_cast<const unsigned long long*>(mine) + b * SLA_MAX_WORLD + t, seq, err);
  }
  __syncthreads();
  if (t < nv) {
    const double* vals = reinterpret_cast<const double*>(mine + P2P_FLAG_BYTES) + (size_t)b * SLA_MAX_WORLD * P2P_MAX_NV;
    double a = 0.0;
    for (int r = 0; r < world; ++r) a += ld_relaxed_sys_f64(vals + (size_t)r * P2P_MAX_NV + t);   // rank order: the same bits on every rank
    sum[t] = a;
  }
  __syncthreads();
  if (t == 0) finalize_scalars(fin, dst, scal, sum, nv);
}

// ---- x exchange ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(P2P_PUSH_THREADS)
p2p_push_kernel(const p2p_item* __restrict__ items, int nitems, char* const* __restrict__ peer, size_t buf_off,
                const double* __restrict__ x_local, int rank, int world, unsigned long long seq, unsigned int* ticket, int* err) {
  if ((int)blockIdx.x < nitems) {
    const p2p_item it = items[blockIdx.x];
    double* dst = reinterpret_cast<double*>(peer[it.peer] + buf_off) + it.goff;
    const double* src = x_local + it.src;
    if ((((uintptr_t)dst | (uintptr_t)src) & 15u) == 0) {
      const int n2 = it.len >> 1;
      const double2* s2 = reinterpret_cast<const double2*>(src);
      double2* d2 = reinterpret_cast<double2*>(dst);
      for (int i = threadIdx.x; i < n2; i += P2P_PUSH_THREADS) d2[i] = s2[i];
      if ((it.len & 1) && threadIdx.x == 0) dst[it.len - 1] = src[it.len - 1];
    } else {
      for (int i = threadIdx.x; i < it.len; i += P2P_PUSH_THREADS) dst[i] = src[i];
    }
  }
  // publish: every thread fences its own peer stores, the last CTA to arrive raises the flags
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence_system();          // cumulative: orders the CTA's peer stores (observed through the barrier) before the ticket
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) *ticket = 0u;
  __threadfence_system();
  const int t = threadIdx.x;
  if (t < world && t != rank) st_release_sys_u64(reinterpret_cast<unsigned long long*>(peer[t]) + rank, seq);
  if (t < world && t != rank) wait_flag(reinterpret_cast<const unsigned long long*>(peer[rank]) + t, seq, err);
}

// mode 2: the flag that follows a copy-engine block on the comm stream, and the per-source wait before a panel kernel
__global__ void p2p_flag_kernel(unsigned long long* dst, unsigned long long seq) {
  __threadfence_system();
  st_release_sys_u64(dst, seq);
}
__global__ void __launch_bounds__(32)
p2p_wait_kernel(const unsigned long long* flags, unsigned int mask, unsigned long long seq, int* err) {
  const int t = threadIdx.x;
  if (t < SLA_MAX_WORLD && ((mask >> t) & 1u)) wait_flag(flags + t, seq, err);
}

sla_status open_peers(sla_ctx* c, const void* handles, char* own, char** peer) {
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank) { peer[p] = own; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)p * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      for (int q = 0; q < p; ++q) if (q != c->rank && peer[q]) { cudaIpcCloseMemHandle(peer[q]); peer[q] = nullptr; }
      snprintf(c->err, sizeof(c->err), "p2p: cudaIpcOpenMemHandle failed for rank %d (%s)", p, cudaGetErrorString(e));
      return SLA_ERR_COMM;
    }
    peer[p] = (char*)ptr;
  }
  return SLA_OK;
}

void close_peers(sla_ctx* c, char** peer) {
  for (int p = 0; p < c->world; ++p) {
    if (p != c->rank && peer[p]) cudaIpcCloseMemHandle(peer[p]);
    peer[p] = nullptr;
  }
}

}  // namespace

// ---- context window ------------------------------------------------------------------------------------------

extern "C" sla_status sla_p2p_export(sla_ctx* c, void* handle64) {
  if (!c || !handle64) return SLA_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  memset(handle64, 0, 64);
  if (c->world < 2 || c->world > SLA_MAX_WORLD) return sla_fail(c, SLA_ERR_INVALID, "p2p: needs 2..16 ranks");
  if (!c->p2p) {
    sla_p2p* P = new (std::nothrow) sla_p2p();
    if (!P) return sla_fail(c, SLA_ERR_ALLOC, "p2p alloc");
    memset(P, 0, sizeof(*P));
    c->p2p = P;
    SLA_CUDA(c, cudaMalloc(&P->win, P2P_AR_WINDOW_BYTES));
    SLA_CUDA(c, cudaMemsetAsync(P->win, 0, P2P_AR_WINDOW_BYTES, c->stream));
    SLA_CUDA(c, cudaMalloc(&P->d_peer, sizeof(char*) * SLA_MAX_WORLD));
    SLA_CUDA(c, cudaMalloc(&P->d_err, sizeof(int)));
    SLA_CUDA(c, cudaMemsetAsync(P->d_err, 0, sizeof(int), c->stream));
    SLA_CUDA(c, cudaMallocHost(&P->h_err, sizeof(int)));
    *P->h_err = 0;
    SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  cudaIpcMemHandle_t h;
  SLA_CUDA(c, cudaIpcGetMemHandle(&h, c->p2p->win));
  memcpy(handle64, &h, sizeof(h));
  return SLA_OK;
}

// handles: world x 64 bytes in rank order (own entry ignored)
extern "C" sla_status sla_p2p_attach(sla_ctx* c, const void* handles) {
  if (!c || !handles || !c->p2p) return SLA_ERR_INVALID;
  sla_p2p* P = c->p2p;
  SLA_TRY(open_peers(c, handles, P->win, P->peer));
  SLA_CUDA(c, cudaMemcpyAsync(P->d_peer, P->peer, sizeof(char*) * SLA_MAX_WORLD, cudaMemcpyHostToDevice, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  return SLA_OK;
}

// collective switch: on = 1 only when EVERY rank attached successfully
extern "C" sla_status sla_p2p_enable(sla_ctx* c, int on) {
  if (!c) return SLA_ERR_INVALID;
  if (!c->p2p) return on ? sla_fail(c, SLA_ERR_INVALID, "p2p: no window exported") : SLA_OK;
  if (on && !c->p2p->peer[c->rank]) return sla_fail(c, SLA_ERR_INVALID, "p2p: windows not attached");
  c->p2p->enabled = on ? 1 : 0;
  return SLA_OK;
}

extern "C" int sla_p2p_enabled(const sla_ctx* c) { return c && c->p2p && c->p2p->enabled; }

bool sla_p2p_active(const sla_ctx* c) { return c->p2p && c->p2p->enabled; }

// all-reduce of scal[src .. src+nv) over the ranks followed by the scalar post-processing `fin` (nv <= P2P_MAX_NV = 32)
sla_status sla_p2p_allreduce(sla_ctx* c, int nv, int src, int fin, int dst) {
  sla_p2p* P = c->p2p;
  P->seq++;
  p2p_allreduce_kernel<<<1, P2P_AR_THREADS, 0, c->stream>>>(P->d_peer, c->rank, c->world, P->seq, nv, src, fin, dst, c->scal, P->d_err);
  SLA_LAUNCH_CHECK(c);
  return SLA_OK;
}

// reports a timed-out wait (called at synchronisation points)
sla_status sla_p2p_check(sla_ctx* c) {
  if (!c->p2p) return SLA_OK;
  sla_p2p* P = c->p2p;
  SLA_CUDA(c, cudaMemcpyAsync(P->h_err, P->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  if (*P->h_err) return sla_fail(c, SLA_ERR_COMM, "p2p: a peer did not arrive within the time-out (rank lost or collective calls out of order)");
  return SLA_OK;
}

void sla_p2p_free(sla_ctx* c) {
  sla_p2p* P = c->p2p;
  if (!P) return;
  close_peers(c, P->peer);
  cudaFree(P->win); cudaFree(P->d_peer); cudaFree(P->d_err); cudaFreeHost(P->h_err);
  delete P;
  c->p2p = nullptr;
}

// ---- matrix windows --------------------------------------------------------------------------------------------

extern "C" sla_status sla_csr_p2p_export(sla_ctx* c, sla_csr* A, void* handle64) {
  if (!c || !A || !handle64) return SLA_ERR_INVALID;
  memset(handle64, 0, 64);
  if (!A->dist) return sla_fail(c, SLA_ERR_INVALID, "p2p: the matrix has no exchange plan (sla_csr_set_dist first)");
  if (!c->p2p) return sla_fail(c, SLA_ERR_INVALID, "p2p: the context window was not exported");
  sla_dist_info* d = A->dist;
  if (!d->xwin) {
    sla_xwin* X = new (std::nothrow) sla_xwin();
    if (!X) return sla_fail(c, SLA_ERR_ALLOC, "p2p alloc");
    memset(X, 0, sizeof(*X));
    d->xwin = X;
    X->buf_bytes = (sizeof(double) * (size_t)(A->n + 2) + 255) & ~(size_t)255;
    const size_t bytes = P2P_FLAG_BYTES + 2 * X->buf_bytes;
    SLA_CUDA(c, cudaMalloc(&X->win, bytes));
    SLA_CUDA(c, cudaMemsetAsync(X->win, 0, bytes, c->stream));
    SLA_CUDA(c, cudaMalloc(&X->d_peer, sizeof(char*) * SLA_MAX_WORLD));
    SLA_CUDA(c, cudaMalloc(&X->d_ticket, sizeof(unsigned int)));
    SLA_CUDA(c, cudaMemsetAsync(X->d_ticket, 0, sizeof(unsigned int), c->stream));
    // work items: every send segment cut into P2P_ITEM_LEN pieces (cuts at even offsets keep the 16-byte path)
    int total = 0;
    for (int s = 0; s < d->nseg; ++s)
      if (d->seg[s].dir == 1) total += (int)((d->seg[s].count + P2P_ITEM_LEN - 1) / P2P_ITEM_LEN);
    p2p_item* items = new (std::nothrow) p2p_item[total > 0 ? total : 1];
    if (!items) return sla_fail(c, SLA_ERR_ALLOC, "p2p alloc");
    int k = 0;
    for (int s = 0; s < d->nseg; ++s) {
      const sla_xseg& g = d->seg[s];
      if (g.dir != 1) continue;
      for (int64_t o = 0; o < g.count; o += P2P_ITEM_LEN) {
        items[k].peer = g.peer;
        items[k].len = (int)(g.count - o < P2P_ITEM_LEN ? g.count - o : P2P_ITEM_LEN);
        items[k].goff = g.goff + o;
        items[k].src = g.goff + o - d->row0;
        ++k;
      }
    }
    X->nitems = total;
    cudaError_t e = cudaMalloc(&X->d_items, sizeof(p2p_item) * (size_t)(total > 0 ? total : 1));
    if (e == cudaSuccess && total > 0) e = cudaMemcpyAsync(X->d_items, items, sizeof(p2p_item) * (size_t)total, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    delete[] items;
    SLA_CUDA(c, e);
  }
  cudaIpcMemHandle_t h;
  SLA_CUDA(c, cudaIpcGetMemHandle(&h, d->xwin->win));
  memcpy(handle64, &h, sizeof(h));
  return SLA_OK;
}

extern "C" sla_status sla_csr_p2p_attach(sla_ctx* c, sla_csr* A, const void* handles) {
  if (!c || !A || !handles || !A->dist || !A->dist->xwin) return SLA_ERR_INVALID;
  sla_xwin* X = A->dist->xwin;
  SLA_TRY(open_peers(c, handles, X->win, X->peer));
  SLA_CUDA(c, cudaMemcpyAsync(X->d_peer, X->peer, sizeof(char*) * SLA_MAX_WORLD, cudaMemcpyHostToDevice, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  return SLA_OK;
}

extern "C" sla_status sla_csr_p2p_enable(sla_ctx* c, sla_csr* A, int on) {
  if (!c || !A || !A->dist) return SLA_ERR_INVALID;
  sla_dist_info* d = A->dist;
  if (!d->xwin) return on ? sla_fail(c, SLA_ERR_INVALID, "p2p: no matrix window exported") : SLA_OK;
  if (on && !d->xwin->peer[c->rank]) return sla_fail(c, SLA_ERR_INVALID, "p2p: matrix windows not attached");
  if (on && !sla_p2p_active(c)) return sla_fail(c, SLA_ERR_INVALID, "p2p: the context-level switch is off");
  int mode = on ? 1 : 0;
  if (on == 2) {
    // arrival-order mode needs one column panel per source rank: equal blocks in rank order, panel width = block size.
    // The test only looks at global quantities, so every rank takes the same branch.
    const int W = c->world;
    const bool eligible = A->n % W == 0 && A->m == A->n / W && d->row0 == (int64_t)c->rank * A->m && A->m % 16 == 0 && A->m > 0 &&
                          W <= SLA_MAX_PANELS && c->comm_stream != nullptr;
    if (eligible) {
      SLA_TRY(sla_csr_force_panels(c, A, W));
      if (A->npanels == W && A->panel_width == A->m) mode = 2;
    }
  }
  d->xwin->enabled = on ? 1 : 0;
  d->xwin->mode = mode;
  if (on) {
    // the kernels now read the remote entries from the window: the private gathered-x buffer is not needed
    SLA_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d->xfull);
    d->xfull = reinterpret_cast<double*>(d->xwin->win + P2P_FLAG_BYTES);
    d->allgather = 0; d->pipelined = 0;
  }
  return SLA_OK;
}

bool sla_xwin_active(const sla_csr* A) { return A->dist && A->dist->xwin && A->dist->xwin->enabled; }
int sla_xwin_mode(const sla_csr* A) { return sla_xwin_active(A) ? A->dist->xwin->mode : 0; }
extern "C" int sla_csr_p2p_mode(const sla_csr* A) { return A ? sla_xwin_mode(A) : 0; }

// mode 2, step 1: once x is final, the local block goes to the peers on the copy engines (comm stream) in the order
// r+1, r+2, ..., each copy followed by this rank's flag on that peer; flips the double buffer.
sla_status sla_p2p_arrival_begin(sla_ctx* c, const sla_csr* A, const double* x_local) {
  sla_dist_info* d = A->dist;
  sla_xwin* X = d->xwin;
  const int W = c->world;
  X->seq++;
  const size_t off = P2P_FLAG_BYTES + (size_t)(X->seq & 1ull) * X->buf_bytes;
  SLA_CUDA(c, cudaEventRecord(c->ev_x0, c->stream));                   // x is final; my earlier panel kernels are done
  SLA_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_x0, 0));
  for (int k = 1; k < W; ++k) {
    const int q = (c->rank + k) % W;
    double* dst = reinterpret_cast<double*>(X->peer[q] + off) + d->row0;
    SLA_CUDA(c, cudaMemcpyAsync(dst, x_local, sizeof(double) * (size_t)A->m, cudaMemcpyDefault, c->comm_stream));
    p2p_flag_kernel<<<1, 1, 0, c->comm_stream>>>(reinterpret_cast<unsigned long long*>(X->peer[q]) + c->rank, X->seq);
    SLA_LAUNCH_CHECK(c);
  }
  SLA_CUDA(c, cudaEventRecord(c->ev_panel[0], c->comm_stream));       // "my outgoing copies have read x_local"
  d->xfull = reinterpret_cast<double*>(X->win + off);
  return SLA_OK;
}

// mode 2, last step (compute stream): whatever follows the (#>) may overwrite x_local, so it must wait until the copy
// engines have finished reading it (the panel kernels only wait for INCOMING blocks)
sla_status sla_p2p_arrival_end(sla_ctx* c) {
  SLA_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_panel[0], 0));
  return SLA_OK;
}

// mode 2, step 2 (compute stream): block of rank `src` has arrived in the current buffer
sla_status sla_p2p_arrival_wait(sla_ctx* c, const sla_csr* A, int src) {
  sla_xwin* X = A->dist->xwin;
  p2p_wait_kernel<<<1, 32, 0, c->stream>>>(reinterpret_cast<const unsigned long long*>(X->win), 1u << src, X->seq, c->p2p->d_err);
  SLA_LAUNCH_CHECK(c);
  return SLA_OK;
}

// pushes the planned pieces of x_local to the peers and waits for theirs; afterwards A->dist->xfull is the buffer
// the next kernel reads
sla_status sla_p2p_exchange_x(sla_ctx* c, const sla_csr* A, const double* x_local) {
  sla_dist_info* d = A->dist;
  sla_xwin* X = d->xwin;
  X->seq++;
  const size_t off = P2P_FLAG_BYTES + (size_t)(X->seq & 1ull) * X->buf_bytes;
  const int grid = X->nitems > 0 ? X->nitems : 1;
  p2p_push_kernel<<<grid, P2P_PUSH_THREADS, 0, c->stream>>>(X->d_items, X->nitems, X->d_peer, off, x_local, c->rank, c->world,
                                                          X->seq, X->d_ticket, c->p2p->d_err);
  SLA_LAUNCH_CHECK(c);
  d->xfull = reinterpret_cast<double*>(X->win + off);
  return SLA_OK;
}

// The window itself is NOT freed here: a peer may still hold a mapping of it (matrices are released by each
// rank's garbage collector at its own time).  It is parked in the context and released by sla_finalize.
void sla_xwin_free(sla_csr* A) {
  sla_dist_info* d = A->dist;
  if (!d || !d->xwin) return;
  sla_xwin* X = d->xwin;
  sla_ctx* c = A->ctx;
  if (c) cudaStreamSynchronize(c->stream);
  if (c && c->comm_stream) cudaStreamSynchronize(c->comm_stream);
  if (c) close_peers(c, X->peer);
  cudaFree(X->d_peer); cudaFree(X->d_items); cudaFree(X->d_ticket);
  if (X->enabled) d->xfull = nullptr;            // it pointed into the window
  if (c && c->n_parked < SLA_MAX_PARKED) c->parked[c->n_parked++] = X->win;   // else: leaked until process exit
  delete X;
  d->xwin = nullptr;
}
