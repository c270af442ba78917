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
#include <stdlib.h>
#include <vector>

#define P2P_AR_THREADS 128
#define P2P_AR_WINDOW_BYTES (P2P_FLAG_BYTES + 2 * SLA_MAX_WORLD * P2P_MAX_NV * 8)   // [2][W] u64 flags | [2][W][P2P_MAX_NV] doubles
#define P2P_ITEM_LEN 4096                        // doubles per push work item (one CTA)
#define P2P_PUSH_THREADS 256

struct p2p_item { int peer; int len; int64_t goff; int64_t src; };

struct sla_p2p {                                 // per context: the all-reduce window
  int enabled;
  int direct;                                    // peers are plain device pointers of THIS process (sla_init_multi), not IPC mappings
  char* win;
  char* peer[SLA_MAX_WORLD];
  char** d_peer;
  unsigned long long seq;
  int* d_err;                                    // device flag: a wait timed out
  int* h_err;                                    // pinned mirror
};

struct sla_xwin {                                // per distributed matrix: [256 B flags][x buffer 0][x buffer 1]
  int enabled;
  int mode;                                      // 1: push kernel + wait for every peer; 2: copy engines, arrival-order panels; 3: LL halo
  int ll;                                        // the window holds LL halo buffers (16-byte tagged entries) instead of gathered-x buffers
  char* win;
  size_t buf_bytes;
  char* peer[SLA_MAX_WORLD];
  char** d_peer;
  p2p_item* d_items;
  int nitems;                                    // send items; LL windows: followed by nrecv unpack items
  int nrecv;
  // mode 5 (phased push): the send items cut into ring-buffer pieces and ordered by phase, the destinations / sources of each phase
  p2p_item* d_items2; int nphase, n_phase[SLA_ROT_MAX]; unsigned dst_mask[SLA_ROT_MAX], src_mask[SLA_ROT_MAX]; unsigned int* d_ticket2; int push_ctas, bulk;
  unsigned int* d_ticket;
  unsigned long long seq;
};

namespace {

// ---- all-reduce + scalar post-processing -------------------------------------------------------------------
// stand-alone form (NCCL-free fallback when the inline path is switched off, SLA_P2P_INLINE=0): the raw sums sit in scal[src ..]
__global__ void __launch_bounds__(P2P_AR_THREADS)
p2p_allreduce_kernel(sla_p2p_args a, int nv, int src, int fin, int dst, double* scal) {
  __shared__ double sum[P2P_MAX_NV];
  if (threadIdx.x < nv) sum[threadIdx.x] = scal[src + threadIdx.x];
  __syncthreads();
  p2p_allreduce_block(a, sum, nv);
  if (threadIdx.x == 0) finalize_scalars(fin, dst, scal, sum, nv);
}

// ---- x exchange ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(P2P_PUSH_THREADS)
p2p_push_kernel(const p2p_item* __restrict__ items, int nitems, char* const* __restrict__ peer, size_t buf_off,
                const double* __restrict__ x_local, int rank, int world, unsigned long long seq, unsigned int* ticket, int* err) {
  // one CTA per item (the launch sizes the grid that way; the loop is the same code as the persistent two-phase kernel below)
  for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
    const p2p_item it = items[w];
    double* dst = reinterpret_cast<double*>(peer[it.peer] + buf_off) + it.goff;
    const double* src = x_local + it.src;
    if ((((uintptr_t)dst | (uintptr_t)src) & 15u) == 0) {
      const int n2 = it.len >> 1;
      const double2* s2 = reinterpret_cast<const double2*>(src);
      double2* d2 = reinterpret_cast<double2*>(dst);
      for (int base = 0; base < n2; base += P2P_PUSH_THREADS * 8) {
        double2 r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = base + u * P2P_PUSH_THREADS + (int)threadIdx.x;
          if (i < n2) r[u] = s2[i];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = base + u * P2P_PUSH_THREADS + (int)threadIdx.x;
          if (i < n2) d2[i] = r[u];
        }
      }
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

// mode 3, the LL halo exchange — ONE kernel per exchange, no fence, no flag, no rendezvous:
//   CTAs [0, nsend)        store the planned pieces of the local slice into the destinations' halo buffers as 16-byte entries
//                          {lo32, tag, hi32, tag}: each 8-byte half carries its own tag, so a reader that sees both tags has the
//                          whole double (8-byte stores are single-copy atomic over NVLink — NCCL's LL protocol, for fp64);
//   CTAs [nsend, nsend+nrecv) poll this rank's halo buffer until the entries of the CURRENT tag are there and unpack them
//                          into the gathered-x buffer the (#>) kernel reads (plain doubles, global column index).
// Every rank stores before it polls, so the kernel cannot deadlock; the cost on the critical path is one launch plus one
// NVLink store latency.  Buffers are double-buffered by tag parity; the plan must be symmetric (dist.py: halo_eligible), which
// bounds a sender to one exchange ahead of any rank that still reads the other buffer.
__global__ void __launch_bounds__(P2P_PUSH_THREADS)
p2p_halo_ll_kernel(const p2p_item* __restrict__ items, int nsend, int nrecv, char* const* __restrict__ peer, size_t buf_off,
                   const double* __restrict__ x_local, double* __restrict__ xfull, int rank, unsigned tag, int* err) {
  const int b = (int)blockIdx.x;
  if (b >= nsend + nrecv) return;
  const p2p_item it = items[b];
  if (b < nsend) {
    uint4* dst = reinterpret_cast<uint4*>(peer[it.peer] + buf_off) + it.goff;
    const double* src = x_local + it.src;
    for (int i = threadIdx.x; i < it.len; i += P2P_PUSH_THREADS) {
      const double v = src[i];
      const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
      asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
    }
    return;
  }
  const uint4* src = reinterpret_cast<const uint4*>(peer[rank] + buf_off) + it.src;
  double* dst = xfull + it.goff;
  const long long t0 = clock64();
  for (int i = threadIdx.x; i < it.len; i += P2P_PUSH_THREADS) {
    uint4 v;
    for (;;) {
      asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + i) : "memory");
      if (v.y == tag && v.w == tag) break;
      if (clock64() - t0 > P2P_TIMEOUT_CYCLES) { atomicExch(err, 1); break; }
    }
    dst[i] = __hiloint2double((int)v.z, (int)v.x);
  }
}

// mode 5: one PHASE of the two-phase exchange — the planned pieces for the peers of this phase, then (last CTA) this rank's
// flag on exactly those peers.  Nobody waits here: the consumer's wait kernel precedes the panel kernel that needs the data.
__global__ void __launch_bounds__(P2P_PUSH_THREADS)
p2p_push_phase_kernel(const p2p_item* __restrict__ items, int nitems, char* const* __restrict__ peer, size_t buf_off,
                      const double* __restrict__ x_local, int rank, int world, unsigned dst_mask, unsigned long long seq, unsigned int* ticket) {
  // persistent: a FEW CTAs walk the items, so that the (#>) panel kernel launched behind this one finds room on every SM
  for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
    const p2p_item it = items[w];
    double* dst = reinterpret_cast<double*>(peer[it.peer] + buf_off) + it.goff;
    const double* src = x_local + it.src;
    if ((((uintptr_t)dst | (uintptr_t)src) & 15u) == 0) {
      const int n2 = it.len >> 1;
      const double2* s2 = reinterpret_cast<const double2*>(src);
      double2* d2 = reinterpret_cast<double2*>(dst);
      for (int base = 0; base < n2; base += P2P_PUSH_THREADS * 8) {
        double2 r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = base + u * P2P_PUSH_THREADS + (int)threadIdx.x;
          if (i < n2) r[u] = s2[i];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = base + u * P2P_PUSH_THREADS + (int)threadIdx.x;
          if (i < n2) d2[i] = r[u];
        }
      }
      if ((it.len & 1) && threadIdx.x == 0) dst[it.len - 1] = src[it.len - 1];
    } else {
      for (int i = threadIdx.x; i < it.len; i += P2P_PUSH_THREADS) dst[i] = src[i];
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) *ticket = 0u;
  __threadfence_system();
  const int t = threadIdx.x;
  if (t < world && ((dst_mask >> t) & 1u)) st_release_sys_u64(reinterpret_cast<unsigned long long*>(peer[t]) + rank, seq);
}

// mode 5, the same phase driven by the TMA unit: ONE thread per CTA moves 16 KB pieces global -> shared -> peer global with bulk
// copies (cp.async.bulk, SASS UBLKCP) through a ring of P2P_BULK_STAGES buffers — two loads and two stores in flight per CTA, no
// LSU instructions and no registers taken from the (#>) kernel that shares the SM.  Every piece is 16-byte aligned (checked when
// the mode is enabled).
#define P2P_BULK_STAGES 4
#define P2P_BULK_BYTES 16384
#define P2P_BULK_LEN (P2P_BULK_BYTES / 8)
__device__ __forceinline__ uint32_t p2p_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32)
p2p_push_bulk_kernel(const p2p_item* __restrict__ items, int nitems, char* const* __restrict__ peer, size_t buf_off,
                     const double* __restrict__ x_local, int rank, int world, unsigned dst_mask, unsigned long long seq, unsigned int* ticket) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ __align__(8) uint64_t bar[P2P_BULK_STAGES];
  const int G = (int)gridDim.x, b = (int)blockIdx.x;
  if (threadIdx.x == 0) {
    for (int i = 0; i < P2P_BULK_STAGES; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(p2p_smem_u32(&bar[i])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int nj = b < nitems ? (nitems - b + G - 1) / G : 0;
    constexpr int L = P2P_BULK_STAGES - 2;                    // loads ahead of the store being issued
    auto load = [&](int j) {
      const p2p_item it = items[b + j * G];
      const int st = j % P2P_BULK_STAGES;
      const uint32_t bytes = (uint32_t)it.len * 8u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(p2p_smem_u32(&bar[st])), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(p2p_smem_u32(ring + (size_t)st * P2P_BULK_BYTES)), "l"(x_local + it.src), "r"(bytes), "r"(p2p_smem_u32(&bar[st])) : "memory");
    };
    for (int j = 0; j < L && j < nj; ++j) load(j);
    for (int j = 0; j < nj; ++j) {
      if (j + L < nj) {
        // the buffer of piece j + L was last read by the store of piece j - 2: at most ONE store (j - 1) may still be reading
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        load(j + L);
      }
      const int st = j % P2P_BULK_STAGES;
      const uint32_t parity = (uint32_t)(j / P2P_BULK_STAGES) & 1u;
      asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                   ::"r"(p2p_smem_u32(&bar[st])), "r"(parity) : "memory");
      const p2p_item it = items[b + j * G];
      double* dst = reinterpret_cast<double*>(peer[it.peer] + buf_off) + it.goff;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   ::"l"(dst), "r"(p2p_smem_u32(ring + (size_t)st * P2P_BULK_BYTES)), "r"((uint32_t)it.len * 8u) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every store of this CTA has been performed
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence_system();
  }
  __syncwarp();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncwarp();
  if (!last) return;
  if (threadIdx.x == 0) *ticket = 0u;
  __threadfence_system();
  const int t = threadIdx.x;
  if (t < world && ((dst_mask >> t) & 1u)) st_release_sys_u64(reinterpret_cast<unsigned long long*>(peer[t]) + rank, seq);
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

void close_peers(sla_ctx* c, char** peer, bool direct = false) {
  for (int p = 0; p < c->world; ++p) {
    if (p != c->rank && peer[p] && !direct) cudaIpcCloseMemHandle(peer[p]);
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

// One process driving several GPUs (multi.cu): the windows of the other ranks are ordinary device pointers of this process,
// reachable once peer access is enabled — no IPC handles.  wins: world pointers in rank order (own entry ignored).
void* sla_p2p_window(sla_ctx* c) { return c->p2p ? (void*)c->p2p->win : nullptr; }
sla_status sla_p2p_attach_direct(sla_ctx* c, void* const* wins) {
  if (!c || !wins || !c->p2p) return SLA_ERR_INVALID;
  sla_p2p* P = c->p2p;
  for (int p = 0; p < c->world; ++p) P->peer[p] = p == c->rank ? P->win : (char*)wins[p];
  P->direct = 1;
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
  p2p_allreduce_kernel<<<1, P2P_AR_THREADS, 0, c->stream>>>(sla_p2p_next(c), nv, src, fin, dst, c->scal);
  SLA_LAUNCH_CHECK(c);
  return SLA_OK;
}

// arguments of the next all-reduce over the context window (inline in the reducing kernel, or the stand-alone kernel):
// one sequence number per reduction, drawn in issue order — identical on every rank because the call sequence is
sla_p2p_args sla_p2p_next(sla_ctx* c) {
  sla_p2p* P = c->p2p;
  sla_p2p_args a;
  a.peer = P->d_peer; a.err = P->d_err; a.seq = ++P->seq; a.rank = c->rank; a.world = c->world;
  return a;
}

// the last CTA of every reducing kernel completes the all-reduce itself (default); SLA_P2P_INLINE=0: separate one-CTA kernel
bool sla_p2p_inline(const sla_ctx* c) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("SLA_P2P_INLINE"); on = e ? atoi(e) != 0 : 1; }
  return on && sla_p2p_active(c);
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
  close_peers(c, P->peer, P->direct != 0);
  cudaFree(P->win); cudaFree(P->d_peer); cudaFree(P->d_err); cudaFreeHost(P->h_err);
  delete P;
  c->p2p = nullptr;
}

// ---- matrix windows --------------------------------------------------------------------------------------------

// LL halo plan: base[s] belongs to segment s of the plan given to sla_csr_set_dist — for a receive segment the compact
// offset of its first entry in this rank's halo buffer, for a send segment the offset in the destination's buffer (the
// host derives both from the global table of column ranges, dist.py).  Call before sla_csr_p2p_export.
extern "C" sla_status sla_csr_set_halo(sla_ctx* c, sla_csr* A, int nseg, const int64_t* base, int64_t capacity) {
  if (!c || !A || !A->dist || nseg < 0 || (nseg > 0 && !base)) return SLA_ERR_INVALID;
  sla_dist_info* d = A->dist;
  if (nseg != d->nseg) return sla_fail(c, SLA_ERR_INVALID, "set_halo: one base per exchange segment expected");
  if (d->xwin) return sla_fail(c, SLA_ERR_INVALID, "set_halo: the exchange window already exists");
  int nrecv = 0;
  int64_t total = 0;
  for (int s = 0; s < nseg; ++s) {
    if (base[s] < 0) return sla_fail(c, SLA_ERR_INVALID, "set_halo: negative offset");
    if (d->seg[s].dir == 0) { ++nrecv; if (base[s] + d->seg[s].count > total) total = base[s] + d->seg[s].count; }
  }
  (void)nrecv;
  if (capacity < total) return sla_fail(c, SLA_ERR_INVALID, "set_halo: capacity is smaller than this rank's halo");
  total = capacity;                             // one buffer size for the whole job: senders address peers' buffers with their own stride
  if (total >= (int64_t)1 << 31) return sla_fail(c, SLA_ERR_INVALID, "set_halo: halo too large");
  delete[] d->seg_base;
  d->seg_base = new (std::nothrow) int64_t[nseg > 0 ? nseg : 1];
  if (!d->seg_base) return sla_fail(c, SLA_ERR_ALLOC, "set_halo alloc");
  for (int s = 0; s < nseg; ++s) d->seg_base[s] = base[s];
  d->halo_total = total;
  return SLA_OK;
}

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
    X->ll = d->seg_base != nullptr;
    X->buf_bytes = X->ll ? ((16 * (size_t)(d->halo_total + 1) + 255) & ~(size_t)255) : ((sizeof(double) * (size_t)(A->n + 2) + 255) & ~(size_t)255);
    const size_t bytes = P2P_FLAG_BYTES + 2 * X->buf_bytes;
    SLA_CUDA(c, cudaMalloc(&X->win, bytes));
    SLA_CUDA(c, cudaMemsetAsync(X->win, 0, bytes, c->stream));
    SLA_CUDA(c, cudaMalloc(&X->d_peer, sizeof(char*) * SLA_MAX_WORLD));
    SLA_CUDA(c, cudaMalloc(&X->d_ticket, sizeof(unsigned int)));
    SLA_CUDA(c, cudaMemsetAsync(X->d_ticket, 0, sizeof(unsigned int), c->stream));
    // work items: every send segment cut into P2P_ITEM_LEN pieces (cuts at even offsets keep the 16-byte path)
    int total = 0, total_recv = 0;
    for (int s = 0; s < d->nseg; ++s) {
      const int pieces = (int)((d->seg[s].count + P2P_ITEM_LEN - 1) / P2P_ITEM_LEN);
      if (d->seg[s].dir == 1) total += pieces; else if (X->ll) total_recv += pieces;
    }
    p2p_item* items = new (std::nothrow) p2p_item[total + total_recv > 0 ? total + total_recv : 1];
    if (!items) return sla_fail(c, SLA_ERR_ALLOC, "p2p alloc");
    int k = 0;
    for (int s = 0; s < d->nseg; ++s) {
      const sla_xseg& g = d->seg[s];
      if (g.dir != 1) continue;
      for (int64_t o = 0; o < g.count; o += P2P_ITEM_LEN) {
        items[k].peer = g.peer;
        items[k].len = (int)(g.count - o < P2P_ITEM_LEN ? g.count - o : P2P_ITEM_LEN);
        items[k].goff = X->ll ? d->seg_base[s] + o : g.goff + o;      // LL: compact index in the destination's halo buffer
        items[k].src = g.goff + o - d->row0;
        ++k;
      }
    }
    // LL windows: the unpack items follow — where a received piece sits in this rank's halo buffer and where it goes in xfull
    for (int s = 0; s < d->nseg && X->ll; ++s) {
      const sla_xseg& g = d->seg[s];
      if (g.dir != 0) continue;
      for (int64_t o = 0; o < g.count; o += P2P_ITEM_LEN) {
        items[k].peer = -1;
        items[k].len = (int)(g.count - o < P2P_ITEM_LEN ? g.count - o : P2P_ITEM_LEN);
        items[k].goff = g.goff + o;
        items[k].src = d->seg_base[s] + o;
        ++k;
      }
    }
    X->nitems = total;
    X->nrecv = total_recv;
    const int all = total + total_recv;
    cudaError_t e = cudaMalloc(&X->d_items, sizeof(p2p_item) * (size_t)(all > 0 ? all : 1));
    if (e == cudaSuccess && all > 0) e = cudaMemcpyAsync(X->d_items, items, sizeof(p2p_item) * (size_t)all, cudaMemcpyHostToDevice, c->stream);
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

// ---- diagnostic (scripts/prof_push_contention.py, ONE GPU) ------------------------------------------------------------------
// The mode-5 push kernels copying n doubles src -> dst inside this GPU's own memory from a high-priority side stream, so that the
// SM-side cost of a push running beside the (#>) panel kernels can be measured without a second GPU (no NVLink in the picture:
// what is left is the CTA slots, shared memory and HBM / L2 bandwidth the push takes from the panel kernels).
struct sla_debug_push {
  cudaStream_t side; cudaEvent_t ev0, ev1;
  p2p_item* d_items; char** d_peer; unsigned int* d_ticket; int nitems; const double* src;
};
extern "C" void sla_debug_push_free(sla_debug_push* h);
extern "C" sla_status sla_debug_push_create(sla_ctx* c, sla_vec* dstv, const sla_vec* srcv, sla_debug_push** out) {
  if (!c || !dstv || !srcv || !out || dstv->n != srcv->n || srcv->n < 2) return SLA_ERR_INVALID;
  SLA_GUARD(c);
  double* dst = dstv->d;
  const double* src = srcv->d;
  const int64_t n = srcv->n & ~(int64_t)1;
  sla_debug_push* h = new sla_debug_push();
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  std::vector<p2p_item> items;
  for (int64_t o = 0; o < n; o += P2P_BULK_LEN) {
    p2p_item it; it.peer = 0; it.len = (int)(n - o < P2P_BULK_LEN ? n - o : P2P_BULK_LEN); it.goff = o; it.src = o;
    items.push_back(it);
  }
  h->nitems = (int)items.size(); h->src = src;
  char* peer0 = reinterpret_cast<char*>(dst);
  cudaError_t e = cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev0, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev1, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_items, sizeof(p2p_item) * items.size());
  if (e == cudaSuccess) e = cudaMalloc(&h->d_peer, sizeof(char*));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_ticket, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_items, items.data(), sizeof(p2p_item) * items.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_peer, &peer0, sizeof(char*), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(h->d_ticket, 0, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(p2p_push_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2P_BULK_STAGES * P2P_BULK_BYTES);
  if (e != cudaSuccess) { cudaGetLastError(); sla_debug_push_free(h); return sla_fail(c, SLA_ERR_CUDA, "debug_push_create: CUDA error"); }
  *out = h;
  return SLA_OK;
}
// starts the copy on the side stream once everything already queued on the ctx stream is done; kind 0 = TMA bulk kernel, 1 = LSU kernel
extern "C" sla_status sla_debug_push_start(sla_ctx* c, sla_debug_push* h, int ctas, int kind) {
  if (!c || !h || ctas <= 0) return SLA_ERR_INVALID;
  SLA_GUARD(c);
  SLA_CUDA(c, cudaEventRecord(h->ev0, c->stream));
  SLA_CUDA(c, cudaStreamWaitEvent(h->side, h->ev0, 0));
  const int grid = ctas < h->nitems ? ctas : h->nitems;
  if (kind == 0)
    p2p_push_bulk_kernel<<<grid, 32, P2P_BULK_STAGES * P2P_BULK_BYTES, h->side>>>(h->d_items, h->nitems, h->d_peer, 0, h->src, 0, 1, 0u, 1ull, h->d_ticket);
  else
    p2p_push_phase_kernel<<<grid, P2P_PUSH_THREADS, 0, h->side>>>(h->d_items, h->nitems, h->d_peer, 0, h->src, 0, 1, 0u, 1ull, h->d_ticket);
  SLA_LAUNCH_CHECK(c);
  SLA_CUDA(c, cudaEventRecord(h->ev1, h->side));
  return SLA_OK;
}
// the ctx stream waits for the copy started last
extern "C" sla_status sla_debug_push_join(sla_ctx* c, sla_debug_push* h) {
  if (!c || !h) return SLA_ERR_INVALID;
  SLA_CUDA(c, cudaStreamWaitEvent(c->stream, h->ev1, 0));
  return SLA_OK;
}
extern "C" void sla_debug_push_free(sla_debug_push* h) {
  if (!h) return;
  if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  cudaFree(h->d_items); cudaFree(h->d_peer); cudaFree(h->d_ticket);
  delete h;
}

// The panel schedule of the phased exchange (mode 5): sizes[p] = how many column blocks panel p holds — panel 0 the own block,
// panel p >= 1 the next sizes[p] predecessors, whose blocks travel in phase p.  spec = "1,1,2"-style override (must start with
// 1 and add up to world), else 1, 1, 2, 4, ...: a phase is never larger than everything multiplied before it.  Pure host
// arithmetic (the Python planner holds the same rule: dist.phase_schedule); returns the number of panels.
extern "C" int sla_p2p_phase_schedule(int world, const char* spec, int* sizes) {
  if (world < 1 || !sizes) return 0;
  int ns = 0, tot = 0;
  const char* e = spec;
  while (e && *e && ns < SLA_ROT_MAX) {
    const int v = atoi(e);
    if (v <= 0) { ns = 0; break; }
    sizes[ns++] = v; tot += v;
    e = strchr(e, ',');
    if (e) ++e;
  }
  if (e && *e) ns = 0;                                      // more entries than panels
  if (ns >= 2 && tot == world && sizes[0] == 1) return ns;
  ns = 0; sizes[ns++] = 1;
  int left = world - 1, next = 1;
  while (left > 0) {
    const int take = (next < left && ns < SLA_ROT_MAX - 1) ? next : left;
    sizes[ns++] = take; left -= take;
    next = ns == 2 ? 2 : next * 2;
  }
  return ns;
}

extern "C" sla_status sla_csr_p2p_enable(sla_ctx* c, sla_csr* A, int on) {
  if (!c || !A || !A->dist) return SLA_ERR_INVALID;
  sla_dist_info* d = A->dist;
  if (!d->xwin) return on ? sla_fail(c, SLA_ERR_INVALID, "p2p: no matrix window exported") : SLA_OK;
  if (on && !d->xwin->peer[c->rank]) return sla_fail(c, SLA_ERR_INVALID, "p2p: matrix windows not attached");
  if (on && !sla_p2p_active(c)) return sla_fail(c, SLA_ERR_INVALID, "p2p: the context-level switch is off");
  if (on && d->xwin->ll && on != 3) return sla_fail(c, SLA_ERR_INVALID, "p2p: this window was built for the LL halo exchange (mode 3)");
  if (on == 3 && !d->xwin->ll) return sla_fail(c, SLA_ERR_INVALID, "p2p: mode 3 needs sla_csr_set_halo before sla_csr_p2p_export");
  int mode = on == 3 ? 3 : on ? 1 : 0;
  if ((on == 2 || on == 4) && d->dense_equal && c->comm_stream != nullptr) {
    // Copy-engine all-gather.  Arrival-order consumption (mode 2) needs one column panel per source rank — equal blocks in
    // rank order, panel width = block size — and pays one extra pass over y and row_ptr per panel, so it is chosen only when
    // x would not stay L2-resident anyway (8 n above the panel threshold: the single-GPU plan panelises such matrices too);
    // otherwise the blocks are waited for as a whole and the matrix keeps its own plan (mode 4).  The tests only look at
    // global quantities (n, world), so every rank takes the same branch.
    const int W = c->world;
    const bool eligible = A->n % W == 0 && A->m == A->n / W && d->row0 == (int64_t)c->rank * A->m && A->m % 16 == 0 && A->m > 0 &&
                          W <= SLA_MAX_PANELS;
    const bool big_x = (uint64_t)A->n * 8u > (56u << 20);
    mode = 4;
    if (on == 2 && eligible && (big_x || getenv("SLA_P2P_ARRIVAL_ALWAYS"))) {
      SLA_TRY(sla_csr_force_panels(c, A, W));
      if (A->npanels == W && A->panel_width == A->m) mode = 2;
    }
  }
  const bool was_in_window = d->xwin->enabled && !d->xwin->ll;
  if (!on && d->xwin->enabled && d->xwin->mode == 5) sla_csr_free_panels(A);        // the rotated panels belong to the two-phase exchange
  if (on == 5 && d->dense_equal && c->comm_stream != nullptr && !d->xwin->ll) {
    // Phased push (dense equal-block plans): the (#>) runs ROTATED column panels — the own block first, then the blocks of the
    // predecessors in growing groups — and the blocks of group p travel (TMA bulk copies issued by a few one-thread CTAs on a
    // high-priority side stream) under the kernels of the panels before it.  Balanced: in each phase every rank sends and
    // receives the same amount.  Rows are folded panel by panel: within the fp64 bound of SURVEY.md 8(d), like arrival order.
    const int W = c->world;
    sla_xwin* X = d->xwin;
    const bool eligible = W >= 2 && W <= 31 && A->n % W == 0 && A->m == A->n / W && d->row0 == (int64_t)c->rank * A->m && A->m % 16 == 0 && A->m > 0;
    const bool big_x = (uint64_t)A->n * 8u > (56u << 20);   // like mode 2: panels pay only when x would not stay L2-resident anyway
    mode = 1;
    if (eligible && (big_x || getenv("SLA_P2P_ARRIVAL_ALWAYS"))) {
      X->push_ctas = 64;                                    // measured at 4 ranks, cfg 2: 32 -> 0.491, 64 -> 0.482, 128 -> 0.502 ms per (#>)
      if (const char* e = getenv("SLA_P2P_PUSH_CTAS")) X->push_ctas = atoi(e) > 0 ? atoi(e) : 64;
      // Panel p = the blocks of predecessors kb[p] .. kb[p + 1] - 1 (panel 0 = own block, nothing to wait for); phase p brings
      // exactly those blocks.  Default: 1, 1, 2, 4, ... blocks — every phase is at most as large as everything computed before
      // it, and a block of columns takes longer to multiply than to receive (cfg 2: ~2x).  SLA_P2P_PANELS="1,1,2" overrides.
      sla_rot_spec rs;
      rs.n = A->n; rs.m = A->m; rs.own_end = d->row0 + A->m; rs.P = 0; rs.kb[0] = 0;
      {
        int sizes[SLA_ROT_MAX];
        rs.P = sla_p2p_phase_schedule(W, getenv("SLA_P2P_PANELS"), sizes);
        for (int i = 0; i < rs.P; ++i) rs.kb[i + 1] = rs.kb[i] + sizes[i];
      }
      X->nphase = rs.P;
      for (int ph = 0; ph < SLA_ROT_MAX; ++ph) X->dst_mask[ph] = X->src_mask[ph] = 0u;
      for (int ph = 1; ph < rs.P; ++ph)
        for (int k = rs.kb[ph]; k < rs.kb[ph + 1]; ++k) {
          X->dst_mask[ph] |= 1u << ((c->rank + k) % W);        // I am predecessor k of rank r + k
          X->src_mask[ph] |= 1u << ((c->rank - k + W) % W);
        }
      // items re-ordered by phase
      std::vector<p2p_item> all((size_t)(X->nitems > 0 ? X->nitems : 1)), ord;
      if (X->nitems > 0) SLA_CUDA(c, cudaMemcpy(all.data(), X->d_items, sizeof(p2p_item) * (size_t)X->nitems, cudaMemcpyDeviceToHost));
      bool aligned = true;
      for (int ph = 0; ph < rs.P; ++ph) {
        X->n_phase[ph] = 0;
        for (int i = 0; i < X->nitems; ++i) {
          if (!((X->dst_mask[ph] >> all[i].peer) & 1u)) continue;
          for (int o = 0; o < all[i].len; o += P2P_BULK_LEN) {      // pieces of one ring buffer
            p2p_item h = all[i];
            h.len = all[i].len - o < P2P_BULK_LEN ? all[i].len - o : P2P_BULK_LEN;
            h.goff += o; h.src += o;
            aligned = aligned && (h.len % 2 == 0) && (h.goff % 2 == 0) && (h.src % 2 == 0);
            ord.push_back(h); X->n_phase[ph]++;
          }
        }
      }
      X->bulk = aligned && !getenv("SLA_P2P_PUSH_LSU");
      if (X->bulk)
        SLA_CUDA(c, cudaFuncSetAttribute(p2p_push_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2P_BULK_STAGES * P2P_BULK_BYTES));
      if (!X->d_items2) SLA_CUDA(c, cudaMalloc(&X->d_items2, sizeof(p2p_item) * (size_t)(ord.size() > 0 ? ord.size() : 1)));
      if (!ord.empty()) SLA_CUDA(c, cudaMemcpy(X->d_items2, ord.data(), sizeof(p2p_item) * ord.size(), cudaMemcpyHostToDevice));
      if (!X->d_ticket2) { SLA_CUDA(c, cudaMalloc(&X->d_ticket2, sizeof(unsigned int) * SLA_ROT_MAX)); SLA_CUDA(c, cudaMemset(X->d_ticket2, 0, sizeof(unsigned int) * SLA_ROT_MAX)); }
      SLA_TRY(sla_csr_force_rot_panels(c, A, &rs));
      if (A->npanels == rs.P) mode = 5;
    }
  }
  d->xwin->enabled = on ? 1 : 0;
  d->xwin->mode = mode;
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  if (on && !d->xwin->ll) {
    // the kernels now read the remote entries from the window: the private gathered-x buffer is not needed
    if (!was_in_window) cudaFree(d->xfull);
    d->xfull = reinterpret_cast<double*>(d->xwin->win + P2P_FLAG_BYTES);
    d->allgather = 0; d->pipelined = 0;
  } else if (!on && was_in_window) {
    // back to NCCL: a private gathered-x buffer again, and the collective all-gather decision as installed
    d->xfull = nullptr;
    SLA_CUDA(c, cudaMalloc(&d->xfull, sizeof(double) * (size_t)((A->n + 2) & ~(int64_t)1)));
    SLA_CUDA(c, cudaMemsetAsync(d->xfull, 0, sizeof(double) * (size_t)A->n, c->stream));
    d->allgather = d->dense_equal;
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

// mode 5, step 1: both phases of the push on the comm stream (phase A first); flips the double buffer
sla_status sla_p2p_twophase_begin(sla_ctx* c, const sla_csr* A, const double* x_local) {
  sla_dist_info* d = A->dist;
  sla_xwin* X = d->xwin;
  X->seq++;
  const size_t off = P2P_FLAG_BYTES + (size_t)(X->seq & 1ull) * X->buf_bytes;
  SLA_CUDA(c, cudaEventRecord(c->ev_x0, c->stream));                   // x is final; my earlier panel kernels are done
  SLA_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_x0, 0));
  int first = 0;
  for (int ph = 0; ph < X->nphase; ++ph) {
    if (X->dst_mask[ph]) {
      int grid = X->n_phase[ph] > 0 ? X->n_phase[ph] : 1;
      if (grid > X->push_ctas) grid = X->push_ctas;
      if (X->bulk)
        p2p_push_bulk_kernel<<<grid, 32, P2P_BULK_STAGES * P2P_BULK_BYTES, c->comm_stream>>>(X->d_items2 + first, X->n_phase[ph], X->d_peer, off, x_local,
                                                                                          c->rank, c->world, X->dst_mask[ph], X->seq, X->d_ticket2 + ph);
      else
        p2p_push_phase_kernel<<<grid, P2P_PUSH_THREADS, 0, c->comm_stream>>>(X->d_items2 + first, X->n_phase[ph], X->d_peer, off, x_local, c->rank, c->world,
                                                                            X->dst_mask[ph], X->seq, X->d_ticket2 + ph);
      SLA_LAUNCH_CHECK(c);
    }
    first += X->n_phase[ph];
  }
  SLA_CUDA(c, cudaEventRecord(c->ev_panel[0], c->comm_stream));       // "my pushes have read x_local"
  d->xfull = reinterpret_cast<double*>(X->win + off);
  return SLA_OK;
}

// mode 5, step 2 (compute stream): the blocks of phase ph have arrived
sla_status sla_p2p_twophase_wait(sla_ctx* c, const sla_csr* A, int ph) {
  sla_xwin* X = A->dist->xwin;
  if (!X->src_mask[ph]) return SLA_OK;
  p2p_wait_kernel<<<1, 32, 0, c->stream>>>(reinterpret_cast<const unsigned long long*>(X->win), X->src_mask[ph], X->seq, c->p2p->d_err);
  SLA_LAUNCH_CHECK(c);
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
  const unsigned mask = src >= 0 ? 1u << src : (((1u << c->world) - 1u) & ~(1u << c->rank));   // src < 0: every other rank
  p2p_wait_kernel<<<1, 32, 0, c->stream>>>(reinterpret_cast<const unsigned long long*>(X->win), mask, X->seq, c->p2p->d_err);
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
  if (X->mode == 3) {
    if (X->nitems + X->nrecv > 0) {
      p2p_halo_ll_kernel<<<X->nitems + X->nrecv, P2P_PUSH_THREADS, 0, c->stream>>>(X->d_items, X->nitems, X->nrecv, X->d_peer, off, x_local,
                                                                                   d->xfull, c->rank, (unsigned)X->seq, c->p2p->d_err);
      SLA_LAUNCH_CHECK(c);
    }
    return SLA_OK;
  }
  if (X->mode == 5) {
    SLA_TRY(sla_p2p_twophase_begin(c, A, x_local));
    for (int ph = 0; ph < X->nphase; ++ph) SLA_TRY(sla_p2p_twophase_wait(c, A, ph));
    return sla_p2p_arrival_end(c);
  }
  if (X->mode == 2 || X->mode == 4) {           // (#>) drives these itself; other callers get the whole gather
    SLA_TRY(sla_p2p_arrival_begin(c, A, x_local));
    SLA_TRY(sla_p2p_arrival_wait(c, A, -1));
    return sla_p2p_arrival_end(c);
  }
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
  cudaFree(X->d_peer); cudaFree(X->d_items); cudaFree(X->d_ticket); cudaFree(X->d_items2); cudaFree(X->d_ticket2);
  if (d->xfull && (char*)d->xfull >= X->win && (char*)d->xfull < X->win + P2P_FLAG_BYTES + 2 * X->buf_bytes) d->xfull = nullptr;   // it pointed into the window
  if (c && c->n_parked < SLA_MAX_PARKED) c->parked[c->n_parked++] = X->win;   // else: leaked until process exit
  delete X;
  d->xwin = nullptr;
}
