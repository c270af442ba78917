// dist.cu — multi-GPU plumbing: one process (one sla_ctx) per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The matrix is row-partitioned: rank p owns a contiguous block of rows (with GLOBAL column indices) and the
// matching slice of every vector.  Per (#>) the ranks exchange exactly the x entries the plan says they need
// (a full all-gather when the column support is dense, neighbour halos for banded / stencil matrices);
// every dot / norm is a local deterministic reduction followed by one small all-reduce, after which a
// one-thread kernel derives alpha / omega / beta on every rank identically.
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>      // types only: the library is bound at run time (see nccl_bind)
#include <new>

// NCCL is bound with dlopen when the first multi-GPU call arrives, not at link time: a process that already
// holds torch's bundled libnccl.so.2 keeps using that copy (RTLD_NOLOAD), a single-GPU process never loads NCCL,
// and importing this library can never shadow the NCCL another framework expects.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*);   // optional (NCCL >= 2.18)
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
  bool ok;
};
static NcclApi g_nccl = {};

static bool nccl_bind() {
  if (g_nccl.ok) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return false;
#define BIND(field, name) *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return false
  BIND(GetUniqueId, "ncclGetUniqueId"); BIND(CommInitRank, "ncclCommInitRank"); BIND(CommDestroy, "ncclCommDestroy");
  BIND(AllReduce, "ncclAllReduce"); BIND(AllGather, "ncclAllGather"); BIND(Send, "ncclSend"); BIND(Recv, "ncclRecv");
  BIND(GroupStart, "ncclGroupStart"); BIND(GroupEnd, "ncclGroupEnd"); BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
  *(void**)(&g_nccl.CommSplit) = dlsym(h, "ncclCommSplit");
  g_nccl.ok = true;
  return true;
}
#define ncclGetUniqueId g_nccl.GetUniqueId
#define ncclCommInitRank g_nccl.CommInitRank
#define ncclCommDestroy g_nccl.CommDestroy
#define ncclAllReduce g_nccl.AllReduce
#define ncclAllGather g_nccl.AllGather
#define ncclSend g_nccl.Send
#define ncclRecv g_nccl.Recv
#define ncclGroupStart g_nccl.GroupStart
#define ncclGroupEnd g_nccl.GroupEnd
#define ncclGetErrorString g_nccl.GetErrorString

#define SLA_NCCL(ctx, call)                                                                          \
  do {                                                                                               \
    ncclResult_t _r = (call);                                                                        \
    if (_r != ncclSuccess) {                                                                         \
      snprintf((ctx)->err, sizeof((ctx)->err), "NCCL error %s at %s:%d (%s)", ncclGetErrorString(_r), \
               __FILE__, __LINE__, #call);                                                           \
      return SLA_ERR_COMM;                                                                           \
    }                                                                                                \
  } while (0)

extern "C" sla_status sla_nccl_unique_id(void* out128) {
  if (!out128) return SLA_ERR_INVALID;
  if (!nccl_bind()) return SLA_ERR_COMM;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return SLA_ERR_COMM;
  memcpy(out128, &id, sizeof(id));
  return SLA_OK;
}

sla_status sla_dist_attach(sla_ctx* c, const void* nccl_id128) {
  if (!nccl_id128) return sla_fail(c, SLA_ERR_INVALID, "sla_init_dist: missing NCCL unique id");
  if (!nccl_bind()) return sla_fail(c, SLA_ERR_COMM, "sla_init_dist: libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  memcpy(&id, nccl_id128, sizeof(id));
  ncclComm_t comm = nullptr;
  SLA_CUDA(c, cudaSetDevice(c->device));
  SLA_NCCL(c, ncclCommInitRank(&comm, c->world, id, c->rank));
  c->nccl = (void*)comm;
  // the x exchange gets its own stream and, when the library offers it, its own communicator, so that it can
  // run under the column-panel kernels without being serialised behind the all-reduces of the compute stream
  c->nccl_x = c->nccl;
  if (g_nccl.CommSplit) {
    ncclComm_t cx = nullptr;
    if (g_nccl.CommSplit(comm, 0, c->rank, &cx, nullptr) == ncclSuccess && cx) c->nccl_x = (void*)cx;
  }
  {
    // the exchange stream outranks the compute stream: its (few) CTAs are scheduled ahead of the panel kernels they run under
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    SLA_CUDA(c, cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
  }
  SLA_CUDA(c, cudaEventCreateWithFlags(&c->ev_x0, cudaEventDisableTiming));
  for (int p = 0; p < SLA_MAX_PANELS; ++p) SLA_CUDA(c, cudaEventCreateWithFlags(&c->ev_panel[p], cudaEventDisableTiming));
  return SLA_OK;
}

void sla_dist_detach(sla_ctx* c) {
  if (c->comm_stream) {
    cudaStreamSynchronize(c->comm_stream);
    cudaStreamDestroy(c->comm_stream); c->comm_stream = nullptr;
    cudaEventDestroy(c->ev_x0);
    for (int p = 0; p < SLA_MAX_PANELS; ++p) cudaEventDestroy(c->ev_panel[p]);
  }
  if (c->nccl_x && c->nccl_x != c->nccl) ncclCommDestroy((ncclComm_t)c->nccl_x);
  c->nccl_x = nullptr;
  if (c->nccl) { ncclCommDestroy((ncclComm_t)c->nccl); c->nccl = nullptr; }
}

// one-thread kernel: the scalar post-processing of a grid reduction, run after the all-reduce
__global__ void finalize_kernel(int fin, int dst, double* scal, int src, int nv) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double sum[32];
    for (int k = 0; k < nv && k < 32; ++k) sum[k] = scal[src + k];
    finalize_scalars(fin, dst, scal, sum, nv);
  }
}

// Completes a reduction whose per-rank sums sit in scal[S_RAW .. S_RAW + nv): all-reduce, then finalize.
sla_status sla_dist_finish_reduction(sla_ctx* c, int nv, int fin, int dst) {
  if (c->world <= 1) return SLA_OK;
  // peer-memory path: all-reduce + post-processing in ONE single-CTA kernel (p2p.cu)
  if (sla_p2p_active(c) && nv <= 32) return sla_p2p_allreduce(c, nv, fin == FIN_STORE ? dst : S_RAW, fin, dst);
  ncclComm_t comm = (ncclComm_t)c->nccl;
  if (fin == FIN_STORE) {
    // raw sums were written straight to their destination slots
    SLA_NCCL(c, ncclAllReduce(c->scal + dst, c->scal + dst, (size_t)nv, ncclDouble, ncclSum, comm, c->stream));
    return SLA_OK;
  }
  SLA_NCCL(c, ncclAllReduce(c->scal + S_RAW, c->scal + S_RAW, (size_t)nv, ncclDouble, ncclSum, comm, c->stream));
  finalize_kernel<<<1, 32, 0, c->stream>>>(fin, dst, c->scal, S_RAW, nv);
  SLA_LAUNCH_CHECK(c);
  return SLA_OK;
}

// ---- helpers of the distributed transpose (dist_transpose.cu keeps no NCCL types of its own) --------------------
sla_status sla_dist_allgather_i32(sla_ctx* c, const int* d_send, int* d_recv, int count) {
  if (c->world <= 1) return SLA_OK;
  SLA_NCCL(c, ncclAllGather(d_send, d_recv, (size_t)count, ncclInt32, (ncclComm_t)c->nccl, c->stream));
  return SLA_OK;
}
sla_status sla_dist_group_begin(sla_ctx* c) { SLA_NCCL(c, ncclGroupStart()); return SLA_OK; }
sla_status sla_dist_group_end(sla_ctx* c) { SLA_NCCL(c, ncclGroupEnd()); return SLA_OK; }
// bytes8 = 0: int32 elements, 1: doubles
sla_status sla_dist_send(sla_ctx* c, const void* p, size_t count, int bytes8, int peer) {
  SLA_NCCL(c, ncclSend(p, count, bytes8 ? ncclDouble : ncclInt32, peer, (ncclComm_t)c->nccl, c->stream));
  return SLA_OK;
}
sla_status sla_dist_recv(sla_ctx* c, void* p, size_t count, int bytes8, int peer) {
  SLA_NCCL(c, ncclRecv(p, count, bytes8 ? ncclDouble : ncclInt32, peer, (ncclComm_t)c->nccl, c->stream));
  return SLA_OK;
}

sla_status sla_dist_allreduce_int(sla_ctx* c, int* d_val, int count) {
  if (c->world <= 1) return SLA_OK;
  SLA_NCCL(c, ncclAllReduce(d_val, d_val, (size_t)count, ncclInt, ncclSum, (ncclComm_t)c->nccl, c->stream));
  return SLA_OK;
}

// ---- x exchange ------------------------------------------------------------------------------------------

void sla_csr_free_dist(sla_csr* A) {
  if (!A->dist) return;
  if (A->ctx && A->ctx->comm_stream) cudaStreamSynchronize(A->ctx->comm_stream);
  sla_xwin_free(A);                   // in peer-memory mode xfull pointed into the window
  cudaFree(A->dist->xfull);
  delete[] A->dist->seg;
  delete[] A->dist->pseg;
  delete[] A->dist->seg_base;
  delete A->dist;
  A->dist = nullptr;
}

// Declares A a row block of a distributed matrix: local rows are global rows [row0, row0 + m); segments say
// which contiguous pieces of x travel (dir 0: receive [goff, goff+count) of the global vector from peer;
// dir 1: send local entries [goff - row0, ...) to peer).  The plan is computed on the host (dist.py).
extern "C" sla_status sla_csr_set_dist(sla_ctx* c, sla_csr* A, int64_t row0, int nseg, const int* dir, const int* peer,
                                       const int64_t* goff, const int64_t* count, int allgather) {
  if (!c || !A || nseg < 0 || (nseg > 0 && (!dir || !peer || !goff || !count))) return SLA_ERR_INVALID;
  if (row0 < 0 || row0 + A->m > A->n) return sla_fail(c, SLA_ERR_INVALID, "set_dist: local row block lies outside the global index range");
  sla_csr_free_dist(A);
  sla_dist_info* d = new (std::nothrow) sla_dist_info();
  if (!d) return sla_fail(c, SLA_ERR_ALLOC, "set_dist alloc");
  d->row0 = row0; d->nseg = nseg; d->xfull = nullptr; d->allgather = 0; d->dense_equal = 0; d->pipelined = 0; d->pseg = nullptr; d->xwin = nullptr; d->seg_base = nullptr; d->halo_total = 0;
  d->seg = new sla_xseg[nseg > 0 ? nseg : 1];
  for (int s = 0; s < nseg; ++s) {
    if (peer[s] < 0 || peer[s] >= c->world || peer[s] == c->rank || goff[s] < 0 || count[s] < 0 || goff[s] + count[s] > A->n ||
        (dir[s] == 1 && (goff[s] < row0 || goff[s] + count[s] > row0 + A->m))) {
      delete[] d->seg; delete d;
      return sla_fail(c, SLA_ERR_INVALID, "set_dist: bad exchange segment");
    }
    d->seg[s].dir = dir[s]; d->seg[s].peer = peer[s]; d->seg[s].goff = goff[s]; d->seg[s].count = count[s];
  }
  if (cudaMalloc(&d->xfull, sizeof(double) * (size_t)((A->n + 2) & ~(int64_t)1)) != cudaSuccess) {
    cudaGetLastError(); delete[] d->seg; delete d;
    return sla_fail(c, SLA_ERR_ALLOC, "set_dist: cudaMalloc failed for the gathered x buffer");
  }
  cudaMemsetAsync(d->xfull, 0, sizeof(double) * (size_t)A->n, c->stream);
  // allgather is a COLLECTIVE decision taken by the host planner from the global needs table (every rank must
  // issue the same NCCL call); it requires equal blocks laid out in rank order.
  if (allgather) {
    if (c->world < 2 || A->n % c->world != 0 || A->m != A->n / c->world || row0 != (int64_t)c->rank * A->m) {
      cudaFree(d->xfull); delete[] d->seg; delete d;
      return sla_fail(c, SLA_ERR_INVALID, "set_dist: all-gather needs equal row blocks in rank order");
    }
    d->allgather = 1;
    d->dense_equal = 1;
  }
  A->dist = d;
  // EXPERIMENTAL (SLA_DIST_PIPELINE=1): pipeline the exchange under the column-panel kernels.  The panel count is
  // derived from global quantities only (n), so every rank clips the same segments at the same boundaries.
  // Measured on B200 (cfg 2, round 1): the grouped ncclSend/ncclRecv per panel are slower than one ncclAllGather
  // (N=2: 0.994 vs 0.950 ms, N=4: 0.694 vs 0.566 ms) and only the exchange of the panels after the first can hide
  // behind compute, so the default stays the single all-gather; see DESIGN.md §5.
  if (allgather && c->world > 1 && getenv("SLA_DIST_PIPELINE")) {
    int P = (int)(((uint64_t)A->n * 8u + (40u << 20) - 1) / (40u << 20));
    if (P < 2) P = 2;
    if (const char* e = getenv("SLA_DIST_PANELS")) P = atoi(e);
    if ((uint64_t)A->n * 8u >= (4u << 20) && P >= 2) {
      sla_status ps = sla_csr_force_panels(c, A, P);
      if (ps != SLA_OK) { sla_csr_free_dist(A); return ps; }
      const int np = A->npanels;
      const int64_t W = A->panel_width;
      // clip every segment to every panel it overlaps
      int total = 0;
      for (int pass = 0; pass < 2; ++pass) {
        int k = 0;
        for (int p = 0; p < np; ++p) {
          if (pass == 1) d->pan_first[p] = k;
          const int64_t lo = (int64_t)p * W, hi = lo + W < A->n ? lo + W : A->n;
          for (int sgi = 0; sgi < nseg; ++sgi) {
            const sla_xseg& g = d->seg[sgi];
            const int64_t a = g.goff > lo ? g.goff : lo, b = g.goff + g.count < hi ? g.goff + g.count : hi;
            if (b <= a) continue;
            if (pass == 1) { d->pseg[k].dir = g.dir; d->pseg[k].peer = g.peer; d->pseg[k].goff = a; d->pseg[k].count = b - a; }
            ++k;
          }
        }
        if (pass == 0) { total = k; d->pseg = new sla_xseg[total > 0 ? total : 1]; }
        else d->pan_first[np] = k;
      }
      d->pipelined = 1;
      d->allgather = 0;      // the per-panel point-to-point groups replace the single all-gather
    }
  }
  A->dist = d;
  return SLA_OK;
}

// (##) with a dense right operand on a row-partitioned matrix (NOT YET RUN ON HARDWARE): brings the rows of B this
// rank's block references into `full` (n x k, row-major) — the x-exchange plan applied to k-wide rows.  `local` holds
// this rank's rows [row0, row0 + m).  Dense equal-block plans use one all-gather, every other plan grouped send/recv.
sla_status sla_dist_gather_rows(sla_ctx* c, const sla_csr* A, const void* local, void* full, int64_t k, int dtype) {
  const sla_dist_info* d = A->dist;
  if (!d || c->world <= 1) return SLA_OK;
  const size_t esz = dtype == SLA_BF16 ? 2 : 8;
  const ncclDataType_t nt = dtype == SLA_BF16 ? ncclBfloat16 : ncclDouble;
  ncclComm_t comm = (ncclComm_t)c->nccl;
  if (d->dense_equal) {
    SLA_NCCL(c, ncclAllGather(local, full, (size_t)A->m * (size_t)k, nt, comm, c->stream));
    return SLA_OK;
  }
  if (A->m > 0)
    SLA_CUDA(c, cudaMemcpyAsync((char*)full + (size_t)d->row0 * (size_t)k * esz, local, (size_t)A->m * (size_t)k * esz,
                                cudaMemcpyDeviceToDevice, c->stream));
  SLA_NCCL(c, ncclGroupStart());
  for (int s = 0; s < d->nseg; ++s) {
    const sla_xseg& g = d->seg[s];
    if (g.count == 0) continue;
    if (g.dir == 0) SLA_NCCL(c, ncclRecv((char*)full + (size_t)g.goff * (size_t)k * esz, (size_t)g.count * (size_t)k, nt, g.peer, comm, c->stream));
    else            SLA_NCCL(c, ncclSend((const char*)local + (size_t)(g.goff - d->row0) * (size_t)k * esz, (size_t)g.count * (size_t)k, nt, g.peer, comm, c->stream));
  }
  SLA_NCCL(c, ncclGroupEnd());
  return SLA_OK;
}

// the exchange restricted to column panel p, issued on comm_stream (the caller orders it with events)
sla_status sla_dist_exchange_panel(sla_ctx* c, const sla_csr* A, const double* x_local, int p) {
  const sla_dist_info* d = A->dist;
  ncclComm_t comm = (ncclComm_t)c->nccl_x;
  const int s0 = d->pan_first[p], s1 = d->pan_first[p + 1];
  if (s1 <= s0) return SLA_OK;
  SLA_NCCL(c, ncclGroupStart());
  for (int s = s0; s < s1; ++s) {
    const sla_xseg& g = d->pseg[s];
    if (g.dir == 0) SLA_NCCL(c, ncclRecv(d->xfull + g.goff, (size_t)g.count, ncclDouble, g.peer, comm, c->comm_stream));
    else            SLA_NCCL(c, ncclSend(x_local + (g.goff - d->row0), (size_t)g.count, ncclDouble, g.peer, comm, c->comm_stream));
  }
  SLA_NCCL(c, ncclGroupEnd());
  return SLA_OK;
}

// brings the remote x entries this rank's rows reference into A->dist->xfull
sla_status sla_dist_exchange_x(sla_ctx* c, const sla_csr* A, const double* x_local) {
  const sla_dist_info* d = A->dist;
  if (!d || c->world <= 1) return SLA_OK;
  if (sla_xwin_active(A)) return sla_p2p_exchange_x(c, A, x_local);      // NVLink stores into the peers' buffers (p2p.cu)
  ncclComm_t comm = (ncclComm_t)c->nccl;
  if (d->allgather) {
    SLA_NCCL(c, ncclAllGather(x_local, d->xfull, (size_t)A->m, ncclDouble, comm, c->stream));
    return SLA_OK;
  }
  SLA_NCCL(c, ncclGroupStart());
  for (int s = 0; s < d->nseg; ++s) {
    const sla_xseg& g = d->seg[s];
    if (g.count == 0) continue;
    if (g.dir == 0) SLA_NCCL(c, ncclRecv(d->xfull + g.goff, (size_t)g.count, ncclDouble, g.peer, comm, c->stream));
    else            SLA_NCCL(c, ncclSend(x_local + (g.goff - d->row0), (size_t)g.count, ncclDouble, g.peer, comm, c->stream));
  }
  SLA_NCCL(c, ncclGroupEnd());
  return SLA_OK;
}
