// multi.cu — ONE host process driving several GPUs ("the caller never sees ranks").
//
// The reference is a single-threaded pure library, so its natural caller is one process.  sla_init_multi(n_gpus, ids) builds
// one per-GPU context (the same sla_ctx the one-process-per-GPU path uses: own stream, own NCCL rank, peer-memory all-reduce
// window) on a worker thread per GPU and exposes GLOBAL objects — a matrix is row-partitioned across the GPUs, a vector is
// the concatenation of the ranks' slices — behind handles of their own.  Every entry point below is the same call issued on
// all ranks AT ONCE (the per-rank calls are collective: a rank's kernel waits for its peers' all-reduce contributions, so
// they cannot be issued one after the other from a single thread) — the worker pool is what lets a plain sequential caller
// (a Haskell `foreign import ccall safe`, a C program) drive the row-partitioned path of dist.cu / p2p.cu.
//
// The exchange plan is computed here the same way sparse_linear_algebra_b200/dist.py computes it for torchrun jobs
// (row_partition, plan_exchange, densify_needs): contiguous row blocks, per-rank column range -> which contiguous pieces of x
// travel; dense equal-block supports use one all-gather.  The x exchange uses NCCL; the all-reduce behind every dot runs over
// peer memory (plain device pointers of this process once peer access is enabled), inlined into the reducing kernels.
#include "common.cuh"

#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

extern "C" sla_status sla_p2p_export(sla_ctx* c, void* handle64);
extern "C" sla_status sla_p2p_enable(sla_ctx* c, int on);

struct sla_mctx {
  int n;
  int dev[SLA_MAX_WORLD];
  sla_ctx* ctx[SLA_MAX_WORLD];
  std::thread th[SLA_MAX_WORLD];
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  std::function<sla_status(int)> task;
  unsigned long long gen;
  int pending;
  bool quit;
  sla_status status[SLA_MAX_WORLD];
  char err[512];
};
struct sla_mcsr { sla_mctx* m; sla_csr* blk[SLA_MAX_WORLD]; int64_t rows, cols, nnz; std::vector<int64_t> starts; };
struct sla_mvec { sla_mctx* m; sla_vec* v[SLA_MAX_WORLD]; int64_t n; std::vector<int64_t> starts; };
struct sla_mkrylov { sla_mctx* m; sla_krylov* st[SLA_MAX_WORLD]; int64_t n; std::vector<int64_t> starts; };
struct sla_mdense { sla_mctx* m; sla_dense* d[SLA_MAX_WORLD]; int64_t rows, cols; std::vector<int64_t> starts; };

namespace {

char g_minit_err[512] = "";

void worker(sla_mctx* m, int rank) {
  cudaSetDevice(m->dev[rank]);
  unsigned long long seen = 0;
  for (;;) {
    std::function<sla_status(int)> job;
    {
      std::unique_lock<std::mutex> lk(m->mu);
      m->cv_go.wait(lk, [&] { return m->quit || m->gen != seen; });
      if (m->quit) return;
      seen = m->gen;
      job = m->task;
    }
    const sla_status s = job(rank);
    {
      std::lock_guard<std::mutex> lk(m->mu);
      m->status[rank] = s;
      if (--m->pending == 0) m->cv_done.notify_all();
    }
  }
}

// the same call on every rank at once; the first failing rank's status and message are reported
sla_status run_all(sla_mctx* m, std::function<sla_status(int)> fn) {
  {
    std::lock_guard<std::mutex> lk(m->mu);
    m->task = std::move(fn);
    m->pending = m->n;
    m->gen++;
  }
  m->cv_go.notify_all();
  std::unique_lock<std::mutex> lk(m->mu);
  m->cv_done.wait(lk, [&] { return m->pending == 0; });
  for (int r = 0; r < m->n; ++r)
    if (m->status[r] != SLA_OK && m->status[r] != SLA_ERR_BREAKDOWN) {
      snprintf(m->err, sizeof(m->err), "gpu %d: %.480s", m->dev[r], m->ctx[r] ? m->ctx[r]->err : "context missing");
      return m->status[r];
    }
  for (int r = 0; r < m->n; ++r) if (m->status[r] != SLA_OK) return m->status[r];
  return SLA_OK;
}

std::vector<int64_t> row_partition(int64_t n, int world) {           // dist.py: row_partition
  std::vector<int64_t> s(world + 1);
  for (int p = 0; p <= world; ++p) s[p] = n * p / world;
  return s;
}

struct Seg { int dir, peer; int64_t goff, count; };

// dist.py: plan_exchange
std::vector<Seg> plan_exchange(int rank, const std::vector<int64_t>& starts, const std::vector<std::pair<int64_t, int64_t>>& needs) {
  const int world = (int)starts.size() - 1;
  std::vector<Seg> segs;
  auto overlap = [](int64_t lo, int64_t hi, int64_t a, int64_t b, int64_t* off, int64_t* cnt) {
    const int64_t s = lo > a ? lo : a, e = hi + 1 < b ? hi + 1 : b;
    if (e <= s) return false;
    *off = s; *cnt = e - s;
    return true;
  };
  for (int q = 0; q < world; ++q) {
    if (q == rank) continue;
    int64_t off, cnt;
    if (needs[rank].second >= needs[rank].first && overlap(needs[rank].first, needs[rank].second, starts[q], starts[q + 1], &off, &cnt))
      segs.push_back({0, q, off, cnt});
    if (needs[q].second >= needs[q].first && overlap(needs[q].first, needs[q].second, starts[rank], starts[rank + 1], &off, &cnt))
      segs.push_back({1, q, off, cnt});
  }
  return segs;
}

// dist.py: densify_needs — a collective decision from the global tables only
bool densify(const std::vector<int64_t>& starts, std::vector<std::pair<int64_t, int64_t>>& needs, bool* allgather) {
  const int world = (int)starts.size() - 1;
  const int64_t n = starts[world];
  bool any = false, dense = true;
  for (int q = 0; q < world; ++q) {
    if (needs[q].second < needs[q].first) continue;
    any = true;
    int64_t vol = 0;
    for (int p = 0; p < world; ++p) {
      if (p == q) continue;
      const int64_t s = needs[q].first > starts[p] ? needs[q].first : starts[p];
      const int64_t e = needs[q].second + 1 < starts[p + 1] ? needs[q].second + 1 : starts[p + 1];
      if (e > s) vol += e - s;
    }
    if (!(2 * vol > n - (starts[q + 1] - starts[q]))) dense = false;
  }
  dense = dense && any;
  *allgather = false;
  if (!dense) return false;
  bool equal = n % world == 0;
  for (int p = 0; p <= world && equal; ++p) equal = starts[p] == p * (n / world);
  for (int q = 0; q < world; ++q) needs[q] = {0, n - 1};
  *allgather = equal;
  return true;
}

// gives every rank's block of a freshly built matrix its exchange plan (collective over the worker pool)
sla_status install_plans(sla_mctx* m, sla_mcsr* A) {
  const int W = m->n;
  if (W == 1) return SLA_OK;
  std::vector<std::pair<int64_t, int64_t>> needs(W);
  SLA_TRY(run_all(m, [&](int r) {
    int64_t lo = 0, hi = -1;
    const sla_status s = sla_csr_col_range(m->ctx[r], A->blk[r], &lo, &hi);
    needs[r] = {lo, hi};
    return s;
  }));
  bool allgather = false;
  densify(A->starts, needs, &allgather);
  return run_all(m, [&](int r) {
    const std::vector<Seg> segs = plan_exchange(r, A->starts, needs);
    const int ns = (int)segs.size();
    std::vector<int> dir(ns > 0 ? ns : 1), peer(ns > 0 ? ns : 1);
    std::vector<int64_t> goff(ns > 0 ? ns : 1), cnt(ns > 0 ? ns : 1);
    for (int s = 0; s < ns; ++s) { dir[s] = segs[s].dir; peer[s] = segs[s].peer; goff[s] = segs[s].goff; cnt[s] = segs[s].count; }
    return sla_csr_set_dist(m->ctx[r], A->blk[r], A->starts[r], ns, dir.data(), peer.data(), goff.data(), cnt.data(), allgather ? 1 : 0);
  });
}

template <class T> T* new_obj() { return new (std::nothrow) T(); }

}  // namespace

// ---- context ------------------------------------------------------------------------------------------------------

extern "C" const char* sla_multi_last_error(const sla_mctx* m) { return m ? m->err : g_minit_err; }
extern "C" int sla_multi_world(const sla_mctx* m) { return m ? m->n : 0; }
extern "C" sla_ctx* sla_multi_ctx(sla_mctx* m, int rank) { return m && rank >= 0 && rank < m->n ? m->ctx[rank] : nullptr; }

extern "C" void sla_finalize_multi(sla_mctx* m) {
  if (!m) return;
  if (m->n > 0 && m->th[0].joinable()) {
    run_all(m, [&](int r) { if (m->ctx[r]) { cudaStreamSynchronize(m->ctx[r]->stream); } return SLA_OK; });
    run_all(m, [&](int r) { sla_finalize(m->ctx[r]); m->ctx[r] = nullptr; return SLA_OK; });
    { std::lock_guard<std::mutex> lk(m->mu); m->quit = true; }
    m->cv_go.notify_all();
    for (int r = 0; r < m->n; ++r) if (m->th[r].joinable()) m->th[r].join();
  }
  delete m;
}

extern "C" sla_status sla_init_multi(int n_gpus, const int* device_ids, sla_mctx** out) {
  if (!out) return SLA_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    snprintf(g_minit_err, sizeof(g_minit_err), "sla_init_multi: no CUDA device available; this library has no CPU path");
    return SLA_ERR_CUDA;
  }
  if (n_gpus < 1 || n_gpus > SLA_MAX_WORLD || n_gpus > ndev) {
    snprintf(g_minit_err, sizeof(g_minit_err), "sla_init_multi: %d GPUs requested, %d visible (at most %d)", n_gpus, ndev, SLA_MAX_WORLD);
    return SLA_ERR_INVALID;
  }
  sla_mctx* m = new_obj<sla_mctx>();
  if (!m) return SLA_ERR_ALLOC;
  m->n = n_gpus; m->gen = 0; m->pending = 0; m->quit = false; m->err[0] = 0;
  for (int r = 0; r < n_gpus; ++r) {
    m->dev[r] = device_ids ? device_ids[r] : r;
    m->ctx[r] = nullptr;
    if (m->dev[r] < 0 || m->dev[r] >= ndev) { snprintf(g_minit_err, sizeof(g_minit_err), "sla_init_multi: device %d out of range", m->dev[r]); delete m; return SLA_ERR_INVALID; }
  }
  unsigned char id[128] = {0};
  if (n_gpus > 1 && sla_nccl_unique_id(id) != SLA_OK) {
    snprintf(g_minit_err, sizeof(g_minit_err), "sla_init_multi: libnccl.so.2 could not be loaded");
    delete m;
    return SLA_ERR_COMM;
  }
  for (int r = 0; r < n_gpus; ++r) m->th[r] = std::thread(worker, m, r);
  // every rank joins the communicator at the same time (ncclCommInitRank is collective)
  sla_status s = run_all(m, [&](int r) {
    sla_ctx* c = nullptr;
    const sla_status st = n_gpus > 1 ? sla_init_dist(m->dev[r], r, n_gpus, id, &c) : sla_init(m->dev[r], &c);
    m->ctx[r] = c;
    return st;
  });
  if (s != SLA_OK) {
    snprintf(g_minit_err, sizeof(g_minit_err), "sla_init_multi: %s", sla_last_error(nullptr));
    sla_finalize_multi(m);
    return s;
  }
  if (n_gpus > 1) {
    // peer-memory all-reduce: windows are plain device pointers of this process once peer access is on
    std::vector<int> ok(n_gpus, 0);
    run_all(m, [&](int r) {
      unsigned char h[64];
      bool good = sla_p2p_export(m->ctx[r], h) == SLA_OK;
      for (int q = 0; q < n_gpus && good; ++q) {
        if (q == r) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, m->dev[r], m->dev[q]) != cudaSuccess || !can) { good = false; break; }
        const cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[q], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) good = false;
        cudaGetLastError();
      }
      ok[r] = good ? 1 : 0;
      return SLA_OK;
    });
    bool all = true;
    for (int r = 0; r < n_gpus; ++r) all = all && ok[r];
    if (all) {
      std::vector<void*> wins(n_gpus);
      for (int r = 0; r < n_gpus; ++r) wins[r] = sla_p2p_window(m->ctx[r]);
      run_all(m, [&](int r) { ok[r] = sla_p2p_attach_direct(m->ctx[r], wins.data()) == SLA_OK; return SLA_OK; });
      for (int r = 0; r < n_gpus; ++r) all = all && ok[r];
    }
    run_all(m, [&](int r) { return sla_p2p_enable(m->ctx[r], all ? 1 : 0); });   // the same value on every rank (else NCCL all-reduce)
  }
  *out = m;
  return SLA_OK;
}

// ---- matrices -----------------------------------------------------------------------------------------------------

extern "C" void sla_multi_csr_free(sla_mcsr* A) {
  if (!A) return;
  sla_mctx* m = A->m;
  run_all(m, [&](int r) { sla_csr_free(A->blk[r]); return SLA_OK; });
  delete A;
}

static sla_mcsr* mcsr_new(sla_mctx* m, int64_t rows, int64_t cols) {
  sla_mcsr* A = new_obj<sla_mcsr>();
  if (!A) return nullptr;
  A->m = m; A->rows = rows; A->cols = cols; A->nnz = 0;
  A->starts = row_partition(rows, m->n);
  for (int r = 0; r < SLA_MAX_WORLD; ++r) A->blk[r] = nullptr;
  return A;
}

static sla_status mcsr_finish(sla_mctx* m, sla_mcsr* A, sla_status s, sla_mcsr** out) {
  if (s == SLA_OK) s = install_plans(m, A);
  if (s != SLA_OK) { sla_multi_csr_free(A); return s; }
  A->nnz = 0;
  for (int r = 0; r < m->n; ++r) A->nnz += A->blk[r]->nnz;
  *out = A;
  return SLA_OK;
}

extern "C" sla_status sla_multi_csr_generate(sla_mctx* m, int kind, int64_t n, int nnz_per_row, uint64_t seed, int64_t band, sla_mcsr** out) {
  if (!m || !out) return SLA_ERR_INVALID;
  *out = nullptr;
  sla_mcsr* A = mcsr_new(m, n, n);
  if (!A) return SLA_ERR_ALLOC;
  const sla_status s = run_all(m, [&](int r) {
    return sla_csr_generate_rows(m->ctx[r], kind, n, nnz_per_row, seed, band, A->starts[r], A->starts[r + 1], &A->blk[r]);
  });
  return mcsr_finish(m, A, s, out);
}

// global CSR in host memory (square: rows and vectors share one partition) -> one row block per GPU
extern "C" sla_status sla_multi_csr_from_csr(sla_mctx* m, int64_t rows, int64_t cols, int64_t nnz, const int32_t* row_ptr, const int32_t* col_idx,
                                             const double* val, sla_mcsr** out) {
  if (!m || !out || !row_ptr || (nnz > 0 && (!col_idx || !val))) return SLA_ERR_INVALID;
  *out = nullptr;
  if (m->n > 1 && rows != cols) { snprintf(m->err, sizeof(m->err), "sla_multi_csr_from_csr: a row-partitioned matrix must be square"); return SLA_ERR_SIZE_MISMATCH; }
  if (row_ptr[rows] != nnz) { snprintf(m->err, sizeof(m->err), "sla_multi_csr_from_csr: row_ptr[m] != nnz"); return SLA_ERR_INVALID; }
  sla_mcsr* A = mcsr_new(m, rows, cols);
  if (!A) return SLA_ERR_ALLOC;
  const sla_status s = run_all(m, [&](int r) {
    const int64_t r0 = A->starts[r], r1 = A->starts[r + 1];
    const int32_t base = row_ptr[r0];
    std::vector<int32_t> rp((size_t)(r1 - r0 + 1));
    for (int64_t i = r0; i <= r1; ++i) rp[(size_t)(i - r0)] = row_ptr[i] - base;
    return sla_csr_from_csr(m->ctx[r], r1 - r0, cols, row_ptr[r1] - base, rp.data(), col_idx + base, val + base, &A->blk[r]);
  });
  return mcsr_finish(m, A, s, out);
}

extern "C" sla_status sla_multi_csr_dims(const sla_mcsr* A, int64_t* rows, int64_t* cols, int64_t* nnz) {
  if (!A) return SLA_ERR_INVALID;
  if (rows) *rows = A->rows;
  if (cols) *cols = A->cols;
  if (nnz) *nnz = A->nnz;
  return SLA_OK;
}

// ---- vectors ------------------------------------------------------------------------------------------------------

extern "C" void sla_multi_vec_free(sla_mvec* x) {
  if (!x) return;
  run_all(x->m, [&](int r) { sla_vec_free(x->v[r]); return SLA_OK; });
  delete x;
}

static sla_mvec* mvec_new(sla_mctx* m, int64_t n) {
  sla_mvec* x = new_obj<sla_mvec>();
  if (!x) return nullptr;
  x->m = m; x->n = n; x->starts = row_partition(n, m->n);
  for (int r = 0; r < SLA_MAX_WORLD; ++r) x->v[r] = nullptr;
  return x;
}

static sla_status mvec_finish(sla_mvec* x, sla_status s, sla_mvec** out) {
  if (s != SLA_OK) { sla_multi_vec_free(x); return s; }
  *out = x;
  return SLA_OK;
}

extern "C" sla_status sla_multi_vec_create(sla_mctx* m, int64_t n, sla_mvec** out) {
  if (!m || !out || n < 0) return SLA_ERR_INVALID;
  *out = nullptr;
  sla_mvec* x = mvec_new(m, n);
  if (!x) return SLA_ERR_ALLOC;
  return mvec_finish(x, run_all(m, [&](int r) { return sla_vec_create(m->ctx[r], x->starts[r + 1] - x->starts[r], &x->v[r]); }), out);
}

extern "C" sla_status sla_multi_vec_from_host(sla_mctx* m, int64_t n, const double* host, sla_mvec** out) {
  if (!m || !out || n < 0 || (n > 0 && !host)) return SLA_ERR_INVALID;
  *out = nullptr;
  sla_mvec* x = mvec_new(m, n);
  if (!x) return SLA_ERR_ALLOC;
  return mvec_finish(x, run_all(m, [&](int r) { return sla_vec_from_host(m->ctx[r], x->starts[r + 1] - x->starts[r], host + x->starts[r], &x->v[r]); }), out);
}

extern "C" sla_status sla_multi_vec_generate(sla_mctx* m, int64_t n, uint64_t seed, sla_mvec** out) {
  if (!m || !out || n < 0) return SLA_ERR_INVALID;
  *out = nullptr;
  sla_mvec* x = mvec_new(m, n);
  if (!x) return SLA_ERR_ALLOC;
  return mvec_finish(x, run_all(m, [&](int r) { return sla_vec_generate_slice(m->ctx[r], x->starts[r], x->starts[r + 1] - x->starts[r], seed, &x->v[r]); }), out);
}

extern "C" sla_status sla_multi_vec_to_host(sla_mctx* m, const sla_mvec* x, double* host) {
  if (!m || !x || (x->n > 0 && !host)) return SLA_ERR_INVALID;
  return run_all(m, [&](int r) { return sla_vec_to_host(m->ctx[r], x->v[r], host + x->starts[r]); });
}

extern "C" sla_status sla_multi_vec_copy(sla_mctx* m, const sla_mvec* src, sla_mvec* dst) {
  if (!m || !src || !dst) return SLA_ERR_INVALID;
  return run_all(m, [&](int r) { return sla_vec_copy(m->ctx[r], src->v[r], dst->v[r]); });
}

extern "C" int64_t sla_multi_vec_dim(const sla_mvec* x) { return x ? x->n : -1; }

// ---- operator surface ---------------------------------------------------------------------------------------------

extern "C" sla_status sla_multi_spmv(sla_mctx* m, const sla_mcsr* A, const sla_mvec* x, sla_mvec* y) {          // (#>)
  if (!m || !A || !x || !y) return SLA_ERR_INVALID;
  if (A->cols != x->n || A->rows != y->n) { snprintf(m->err, sizeof(m->err), "matVec : mismatched dimensions (%lld,%lld)", (long long)A->cols, (long long)x->n); return SLA_ERR_SIZE_MISMATCH; }
  return run_all(m, [&](int r) { return sla_spmv(m->ctx[r], A->blk[r], x->v[r], y->v[r]); });
}

extern "C" sla_status sla_multi_dot(sla_mctx* m, const sla_mvec* x, const sla_mvec* y, double* out) {           // (<.>)
  if (!m || !x || !y || !out) return SLA_ERR_INVALID;
  std::vector<double> d(m->n, 0.0);
  SLA_TRY(run_all(m, [&](int r) { return sla_dot(m->ctx[r], x->v[r], y->v[r], &d[r]); }));
  *out = d[0];                                   // every rank holds the same all-reduced bits
  return SLA_OK;
}

extern "C" sla_status sla_multi_norm2(sla_mctx* m, const sla_mvec* x, double* out) {
  if (!m || !x || !out) return SLA_ERR_INVALID;
  std::vector<double> d(m->n, 0.0);
  SLA_TRY(run_all(m, [&](int r) { return sla_norm2(m->ctx[r], x->v[r], &d[r]); }));
  *out = d[0];
  return SLA_OK;
}

extern "C" sla_status sla_multi_vec_axpy(sla_mctx* m, double a, const sla_mvec* x, const sla_mvec* y, sla_mvec* z) {   // z = y ^+^ (a .* x)
  if (!m || !x || !y || !z) return SLA_ERR_INVALID;
  return run_all(m, [&](int r) { return sla_vec_axpy(m->ctx[r], a, x->v[r], y->v[r], z->v[r]); });
}

extern "C" sla_status sla_multi_vec_scale(sla_mctx* m, double a, const sla_mvec* x, sla_mvec* z) {                      // z = a .* x
  if (!m || !x || !z) return SLA_ERR_INVALID;
  return run_all(m, [&](int r) { return sla_vec_scale(m->ctx[r], a, x->v[r], z->v[r]); });
}

// ---- Krylov -------------------------------------------------------------------------------------------------------

extern "C" void sla_multi_krylov_free(sla_mkrylov* st) {
  if (!st) return;
  run_all(st->m, [&](int r) { sla_krylov_free(st->st[r]); return SLA_OK; });
  delete st;
}

static sla_status mkrylov_make(sla_mctx* m, int64_t n, std::function<sla_status(int, sla_krylov**)> fn, sla_mkrylov** out) {
  *out = nullptr;
  sla_mkrylov* st = new_obj<sla_mkrylov>();
  if (!st) return SLA_ERR_ALLOC;
  st->m = m; st->n = n; st->starts = row_partition(n, m->n);
  for (int r = 0; r < SLA_MAX_WORLD; ++r) st->st[r] = nullptr;
  const sla_status s = run_all(m, [&](int r) { return fn(r, &st->st[r]); });
  if (s != SLA_OK) { sla_multi_krylov_free(st); return s; }
  *out = st;
  return SLA_OK;
}

extern "C" sla_status sla_multi_bicgstab_init(sla_mctx* m, const sla_mcsr* A, const sla_mvec* b, const sla_mvec* x0, sla_mkrylov** out) {
  if (!m || !A || !b || !x0 || !out) return SLA_ERR_INVALID;
  return mkrylov_make(m, A->rows, [&](int r, sla_krylov** o) { return sla_bicgstab_init(m->ctx[r], A->blk[r], b->v[r], x0->v[r], o); }, out);
}
extern "C" sla_status sla_multi_bicgstab_step(sla_mctx* m, const sla_mcsr* A, const sla_mvec* r0hat, sla_mkrylov* st) {
  if (!m || !A || !r0hat || !st) return SLA_ERR_INVALID;
  return run_all(m, [&](int r) { return sla_bicgstab_step(m->ctx[r], A->blk[r], r0hat->v[r], st->st[r]); });
}
extern "C" sla_status sla_multi_cgs_init(sla_mctx* m, const sla_mcsr* A, const sla_mvec* b, const sla_mvec* x0, sla_mkrylov** out) {
  if (!m || !A || !b || !x0 || !out) return SLA_ERR_INVALID;
  return mkrylov_make(m, A->rows, [&](int r, sla_krylov** o) { return sla_cgs_init(m->ctx[r], A->blk[r], b->v[r], x0->v[r], o); }, out);
}
extern "C" sla_status sla_multi_cgs_step(sla_mctx* m, const sla_mcsr* A, const sla_mvec* rhat, sla_mkrylov* st) {
  if (!m || !A || !rhat || !st) return SLA_ERR_INVALID;
  return run_all(m, [&](int r) { return sla_cgs_step(m->ctx[r], A->blk[r], rhat->v[r], st->st[r]); });
}
extern "C" sla_status sla_multi_krylov_clone(sla_mctx* m, const sla_mkrylov* st, sla_mkrylov** out) {
  if (!m || !st || !out) return SLA_ERR_INVALID;
  return mkrylov_make(m, st->n, [&](int r, sla_krylov** o) { return sla_krylov_clone(m->ctx[r], st->st[r], o); }, out);
}
// field of the record (SLA_FIELD_X / _R / _P / _U), gathered into host memory (n doubles)
extern "C" sla_status sla_multi_krylov_get(sla_mctx* m, const sla_mkrylov* st, int field, double* host) {
  if (!m || !st || !host) return SLA_ERR_INVALID;
  return run_all(m, [&](int r) { return sla_krylov_get(m->ctx[r], st->st[r], field, host + st->starts[r]); });
}

extern "C" sla_status sla_multi_linsolve0(sla_mctx* m, int method, const sla_mcsr* A, const sla_mvec* b, const sla_mvec* x0, const sla_solve_opts* opts,
                                          sla_mvec* x, int* iters, double* resnorm) {
  if (!m || !A || !b || !x0 || !x) return SLA_ERR_INVALID;
  std::vector<int> it(m->n, 0);
  std::vector<double> rs(m->n, 0.0);
  const sla_status s = run_all(m, [&](int r) { return sla_linsolve0(m->ctx[r], method, A->blk[r], b->v[r], x0->v[r], opts, x->v[r], &it[r], &rs[r]); });
  if (iters) *iters = it[0];
  if (resnorm) *resnorm = rs[0];
  return s;
}

extern "C" sla_status sla_multi_gmres(sla_mctx* m, const sla_mcsr* A, const sla_mvec* b, const sla_mvec* x0, int restart, const sla_solve_opts* opts,
                                      sla_mvec* x, int* iters, double* resnorm) {
  if (!m || !A || !b || !x0 || !x) return SLA_ERR_INVALID;
  std::vector<int> it(m->n, 0);
  std::vector<double> rs(m->n, 0.0);
  const sla_status s = run_all(m, [&](int r) { return sla_gmres(m->ctx[r], A->blk[r], b->v[r], x0->v[r], restart, opts, x->v[r], &it[r], &rs[r]); });
  if (iters) *iters = it[0];
  if (resnorm) *resnorm = rs[0];
  return s;
}

// arnoldi aa b kn: H (replicated) goes to h_host, the basis Q stays on the GPUs, row-partitioned like every vector
extern "C" void sla_multi_dense_free(sla_mdense* Q) {
  if (!Q) return;
  run_all(Q->m, [&](int r) { sla_dense_free(Q->d[r]); return SLA_OK; });
  delete Q;
}

extern "C" sla_status sla_multi_arnoldi(sla_mctx* m, const sla_mcsr* A, const sla_mvec* b, int kn, sla_mdense** Qout, double* h_host, int* nmax) {
  if (!m || !A || !b || !Qout || !h_host || !nmax || kn < 2) return SLA_ERR_INVALID;
  *Qout = nullptr;
  sla_mdense* Q = new_obj<sla_mdense>();
  if (!Q) return SLA_ERR_ALLOC;
  Q->m = m; Q->rows = A->rows; Q->cols = 0; Q->starts = A->starts;
  for (int r = 0; r < SLA_MAX_WORLD; ++r) Q->d[r] = nullptr;
  std::vector<int> nm(m->n, 0);
  std::vector<std::vector<double>> h(m->n, std::vector<double>((size_t)(kn + 1) * kn, 0.0));
  const sla_status s = run_all(m, [&](int r) { return sla_arnoldi(m->ctx[r], A->blk[r], b->v[r], kn, &Q->d[r], h[r].data(), &nm[r]); });
  if (s != SLA_OK && s != SLA_ERR_BREAKDOWN) { sla_multi_dense_free(Q); return s; }
  *nmax = nm[0];
  Q->cols = nm[0] + 1;
  memcpy(h_host, h[0].data(), sizeof(double) * (size_t)(nm[0] + 1) * (size_t)nm[0]);
  *Qout = Q;
  return s;
}

// Q as one column-major rows x cols array in host memory
extern "C" sla_status sla_multi_dense_to_host(sla_mctx* m, const sla_mdense* Q, double* host_colmajor) {
  if (!m || !Q || !host_colmajor) return SLA_ERR_INVALID;
  return run_all(m, [&](int r) {
    const int64_t rl = Q->starts[r + 1] - Q->starts[r];
    std::vector<double> loc((size_t)rl * (size_t)Q->cols);
    SLA_TRY(sla_dense_to_host(m->ctx[r], Q->d[r], loc.data()));
    for (int64_t j = 0; j < Q->cols; ++j) memcpy(host_colmajor + j * Q->rows + Q->starts[r], loc.data() + j * rl, sizeof(double) * (size_t)rl);
    return SLA_OK;
  });
}
