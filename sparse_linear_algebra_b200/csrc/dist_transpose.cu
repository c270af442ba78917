// dist_transpose.cu — transposeSM (SpMatrix.hs:717-718) of a ROW-PARTITIONED matrix: an all-to-all of entries.
// NOT YET RUN ON HARDWARE (written after the round-1 GPU budget was spent; tests/dist_check.py checks it behind
// SLA_DIST_CHECK_EXPERIMENTAL=1).
//
// Rank r holds rows [starts[r], starts[r+1]) of the square n x n matrix with GLOBAL column indices and receives rows
// [starts[r], starts[r+1]) of the transpose.  Steps:
//   1. local transpose of the block (the bit-exact single-GPU kernel): T_loc is n x m_r, sorted by (new row, old local row);
//      the entries destined to rank q — new rows [starts[q], starts[q+1]) — are therefore one contiguous range;
//   2. the ranks all-gather the range lengths, then exchange (row_ptr slice, col, val) triples in one NCCL group;
//   3. the receiver concatenates, per row, the pieces of the sources IN RANK ORDER — ascending old row = ascending new
//      column — adding starts[s] to turn the source's local row numbers into global column indices.
// The result is bit-identical to the single-process transpose restricted to the local rows (integer + copy work).
#include "common.cuh"

#include <cub/device/device_scan.cuh>
#include <new>

namespace {

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <class T> T* as() { return (T*)p; }
};

struct TdSources {                 // per source rank: the received (or local) pieces
  const int* rp[SLA_MAX_WORLD];    // m + 1 raw row_ptr values of the source's T_loc for my rows (base = rp[s][0])
  const int* col[SLA_MAX_WORLD];   // the source's local row numbers, starting at the base
  const double* val[SLA_MAX_WORLD];
  int col_off[SLA_MAX_WORLD];      // starts[s]
};

__global__ void td_len_kernel(int m, int W, TdSources S, int* __restrict__ len) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= m; j += gridDim.x * blockDim.x) {
    int t = 0;
    if (j < m)
      for (int s = 0; s < W; ++s) t += S.rp[s][j + 1] - S.rp[s][j];
    len[j] = t;                    // len[m] = 0: the exclusive scan over m + 1 entries ends with the total
  }
}

__global__ void td_fill_kernel(int m, int W, TdSources S, const int* __restrict__ out_ptr, int* __restrict__ out_col,
                               double* __restrict__ out_val) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
    int o = out_ptr[j];
    for (int s = 0; s < W; ++s) {
      const int base = S.rp[s][0];
      for (int k = S.rp[s][j]; k < S.rp[s][j + 1]; ++k, ++o) {
        out_col[o] = S.col[s][k - base] + S.col_off[s];
        out_val[o] = S.val[s][k - base];
      }
    }
  }
}

inline unsigned blocks_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > SLA_NUM_SMS * 16) b = SLA_NUM_SMS * 16;
  return (unsigned)b;
}

}  // namespace

// starts: world + 1 global row offsets (the same on every rank).  *out is this rank's row block of the transpose with
// GLOBAL column indices; the caller installs its exchange plan (sla_csr_col_range + sla_csr_set_dist) like for any block.
extern "C" sla_status sla_csr_transpose_dist(sla_ctx* c, const sla_csr* A, const int64_t* starts, sla_csr** out) {
  if (!c || !A || !starts || !out) return SLA_ERR_INVALID;
  *out = nullptr;
  const int W = c->world, me = c->rank;
  if (W < 2 || W > SLA_MAX_WORLD) return sla_fail(c, SLA_ERR_INVALID, "transpose (distributed): needs 2..16 ranks");
  const int64_t n = starts[W];
  if (A->n != n || A->m != starts[me + 1] - starts[me] || starts[0] != 0)
    return sla_fail(c, SLA_ERR_SIZE_MISMATCH, "transpose (distributed): the block does not match the row partition of a square matrix");
  const int m = (int)A->m;

  // 1. local transpose: n x m, columns = local row numbers
  sla_csr* T = nullptr;
  SLA_TRY(sla_csr_transpose(c, A, &T));
  struct Guard { sla_csr* t; ~Guard() { if (t) sla_csr_free(t); } } guard{T};

  // 2. range of every destination inside T, lengths all-gathered
  int h_off[SLA_MAX_WORLD + 1];
  for (int q = 0; q <= W; ++q)
    SLA_CUDA(c, cudaMemcpyAsync(&h_off[q], T->row_ptr + starts[q], sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  int h_send[SLA_MAX_WORLD], h_all[SLA_MAX_WORLD * SLA_MAX_WORLD];
  for (int q = 0; q < W; ++q) h_send[q] = h_off[q + 1] - h_off[q];
  DevBuf d_send, d_all;
  SLA_CUDA(c, d_send.alloc(sizeof(int) * W)); SLA_CUDA(c, d_all.alloc(sizeof(int) * W * W));
  SLA_CUDA(c, cudaMemcpyAsync(d_send.p, h_send, sizeof(int) * W, cudaMemcpyHostToDevice, c->stream));
  SLA_TRY(sla_dist_allgather_i32(c, d_send.as<int>(), d_all.as<int>(), W));
  SLA_CUDA(c, cudaMemcpyAsync(h_all, d_all.p, sizeof(int) * W * W, cudaMemcpyDeviceToHost, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  // h_all[s * W + d] = entries rank s sends to rank d

  // receive buffers
  DevBuf rp[SLA_MAX_WORLD], rc[SLA_MAX_WORLD], rv[SLA_MAX_WORLD];
  for (int s = 0; s < W; ++s) {
    if (s == me) continue;
    const int cnt = h_all[s * W + me];
    SLA_CUDA(c, rp[s].alloc(sizeof(int) * (size_t)(m + 1)));
    SLA_CUDA(c, rc[s].alloc(sizeof(int) * (size_t)cnt));
    SLA_CUDA(c, rv[s].alloc(sizeof(double) * (size_t)cnt));
  }
  SLA_TRY(sla_dist_group_begin(c));
  for (int q = 0; q < W; ++q) {
    if (q == me) continue;
    const int mq = (int)(starts[q + 1] - starts[q]);
    SLA_TRY(sla_dist_send(c, T->row_ptr + starts[q], (size_t)(mq + 1), 0, q));
    if (h_send[q] > 0) {
      SLA_TRY(sla_dist_send(c, T->col + h_off[q], (size_t)h_send[q], 0, q));
      SLA_TRY(sla_dist_send(c, T->val + h_off[q], (size_t)h_send[q], 1, q));
    }
    const int cnt = h_all[q * W + me];
    SLA_TRY(sla_dist_recv(c, rp[q].p, (size_t)(m + 1), 0, q));
    if (cnt > 0) {
      SLA_TRY(sla_dist_recv(c, rc[q].p, (size_t)cnt, 0, q));
      SLA_TRY(sla_dist_recv(c, rv[q].p, (size_t)cnt, 1, q));
    }
  }
  SLA_TRY(sla_dist_group_end(c));

  // 3. assemble
  TdSources S;
  int64_t total = 0;
  for (int s = 0; s < W; ++s) {
    S.col_off[s] = (int)starts[s];
    if (s == me) {
      S.rp[s] = T->row_ptr + starts[me]; S.col[s] = T->col + h_off[me]; S.val[s] = T->val + h_off[me];
      total += h_send[me];
    } else {
      S.rp[s] = rp[s].as<int>(); S.col[s] = rc[s].as<int>(); S.val[s] = rv[s].as<double>();
      total += h_all[s * W + me];
    }
  }
  for (int s = W; s < SLA_MAX_WORLD; ++s) { S.rp[s] = nullptr; S.col[s] = nullptr; S.val[s] = nullptr; S.col_off[s] = 0; }
  sla_csr* R = nullptr;
  SLA_TRY(sla_csr_alloc(c, m, n, total, &R));
  struct Guard2 { sla_csr* r; ~Guard2() { if (r) sla_csr_free(r); } } guard2{R};
  DevBuf len, tmp;
  SLA_CUDA(c, len.alloc(sizeof(int) * (size_t)(m + 1)));
  td_len_kernel<<<blocks_for(m + 1), 256, 0, c->stream>>>(m, W, S, len.as<int>());
  SLA_LAUNCH_CHECK(c);
  size_t tb = 0;
  SLA_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tb, len.as<int>(), R->row_ptr, m + 1, c->stream));
  SLA_CUDA(c, tmp.alloc(tb));
  SLA_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, tb, len.as<int>(), R->row_ptr, m + 1, c->stream));
  c->launches += 2;
  if (m > 0) {
    td_fill_kernel<<<blocks_for(m), 256, 0, c->stream>>>(m, W, S, R->row_ptr, R->col, R->val);
    SLA_LAUNCH_CHECK(c);
  }
  SLA_TRY(sla_csr_build_plan(c, R));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));       // the receive buffers and T are released on return
  guard2.r = nullptr;
  *out = R;
  return SLA_OK;
}

// Hands a distributed transpose (built with sla_csr_transpose_dist and given its exchange plan) to A as its cached
// transpose: (<#) and cgneInit / cgneStep on the row-partitioned A use it.  A owns T from here on.
extern "C" sla_status sla_csr_attach_transpose(sla_ctx* c, sla_csr* A, sla_csr* T) {
  if (!c || !A || !T) return SLA_ERR_INVALID;
  if (T->m != A->m || T->n != A->n) return sla_fail(c, SLA_ERR_SIZE_MISMATCH, "attach_transpose: shapes differ");
  if (A->T && A->T != T) sla_csr_free(A->T);
  A->T = T;
  return SLA_OK;
}
