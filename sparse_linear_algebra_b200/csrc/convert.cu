// convert.cu — integer / layout work either side of the hot path, all bit-exact:
//   fromListSM  (COO -> CSR, later duplicates overwrite, out-of-bounds is an error)   SpMatrix.hs:205-224
//   transposeSM (CSR -> CSR of the transpose, explicit zeros kept)                     SpMatrix.hs:717-718
//   isDiagonalSM                                                                       SpMatrix.hs:411-415
//   synthetic generators of SURVEY.md §8(d) (include/sla_synth.h)
// Sorting and scans use CUB device primitives (setup path, not the timed path); every key is a unique
// 64-bit (major, minor) pair or the sort is stable, so the output order is fully determined.
#include "common.cuh"
#include "../../include/sla_synth.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <new>

static int64_t padded_nnz(int64_t nnz) {
  const int64_t nt = nnz / SLA_SPMV_TILE + 1;
  return nt * SLA_SPMV_TILE;
}

sla_status sla_csr_alloc(sla_ctx* c, int64_t m, int64_t n, int64_t nnz, sla_csr** out) {
  if (!c || !out || m < 0 || n < 0 || nnz < 0) return SLA_ERR_INVALID;
  SLA_GUARD(c);
  if (m >= (1LL << 31) - 1 || n >= (1LL << 31) - 1 || nnz >= (1LL << 31) - 2 * SLA_SPMV_TILE)
    return sla_fail(c, SLA_ERR_INVALID, "csr: dimensions or nnz exceed the int32 index range of this build");
  sla_csr* A = new (std::nothrow) sla_csr();
  if (!A) return sla_fail(c, SLA_ERR_ALLOC, "csr alloc");
  memset(A, 0, sizeof(*A));
  A->ctx = c; A->m = m; A->n = n; A->nnz = nnz; A->is_diag = -1;
  A->ntiles = (int)(nnz / SLA_SPMV_TILE + 1);
  const int64_t pn = padded_nnz(nnz);
  cudaError_t e1 = cudaMalloc(&A->row_ptr, sizeof(int32_t) * (size_t)(m + 1));
  cudaError_t e2 = cudaMalloc(&A->col, sizeof(int32_t) * (size_t)pn);
  cudaError_t e3 = cudaMalloc(&A->val, sizeof(double) * (size_t)pn);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
    cudaGetLastError();
    sla_csr_free(A);
    return sla_fail(c, SLA_ERR_ALLOC, "cudaMalloc failed for a CSR matrix");
  }
  // zero the padding (and everything else) so that tile loads past nnz read col 0 / val 0.0
  if (cudaMemsetAsync(A->col + nnz, 0, sizeof(int32_t) * (size_t)(pn - nnz), c->stream) != cudaSuccess ||
      cudaMemsetAsync(A->val + nnz, 0, sizeof(double) * (size_t)(pn - nnz), c->stream) != cudaSuccess) {
    cudaGetLastError();
    sla_csr_free(A);
    return sla_fail(c, SLA_ERR_CUDA, "cudaMemsetAsync failed for a CSR matrix");
  }
  *out = A;
  return SLA_OK;
}

extern "C" void sla_csr_free(sla_csr* A) {
  if (!A) return;
  if (A->ctx) cudaStreamSynchronize(A->ctx->stream);
  if (A->T) sla_csr_free(A->T);
  sla_csr_free_panels(A);
  sla_csr_free_band(A);
  sla_csr_free_dist(A);
  sla_csr_free_bsr(A);
  sla_csr_free_tri(A);
  cudaFree(A->row_ptr); cudaFree(A->col); cudaFree(A->val); cudaFree(A->tile_row); cudaFree(A->val_bf16);
  delete A;
}

extern "C" sla_status sla_csr_dims(const sla_csr* A, int64_t* m, int64_t* n, int64_t* nnz) {
  if (!A) return SLA_ERR_INVALID;
  if (m) *m = A->m;
  if (n) *n = A->n;
  if (nnz) *nnz = A->nnz;
  return SLA_OK;
}

// B_spmv(n, nnz) = 12 nnz + 4 (m+1) + 8 n (x) + 8 m (y)   SURVEY.md §8(d)
extern "C" int64_t sla_csr_spmv_bytes(const sla_csr* A) {
  return A ? 12 * A->nnz + 4 * (A->m + 1) + 8 * A->n + 8 * A->m : 0;
}

extern "C" sla_status sla_csr_to_host(sla_ctx* c, const sla_csr* A, int32_t* row_ptr, int32_t* col_idx, double* val) {
  if (!c || !A) return SLA_ERR_INVALID;
  if (row_ptr) SLA_CUDA(c, cudaMemcpyAsync(row_ptr, A->row_ptr, sizeof(int32_t) * (size_t)(A->m + 1), cudaMemcpyDeviceToHost, c->stream));
  if (col_idx && A->nnz) SLA_CUDA(c, cudaMemcpyAsync(col_idx, A->col, sizeof(int32_t) * (size_t)A->nnz, cudaMemcpyDeviceToHost, c->stream));
  if (val && A->nnz) SLA_CUDA(c, cudaMemcpyAsync(val, A->val, sizeof(double) * (size_t)A->nnz, cudaMemcpyDeviceToHost, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  return SLA_OK;
}

// ---- kernels ----------------------------------------------------------------------------------------

#define GS_LOOP(i, n) for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)
static inline unsigned gs_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > SLA_NUM_SMS * 16) b = SLA_NUM_SMS * 16;
  return (unsigned)b;
}

// keys = (major << 32) | minor ; flags out-of-bounds entries   (inBounds02, Utils.hs:109-110)
__global__ void coo_keys_kernel(const int64_t* __restrict__ i, const int64_t* __restrict__ j, int64_t nnz,
                                int64_t m, int64_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ idx,
                                int* __restrict__ oob) {
  GS_LOOP(q, nnz) {
    const int64_t a = i[q], b = j[q];
    if (a < 0 || a >= m || b < 0 || b >= n) { *oob = 1; keys[q] = ~0ULL; }
    else keys[q] = ((uint64_t)a << 32) | (uint64_t)b;
    idx[q] = (uint32_t)q;
  }
}

// keep[q] = 1 when q is the LAST entry of its run of equal keys (later duplicates overwrite)
__global__ void mark_last_kernel(const uint64_t* __restrict__ keys, int64_t nnz, int32_t* __restrict__ keep) {
  GS_LOOP(q, nnz) keep[q] = (q == nnz - 1 || keys[q] != keys[q + 1]) ? 1 : 0;
}

__global__ void compact_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx,
                               const int32_t* __restrict__ keep, const int32_t* __restrict__ pos, int64_t nnz,
                               const double* __restrict__ v_in, uint64_t* __restrict__ ukeys,
                               int32_t* __restrict__ col, double* __restrict__ val) {
  GS_LOOP(q, nnz) {
    if (keep[q]) {
      const int32_t o = pos[q];
      ukeys[o] = keys[q];
      col[o] = (int32_t)(keys[q] & 0xffffffffu);
      val[o] = v_in[idx[q]];
    }
  }
}

// row_ptr[r] = first position whose major index >= r, for r in [0, m]
__global__ void rowptr_from_keys_kernel(const uint64_t* __restrict__ ukeys, int64_t nu, int64_t m, int32_t* __restrict__ row_ptr) {
  GS_LOOP(r, m + 1) {
    const uint64_t target = (uint64_t)r << 32;
    int64_t lo = 0, hi = nu;
    while (lo < hi) {
      const int64_t mid = lo + ((hi - lo) >> 1);
      if (ukeys[mid] < target) lo = mid + 1; else hi = mid;
    }
    row_ptr[r] = (int32_t)lo;
  }
}

// transpose keys: (col << 32) | row for every stored entry; row found by binary search in row_ptr
__global__ void transpose_keys_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col, int64_t m,
                                      int64_t nnz, uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
  GS_LOOP(q, nnz) {
    int64_t lo = 0, hi = m;                 // last row r with row_ptr[r] <= q
    while (lo < hi) {
      const int64_t mid = lo + ((hi - lo + 1) >> 1);
      if (row_ptr[mid] <= q) lo = mid; else hi = mid - 1;
    }
    keys[q] = ((uint64_t)(uint32_t)col[q] << 32) | (uint64_t)lo;
    idx[q] = (uint32_t)q;
  }
}

__global__ void gather_transposed_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx, int64_t nnz,
                                         const double* __restrict__ v_in, int32_t* __restrict__ col, double* __restrict__ val) {
  GS_LOOP(q, nnz) {
    col[q] = (int32_t)(keys[q] & 0xffffffffu);
    val[q] = v_in[idx[q]];
  }
}

// validates a caller-supplied CSR: row_ptr monotone from 0 to nnz, columns in range and strictly ascending
__global__ void validate_csr_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col, int64_t m,
                                    int64_t n, int64_t nnz, int* __restrict__ bad) {
  GS_LOOP(r, m) {
    const int s = row_ptr[r], e = row_ptr[r + 1];
    if (s > e || s < 0 || e > nnz) { *bad = 1; continue; }
    for (int k = s; k < e; ++k) {
      if (col[k] < 0 || col[k] >= n) *bad = 2;
      if (k > s && col[k - 1] >= col[k]) *bad = 3;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (row_ptr[0] != 0 || row_ptr[m] != nnz)) *bad = 1;
}

// isDiagonalSM: every one of the nrows rows is stored with exactly one entry, on the diagonal
__global__ void is_diag_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col, int64_t m, int64_t row0,
                               int* __restrict__ notdiag) {
  GS_LOOP(r, m) {
    const int s = row_ptr[r], e = row_ptr[r + 1];
    if (e - s != 1 || col[s] != (int32_t)(r + row0)) *notdiag = 1;
  }
}

// min / max column index stored in the local block (multi-GPU halo planning)
__global__ void col_range_kernel(const int32_t* __restrict__ col, int64_t nnz, int* __restrict__ lohi) {
  int mn = 0x7fffffff, mx = -1;
  GS_LOOP(q, nnz) { const int cidx = col[q]; mn = min(mn, cidx); mx = max(mx, cidx); }
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0 && mx >= 0) { atomicMin(&lohi[0], mn); atomicMax(&lohi[1], mx); }
}

// rows [row0, row0 + m) of the n x n synthetic family
__global__ void synth_len_kernel(int kind, int64_t n, int k, int64_t band, int64_t row0, int64_t m, int32_t* __restrict__ len) {
  GS_LOOP(i, m) len[i] = sla_synth_row_len(kind, n, k, band, row0 + i);
}

__global__ void synth_fill_kernel(int kind, int64_t n, int k, uint64_t seed, int64_t band, int64_t row0, int64_t m,
                                  const int32_t* __restrict__ row_ptr, int32_t* __restrict__ col, double* __restrict__ val) {
  GS_LOOP(i, m) {
    int64_t cols[SLA_SYNTH_MAX_K];
    double vals[SLA_SYNTH_MAX_K];
    const int cnt = sla_synth_row(kind, n, k, seed, band, row0 + i, cols, vals);
    const int s = row_ptr[i];
    for (int q = 0; q < cnt; ++q) { col[s + q] = (int32_t)cols[q]; val[s + q] = vals[q]; }
  }
}

__global__ void synth_vec_kernel(uint64_t seed, int64_t i0, int64_t n, double* __restrict__ x) {
  GS_LOOP(i, n) x[i] = sla_synth_vec(seed, i0 + i);
}

__global__ void set_last_rowptr_kernel(int32_t* row_ptr, const int32_t* len, int64_t m) {
  if (threadIdx.x == 0 && blockIdx.x == 0) row_ptr[m] = (m > 0) ? row_ptr[m - 1] + len[m - 1] : 0;
}

// ---- host drivers ------------------------------------------------------------------------------------

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <class T> T* as() { return (T*)p; }
};

static sla_status sort_pairs(sla_ctx* c, uint64_t* keys_in, uint64_t* keys_out, uint32_t* idx_in, uint32_t* idx_out,
                             int64_t n, int end_bit) {
  size_t tmp_bytes = 0;
  SLA_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, idx_in, idx_out, (int)n, 0, end_bit, c->stream));
  DevBuf tmp;
  SLA_CUDA(c, tmp.alloc(tmp_bytes));
  SLA_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_in, keys_out, idx_in, idx_out, (int)n, 0, end_bit, c->stream));
  c->launches += 8;
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  return SLA_OK;
}

static int bits_for(int64_t v) { int b = 1; while ((1LL << b) < v && b < 32) ++b; return b; }

extern "C" sla_status sla_csr_from_coo(sla_ctx* c, int64_t m, int64_t n, int64_t nnz, const int64_t* i, const int64_t* j,
                                       const double* v, sla_csr** out) {
  if (!c || !out || m < 0 || n < 0 || nnz < 0 || (nnz > 0 && (!i || !j || !v))) return SLA_ERR_INVALID;
  *out = nullptr;
  if (nnz >= (1LL << 31) - 2 * SLA_SPMV_TILE) return sla_fail(c, SLA_ERR_INVALID, "fromListSM: too many entries for int32 indexing");
  DevBuf di, dj, dv, k0, k1, x0, x1, keep, pos, oob, uk;
  SLA_CUDA(c, di.alloc(sizeof(int64_t) * nnz)); SLA_CUDA(c, dj.alloc(sizeof(int64_t) * nnz)); SLA_CUDA(c, dv.alloc(sizeof(double) * nnz));
  SLA_CUDA(c, k0.alloc(sizeof(uint64_t) * nnz)); SLA_CUDA(c, k1.alloc(sizeof(uint64_t) * nnz));
  SLA_CUDA(c, x0.alloc(sizeof(uint32_t) * nnz)); SLA_CUDA(c, x1.alloc(sizeof(uint32_t) * nnz));
  SLA_CUDA(c, keep.alloc(sizeof(int32_t) * (nnz + 1))); SLA_CUDA(c, pos.alloc(sizeof(int32_t) * (nnz + 1)));
  SLA_CUDA(c, oob.alloc(sizeof(int))); SLA_CUDA(c, uk.alloc(sizeof(uint64_t) * nnz));
  SLA_CUDA(c, cudaMemsetAsync(oob.p, 0, sizeof(int), c->stream));
  int64_t nu = 0;
  if (nnz > 0) {
    SLA_CUDA(c, cudaMemcpyAsync(di.p, i, sizeof(int64_t) * nnz, cudaMemcpyHostToDevice, c->stream));
    SLA_CUDA(c, cudaMemcpyAsync(dj.p, j, sizeof(int64_t) * nnz, cudaMemcpyHostToDevice, c->stream));
    SLA_CUDA(c, cudaMemcpyAsync(dv.p, v, sizeof(double) * nnz, cudaMemcpyHostToDevice, c->stream));
    coo_keys_kernel<<<gs_blocks(nnz), 256, 0, c->stream>>>(di.as<int64_t>(), dj.as<int64_t>(), nnz, m, n, k0.as<uint64_t>(), x0.as<uint32_t>(), oob.as<int>());
    SLA_LAUNCH_CHECK(c);
    int h_oob = 0;
    SLA_CUDA(c, cudaMemcpyAsync(&h_oob, oob.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SLA_CUDA(c, cudaStreamSynchronize(c->stream));
    if (h_oob) return sla_fail(c, SLA_ERR_OOB_INDEX, "insertSpMatrix : index out of bounds");
    // stable LSD radix sort on (row, col): equal keys keep their list order, so the last one is the last written
    SLA_TRY(sort_pairs(c, k0.as<uint64_t>(), k1.as<uint64_t>(), x0.as<uint32_t>(), x1.as<uint32_t>(), nnz, 32 + bits_for(m)));
    mark_last_kernel<<<gs_blocks(nnz), 256, 0, c->stream>>>(k1.as<uint64_t>(), nnz, keep.as<int32_t>());
    SLA_LAUNCH_CHECK(c);
    size_t tb = 0;
    SLA_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tb, keep.as<int32_t>(), pos.as<int32_t>(), (int)nnz, c->stream));
    DevBuf tmp; SLA_CUDA(c, tmp.alloc(tb));
    SLA_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, tb, keep.as<int32_t>(), pos.as<int32_t>(), (int)nnz, c->stream));
    c->launches += 2;
    int32_t last_pos = 0, last_keep = 0;
    SLA_CUDA(c, cudaMemcpyAsync(&last_pos, pos.as<int32_t>() + (nnz - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    SLA_CUDA(c, cudaMemcpyAsync(&last_keep, keep.as<int32_t>() + (nnz - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    SLA_CUDA(c, cudaStreamSynchronize(c->stream));
    nu = (int64_t)last_pos + last_keep;
  }
  sla_csr* A = nullptr;
  SLA_TRY(sla_csr_alloc(c, m, n, nu, &A));
  if (nu > 0) {
    compact_kernel<<<gs_blocks(nnz), 256, 0, c->stream>>>(k1.as<uint64_t>(), x1.as<uint32_t>(), keep.as<int32_t>(), pos.as<int32_t>(), nnz,
                                                         dv.as<double>(), uk.as<uint64_t>(), A->col, A->val);
    SLA_LAUNCH_CHECK(c);
  }
  rowptr_from_keys_kernel<<<gs_blocks(m + 1), 256, 0, c->stream>>>(uk.as<uint64_t>(), nu, m, A->row_ptr);
  SLA_LAUNCH_CHECK(c);
  sla_status s = sla_csr_build_plan(c, A);
  if (s != SLA_OK) { sla_csr_free(A); return s; }
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  *out = A;
  return SLA_OK;
}

extern "C" sla_status sla_csr_from_csr(sla_ctx* c, int64_t m, int64_t n, int64_t nnz, const int32_t* row_ptr,
                                       const int32_t* col_idx, const double* val, sla_csr** out) {
  if (!c || !out || !row_ptr || (nnz > 0 && (!col_idx || !val))) return SLA_ERR_INVALID;
  *out = nullptr;
  sla_csr* A = nullptr;
  SLA_TRY(sla_csr_alloc(c, m, n, nnz, &A));
  DevBuf bad;
  SLA_CUDA(c, bad.alloc(sizeof(int)));
  SLA_CUDA(c, cudaMemsetAsync(bad.p, 0, sizeof(int), c->stream));
  SLA_CUDA(c, cudaMemcpyAsync(A->row_ptr, row_ptr, sizeof(int32_t) * (size_t)(m + 1), cudaMemcpyHostToDevice, c->stream));
  if (nnz > 0) {
    SLA_CUDA(c, cudaMemcpyAsync(A->col, col_idx, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    SLA_CUDA(c, cudaMemcpyAsync(A->val, val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
  }
  validate_csr_kernel<<<gs_blocks(m), 256, 0, c->stream>>>(A->row_ptr, A->col, m, n, nnz, bad.as<int>());
  SLA_LAUNCH_CHECK(c);
  int h_bad = 0;
  SLA_CUDA(c, cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  if (h_bad) {
    sla_csr_free(A);
    return sla_fail(c, h_bad == 2 ? SLA_ERR_OOB_INDEX : SLA_ERR_INVALID,
                    h_bad == 2 ? "csr: column index out of bounds" : h_bad == 3 ? "csr: columns must be strictly ascending within a row"
                                                                                : "csr: row_ptr is not a monotone 0..nnz sequence");
  }
  sla_status s = sla_csr_build_plan(c, A);
  if (s != SLA_OK) { sla_csr_free(A); return s; }
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  *out = A;
  return SLA_OK;
}

extern "C" sla_status sla_csr_transpose(sla_ctx* c, const sla_csr* A, sla_csr** out) {
  if (!c || !A || !out) return SLA_ERR_INVALID;
  *out = nullptr;
  const int64_t nnz = A->nnz;
  sla_csr* T = nullptr;
  SLA_TRY(sla_csr_alloc(c, A->n, A->m, nnz, &T));
  DevBuf k0, k1, x0, x1;
  SLA_CUDA(c, k0.alloc(sizeof(uint64_t) * nnz)); SLA_CUDA(c, k1.alloc(sizeof(uint64_t) * nnz));
  SLA_CUDA(c, x0.alloc(sizeof(uint32_t) * nnz)); SLA_CUDA(c, x1.alloc(sizeof(uint32_t) * nnz));
  if (nnz > 0) {
    transpose_keys_kernel<<<gs_blocks(nnz), 256, 0, c->stream>>>(A->row_ptr, A->col, A->m, nnz, k0.as<uint64_t>(), x0.as<uint32_t>());
    SLA_LAUNCH_CHECK(c);
    sla_status s = sort_pairs(c, k0.as<uint64_t>(), k1.as<uint64_t>(), x0.as<uint32_t>(), x1.as<uint32_t>(), nnz, 32 + bits_for(A->n));
    if (s != SLA_OK) { sla_csr_free(T); return s; }
    gather_transposed_kernel<<<gs_blocks(nnz), 256, 0, c->stream>>>(k1.as<uint64_t>(), x1.as<uint32_t>(), nnz, A->val, T->col, T->val);
    SLA_LAUNCH_CHECK(c);
  }
  rowptr_from_keys_kernel<<<gs_blocks(T->m + 1), 256, 0, c->stream>>>(k1.as<uint64_t>(), nnz, T->m, T->row_ptr);
  SLA_LAUNCH_CHECK(c);
  sla_status s = sla_csr_build_plan(c, T);
  if (s != SLA_OK) { sla_csr_free(T); return s; }
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  *out = T;
  return SLA_OK;
}

extern "C" sla_status sla_csr_is_diagonal(sla_ctx* c, const sla_csr* A, int* out) {
  if (!c || !A || !out) return SLA_ERR_INVALID;
  if (A->is_diag < 0) {
    DevBuf nd;
    SLA_CUDA(c, nd.alloc(sizeof(int)));
    // size d == nrows m needs exactly one entry per row; every rank of a distributed matrix votes
    SLA_CUDA(c, cudaMemsetAsync(nd.p, A->nnz == A->m ? 0 : 1, sizeof(int), c->stream));
    if (A->nnz == A->m && A->m > 0) {
      is_diag_kernel<<<gs_blocks(A->m), 256, 0, c->stream>>>(A->row_ptr, A->col, A->m, A->dist ? A->dist->row0 : 0, nd.as<int>());
      SLA_LAUNCH_CHECK(c);
    }
    if (A->dist) SLA_TRY(sla_dist_allreduce_int(c, nd.as<int>(), 1));
    int h = 0;
    SLA_CUDA(c, cudaMemcpyAsync(&h, nd.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SLA_CUDA(c, cudaStreamSynchronize(c->stream));
    const_cast<sla_csr*>(A)->is_diag = h ? 0 : 1;
  }
  *out = A->is_diag;
  return SLA_OK;
}

extern "C" sla_status sla_csr_col_range(sla_ctx* c, const sla_csr* A, int64_t* lo, int64_t* hi) {
  if (!c || !A || !lo || !hi) return SLA_ERR_INVALID;
  DevBuf b;
  SLA_CUDA(c, b.alloc(2 * sizeof(int)));
  const int init[2] = {0x7fffffff, -1};
  SLA_CUDA(c, cudaMemcpyAsync(b.p, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
  if (A->nnz > 0) {
    col_range_kernel<<<gs_blocks(A->nnz), 256, 0, c->stream>>>(A->col, A->nnz, b.as<int>());
    SLA_LAUNCH_CHECK(c);
  }
  int h[2];
  SLA_CUDA(c, cudaMemcpyAsync(h, b.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  *lo = h[1] < 0 ? 0 : h[0];
  *hi = h[1] < 0 ? -1 : h[1];      // empty block: hi < lo
  return SLA_OK;
}

// rows [row_lo, row_hi) of the n x n synthetic matrix, columns GLOBAL (the local block of a row partition)
extern "C" sla_status sla_csr_generate_rows(sla_ctx* c, int kind, int64_t n, int nnz_per_row, uint64_t seed, int64_t band,
                                            int64_t row_lo, int64_t row_hi, sla_csr** out) {
  if (!c || !out || n <= 0 || nnz_per_row < 1 || row_lo < 0 || row_hi < row_lo || row_hi > n) return SLA_ERR_INVALID;
  if (kind != SLA_GEN_UNIFORM && kind != SLA_GEN_BANDED && kind != SLA_GEN_LAPLACE2D && kind != SLA_GEN_BLOCK16) return sla_fail(c, SLA_ERR_INVALID, "generate: unknown kind");
  if (kind == SLA_GEN_BLOCK16 && n % 16 != 0) return sla_fail(c, SLA_ERR_INVALID, "generate: block16 needs n to be a multiple of 16");
  if (kind == SLA_GEN_LAPLACE2D && band * band != n) return sla_fail(c, SLA_ERR_INVALID, "generate: laplace2d needs n = band*band");
  if (kind == SLA_GEN_BANDED && band < 1) return sla_fail(c, SLA_ERR_INVALID, "generate: banded needs band >= 1");
  *out = nullptr;
  const int64_t m = row_hi - row_lo;
  if ((int64_t)nnz_per_row * m >= (1LL << 31) - 2 * SLA_SPMV_TILE) return sla_fail(c, SLA_ERR_INVALID, "generate: nnz exceeds int32 indexing");
  DevBuf len, rp;
  SLA_CUDA(c, len.alloc(sizeof(int32_t) * (size_t)(m + 1)));
  SLA_CUDA(c, rp.alloc(sizeof(int32_t) * (size_t)(m + 1)));
  int32_t nnz32 = 0;
  if (m > 0) {
    synth_len_kernel<<<gs_blocks(m), 256, 0, c->stream>>>(kind, n, nnz_per_row, band, row_lo, m, len.as<int32_t>());
    SLA_LAUNCH_CHECK(c);
    size_t tb = 0;
    SLA_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tb, len.as<int32_t>(), rp.as<int32_t>(), (int)m, c->stream));
    DevBuf tmp; SLA_CUDA(c, tmp.alloc(tb));
    SLA_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, tb, len.as<int32_t>(), rp.as<int32_t>(), (int)m, c->stream));
    c->launches += 2;
    set_last_rowptr_kernel<<<1, 32, 0, c->stream>>>(rp.as<int32_t>(), len.as<int32_t>(), m);
    SLA_LAUNCH_CHECK(c);
    SLA_CUDA(c, cudaMemcpyAsync(&nnz32, rp.as<int32_t>() + m, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  } else {
    SLA_CUDA(c, cudaMemsetAsync(rp.p, 0, sizeof(int32_t), c->stream));
  }
  sla_csr* A = nullptr;
  SLA_TRY(sla_csr_alloc(c, m, n, (int64_t)nnz32, &A));
  SLA_CUDA(c, cudaMemcpyAsync(A->row_ptr, rp.p, sizeof(int32_t) * (size_t)(m + 1), cudaMemcpyDeviceToDevice, c->stream));
  if (m > 0) {
    synth_fill_kernel<<<gs_blocks(m), 256, 0, c->stream>>>(kind, n, nnz_per_row, seed, band, row_lo, m, A->row_ptr, A->col, A->val);
    SLA_LAUNCH_CHECK(c);
  }
  sla_status s = sla_csr_build_plan(c, A);
  if (s != SLA_OK) { sla_csr_free(A); return s; }
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  *out = A;
  return SLA_OK;
}

extern "C" sla_status sla_csr_generate(sla_ctx* c, int kind, int64_t n, int nnz_per_row, uint64_t seed, int64_t band,
                                       sla_csr** out) {
  return sla_csr_generate_rows(c, kind, n, nnz_per_row, seed, band, 0, n, out);
}

// entries [i0, i0 + n) of the synthetic vector with this seed (i0 = 0 on a single GPU)
extern "C" sla_status sla_vec_generate_slice(sla_ctx* c, int64_t i0, int64_t n, uint64_t seed, sla_vec** out) {
  SLA_TRY(sla_vec_create(c, n, out));
  if (n > 0) {
    synth_vec_kernel<<<gs_blocks(n), 256, 0, c->stream>>>(seed, i0, n, (*out)->d);
    SLA_LAUNCH_CHECK(c);
  }
  return SLA_OK;
}

extern "C" sla_status sla_vec_generate(sla_ctx* c, int64_t n, uint64_t seed, sla_vec** out) {
  return sla_vec_generate_slice(c, 0, n, seed, out);
}
