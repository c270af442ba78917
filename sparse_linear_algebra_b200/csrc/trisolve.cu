// trisolve.cu — the step either side of the Krylov loop:
//   diagPartitions / extractSubDiag / extractDiag / extractSuperDiag     Sparse.hs:673-679, SpMatrix.hs:306-315
//   jacobiPre                                                            Sparse.hs:686-687
//   mSsorPre                                                             Sparse.hs:713-721
//   triLowerSolve / triUpperSolve                                        Sparse.hs:750-811
//
// The triangular solves are the part with a data dependency: row i needs every w_j it references.  The schedule:
//   analysis (once per matrix and direction, cached in the matrix)
//     level(i) = 1 + max level(j) over the referenced j; computed by ONE launch in which a thread owns a row and
//     polls the levels it depends on (0 = not known yet).  CTAs take 256-row chunks in sweep order from a ticket,
//     so every dependency belongs to a CTA that is already running: no deadlock, no host loop over levels.
//     Dependencies inside the CTA's own chunk (the i-1 neighbour of a stencil) are polled in shared memory.
//     Rows are then sorted by (level, row) with a stable radix sort.
//   solve (two launches: sentinel fill, sweep)
//     thread p owns row order[p]; a fixed set of persistent warps takes 32-position chunks from a ticket, again in
//     dependency order; the number of rows in flight is ~4 levels of average width (measured on the 4096^2 stencil:
//     7.8 ms against 22.5 ms for one CTA per 128 positions, profiles/r01_sptrsv_sweep.txt).  The
//     partial solution itself is the ready flag: w starts as a NaN pattern no arithmetic produces, a consumer polls
//     w_j with ld.relaxed.gpu until it changes.  Up to four dependencies are polled per round so that a row whose
//     dependencies are ready costs one L2 round trip, and they are consumed strictly in ascending column order
//     with __dmul_rn / __dadd_rn — the reference's left fold — so the result is bit-identical to the reference's evaluation order.
//   Every poll round is non-blocking and a row is published inside the round that completes it, so lanes of one
//   warp may depend on each other (level boundaries fall anywhere).
// The sweep is latency-bound by construction: (number of levels) x (store -> L2 -> poll), see DESIGN.md.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <new>
#include <stdlib.h>

#define TRI_LVL_THREADS 256
#define TRI_SOLVE_THREADS 128
#define TRI_POLL 4
#ifndef TRI_MODE
#define TRI_MODE 1              // solve kernel: 0 = one CTA per 128 positions, 1 = persistent warps (env SLA_TRI_MODE overrides)
#endif
#define TRI_LEVELS_IN_FLIGHT 4  // persistent mode: rows in flight = this many levels of average width (env SLA_TRI_LIF)
#ifndef TRI_BACKOFF_NS
#define TRI_BACKOFF_NS 0        // nanosleep of a warp that made no progress in a poll round (env SLA_TRI_BACKOFF overrides)
#endif
#define TRI_SENTINEL 0xFFF75EEDDEADBEEFULL
#define TRI_CANONICAL_NAN 0x7FF8000000000000ULL

struct sla_tri_plan {
  int nlevels;
  int64_t nnz_tri;       // stored entries of the triangle, diagonal included
  int64_t bad_row;       // first row (in sweep order) whose diagonal is missing or nearZero, else -1
  int32_t* order;        // rows sorted by (level, row)
};

namespace {

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <class T> T* as() { return (T*)p; }
};

#define GS_LOOP(i, n) for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)
inline unsigned gs_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > SLA_NUM_SMS * 16) b = SLA_NUM_SMS * 16;
  return (unsigned)b;
}
inline int bits_for(int64_t v) { int b = 1; while ((1LL << b) < v && b < 32) ++b; return b; }

__device__ __forceinline__ int ld_relaxed_i32(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_i32(int* p, int v) {
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const double* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(double* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// first k in [lo, hi) with col[k] >= row
__device__ __forceinline__ int diag_lower_bound(const int32_t* __restrict__ col, int lo, int hi, int row) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (col[mid] < row) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// nearZero a = abs a <= 1e-12   Eps.hs:41-42
__device__ __forceinline__ bool near_zero(double a) { return fabs(a) <= 1e-12; }

// ---- analysis ------------------------------------------------------------------------------------------

// stats[0] = max level, stats[1] = bad row (min for the forward sweep, max for the backward one)
template <bool UPPER>
__global__ void __launch_bounds__(TRI_LVL_THREADS)
tri_levels_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col, const double* __restrict__ val,
                  int n, const int32_t* __restrict__ diag_idx, int* level, unsigned int* ticket, int* stats,
                  unsigned long long* nnz_tri) {
  __shared__ unsigned int s_chunk;
  __shared__ volatile int s_lvl[TRI_LVL_THREADS];
  if (threadIdx.x == 0) s_chunk = atomicAdd(ticket, 1u);
  s_lvl[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)s_chunk * TRI_LVL_THREADS;
  const int64_t p = base + threadIdx.x;          // position in sweep order
  bool done = p >= n;
  int row = 0, k = 0, k1 = 0, lvl = 0, cnt = 0;
  if (!done) {
    row = UPPER ? (int)(n - 1 - p) : (int)p;
    const int lo = row_ptr[row], hi = row_ptr[row + 1];
    const int d = diag_idx[row];
    const bool has = d < hi && col[d] == row;
    if (!has || near_zero(val[d])) { if (UPPER) atomicMax(&stats[1], row); else atomicMin(&stats[1], row); }
    k = UPPER ? d + (has ? 1 : 0) : lo;
    k1 = UPPER ? hi : d;
    cnt = (k1 - k) + (has ? 1 : 0);
  }
  while (__any_sync(0xffffffffu, !done)) {
    if (!done) {
      if (k < k1) {
        const int j = col[k];
        const int64_t pj = UPPER ? (int64_t)(n - 1 - j) : (int64_t)j;
        const int lj = pj >= base ? s_lvl[pj - base] : ld_relaxed_i32(level + j);
        if (lj) { lvl = max(lvl, lj); ++k; }
      }
      if (k >= k1) {
        s_lvl[threadIdx.x] = lvl + 1;
        st_relaxed_i32(level + row, lvl + 1);
        done = true;
      }
    }
  }
  int mx = p < n ? lvl + 1 : 0;
  unsigned long long sum = (unsigned long long)cnt;
  for (int o = 16; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicMax(&stats[0], mx); atomicAdd(nnz_tri, sum); }
}

__global__ void tri_diag_idx_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col, int64_t n,
                                    int32_t* __restrict__ diag_idx) {
  GS_LOOP(i, n) diag_idx[i] = diag_lower_bound(col, row_ptr[i], row_ptr[i + 1], (int)i);
}

__global__ void tri_iota_kernel(int32_t* v, int64_t n) { GS_LOOP(i, n) v[i] = (int32_t)i; }

// ---- solve ---------------------------------------------------------------------------------------------

__global__ void tri_fill_kernel(double* w, int64_t n, unsigned int* ticket) {
  GS_LOOP(i, n) reinterpret_cast<unsigned long long*>(w)[i] = TRI_SENTINEL;
  if (blockIdx.x == 0 && threadIdx.x == 0) *ticket = 0u;
}

template <bool UPPER>
__global__ void __launch_bounds__(TRI_SOLVE_THREADS)
tri_solve_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col, const double* __restrict__ val,
                 const int32_t* __restrict__ diag_idx, const int32_t* __restrict__ order, const double* b,
                 double* wraw, double* out, int64_t n, unsigned int* ticket, unsigned int backoff_ns) {
  __shared__ unsigned int s_chunk;
  if (threadIdx.x == 0) s_chunk = atomicAdd(ticket, 1u);
  __syncthreads();
  const int64_t p = (int64_t)s_chunk * TRI_SOLVE_THREADS + threadIdx.x;
  bool done = p >= n;
  int row = 0, k = 0, k1 = 0;
  double acc = 0.0, dv = 1.0, bi = 0.0;
  if (!done) {
    row = order[p];
    const int lo = row_ptr[row], hi = row_ptr[row + 1], d = diag_idx[row];
    dv = val[d];                         // the plan verified that the diagonal is stored and not nearZero
    bi = b[row];
    k = UPPER ? d + 1 : lo;
    k1 = UPPER ? hi : d;
  }
  unsigned int sleep_ns = backoff_ns;
  while (__any_sync(0xffffffffu, !done)) {
    bool progressed = false;
    if (!done) {
      if (k < k1) {
        unsigned long long bits[TRI_POLL];
        double a[TRI_POLL];
#pragma unroll
        for (int q = 0; q < TRI_POLL; ++q) {
          bits[q] = TRI_SENTINEL;
          a[q] = 0.0;
          if (k + q < k1) {
            a[q] = val[k + q];
            bits[q] = ld_relaxed_u64(wraw + col[k + q]);
          }
        }
        bool go = true;
#pragma unroll
        for (int q = 0; q < TRI_POLL; ++q) {
          go = go && bits[q] != TRI_SENTINEL;
          if (go) {                      // r = sum of l_ij * w_j, ascending j, strict left fold from 0   Sparse.hs:762, 795
            acc = __dadd_rn(acc, __dmul_rn(a[q], __longlong_as_double((long long)bits[q])));
            ++k;
            progressed = true;
          }
        }
      }
      if (k >= k1) {
        const double w = __ddiv_rn(__dsub_rn(bi, acc), dv);        // wi = (bi - r) / lii   Sparse.hs:761, 794
        unsigned long long wb = (unsigned long long)__double_as_longlong(w);
        if (wb == TRI_SENTINEL) wb = TRI_CANONICAL_NAN;
        st_relaxed_u64(wraw + row, wb);
        out[row] = near_zero(w) ? 0.0 : w;                         // sparsifySV   Sparse.hs:777, 811
        done = true;
        progressed = true;
      }
    }
    // a warp whose lanes all wait for rows of other CTAs steps back instead of hammering the L2 request port
    // (ncu, 4096^2 Laplacian: the polls alone kept l1tex2xbar 83 % busy and slowed the runnable level down)
    // exponentially, up to 16 x the base, so that rows far ahead of the wavefront poll rarely.
    if (backoff_ns) {
      if (__any_sync(0xffffffffu, progressed)) sleep_ns = backoff_ns;
      else { __nanosleep(sleep_ns); sleep_ns = min(sleep_ns * 2u, backoff_ns * 16u); }
    }
  }
}

// Persistent variant: a fixed number of warps, each taking 32-position chunks from the ticket until the sweep is
// done.  The number of rows in flight is chosen by the host from the average level width (a few levels' worth):
// on a deep, narrow schedule (the 4096^2 stencil: 8191 levels of <= 4096 rows) the non-persistent kernel keeps
// ~55 levels of rows resident, all polling, and the polls of the 54 levels that cannot run yet saturate the
// L1 -> L2 request port (ncu: l1tex2xbar 83 % busy, 168 GB of poll traffic) and slow the one level that can.
// Deadlock-free for any grid size: a position is only waited for by higher positions, and it was claimed
// (ticket) by a warp that is running.
template <bool UPPER>
__global__ void __launch_bounds__(TRI_SOLVE_THREADS)
tri_solve_persist_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col, const double* __restrict__ val,
                         const int32_t* __restrict__ diag_idx, const int32_t* __restrict__ order, const double* b,
                         double* wraw, double* out, int64_t n, unsigned int* ticket, unsigned int backoff_ns) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    unsigned int chunk = 0;
    if (lane == 0) chunk = atomicAdd(ticket, 1u);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if ((int64_t)chunk * 32 >= n) break;
    const int64_t p = (int64_t)chunk * 32 + lane;
    bool done = p >= n;
    int row = 0, k = 0, k1 = 0;
    double acc = 0.0, dv = 1.0, bi = 0.0;
    if (!done) {
      row = order[p];
      const int lo = row_ptr[row], hi = row_ptr[row + 1], d = diag_idx[row];
      dv = val[d];
      bi = b[row];
      k = UPPER ? d + 1 : lo;
      k1 = UPPER ? hi : d;
    }
    unsigned int sleep_ns = backoff_ns;
    while (__any_sync(0xffffffffu, !done)) {
      bool progressed = false;
      if (!done) {
        if (k < k1) {
          unsigned long long bits[TRI_POLL];
          double a[TRI_POLL];
#pragma unroll
          for (int q = 0; q < TRI_POLL; ++q) {
            bits[q] = TRI_SENTINEL;
            a[q] = 0.0;
            if (k + q < k1) {
              a[q] = val[k + q];
              bits[q] = ld_relaxed_u64(wraw + col[k + q]);
            }
          }
          bool go = true;
#pragma unroll
          for (int q = 0; q < TRI_POLL; ++q) {
            go = go && bits[q] != TRI_SENTINEL;
            if (go) {                    // ascending j, strict left fold from 0   Sparse.hs:762, 795
              acc = __dadd_rn(acc, __dmul_rn(a[q], __longlong_as_double((long long)bits[q])));
              ++k;
              progressed = true;
            }
          }
        }
        if (k >= k1) {
          const double w = __ddiv_rn(__dsub_rn(bi, acc), dv);      // wi = (bi - r) / lii   Sparse.hs:761, 794
          unsigned long long wb = (unsigned long long)__double_as_longlong(w);
          if (wb == TRI_SENTINEL) wb = TRI_CANONICAL_NAN;
          st_relaxed_u64(wraw + row, wb);
          out[row] = near_zero(w) ? 0.0 : w;                       // sparsifySV   Sparse.hs:777, 811
          done = true;
          progressed = true;
        }
      }
      if (backoff_ns) {
        if (__any_sync(0xffffffffu, progressed)) sleep_ns = backoff_ns;
        else { __nanosleep(sleep_ns); sleep_ns = min(sleep_ns * 2u, backoff_ns * 16u); }
      }
    }
  }
}

// ---- partitions and preconditioners ----------------------------------------------------------------------

// cnt[0..m] sub-diagonal, cnt[m+1 ..] diagonal, cnt[2(m+1) ..] super-diagonal entries per row (slot m of each = 0)
__global__ void part_count_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col, int64_t m,
                                  int32_t* __restrict__ cnt) {
  GS_LOOP(i, m + 1) {
    int e = 0, dd = 0, f = 0;
    if (i < m) {
      const int lo = row_ptr[i], hi = row_ptr[i + 1];
      const int d = diag_lower_bound(col, lo, hi, (int)i);
      dd = (d < hi && col[d] == (int)i) ? 1 : 0;
      e = d - lo;
      f = hi - d - dd;
    }
    cnt[i] = e; cnt[(m + 1) + i] = dd; cnt[2 * (m + 1) + i] = f;
  }
}

// which: -1 sub, 0 diagonal, +1 super.  op: 0 copy, 1 reciprocal (jacobiPre: recip <$> extractDiag)
__global__ void part_fill_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                 const double* __restrict__ val, int64_t m, int which, int op,
                                 const int32_t* __restrict__ out_ptr, int32_t* __restrict__ out_col, double* __restrict__ out_val) {
  GS_LOOP(i, m) {
    const int lo = row_ptr[i], hi = row_ptr[i + 1];
    const int d = diag_lower_bound(col, lo, hi, (int)i);
    const int dd = (d < hi && col[d] == (int)i) ? 1 : 0;
    const int s = which < 0 ? lo : which == 0 ? d : d + dd;
    const int e = which < 0 ? d : which == 0 ? d + dd : hi;
    int o = out_ptr[i];
    for (int k = s; k < e; ++k, ++o) {
      out_col[o] = col[k];
      out_val[o] = op == 1 ? __ddiv_rn(1.0, val[k]) : val[k];
    }
  }
}

// reciprocal d as a dense array: rd[i] = recip d_ii, has[i] = the diagonal entry is stored
__global__ void diag_recip_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                  const double* __restrict__ val, int64_t n, double* __restrict__ rd, unsigned char* __restrict__ has) {
  GS_LOOP(i, n) {
    const int lo = row_ptr[i], hi = row_ptr[i + 1];
    const int d = diag_lower_bound(col, lo, hi, (int)i);
    const bool h = d < hi && col[d] == (int)i;
    has[i] = h ? 1 : 0;
    rd[i] = h ? __ddiv_rn(1.0, val[d]) : 0.0;
  }
}

// l = (eye n ^-^ scale omega e) ## reciprocal d : entries (i, c) with c <= i whose column has a stored d_cc
// r = d ^-^ scale omega f                       : the stored diagonal and the super-diagonal entries
__global__ void mssor_count_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col, int64_t n,
                                   const unsigned char* __restrict__ has, int32_t* __restrict__ cnt) {
  GS_LOOP(i, n + 1) {
    int l = 0, r = 0;
    if (i < n) {
      const int lo = row_ptr[i], hi = row_ptr[i + 1];
      const int d = diag_lower_bound(col, lo, hi, (int)i);
      for (int k = lo; k < d; ++k) l += has[col[k]];
      l += has[i];
      r = hi - d;
    }
    cnt[i] = l; cnt[(n + 1) + i] = r;
  }
}

__global__ void mssor_fill_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                  const double* __restrict__ val, int64_t n, double omega, const double* __restrict__ rd,
                                  const unsigned char* __restrict__ has, const int32_t* __restrict__ l_ptr,
                                  int32_t* __restrict__ l_col, double* __restrict__ l_val, const int32_t* __restrict__ r_ptr,
                                  int32_t* __restrict__ r_col, double* __restrict__ r_val) {
  GS_LOOP(i, n) {
    const int lo = row_ptr[i], hi = row_ptr[i + 1];
    const int d = diag_lower_bound(col, lo, hi, (int)i);
    int o = l_ptr[i];
    for (int k = lo; k < d; ++k) {
      const int cidx = col[k];
      if (!has[cidx]) continue;
      // row entry of (eye ^+^ negateV (scale omega e)) = negate (e_ic * omega) ; dott = sum [b_cc * a_ic] from 0
      const double a = -__dmul_rn(val[k], omega);
      l_col[o] = cidx;
      l_val[o] = __dadd_rn(0.0, __dmul_rn(rd[cidx], a));
      ++o;
    }
    if (has[i]) { l_col[o] = (int)i; l_val[o] = __dadd_rn(0.0, __dmul_rn(rd[i], 1.0)); }
    o = r_ptr[i];
    for (int k = d; k < hi; ++k, ++o) {
      r_col[o] = col[k];
      r_val[o] = col[k] == (int)i ? val[k] : -__dmul_rn(val[k], omega);
    }
  }
}

sla_status scan_counts(sla_ctx* c, const int32_t* cnt, int32_t* out, int64_t len) {
  size_t tmp_bytes = 0;
  SLA_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, out, (int)len, c->stream));
  DevBuf tmp;
  SLA_CUDA(c, tmp.alloc(tmp_bytes));
  SLA_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cnt, out, (int)len, c->stream));
  c->launches += 2;
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));     // tmp is freed on return
  return SLA_OK;
}

sla_status single_gpu_only(sla_ctx* c, const sla_csr* A, const char* what) {
  if (!c || !A) return SLA_ERR_INVALID;
  if (A->dist || c->world > 1) {
    snprintf(c->err, sizeof(c->err), "%s: row-partitioned matrices are not supported (single GPU only)", what);
    return SLA_ERR_INVALID;
  }
  return SLA_OK;
}

// allocates a matrix whose row_ptr is the scanned count array `ptr` (device, m + 1 entries)
sla_status alloc_from_ptr(sla_ctx* c, int64_t m, int64_t n, const int32_t* ptr, sla_csr** out) {
  int32_t total = 0;
  SLA_CUDA(c, cudaMemcpyAsync(&total, ptr + m, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  SLA_TRY(sla_csr_alloc(c, m, n, total, out));
  SLA_CUDA(c, cudaMemcpyAsync((*out)->row_ptr, ptr, sizeof(int32_t) * (size_t)(m + 1), cudaMemcpyDeviceToDevice, c->stream));
  return SLA_OK;
}

sla_status finish_matrix(sla_ctx* c, sla_csr* M) {
  SLA_TRY(sla_csr_build_plan(c, M));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  return SLA_OK;
}

// extracts one partition (which = -1, 0, +1) with op applied to the values
sla_status extract_part(sla_ctx* c, const sla_csr* A, const int32_t* ptrs, int which, int op, sla_csr** out) {
  const int64_t m = A->m;
  const int32_t* ptr = ptrs + (size_t)(which + 1) * (size_t)(m + 1);
  sla_csr* M = nullptr;
  SLA_TRY(alloc_from_ptr(c, m, A->n, ptr, &M));
  if (m > 0) {
    part_fill_kernel<<<gs_blocks(m), 256, 0, c->stream>>>(A->row_ptr, A->col, A->val, m, which, op, M->row_ptr, M->col, M->val);
    c->launches++;
    if (cudaGetLastError() != cudaSuccess) { sla_csr_free(M); return sla_fail(c, SLA_ERR_CUDA, "part_fill_kernel launch failed"); }
  }
  sla_status s = finish_matrix(c, M);
  if (s != SLA_OK) { sla_csr_free(M); return s; }
  *out = M;
  return SLA_OK;
}

sla_status partition_ptrs(sla_ctx* c, const sla_csr* A, DevBuf& ptrs) {
  const int64_t m = A->m;
  DevBuf cnt;
  SLA_CUDA(c, cnt.alloc(sizeof(int32_t) * 3 * (size_t)(m + 1)));
  SLA_CUDA(c, ptrs.alloc(sizeof(int32_t) * 3 * (size_t)(m + 1)));
  part_count_kernel<<<gs_blocks(m + 1), 256, 0, c->stream>>>(A->row_ptr, A->col, m, cnt.as<int32_t>());
  SLA_LAUNCH_CHECK(c);
  for (int q = 0; q < 3; ++q)
    SLA_TRY(scan_counts(c, cnt.as<int32_t>() + (size_t)q * (m + 1), ptrs.as<int32_t>() + (size_t)q * (m + 1), m + 1));
  return SLA_OK;
}

sla_status build_tri_plan(sla_ctx* c, sla_csr* A, int upper) {
  if (A->tri[upper]) return SLA_OK;
  const int64_t n = A->m;
  if (!A->tri_diag) {
    SLA_CUDA(c, cudaMalloc(&A->tri_diag, sizeof(int32_t) * (size_t)(n > 0 ? n : 1)));
    SLA_CUDA(c, cudaMalloc(&A->tri_w, sizeof(double) * (size_t)(n > 0 ? n : 1)));
    SLA_CUDA(c, cudaMalloc(&A->tri_ticket, sizeof(unsigned int)));
    if (n > 0) {
      tri_diag_idx_kernel<<<gs_blocks(n), 256, 0, c->stream>>>(A->row_ptr, A->col, n, A->tri_diag);
      SLA_LAUNCH_CHECK(c);
    }
  }
  DevBuf level, rows, keys_out, stats, cnt;
  SLA_CUDA(c, level.alloc(sizeof(int) * (size_t)n)); SLA_CUDA(c, rows.alloc(sizeof(int32_t) * (size_t)n));
  SLA_CUDA(c, keys_out.alloc(sizeof(int) * (size_t)n));
  SLA_CUDA(c, stats.alloc(sizeof(int) * 2)); SLA_CUDA(c, cnt.alloc(sizeof(unsigned long long)));
  sla_tri_plan* P = new (std::nothrow) sla_tri_plan();
  if (!P) return sla_fail(c, SLA_ERR_ALLOC, "tri plan alloc");
  P->nlevels = 0; P->nnz_tri = 0; P->bad_row = -1; P->order = nullptr;
  cudaError_t e = cudaMalloc(&P->order, sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  if (e != cudaSuccess) { delete P; cudaGetLastError(); return sla_fail(c, SLA_ERR_ALLOC, "cudaMalloc failed for a triangular-solve schedule"); }
  sla_status s = SLA_OK;
  int h_stats[2] = {0, upper ? -1 : 0x7fffffff};
  unsigned long long h_cnt = 0;
  do {
    if (n == 0) break;
    if ((e = cudaMemsetAsync(level.p, 0, sizeof(int) * (size_t)n, c->stream)) != cudaSuccess) break;
    if ((e = cudaMemsetAsync(A->tri_ticket, 0, sizeof(unsigned int), c->stream)) != cudaSuccess) break;
    if ((e = cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), c->stream)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(stats.p, h_stats, sizeof(h_stats), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) break;
    const unsigned grid = (unsigned)((n + TRI_LVL_THREADS - 1) / TRI_LVL_THREADS);
    if (upper)
      tri_levels_kernel<true><<<grid, TRI_LVL_THREADS, 0, c->stream>>>(A->row_ptr, A->col, A->val, (int)n, A->tri_diag, level.as<int>(),
                                                                       A->tri_ticket, stats.as<int>(), cnt.as<unsigned long long>());
    else
      tri_levels_kernel<false><<<grid, TRI_LVL_THREADS, 0, c->stream>>>(A->row_ptr, A->col, A->val, (int)n, A->tri_diag, level.as<int>(),
                                                                        A->tri_ticket, stats.as<int>(), cnt.as<unsigned long long>());
    c->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    tri_iota_kernel<<<gs_blocks(n), 256, 0, c->stream>>>(rows.as<int32_t>(), n);
    c->launches++;
    if ((e = cudaMemcpyAsync(h_stats, stats.p, sizeof(h_stats), cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(&h_cnt, cnt.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) break;
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) break;
    // stable sort by level keeps rows ascending inside a level
    size_t tmp_bytes = 0;
    const int end_bit = bits_for((int64_t)h_stats[0] + 1);
    if ((e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, level.as<int>(), keys_out.as<int>(), rows.as<int32_t>(), P->order,
                                             (int)n, 0, end_bit, c->stream)) != cudaSuccess) break;
    DevBuf tmp;
    if ((e = tmp.alloc(tmp_bytes)) != cudaSuccess) break;
    if ((e = cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, level.as<int>(), keys_out.as<int>(), rows.as<int32_t>(), P->order,
                                             (int)n, 0, end_bit, c->stream)) != cudaSuccess) break;
    c->launches += 8;
    e = cudaStreamSynchronize(c->stream);
  } while (0);
  if (e != cudaSuccess) {
    snprintf(c->err, sizeof(c->err), "CUDA error %s while building a triangular-solve schedule", cudaGetErrorString(e));
    cudaGetLastError();
    cudaFree(P->order); delete P;
    return SLA_ERR_CUDA;
  }
  P->nlevels = h_stats[0];
  P->nnz_tri = (int64_t)h_cnt;
  P->bad_row = upper ? (int64_t)h_stats[1] : (h_stats[1] == 0x7fffffff ? -1 : (int64_t)h_stats[1]);
  A->tri[upper] = P;
  return s;
}

sla_status tri_solve(sla_ctx* c, const sla_csr* A_, const sla_vec* b, sla_vec* x, int upper) {
  const char* name = upper ? "triUpperSolve" : "triLowerSolve";
  if (!c || !A_ || !b || !x) return SLA_ERR_INVALID;
  SLA_TRY(single_gpu_only(c, A_, name));
  sla_csr* A = const_cast<sla_csr*>(A_);
  const int64_t n = A->m;
  if (A->m != A->n || b->n != n || x->n != n) {
    snprintf(c->err, sizeof(c->err), "%s : mismatched dimensions (%lld x %lld) vs %lld", name, (long long)A->m, (long long)A->n, (long long)b->n);
    return SLA_ERR_SIZE_MISMATCH;
  }
  if (n == 0) return sla_fail(c, SLA_ERR_OOB_INDEX, "@@ : incompatible indices : matrix size is (0,0), but user looked up (0,0)");
  SLA_TRY(build_tri_plan(c, A, upper));
  const sla_tri_plan* P = (const sla_tri_plan*)A->tri[upper];
  if (P->bad_row >= 0) {
    // NeedsPivoting "triLowerSolve" "L (i,i)" ; the backward sweep's first test reports (0,0) whatever the row (Sparse.hs:802)
    const long long shown = (upper && P->bad_row == n - 1) ? 0 : (long long)P->bad_row;
    snprintf(c->err, sizeof(c->err), "%s : %s (%lld,%lld) is close to 0. Permute the rows to obtain a nonzero diagonal",
             name, upper ? "U" : "L", shown, shown);
    return SLA_ERR_NEEDS_PIVOTING;
  }
  if (n == 1) {
    // modifyUntilM' steps before it tests (Iterative.hs:272-282): the step after the initial one looks up (1,1) / (-1,-1)
    snprintf(c->err, sizeof(c->err), "@@ : incompatible indices : matrix size is (1,1), but user looked up %s", upper ? "(-1,-1)" : "(1,1)");
    return SLA_ERR_OOB_INDEX;
  }
  tri_fill_kernel<<<gs_blocks(n), 256, 0, c->stream>>>(A->tri_w, n, A->tri_ticket);
  SLA_LAUNCH_CHECK(c);
  const unsigned grid = (unsigned)((n + TRI_SOLVE_THREADS - 1) / TRI_SOLVE_THREADS);
  unsigned int backoff = TRI_BACKOFF_NS;
  if (const char* e = getenv("SLA_TRI_BACKOFF")) backoff = (unsigned int)atoi(e);
  int mode = TRI_MODE;                   // 0: one CTA per 128 positions; 1: persistent warps sized from the level width
  if (const char* e = getenv("SLA_TRI_MODE")) mode = atoi(e);
  if (mode == 1) {
    // rows in flight ~ TRI_LEVELS_IN_FLIGHT levels of average width, at least one CTA per SM, at most full residency
    int lif = TRI_LEVELS_IN_FLIGHT;
    if (const char* e = getenv("SLA_TRI_LIF")) lif = atoi(e);
    const int64_t width = (n + P->nlevels - 1) / (P->nlevels > 0 ? P->nlevels : 1);
    int64_t ctas = (width * lif + TRI_SOLVE_THREADS - 1) / TRI_SOLVE_THREADS;
    if (ctas < SLA_NUM_SMS) ctas = SLA_NUM_SMS;
    if (ctas > SLA_NUM_SMS * 12) ctas = SLA_NUM_SMS * 12;
    if (ctas > grid) ctas = grid;
    if (upper)
      tri_solve_persist_kernel<true><<<(unsigned)ctas, TRI_SOLVE_THREADS, 0, c->stream>>>(A->row_ptr, A->col, A->val, A->tri_diag, P->order,
                                                                                         b->d, A->tri_w, x->d, n, A->tri_ticket, backoff);
    else
      tri_solve_persist_kernel<false><<<(unsigned)ctas, TRI_SOLVE_THREADS, 0, c->stream>>>(A->row_ptr, A->col, A->val, A->tri_diag, P->order,
                                                                                          b->d, A->tri_w, x->d, n, A->tri_ticket, backoff);
  } else if (upper)
    tri_solve_kernel<true><<<grid, TRI_SOLVE_THREADS, 0, c->stream>>>(A->row_ptr, A->col, A->val, A->tri_diag, P->order, b->d, A->tri_w,
                                                                      x->d, n, A->tri_ticket, backoff);
  else
    tri_solve_kernel<false><<<grid, TRI_SOLVE_THREADS, 0, c->stream>>>(A->row_ptr, A->col, A->val, A->tri_diag, P->order, b->d, A->tri_w,
                                                                       x->d, n, A->tri_ticket, backoff);
  SLA_LAUNCH_CHECK(c);
  sla_touch(x);
  return SLA_OK;
}

}  // namespace

void sla_csr_free_tri(sla_csr* A) {
  for (int q = 0; q < 2; ++q) {
    sla_tri_plan* P = (sla_tri_plan*)A->tri[q];
    if (P) { cudaFree(P->order); delete P; }
    A->tri[q] = nullptr;
  }
  cudaFree(A->tri_diag); cudaFree(A->tri_w); cudaFree(A->tri_ticket);
  A->tri_diag = nullptr; A->tri_w = nullptr; A->tri_ticket = nullptr;
}

extern "C" sla_status sla_tri_lower_solve(sla_ctx* c, const sla_csr* L, const sla_vec* b, sla_vec* w) { return tri_solve(c, L, b, w, 0); }
extern "C" sla_status sla_tri_upper_solve(sla_ctx* c, const sla_csr* U, const sla_vec* w, sla_vec* x) { return tri_solve(c, U, w, x, 1); }

extern "C" sla_status sla_tri_analysis(sla_ctx* c, const sla_csr* A, int upper, int* nlevels, int64_t* nnz_tri) {
  if (!c || !A) return SLA_ERR_INVALID;
  SLA_TRY(single_gpu_only(c, A, "sla_tri_analysis"));
  if (A->m != A->n) return sla_fail(c, SLA_ERR_SIZE_MISMATCH, "sla_tri_analysis: the matrix must be square");
  upper = upper ? 1 : 0;
  SLA_TRY(build_tri_plan(c, const_cast<sla_csr*>(A), upper));
  const sla_tri_plan* P = (const sla_tri_plan*)A->tri[upper];
  if (nlevels) *nlevels = P->nlevels;
  if (nnz_tri) *nnz_tri = P->nnz_tri;
  return SLA_OK;
}

extern "C" sla_status sla_csr_diag_partitions(sla_ctx* c, const sla_csr* A, sla_csr** E, sla_csr** D, sla_csr** F) {
  if (!c || !A || !E || !D || !F) return SLA_ERR_INVALID;
  *E = *D = *F = nullptr;
  SLA_TRY(single_gpu_only(c, A, "diagPartitions"));
  DevBuf ptrs;
  SLA_TRY(partition_ptrs(c, A, ptrs));
  sla_csr* out[3] = {nullptr, nullptr, nullptr};
  for (int q = 0; q < 3; ++q) {
    sla_status s = extract_part(c, A, ptrs.as<int32_t>(), q - 1, 0, &out[q]);
    if (s != SLA_OK) { for (int r = 0; r < q; ++r) sla_csr_free(out[r]); return s; }
  }
  *E = out[0]; *D = out[1]; *F = out[2];
  return SLA_OK;
}

extern "C" sla_status sla_jacobi_pre(sla_ctx* c, const sla_csr* A, sla_csr** M) {
  if (!c || !A || !M) return SLA_ERR_INVALID;
  *M = nullptr;
  SLA_TRY(single_gpu_only(c, A, "jacobiPre"));
  DevBuf ptrs;
  SLA_TRY(partition_ptrs(c, A, ptrs));
  return extract_part(c, A, ptrs.as<int32_t>(), 0, 1, M);
}

extern "C" sla_status sla_mssor_pre(sla_ctx* c, const sla_csr* A, double omega, sla_csr** L, sla_csr** R) {
  if (!c || !A || !L || !R) return SLA_ERR_INVALID;
  *L = *R = nullptr;
  SLA_TRY(single_gpu_only(c, A, "mSsorPre"));
  if (A->m != A->n) {
    snprintf(c->err, sizeof(c->err), "matMat : incompatible matrix sizes((%lld,%lld),(%lld,%lld))", (long long)A->m, (long long)A->m,
             (long long)A->m, (long long)A->n);
    return SLA_ERR_SIZE_MISMATCH;
  }
  const int64_t n = A->m;
  DevBuf rd, has, cnt, ptrs;
  SLA_CUDA(c, rd.alloc(sizeof(double) * (size_t)n)); SLA_CUDA(c, has.alloc((size_t)n));
  SLA_CUDA(c, cnt.alloc(sizeof(int32_t) * 2 * (size_t)(n + 1))); SLA_CUDA(c, ptrs.alloc(sizeof(int32_t) * 2 * (size_t)(n + 1)));
  if (n > 0) {
    diag_recip_kernel<<<gs_blocks(n), 256, 0, c->stream>>>(A->row_ptr, A->col, A->val, n, rd.as<double>(), has.as<unsigned char>());
    SLA_LAUNCH_CHECK(c);
  }
  mssor_count_kernel<<<gs_blocks(n + 1), 256, 0, c->stream>>>(A->row_ptr, A->col, n, has.as<unsigned char>(), cnt.as<int32_t>());
  SLA_LAUNCH_CHECK(c);
  for (int q = 0; q < 2; ++q)
    SLA_TRY(scan_counts(c, cnt.as<int32_t>() + (size_t)q * (n + 1), ptrs.as<int32_t>() + (size_t)q * (n + 1), n + 1));
  sla_csr *l = nullptr, *r = nullptr;
  SLA_TRY(alloc_from_ptr(c, n, n, ptrs.as<int32_t>(), &l));
  sla_status s = alloc_from_ptr(c, n, n, ptrs.as<int32_t>() + (n + 1), &r);
  if (s != SLA_OK) { sla_csr_free(l); return s; }
  if (n > 0) {
    mssor_fill_kernel<<<gs_blocks(n), 256, 0, c->stream>>>(A->row_ptr, A->col, A->val, n, omega, rd.as<double>(), has.as<unsigned char>(),
                                                           l->row_ptr, l->col, l->val, r->row_ptr, r->col, r->val);
    c->launches++;
    if (cudaGetLastError() != cudaSuccess) s = sla_fail(c, SLA_ERR_CUDA, "mssor_fill_kernel launch failed");
  }
  if (s == SLA_OK) s = finish_matrix(c, l);
  if (s == SLA_OK) s = finish_matrix(c, r);
  if (s != SLA_OK) { sla_csr_free(l); sla_csr_free(r); return s; }
  *L = l; *R = r;
  return SLA_OK;
}
