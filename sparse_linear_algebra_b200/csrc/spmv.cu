// spmv.cu — fp64 CSR sparse matrix-vector product, the (#>) of the reference
// (matVecSD / dotu, src/Data/Sparse/Common.hs:242-260), as a tile-streamed sm_100a kernel.
//
// Decomposition.  The nnz stream (col_idx | val) is cut into fixed tiles of SLA_SPMV_TILE entries; tile t
// OWNS the rows whose first entry lies inside it (tile_row[t] .. tile_row[t+1], found once per matrix by
// sla_csr_build_plan).  A CTA
//   1. streams its tile with 128-bit loads (int4 of col_idx, 2 x double2 of val; L1 no-allocate, L2
//      evict-first so the matrix stream does not push x out of L2), gathers x[col] through the read-only
//      path and writes the products a_ij * x_j (__dmul_rn, no FMA) into shared memory,
//   2. sums each owned row from shared memory IN ASCENDING COLUMN ORDER with __dadd_rn from a 0.0 seed —
//      the reference's strict left fold — so rows of up to SLA_LONG_ROW entries are bit-identical to
//      the Haskell result; longer rows are summed by one warp (lane-strided partials + shuffle tree).
//      A row that runs past the tile end reads its tail straight from global memory.
//   3. optionally folds the row results into up to two dot products / a residual norm (the Krylov
//      epilogues), reduced over the grid deterministically by the last CTA to finish.
// The shared-memory product buffer is skewed by 2 doubles per 32 so that the per-row sequential reads of
// equal-length rows do not pile onto one bank while the 16-byte product stores stay aligned.
#include "common.cuh"

#define SLA_LONG_ROW 256
#define SLA_LONG_CAP 16
#define SPMV_THREADS 256

__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ int4 ld_stream_int4(const int* p, uint64_t pol) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ double2 ld_stream_double2(const double* p, uint64_t pol) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
               : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ double ld_keep_double(const double* p, uint64_t pol) {
  double r;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
  return r;
}

__device__ __forceinline__ int skew(int k) { return k + 2 * (k >> 5); }

template <int EPI>
__device__ __forceinline__ void row_epilogue(int r, double acc, double* __restrict__ y,
                                             const double* __restrict__ u0, double& e0, double& e1) {
  if (EPI == EPI_RESNORM) {
    double d = __dsub_rn(acc, u0[r]);          // (aa #> x) ^-^ b, then (**2)   Sparse.hs:1041
    e0 += d * d;
    return;
  }
  y[r] = acc;
  if (EPI == EPI_DOT1 || EPI == EPI_DOT2_YY) e0 += acc * u0[r];
  if (EPI == EPI_DOT2_YY) e1 += acc * acc;
}

template <int TILE, int EPI>
__global__ void __launch_bounds__(SPMV_THREADS)
spmv_tile_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, const double* __restrict__ val,
                 const double* __restrict__ x, double* __restrict__ y, const int* __restrict__ tile_row,
                 const double* __restrict__ u0, double* partials, unsigned int* counter, double* scal,
                 int fin, int dst) {
  constexpr int PER = TILE / (SPMV_THREADS * 4);          // 128-bit column groups per thread
  __shared__ __align__(16) double prod[TILE + 2 * (TILE / 32)];
  __shared__ double red[2 * 32];
  __shared__ int long_rows[SLA_LONG_CAP];
  __shared__ int n_long;

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int base = tile * TILE;
  const int row_lo = tile_row[tile], row_hi = tile_row[tile + 1];
  const int nrows = row_hi - row_lo;
  if (tid == 0) n_long = 0;

  // ---- phase 1: stream the tile, gather x, write products --------------------------------------
  if (nrows > 0) {
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_keep = policy_evict_last();
    int4 c[PER];
    double2 v[PER][2];
#pragma unroll
    for (int it = 0; it < PER; ++it) {
      const int p = (it * SPMV_THREADS + tid) * 4;
      c[it] = ld_stream_int4(col + base + p, pol_stream);
      v[it][0] = ld_stream_double2(val + base + p, pol_stream);
      v[it][1] = ld_stream_double2(val + base + p + 2, pol_stream);
    }
    double xv[PER][4];
#pragma unroll
    for (int it = 0; it < PER; ++it) {
      xv[it][0] = ld_keep_double(x + c[it].x, pol_keep);
      xv[it][1] = ld_keep_double(x + c[it].y, pol_keep);
      xv[it][2] = ld_keep_double(x + c[it].z, pol_keep);
      xv[it][3] = ld_keep_double(x + c[it].w, pol_keep);
    }
#pragma unroll
    for (int it = 0; it < PER; ++it) {
      const int p = (it * SPMV_THREADS + tid) * 4;
      double2 p0, p1;
      p0.x = __dmul_rn(v[it][0].x, xv[it][0]);             // dotu: a_ij * x_j, matrix entry on the left
      p0.y = __dmul_rn(v[it][0].y, xv[it][1]);
      p1.x = __dmul_rn(v[it][1].x, xv[it][2]);
      p1.y = __dmul_rn(v[it][1].y, xv[it][3]);
      double2* dstp = reinterpret_cast<double2*>(prod + skew(p));
      dstp[0] = p0;
      dstp[1] = p1;
    }
  }
  __syncthreads();

  // ---- phase 2: one thread per owned row, sequential ascending sum --------------------------------
  double e0 = 0.0, e1 = 0.0;
  for (int j = tid; j < nrows; j += SPMV_THREADS) {
    const int r = row_lo + j;
    const int s = row_ptr[r], e = row_ptr[r + 1];
    if (e - s > SLA_LONG_ROW) {
      const int slot = atomicAdd(&n_long, 1);
      long_rows[slot] = r;
      continue;
    }
    const int ks = s - base, ke = e - base;
    const int kin = ke < TILE ? ke : TILE;
    double acc = 0.0;                                        // sum = strict left fold from 0
    for (int k = ks; k < kin; ++k) acc = __dadd_rn(acc, prod[skew(k)]);
    for (int k = (ks > TILE ? ks : TILE); k < ke; ++k) {     // tail beyond the tile (last owned row only)
      const int g = base + k;
      acc = __dadd_rn(acc, __dmul_rn(val[g], x[col[g]]));
    }
    row_epilogue<EPI>(r, acc, y, u0, e0, e1);
  }

  // ---- long rows: one warp per row ---------------------------------------------------------------
  __syncthreads();
  const int nl = n_long;
  if (nl > 0) {
    const int lane = tid & 31, warp = tid >> 5;
    for (int q = warp; q < nl; q += SPMV_THREADS / 32) {
      const int r = long_rows[q];
      const int ks = row_ptr[r] - base, ke = row_ptr[r + 1] - base;
      double acc = 0.0;
      for (int k = ks + lane; k < ke; k += 32) {
        double t;
        if (k < TILE) t = prod[skew(k)];
        else { const int g = base + k; t = __dmul_rn(val[g], x[col[g]]); }
        acc += t;
      }
      acc = warp_sum(acc);
      if (lane == 0) row_epilogue<EPI>(r, acc, y, u0, e0, e1);
    }
  }

  if (EPI != EPI_NONE) {
    double sums[2] = {e0, e1};
    block_sum<2>(sums, red);
    grid_reduce_finish<2>(sums, partials, counter, scal, fin, dst, red);
  }
}

// tile_row[t] = first row r in [0, m] with row_ptr[r] >= t * TILE ; tile_row[ntiles] = m
__global__ void spmv_plan_kernel(const int* __restrict__ row_ptr, int m, int ntiles, int tile, int* __restrict__ tile_row) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles) return;
  if (t == ntiles) { tile_row[t] = m; return; }
  const long long target = (long long)t * tile;
  int lo = 0, hi = m;                       // answer in [0, m]
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if ((long long)row_ptr[mid] < target) lo = mid + 1; else hi = mid;
  }
  tile_row[t] = lo;
}

sla_status sla_csr_build_plan(sla_ctx* c, sla_csr* A) {
  const int nt = A->ntiles;
  if (A->tile_row == nullptr) SLA_CUDA(c, cudaMalloc(&A->tile_row, sizeof(int32_t) * (size_t)(nt + 1)));
  spmv_plan_kernel<<<(nt + 1 + 255) / 256, 256, 0, c->stream>>>(A->row_ptr, (int)A->m, nt, SLA_SPMV_TILE, A->tile_row);
  SLA_LAUNCH_CHECK(c);
  return SLA_OK;
}

template <int EPI>
static sla_status launch_epi(sla_ctx* c, const sla_csr* A, const double* x, double* y, const double* u0, int fin, int dst) {
  spmv_tile_kernel<SLA_SPMV_TILE, EPI><<<A->ntiles, SPMV_THREADS, 0, c->stream>>>(
      A->row_ptr, A->col, A->val, x, y, A->tile_row, u0, c->partials, c->counter, c->scal, fin, dst);
  SLA_LAUNCH_CHECK(c);
  return SLA_OK;
}

// y = A x with an optional fused epilogue.  u1 is reserved (EPI_DOT2_YY uses y itself).
sla_status sla_spmv_launch(sla_ctx* c, const sla_csr* A, const double* x, double* y, int epi,
                           const double* u0, const double* u1, int fin, int dst) {
  (void)u1;
  if (A->ntiles > SLA_MAX_PARTIALS) return sla_fail(c, SLA_ERR_INVALID, "spmv: matrix has too many tiles");
  if (A->m == 0) return SLA_OK;
  switch (epi) {
    case EPI_NONE:    return launch_epi<EPI_NONE>(c, A, x, y, u0, fin, dst);
    case EPI_DOT1:    return launch_epi<EPI_DOT1>(c, A, x, y, u0, fin, dst);
    case EPI_DOT2_YY: return launch_epi<EPI_DOT2_YY>(c, A, x, y, u0, fin, dst);
    case EPI_RESNORM: return launch_epi<EPI_RESNORM>(c, A, x, y, u0, fin, dst);
  }
  return sla_fail(c, SLA_ERR_INVALID, "spmv: unknown epilogue");
}
