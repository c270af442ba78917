// spmv.cu — fp64 CSR sparse matrix-vector product, the (#>) of the reference
// (matVecSD / dotu, src/Data/Sparse/Common.hs:242-260), as a tile-streamed sm_100a kernel.
//
// Decomposition.  The nnz stream (col_idx | val) is cut into fixed tiles of SLA_SPMV_TILE entries; tile t
// OWNS the rows whose first entry lies inside it (tile_row[t] .. tile_row[t+1], found once per matrix by
// sla_csr_build_plan).  A CTA
//   1. streams its tile with fully coalesced lane-contiguous loads (each warp instruction reads 128 B of
//      col_idx / 256 B of val; L1 no-allocate, L2 evict-first so the matrix stream does not push x out of
//      L2), gathers x[col] through the read-only path (L2 evict-last) and writes the products a_ij * x_j
//      (__dmul_rn, no FMA) into shared memory with conflict-free 8-byte stores,
//   2. sums each owned row from shared memory IN ASCENDING COLUMN ORDER with __dadd_rn from a 0.0 seed —
//      the reference's strict left fold — so rows of up to SLA_LONG_ROW entries are bit-identical to
//      the Haskell result; longer rows are summed by one warp (lane-strided partials + shuffle tree).
//      A row that runs past the tile end reads its tail straight from global memory.
//   3. optionally folds the row results into up to two dot products / a residual norm (the Krylov
//      epilogues), reduced over the grid deterministically by the last CTA to finish.
// The shared-memory product buffer is skewed, loc(k) = k + (k >> a), with a chosen per matrix from the
// typical row length L (2^a = largest power of two dividing L; no skew for odd L) so that the per-row
// sequential reads of equal-length rows hit distinct banks: thread t reads (L + L/2^a) t + i, an odd stride.
// (ncu, profiles/r01_*: with 16-byte stores and a fixed skew 45-60 % of the shared-memory wavefronts were
// bank conflicts and the L1TEX data pipe, not HBM, bounded the kernel.)
//
// Column panels.  When x is larger than the L2 can keep resident (measured on B200: the gather rate
// collapses once 8 n > ~48 MB, profiles/r01_l2_sweep.md) and the matrix has no column locality, the plan
// stores a second copy of the matrix split into column panels of <= SLA_PANEL_BYTES of x each (each panel is
// itself a CSR matrix over all rows).  (#>) then runs the same kernel once per panel in ascending column
// order, the later passes continuing the row sums from y (ACC = true) — still the same left fold, so the
// result stays bit-identical — while every pass gathers from an L2-resident slice of x.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <new>
#include <vector>
#include <stdlib.h>

#define SLA_LONG_ROW 256
#define SLA_LONG_CAP 16
#ifndef SPMV_THREADS
#define SPMV_THREADS 128
#endif
#define SLA_PANEL_BYTES (40u << 20)     // x bytes per column panel
#define SLA_PANEL_MIN_X (56u << 20)     // panelise only when 8 n exceeds this ...
#define SLA_PANEL_MIN_SPAN (24u << 20)  // ... and an average tile touches a wider stretch of x than this
#define SLA_GATHER_NA_SPAN (512u << 10) // x gathers bypass L1 allocation when a tile's column span exceeds this

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
__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ int ld_stream_int(const int* p, uint64_t pol) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ double ld_stream_double(const double* p, uint64_t pol) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ double ld_keep_double(const double* p, uint64_t pol) {
  double r;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
  return r;
}

__device__ __forceinline__ double ld_keep_double_na(const double* p, uint64_t pol) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
  return r;
}

__device__ __forceinline__ int skew(int k, int a) { return k + (k >> a); }   // a = 31: no skew

// A row block of a multi-GPU matrix carries GLOBAL column indices: columns inside [col0, col0 + ncl) are read from the local
// slice x, every other column from the gathered-x buffer xr (indexed by global column), which the exchange that precedes the
// kernel has filled — an all-gather, a peer-memory push, copy-engine blocks consumed in arrival order, or the LL halo
// exchange (p2p.cu modes 1-4, dist.cu for NCCL).
// x_j for the out-of-tile paths (row tails, long rows)
template <int DIST>
__device__ __forceinline__ double fetch_x(const double* __restrict__ x, int c, const SpmvDist& dx) {
  if (DIST == 0) return x[c];
  if ((unsigned)(c - dx.col0) < (unsigned)dx.ncl) return x[c - dx.col0];
  return dx.xr[c];
}

template <int EPI>
__device__ __forceinline__ void row_epilogue(int r, double acc, double* y, const double* __restrict__ u0,
                                             double& e0, double& e1) {
  if (EPI == EPI_RESNORM) {
    double d = __dsub_rn(acc, u0[r]);          // (aa #> x) ^-^ b, then (**2)   Sparse.hs:1041
    e0 += d * d;
    return;
  }
  y[r] = acc;
  if (EPI == EPI_DOT1 || EPI == EPI_DOT2_YY) e0 += acc * u0[r];
  if (EPI == EPI_DOT2_YY) e1 += acc * acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// the same wait with a watchdog: a protocol error traps (the launch fails) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// ACC = false: row sums start from 0.0.  ACC = true: they continue from yin[r] (a later column panel).
// DIST > 0: row block of a multi-GPU matrix, see SpmvDist.
// STAGE = 0: every thread loads its 8 (col, val) pairs straight into registers (lane-contiguous LDG).
// STAGE = 1: the tile's 12 KB arrive through TWO BULK COPIES (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier) issued by
//            thread 0 at CTA start; the threads then read their pairs from shared memory.  The point is the L1TEX request port:
//            on random columns the kernel is bound by ~0.95 requests per SM-cycle (profiles/r02_gather_paths.jsonl), 9 % of which
//            were the stream's own line requests; the bulk-copy engine does not use that port.  The product buffer overlays the
//            staging area (every thread has its pairs in registers before the first product is written), so shared memory per
//            CTA stays at 12 KB and occupancy at 16 CTAs per SM — unlike the persistent 3-stage ring of spmv_tma_kernel below.
template <int TILE, int EPI, bool ACC, int DIST, int STAGE>
__global__ void __launch_bounds__(SPMV_THREADS)
spmv_tile_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, const double* __restrict__ val,
                 const double* __restrict__ x, const double* yin, double* y, const int* __restrict__ tile_row,
                 const double* __restrict__ u0, double* partials, int hints, const SpmvDist dx, int tile0) {
  constexpr int PER = TILE / SPMV_THREADS;                // entries per thread, lane-contiguous
  // keep CTA smem small: the L1 left over holds the in-flight gathers.  STAGE 1: [val 8 KB | col 4 KB], products overlay it.
  __shared__ __align__(128) double prod_store[STAGE ? (TILE * 12) / 8 : TILE + TILE / 8];
  __shared__ uint64_t stage_bar;
  double* prod = prod_store;
  __shared__ double red[2 * 32];
  __shared__ int long_rows[SLA_LONG_CAP];
  __shared__ int n_long;

  const int tid = threadIdx.x;
  const int tile = blockIdx.x + tile0;                    // tile0 > 0: a row chunk of the host-pipelined (#>)
  const int base = tile * TILE;
  const int row_lo = tile_row[tile], row_hi = tile_row[tile + 1];
  const int nrows = row_hi - row_lo;
  const int sk = hints >> 8;                              // skew shift a
  if (tid == 0) n_long = 0;
  if (STAGE && nrows > 0) {
    if (tid == 0) {
      const uint64_t pol = (hints & 1) ? policy_evict_first() : policy_evict_normal();
      mbar_init(&stage_bar, 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the init must be visible to the async proxy
      mbar_expect_tx(&stage_bar, TILE * 12);
      bulk_g2s(prod_store, val + base, TILE * 8, &stage_bar, pol);
      bulk_g2s(reinterpret_cast<unsigned char*>(prod_store) + TILE * 8, col + base, TILE * 4, &stage_bar, pol);
    }
    __syncthreads();                                      // nobody waits on the barrier before it is initialised
  }

  // ---- phase 1: stream the tile, gather x, write products --------------------------------------
  if (nrows > 0) {
    const uint64_t pol_stream = (hints & 1) ? policy_evict_first() : policy_evict_normal();
    const uint64_t pol_keep = (hints & 2) ? policy_evict_last() : policy_evict_normal();
    int c[PER];
    double v[PER];
    if (STAGE) {
      mbar_wait(&stage_bar, 0);
      const int* sc = reinterpret_cast<const int*>(reinterpret_cast<const unsigned char*>(prod_store) + TILE * 8);
#pragma unroll
      for (int it = 0; it < PER; ++it) {
        const int k = it * SPMV_THREADS + tid;
        c[it] = sc[k];
        v[it] = prod_store[k];
      }
    } else {
#pragma unroll
      for (int it = 0; it < PER; ++it) {
        const int k = it * SPMV_THREADS + tid;
        c[it] = ld_stream_int(col + base + k, pol_stream);
        v[it] = ld_stream_double(val + base + k, pol_stream);
      }
    }
    double xv[PER];
#pragma unroll
    for (int it = 0; it < PER; ++it) {
      const double* src = x + c[it];
      if (DIST == 1) src = (unsigned)(c[it] - dx.col0) < (unsigned)dx.ncl ? x + (c[it] - dx.col0) : dx.xr + c[it];
      xv[it] = (hints & 4) ? ld_keep_double_na(src, pol_keep) : ld_keep_double(src, pol_keep);
    }
    if (STAGE) __syncthreads();                             // every thread holds its pairs: the staging area becomes the product buffer
#pragma unroll
    for (int it = 0; it < PER; ++it) {
      const int k = it * SPMV_THREADS + tid;
      prod[skew(k, sk)] = __dmul_rn(v[it], xv[it]);         // dotu: a_ij * x_j, matrix entry on the left
    }
  }
  __syncthreads();

  // ---- phase 2: one thread per owned row, sequential ascending sum --------------------------------
  double e0 = 0.0, e1 = 0.0;
  for (int j = tid; j < nrows; j += SPMV_THREADS) {
    const int r = row_lo + j;
    const int s = row_ptr[r], e = row_ptr[r + 1];   // (requesting these before phase 1 was measured slower: 1.76 vs 1.49 ms on cfg 2)
    if (e - s > SLA_LONG_ROW) {
      const int slot = atomicAdd(&n_long, 1);
      long_rows[slot] = r;
      continue;
    }
    const int ks = s - base, ke = e - base;
    const int kin = ke < TILE ? ke : TILE;
    double acc = ACC ? yin[r] : 0.0;                         // sum = strict left fold from 0
    for (int k = ks; k < kin; ++k) acc = __dadd_rn(acc, prod[skew(k, sk)]);
    for (int k = (ks > TILE ? ks : TILE); k < ke; ++k) {     // tail beyond the tile (last owned row only)
      const int g = base + k;
      const int cg = col[g];
      const double xg = fetch_x<DIST>(x, cg, dx);
      acc = __dadd_rn(acc, __dmul_rn(val[g], xg));
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
        if (k < TILE) t = prod[skew(k, sk)];
        else {
          const int g = base + k;
          const int cg = col[g];
          const double xg = fetch_x<DIST>(x, cg, dx);
          t = __dmul_rn(val[g], xg);
        }
        acc += t;
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        if (ACC) acc = __dadd_rn(yin[r], acc);
        row_epilogue<EPI>(r, acc, y, u0, e0, e1);
      }
    }
  }

  if (EPI != EPI_NONE) {
    double sums[2] = {e0, e1};
    block_sum<2>(sums, red);
    // per-CTA partials only: a one-CTA kernel sums them afterwards (partials_reduce_kernel).  The ticket +
    // __threadfence "last CTA" scheme used by the vector kernels cost ~45 % here (ncu: 381 vs 252 us on the
    // Laplacian): 41 k CTAs each ended with a fence, a same-address atomic and a barrier while holding their slot.
    if (tid == 0) {
      partials[blockIdx.x] = sums[0];
      partials[(size_t)gridDim.x + blockIdx.x] = sums[1];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// TMA variant: persistent CTAs, the (col, val) tile stream staged into shared memory by the bulk-copy engine
// (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP) through a ring of SPMV_STAGES buffers.  One elected
// thread issues the copies for tile i + SPMV_STAGES as soon as tile i's stage has been consumed, so the
// stream of the next tiles is in flight while the CTA gathers x, writes the products and sums its rows.
// Arithmetic, ownership rule, epilogues and bit-exactness are those of spmv_tile_kernel.
#define SPMV_STAGES 3

template <int TILE, int EPI, bool ACC, int DIST>
__global__ void __launch_bounds__(SPMV_THREADS)
spmv_tma_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, const double* __restrict__ val,
                const double* __restrict__ x, const double* yin, double* y, const int* __restrict__ tile_row,
                const double* __restrict__ u0, double* partials, int ntiles, int hints, const SpmvDist dx) {
  constexpr int PER = TILE / SPMV_THREADS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sval = reinterpret_cast<double*>(smem_raw);                                   // STAGES x TILE doubles
  int* scol = reinterpret_cast<int*>(smem_raw + (size_t)SPMV_STAGES * TILE * 8);        // STAGES x TILE ints
  double* prod = reinterpret_cast<double*>(smem_raw + (size_t)SPMV_STAGES * TILE * 12); // TILE + TILE/8 doubles
  __shared__ uint64_t full_bar[SPMV_STAGES];
  __shared__ double red[2 * 32];
  __shared__ int long_rows[SLA_LONG_CAP];
  __shared__ int n_long;

  const int tid = threadIdx.x;
  const int sk = hints >> 8;
  const uint64_t pol_stream = (hints & 1) ? policy_evict_first() : policy_evict_normal();
  const uint64_t pol_keep = (hints & 2) ? policy_evict_last() : policy_evict_normal();
  if (tid == 0) {
    for (int s = 0; s < SPMV_STAGES; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // make the inits visible to the async proxy
  }
  __syncthreads();

  // prologue: fill the ring
  if (tid == 0) {
    for (int s = 0; s < SPMV_STAGES; ++s) {
      const int t = blockIdx.x + s * gridDim.x;
      if (t < ntiles) {
        mbar_expect_tx(&full_bar[s], TILE * 12);
        bulk_g2s(scol + s * TILE, col + (size_t)t * TILE, TILE * 4, &full_bar[s], pol_stream);
        bulk_g2s(sval + s * TILE, val + (size_t)t * TILE, TILE * 8, &full_bar[s], pol_stream);
      }
    }
  }

  double e0 = 0.0, e1 = 0.0;
  int it = 0;
  int nxt_lo = 0, nxt_hi = 0;          // row range of the next tile, requested one iteration ahead
  if ((int)blockIdx.x < ntiles) { nxt_lo = tile_row[blockIdx.x]; nxt_hi = tile_row[blockIdx.x + 1]; }
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int stage = it % SPMV_STAGES;
    const uint32_t parity = (uint32_t)((it / SPMV_STAGES) & 1);
    const int base = tile * TILE;
    const int row_lo = nxt_lo, row_hi = nxt_hi;
    const int nrows = row_hi - row_lo;
    if (tid == 0) n_long = 0;
    if (tile + (int)gridDim.x < ntiles) { nxt_lo = tile_row[tile + gridDim.x]; nxt_hi = tile_row[tile + gridDim.x + 1]; }

    // ---- phase 1: products from the staged tile ---------------------------------------------------------
    mbar_wait(&full_bar[stage], parity);
    if (nrows > 0) {
      const int* sc = scol + stage * TILE;
      const double* sv = sval + stage * TILE;
      int c[PER];
#pragma unroll
      for (int q = 0; q < PER; ++q) c[q] = sc[q * SPMV_THREADS + tid];
      double xv[PER];
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const double* src = x + c[q];
        if (DIST == 1) src = (unsigned)(c[q] - dx.col0) < (unsigned)dx.ncl ? x + (c[q] - dx.col0) : dx.xr + c[q];
        xv[q] = (hints & 4) ? ld_keep_double_na(src, pol_keep) : ld_keep_double(src, pol_keep);
      }
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int k = q * SPMV_THREADS + tid;
        prod[skew(k, sk)] = __dmul_rn(sv[k], xv[q]);
      }
    }
    __syncthreads();                 // products complete; every thread is done reading this stage

    // refill this stage with the tile SPMV_STAGES iterations ahead
    if (tid == 0) {
      const int t2 = tile + SPMV_STAGES * gridDim.x;
      if (t2 < ntiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // order the generic-proxy reads before the async writes
        mbar_expect_tx(&full_bar[stage], TILE * 12);
        bulk_g2s(scol + stage * TILE, col + (size_t)t2 * TILE, TILE * 4, &full_bar[stage], pol_stream);
        bulk_g2s(sval + stage * TILE, val + (size_t)t2 * TILE, TILE * 8, &full_bar[stage], pol_stream);
      }
    }

    // ---- phase 2: one thread per owned row, sequential ascending sum ------------------------------------
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
      double acc = ACC ? yin[r] : 0.0;
      for (int k = ks; k < kin; ++k) acc = __dadd_rn(acc, prod[skew(k, sk)]);
      for (int k = (ks > TILE ? ks : TILE); k < ke; ++k) {
        const int g = base + k;
        const int cg = col[g];
        const double xg = fetch_x<DIST>(x, cg, dx);
        acc = __dadd_rn(acc, __dmul_rn(val[g], xg));
      }
      row_epilogue<EPI>(r, acc, y, u0, e0, e1);
    }
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
          if (k < TILE) t = prod[skew(k, sk)];
          else {
            const int g = base + k;
            const int cg = col[g];
            const double xg = fetch_x<DIST>(x, cg, dx);
            t = __dmul_rn(val[g], xg);
          }
          acc += t;
        }
        acc = warp_sum(acc);
        if (lane == 0) {
          if (ACC) acc = __dadd_rn(yin[r], acc);
          row_epilogue<EPI>(r, acc, y, u0, e0, e1);
        }
      }
      __syncthreads();               // long-row readers are done with prod before the next tile overwrites it
    }
  }

  if (EPI != EPI_NONE) {
    double sums[2] = {e0, e1};
    block_sum<2>(sums, red);
    if (tid == 0) {
      partials[blockIdx.x] = sums[0];
      partials[(size_t)gridDim.x + blockIdx.x] = sums[1];
    }
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

// mean over tiles of (max col - min col): how wide a stretch of x one CTA gathers from
__global__ void tile_span_kernel(const int* __restrict__ col, int64_t nnz, int tile, unsigned long long* __restrict__ span_sum) {
  __shared__ int smin[32], smax[32];
  const int64_t base = (int64_t)blockIdx.x * tile;
  int mn = 0x7fffffff, mx = -1;
  for (int k = threadIdx.x; k < tile && base + k < nnz; k += blockDim.x) {
    const int cidx = col[base + k];
    mn = min(mn, cidx); mx = max(mx, cidx);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = min(mn, smin[w]); mx = max(mx, smax[w]); }
    if (mx >= mn) atomicAdd(span_sum, (unsigned long long)(mx - mn));
  }
}

// start[p * m + r] = absolute offset of the first entry of row r with col >= p * width   (p = 0 .. P)
// len  [p * m + r] = entries of row r that fall in panel p
__global__ void panel_split_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int m, int P, int width,
                                   int* __restrict__ start, int* __restrict__ len) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  const int s = row_ptr[r], e = row_ptr[r + 1];
  int prev = s;
  for (int p = 0; p < P; ++p) {
    int nxt = e;
    if (p + 1 < P) {
      const long long bound = (long long)(p + 1) * width;
      int lo = prev, hi = e;
      while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if ((long long)col[mid] < bound) lo = mid + 1; else hi = mid;
      }
      nxt = lo;
    }
    start[(size_t)p * m + r] = prev;
    len[(size_t)p * m + r] = nxt - prev;
    prev = nxt;
  }
}

// scatter every stored entry into its panel: dest = panel row_ptr[r] + (q - start[p][r])
__global__ void panel_fill_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, const double* __restrict__ val,
                                  int m, int64_t nnz, int width, const int* __restrict__ start, sla_panel* __restrict__ panels) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = m;                     // last row r with row_ptr[r] <= q
    while (lo < hi) {
      const int mid = lo + ((hi - lo + 1) >> 1);
      if (row_ptr[mid] <= q) lo = mid; else hi = mid - 1;
    }
    const int cidx = col[q];
    const int p = cidx / width;
    const sla_panel pn = panels[p];
    const int d = pn.row_ptr[lo] + (int)(q - start[(size_t)p * m + lo]);
    pn.col[d] = cidx;
    pn.val[d] = val[q];
  }
}

// ROTATED panels (p2p.cu mode 5): panel p = the column blocks of this rank's predecessors kb[p] .. kb[p + 1] - 1 (0 = own block).
// Inside a panel a row keeps its entries in ascending column order.
__device__ __forceinline__ int rot_panel_of(int c, const sla_rot_spec& rs) {
  long long d = rs.own_end - 1 - (long long)c;
  if (d < 0) d += rs.n;
  const int k = (int)(d / rs.m);
  int p = 0;
  while (p + 1 < rs.P && k >= rs.kb[p + 1]) ++p;
  return p;
}
__global__ void rot_count_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int m, const sla_rot_spec rs, int* __restrict__ len) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  for (int p = 0; p < rs.P; ++p) len[(size_t)p * m + r] = 0;
  for (int q = row_ptr[r]; q < row_ptr[r + 1]; ++q) len[(size_t)rot_panel_of(col[q], rs) * m + r] += 1;
}
__global__ void rot_fill_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, const double* __restrict__ val, int m,
                                const sla_rot_spec rs, const sla_panel* __restrict__ panels) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  int cur[SLA_ROT_MAX];
#pragma unroll
  for (int p = 0; p < SLA_ROT_MAX; ++p) cur[p] = p < rs.P ? panels[p].row_ptr[r] : 0;
  for (int q = row_ptr[r]; q < row_ptr[r + 1]; ++q) {
    const int cidx = col[q];
    const int p = rot_panel_of(cidx, rs);
    int at = 0;
#pragma unroll
    for (int t = 0; t < SLA_ROT_MAX; ++t) if (t == p) { at = cur[t]; cur[t] = at + 1; }
    panels[p].col[at] = cidx;
    panels[p].val[at] = val[q];
  }
}

__global__ void set_last_kernel(int32_t* row_ptr, const int32_t* len, int m) {
  if (threadIdx.x == 0 && blockIdx.x == 0) row_ptr[m] = m > 0 ? row_ptr[m - 1] + len[m - 1] : 0;
}

// skew shift a for a typical row length L: 2^a = largest power of two dividing L (odd L: no skew),
// clamped to a >= 3 so the skewed buffer fits TILE + TILE / 8 doubles (a larger buffer pushes the
// shared-memory carve-out to 228 KB, leaves no L1 for outstanding gathers and costs 3x on random columns).
static int choose_skew(int64_t nnz, int64_t m) {
  if (m <= 0 || nnz <= 0) return 31;
  int64_t L = (nnz + m / 2) / m;
  if (L < 1) L = 1;
  if (L & 1) return 31;
  int a = 0;
  while (((L >> a) & 1) == 0) ++a;
  return a < 3 ? 3 : a;
}

static sla_status build_tile_plan(sla_ctx* c, const int32_t* row_ptr, int64_t m, int ntiles, int32_t** tile_row) {
  if (*tile_row == nullptr) SLA_CUDA(c, cudaMalloc(tile_row, sizeof(int32_t) * (size_t)(ntiles + 1)));
  spmv_plan_kernel<<<(ntiles + 1 + 255) / 256, 256, 0, c->stream>>>(row_ptr, (int)m, ntiles, SLA_SPMV_TILE, *tile_row);
  SLA_LAUNCH_CHECK(c);
  return SLA_OK;
}

void sla_csr_free_panels(sla_csr* A) {
  A->chunk_ready = 0;
  if (!A->panels) return;
  for (int p = 0; p < A->npanels; ++p) {
    cudaFree(A->panels[p].row_ptr); cudaFree(A->panels[p].col); cudaFree(A->panels[p].val); cudaFree(A->panels[p].tile_row);
  }
  delete[] A->panels;
  A->panels = nullptr; A->npanels = 0; A->chunk_ready = 0;
}

// rs != nullptr: rotated panels (blocks of predecessors) instead of P equal column ranges
static sla_status build_panels(sla_ctx* c, sla_csr* A, int P, const sla_rot_spec* rs = nullptr) {
  const int m = (int)A->m;
  const bool rot = rs != nullptr;
  int width = (int)((A->n + P - 1) / P);
  width = (width + 15) & ~15;
  P = rot ? rs->P : (int)((A->n + width - 1) / width);
  if (rot) width = -1;                 // no uniform width: sla_csr_force_panels never mistakes these for its own
  if (P < 2) return SLA_OK;
  int *start = nullptr, *len = nullptr;
  sla_panel* d_panels = nullptr;
  void* tmp = nullptr;
  A->panels = new sla_panel[P]();
  A->npanels = P;
  A->panel_width = width;
  if (m == 0) return SLA_OK;          // a rank without rows still takes part in the per-panel exchange
  sla_status s = SLA_OK;
  do {
    if (cudaMalloc(&start, sizeof(int) * (size_t)P * m) != cudaSuccess || cudaMalloc(&len, sizeof(int) * (size_t)P * m) != cudaSuccess ||
        cudaMalloc(&d_panels, sizeof(sla_panel) * P) != cudaSuccess) { s = sla_fail(c, SLA_ERR_ALLOC, "spmv plan: cudaMalloc failed"); break; }
    if (rot) rot_count_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(A->row_ptr, A->col, m, *rs, len);
    else panel_split_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(A->row_ptr, A->col, m, P, width, start, len);
    c->launches++;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, len, len, m, c->stream);
    if (cudaMalloc(&tmp, tb ? tb : 1) != cudaSuccess) { s = sla_fail(c, SLA_ERR_ALLOC, "spmv plan: cudaMalloc failed"); break; }
    for (int p = 0; p < P && s == SLA_OK; ++p) {
      sla_panel& pn = A->panels[p];
      if (cudaMalloc(&pn.row_ptr, sizeof(int32_t) * (size_t)(m + 1)) != cudaSuccess) { s = sla_fail(c, SLA_ERR_ALLOC, "spmv plan: cudaMalloc failed"); break; }
      cub::DeviceScan::ExclusiveSum(tmp, tb, len + (size_t)p * m, pn.row_ptr, m, c->stream);
      set_last_kernel<<<1, 32, 0, c->stream>>>(pn.row_ptr, len + (size_t)p * m, m);
      c->launches += 3;
      int32_t nz = 0;
      cudaMemcpyAsync(&nz, pn.row_ptr + m, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
      if (cudaStreamSynchronize(c->stream) != cudaSuccess) { s = sla_fail(c, SLA_ERR_CUDA, "spmv plan: CUDA error"); break; }
      pn.nnz = nz;
      pn.skew_a = choose_skew(nz, m);
      pn.ntiles = nz / SLA_SPMV_TILE + 1;
      const size_t pn_pad = (size_t)pn.ntiles * SLA_SPMV_TILE;
      if (cudaMalloc(&pn.col, sizeof(int32_t) * pn_pad) != cudaSuccess || cudaMalloc(&pn.val, sizeof(double) * pn_pad) != cudaSuccess) {
        s = sla_fail(c, SLA_ERR_ALLOC, "spmv plan: cudaMalloc failed for a column panel"); break;
      }
      cudaMemsetAsync(pn.col + nz, 0, sizeof(int32_t) * (pn_pad - nz), c->stream);
      cudaMemsetAsync(pn.val + nz, 0, sizeof(double) * (pn_pad - nz), c->stream);
      s = build_tile_plan(c, pn.row_ptr, m, pn.ntiles, &pn.tile_row);
    }
    if (s != SLA_OK) break;
    cudaMemcpyAsync(d_panels, A->panels, sizeof(sla_panel) * P, cudaMemcpyHostToDevice, c->stream);
    if (A->nnz > 0 && rot) {
      rot_fill_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(A->row_ptr, A->col, A->val, m, *rs, d_panels);
      c->launches++;
    } else if (A->nnz > 0) {
      int64_t blocks = (A->nnz + 255) / 256;
      if (blocks > SLA_NUM_SMS * 32) blocks = SLA_NUM_SMS * 32;
      panel_fill_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(A->row_ptr, A->col, A->val, m, A->nnz, width, start, d_panels);
      c->launches++;
    }
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) s = sla_fail(c, SLA_ERR_CUDA, "spmv plan: CUDA error while building column panels");
  } while (0);
  cudaFree(start); cudaFree(len); cudaFree(d_panels); cudaFree(tmp);
  if (s != SLA_OK) { cudaGetLastError(); sla_csr_free_panels(A); }
  return s;
}

static sla_status build_band_plan(sla_ctx* c, sla_csr* A);      // spmv_band.cuh: x staged through shared memory for matrices with column locality

// Builds the tile plan, picks the gather cache policy and, when x cannot stay L2-resident and the matrix has
// no column locality, the column-panel copy.
//   SLA_SPMV_PANELS=1 disables panels, =P (>1) forces P panels, unset/0 = automatic.
//   SLA_SPMV_HINTS overrides the cache-hint bits (1: stream evict_first, 2: x evict_last, 4: x L1 no_allocate).
sla_status sla_csr_build_plan(sla_ctx* c, sla_csr* A) {
  SLA_TRY(build_tile_plan(c, A->row_ptr, A->m, A->ntiles, &A->tile_row));
  A->skew_a = choose_skew(A->nnz, A->m);
  if (const char* e = getenv("SLA_SPMV_SKEW")) { A->skew_a = atoi(e); if (A->skew_a < 3) A->skew_a = 3; }
  sla_csr_free_panels(A);
  A->hints = c->spmv_hints & 3;
  if (A->nnz == 0 || A->m == 0) return SLA_OK;
  SLA_TRY(build_band_plan(c, A));
  // how wide a stretch of x does one CTA gather from, on average?
  double mean_span_bytes = 0.0;
  {
    unsigned long long* d_span = nullptr;
    unsigned long long h_span = 0;
    SLA_CUDA(c, cudaMalloc(&d_span, sizeof(unsigned long long)));
    cudaMemsetAsync(d_span, 0, sizeof(unsigned long long), c->stream);
    const int nt = (int)((A->nnz + SLA_SPMV_TILE - 1) / SLA_SPMV_TILE);
    tile_span_kernel<<<nt, 256, 0, c->stream>>>(A->col, A->nnz, SLA_SPMV_TILE, d_span);
    c->launches++;
    cudaMemcpyAsync(&h_span, d_span, sizeof(h_span), cudaMemcpyDeviceToHost, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(d_span);
    SLA_CUDA(c, e);
    mean_span_bytes = 8.0 * (double)h_span / (double)nt;
  }
  // Gathers with no reuse inside an SM should not allocate in L1: the L1 lines they would occupy are what
  // tracks outstanding misses (measured: +18 % on random columns, -3 % on the 5-point stencil).
  if (mean_span_bytes > (double)SLA_GATHER_NA_SPAN) A->hints |= 4;
  if (getenv("SLA_SPMV_HINTS")) A->hints = c->spmv_hints & 7;
  int want = 0;
  if (const char* e = getenv("SLA_SPMV_PANELS")) want = atoi(e);
  if (want == 1) return SLA_OK;
  if (A->band && want == 0) return SLA_OK;          // the band plan keeps x in shared memory: no column panels needed
  if (want == 0) {
    if ((uint64_t)A->n * 8u <= SLA_PANEL_MIN_X || mean_span_bytes <= (double)SLA_PANEL_MIN_SPAN) return SLA_OK;
    want = (int)(((uint64_t)A->n * 8u + SLA_PANEL_BYTES - 1) / SLA_PANEL_BYTES);
  }
  if (want > SLA_MAX_PANELS) want = SLA_MAX_PANELS;
  return build_panels(c, A, want);
}

// Multi-GPU: the panel count of a densely-coupled distributed matrix is a collective choice (every rank clips
// the same exchange segments with the same panel boundaries), so dist.cu imposes it here.
sla_status sla_csr_force_panels(sla_ctx* c, sla_csr* A, int P) {
  if (P > SLA_MAX_PANELS) P = SLA_MAX_PANELS;
  int width = (int)((A->n + P - 1) / P);
  width = (width + 15) & ~15;
  if (A->npanels >= 2 && A->panel_width == width) return SLA_OK;
  sla_csr_free_panels(A);
  if (P < 2) return SLA_OK;
  return build_panels(c, A, P);
}

// p2p.cu mode 5: rotated panels of a row-partitioned matrix (collective: every rank calls it with its own rotation)
sla_status sla_csr_force_rot_panels(sla_ctx* c, sla_csr* A, const sla_rot_spec* spec) {
  sla_csr_free_panels(A);
  if (!spec || spec->P < 2 || spec->P > SLA_ROT_MAX || spec->m <= 0) return SLA_OK;
  return build_panels(c, A, spec->P, spec);
}

// Sums the two per-CTA partial arrays of an SpMV epilogue in a fixed order and post-processes the Krylov scalars
// (or stores the raw sums for the multi-GPU all-reduce).  Two levels: up to 128 CTAs each sum a contiguous slice
// (thread-strided, then the block tree), the last of them to finish sums the slice results in order — a single
// CTA needed ~46 us (ncu) for the 2 x 82 k partials of the 4096^2 Laplacian.
#define PRED_THREADS 256
__global__ void __launch_bounds__(PRED_THREADS)
partials_reduce_kernel(const double* __restrict__ partials, int nblk, double* partials2, unsigned int* counter,
                       double* scal, int fin, int dst, sla_p2p_args pa) {
  __shared__ double red[2 * 32];
  const int per = (nblk + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per, hi = min(nblk, lo + per);
  double acc[2] = {0.0, 0.0};
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    acc[0] += partials[i];
    acc[1] += partials[(size_t)nblk + i];
  }
  block_sum<2>(acc, red);
  __syncthreads();
  grid_reduce_finish<2>(acc, partials2, counter, scal, fin, dst, red, pa);
}

#include "spmv_band.cuh"

// everything one launch needs besides the epilogue selection
struct SpmvArgs {
  const int32_t *row_ptr, *col, *tile_row; const double* val;
  int ntiles, skew_a, hints;
  const double *x, *yin; double* y; const double* u0;
  int fin, dst;
  SpmvDist dx;                         // DIST > 0 only
  int tile0;                           // first tile of this launch (row-chunked launches of sla_spmv_host)
};

template <int EPI, bool ACC, int DIST>
static sla_status launch_one(sla_ctx* c, const SpmvArgs& a) {
  static char carve_set[64] = {0};          // function attributes are per device (one process may drive several GPUs)
  if (!carve_set[c->device & 63]) {
    carve_set[c->device & 63] = 1;
    if (const char* e = getenv("SLA_SPMV_CARVEOUT"))
      cudaFuncSetAttribute(spmv_tile_kernel<SLA_SPMV_TILE, EPI, ACC, DIST, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
  }
  int nblk = a.ntiles;
  if (c->spmv_tma) {
    // TMA-staged persistent variant: SPMV_CTAS_PER_SM CTAs per SM, each looping over tiles
    constexpr size_t smem = (size_t)SPMV_STAGES * SLA_SPMV_TILE * 12 + (size_t)(SLA_SPMV_TILE + SLA_SPMV_TILE / 8) * 8;
    static char attr_set[64] = {0};
    if (!attr_set[c->device & 63]) {
      attr_set[c->device & 63] = 1;
      SLA_CUDA(c, cudaFuncSetAttribute(spmv_tma_kernel<SLA_SPMV_TILE, EPI, ACC, DIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    nblk = a.ntiles < SLA_NUM_SMS * c->spmv_tma ? a.ntiles : SLA_NUM_SMS * c->spmv_tma;
    if (a.tile0 != 0) return sla_fail(c, SLA_ERR_INVALID, "spmv: the TMA-staged variant does not take row-chunked launches");
    spmv_tma_kernel<SLA_SPMV_TILE, EPI, ACC, DIST><<<nblk, SPMV_THREADS, smem, c->stream>>>(
        a.row_ptr, a.col, a.val, a.x, a.yin, a.y, a.tile_row, a.u0, c->partials, a.ntiles,
        (a.hints & 0xff) | (a.skew_a << 8), a.dx);
  } else {
    if (c->spmv_bulk && DIST == 0)
      spmv_tile_kernel<SLA_SPMV_TILE, EPI, ACC, 0, 1><<<a.ntiles, SPMV_THREADS, 0, c->stream>>>(
          a.row_ptr, a.col, a.val, a.x, a.yin, a.y, a.tile_row, a.u0, c->partials, (a.hints & 0xff) | (a.skew_a << 8), a.dx, a.tile0);
    else
      spmv_tile_kernel<SLA_SPMV_TILE, EPI, ACC, DIST, 0><<<a.ntiles, SPMV_THREADS, 0, c->stream>>>(
          a.row_ptr, a.col, a.val, a.x, a.yin, a.y, a.tile_row, a.u0, c->partials, (a.hints & 0xff) | (a.skew_a << 8), a.dx, a.tile0);
  }
  SLA_LAUNCH_CHECK(c);
  if (EPI != EPI_NONE) {
    int g = nblk / 2048;
    g = g < 1 ? 1 : (g > 128 ? 128 : g);
    const sla_red_plan rp = sla_red_begin(c, a.fin, 2);
    partials_reduce_kernel<<<g, PRED_THREADS, 0, c->stream>>>(c->partials, nblk, c->partials + 2 * (size_t)SLA_MAX_PARTIALS, c->counter,
                                                             c->scal, rp.fin, a.dst, rp.pa);
    SLA_LAUNCH_CHECK(c);
    SLA_TRY(sla_red_end(c, rp, 2, a.fin, a.dst));
  }
  return SLA_OK;
}

template <bool ACC, int DIST>
static sla_status launch_epi(sla_ctx* c, int epi, const SpmvArgs& a) {
  switch (epi) {
    case EPI_NONE:    return launch_one<EPI_NONE, ACC, DIST>(c, a);
    case EPI_DOT1:    return launch_one<EPI_DOT1, ACC, DIST>(c, a);
    case EPI_DOT2_YY: return launch_one<EPI_DOT2_YY, ACC, DIST>(c, a);
    case EPI_RESNORM: return launch_one<EPI_RESNORM, ACC, DIST>(c, a);
  }
  return sla_fail(c, SLA_ERR_INVALID, "spmv: unknown epilogue");
}

static sla_status launch_any(sla_ctx* c, int epi, bool acc, int dist, const SpmvArgs& a) {
  if (dist == 1) return acc ? launch_epi<true, 1>(c, epi, a) : launch_epi<false, 1>(c, epi, a);
  return acc ? launch_epi<true, 0>(c, epi, a) : launch_epi<false, 0>(c, epi, a);
}

static void dist_args(const sla_csr* A, SpmvDist* dx) {
  memset(dx, 0, sizeof(*dx));
  if (!A->dist) return;
  dx->xr = A->dist->xfull; dx->col0 = (int)A->dist->row0; dx->ncl = (int)A->m;
}

// Row-partitioned (#>) in ARRIVAL order (p2p.cu mode 2, SLA_P2P_X=2; not yet run on hardware): one column panel per
// source rank; the own block first, then the blocks in the order the staggered copy-engine all-gather delivers them,
// each panel kernel preceded by a one-warp wait for that source's flag.  The fold over a row's entries is in rotated
// column order: a valid summation of the same products (SURVEY.md §8(d) bound), not the bit-exact ascending fold.
static sla_status spmv_launch_arrival(sla_ctx* c, const sla_csr* A, const double* x, double* y, int epi, const double* u0, int fin, int dst) {
  const int W = c->world;
  const bool skip = c->skip_exchange != 0;                // diagnostic: kernels only
  if (!skip) SLA_TRY(sla_p2p_arrival_begin(c, A, x));
  SpmvArgs a;
  a.hints = A->hints; a.x = x; a.u0 = u0; a.fin = fin; a.dst = dst;
  dist_args(A, &a.dx); a.tile0 = 0;
  double* ybuf = y;
  if (epi == EPI_RESNORM) {
    if (!c->scratch_r || c->scratch_r->n != A->m) {
      sla_vec_free(c->scratch_r); c->scratch_r = nullptr;
      SLA_TRY(sla_vec_alloc(c, A->m, &c->scratch_r));
    }
    ybuf = c->scratch_r->d;
  }
  for (int k = 0; k < W; ++k) {
    const int q = (c->rank - k + W) % W;                 // rank r-1 sends to me first, r-2 second, ...
    if (k > 0 && !skip) SLA_TRY(sla_p2p_arrival_wait(c, A, q));
    const sla_panel& pn = A->panels[q];
    a.row_ptr = pn.row_ptr; a.col = pn.col; a.val = pn.val; a.tile_row = pn.tile_row;
    a.ntiles = pn.ntiles; a.skew_a = pn.skew_a;
    a.yin = k == 0 ? nullptr : ybuf; a.y = ybuf;
    SLA_TRY(launch_any(c, k + 1 == W ? epi : EPI_NONE, k > 0, 1, a));
  }
  return skip ? SLA_OK : sla_p2p_arrival_end(c);
}

// Row-partitioned (#>) with the PHASED push (p2p.cu mode 5): panel 0 (own block) runs at once, panel p after phase p has arrived —
// which travelled under the kernels of the panels before it.
static sla_status spmv_launch_twophase(sla_ctx* c, const sla_csr* A, const double* x, double* y, int epi, const double* u0, int fin, int dst) {
  const bool skip = c->skip_exchange != 0;                // diagnostic: kernels only
  if (!skip) SLA_TRY(sla_p2p_twophase_begin(c, A, x));
  SpmvArgs a;
  a.hints = A->hints; a.x = x; a.u0 = u0; a.fin = fin; a.dst = dst;
  dist_args(A, &a.dx); a.tile0 = 0;
  double* ybuf = y;
  if (epi == EPI_RESNORM) {
    if (!c->scratch_r || c->scratch_r->n != A->m) {
      sla_vec_free(c->scratch_r); c->scratch_r = nullptr;
      SLA_TRY(sla_vec_alloc(c, A->m, &c->scratch_r));
    }
    ybuf = c->scratch_r->d;
  }
  for (int ph = 0; ph < A->npanels; ++ph) {
    if (!skip) SLA_TRY(sla_p2p_twophase_wait(c, A, ph));
    const sla_panel& pn = A->panels[ph];
    a.row_ptr = pn.row_ptr; a.col = pn.col; a.val = pn.val; a.tile_row = pn.tile_row;
    a.ntiles = pn.ntiles; a.skew_a = pn.skew_a;
    a.yin = ph == 0 ? nullptr : ybuf; a.y = ybuf;
    SLA_TRY(launch_any(c, ph == A->npanels - 1 ? epi : EPI_NONE, ph > 0, 1, a));
  }
  return skip ? SLA_OK : sla_p2p_arrival_end(c);
}

// y = A x with an optional fused epilogue.  u1 is reserved (EPI_DOT2_YY uses y itself).
// For a row block of a distributed matrix, x is the LOCAL slice; the remote entries are exchanged first.
sla_status sla_spmv_launch(sla_ctx* c, const sla_csr* A, const double* x, double* y, int epi,
                           const double* u0, const double* u1, int fin, int dst) {
  (void)u1;
  if (A->ntiles > SLA_MAX_PARTIALS) return sla_fail(c, SLA_ERR_INVALID, "spmv: matrix has too many tiles");
  const bool dist = A->dist != nullptr;
  if (A->band && !dist) return band_launch(c, A, x, y, epi, u0, fin, dst);
  if (dist && c->world > 1 && sla_xwin_mode(A) == 2 && A->npanels == c->world && A->m > 0)
    return spmv_launch_arrival(c, A, x, y, epi, u0, fin, dst);
  if (dist && c->world > 1 && sla_xwin_mode(A) == 5 && A->npanels >= 2 && A->m > 0)
    return spmv_launch_twophase(c, A, x, y, epi, u0, fin, dst);
  // mode 4: the same copy-engine all-gather, waited for as a whole, then the matrix's own plan (no per-source panels)
  const bool ce_gather = dist && c->world > 1 && sla_xwin_mode(A) == 4 && !c->skip_exchange;
  if (ce_gather) {
    SLA_TRY(sla_p2p_arrival_begin(c, A, x));
    SLA_TRY(sla_p2p_arrival_wait(c, A, -1));
  }
  // Dense multi-GPU plans are pipelined: the exchange of panel p+1 (comm stream) overlaps the kernel of panel p.
  const bool pipelined = dist && A->dist->pipelined && A->npanels >= 2 && c->world > 1 && !c->skip_exchange;
  if (pipelined) {
    SLA_CUDA(c, cudaEventRecord(c->ev_x0, c->stream));                  // x is final, earlier readers of xfull are done
    SLA_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_x0, 0));
    for (int p = 0; p < A->npanels; ++p) {
      SLA_TRY(sla_dist_exchange_panel(c, A, x, p));
      SLA_CUDA(c, cudaEventRecord(c->ev_panel[p], c->comm_stream));
    }
  } else if (dist && !c->skip_exchange && !ce_gather) {
    SLA_TRY(sla_dist_exchange_x(c, A, x));                              // every rank takes part, even with no local rows
  }
  SpmvArgs a;
  a.row_ptr = A->row_ptr; a.col = A->col; a.val = A->val; a.tile_row = A->tile_row;
  a.ntiles = A->ntiles; a.skew_a = A->skew_a; a.hints = A->hints;
  a.x = x; a.yin = nullptr; a.y = y; a.u0 = u0; a.fin = fin; a.dst = dst;
  dist_args(A, &a.dx);                                                  // after the exchange: it may flip the double buffer
  a.tile0 = 0;
  const int dmode = dist ? 1 : 0;
  if (A->npanels < 2) {
    SLA_TRY(launch_any(c, epi, false, dmode, a));
    return ce_gather ? sla_p2p_arrival_end(c) : SLA_OK;
  }
  // column panels in ascending order; the epilogue rides on the last pass
  double* ybuf = y;
  if (epi == EPI_RESNORM) {          // y is not an output of this mode: keep the partial sums in a scratch vector
    if (!c->scratch_r || c->scratch_r->n != A->m) {
      sla_vec_free(c->scratch_r); c->scratch_r = nullptr;
      SLA_TRY(sla_vec_alloc(c, A->m, &c->scratch_r));
    }
    ybuf = c->scratch_r->d;
  }
  for (int p = 0; p < A->npanels; ++p) {
    const sla_panel& pn = A->panels[p];
    const bool last = p + 1 == A->npanels;
    a.row_ptr = pn.row_ptr; a.col = pn.col; a.val = pn.val; a.tile_row = pn.tile_row;
    a.ntiles = pn.ntiles; a.skew_a = pn.skew_a;
    a.yin = p == 0 ? nullptr : ybuf; a.y = ybuf;
    if (pipelined) SLA_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_panel[p], 0));
    if (A->m == 0) {
      // a rank without rows has no panel arrays; it still joins the epilogue's all-reduce through the (empty) base plan
      if (last) {
        a.row_ptr = A->row_ptr; a.col = A->col; a.val = A->val; a.tile_row = A->tile_row; a.ntiles = A->ntiles; a.skew_a = A->skew_a;
        a.yin = nullptr;
        SLA_TRY(launch_any(c, epi, false, dmode, a));
      }
      continue;
    }
    SLA_TRY(launch_any(c, last ? epi : EPI_NONE, p > 0, dmode, a));
  }
  return ce_gather ? sla_p2p_arrival_end(c) : SLA_OK;
}

// ---- (#>) on host buffers, pipelined ------------------------------------------------------------------------
// sla_spmv_host: x arrives over PCIe panel by panel on a copy stream while the kernels of the earlier column
// panels run; the last pass is launched in SLA_HOST_CHUNKS row chunks and every finished chunk of y starts its
// way back to the host while the next chunk computes.  Same kernels, same bits as sla_spmv.
#define SLA_HOST_CHUNKS 4

static sla_status host_pipe_setup(sla_ctx* c) {
  if (c->copy_stream) return SLA_OK;
  SLA_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < SLA_MAX_PANELS + 2 * SLA_HOST_CHUNKS; ++k) SLA_CUDA(c, cudaEventCreateWithFlags(&c->ev_copy[k], cudaEventDisableTiming));
  return SLA_OK;
}

sla_status sla_spmv_host_pipelined(sla_ctx* c, const sla_csr* A, const double* x_host, double* y_host, double* dx, double* dy) {
  SLA_TRY(host_pipe_setup(c));
  const int P = A->npanels >= 2 ? A->npanels : 1;
  const int64_t W = P > 1 ? A->panel_width : A->n;
  // which tiles / rows form the chunks of the last pass (cached on the matrix)
  const int32_t* last_tile_row = P > 1 ? A->panels[P - 1].tile_row : A->tile_row;
  const int last_ntiles = P > 1 ? A->panels[P - 1].ntiles : A->ntiles;
  sla_csr* Am = const_cast<sla_csr*>(A);
  if (!Am->chunk_ready) {
    for (int q = 0; q <= SLA_HOST_CHUNKS; ++q) {
      const int t = (int)((int64_t)last_ntiles * q / SLA_HOST_CHUNKS);
      Am->chunk_tile[q] = t;
      SLA_CUDA(c, cudaMemcpyAsync(&Am->chunk_row[q], last_tile_row + t, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    }
    SLA_CUDA(c, cudaStreamSynchronize(c->stream));
    Am->chunk_ready = 1;
  }
  // 1. uploads, one per column panel, on the copy stream (after everything previously queued on the compute stream)
  SLA_CUDA(c, cudaEventRecord(c->ev_copy[SLA_MAX_PANELS], c->stream));
  SLA_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_copy[SLA_MAX_PANELS], 0));
  for (int p = 0; p < P; ++p) {
    const int64_t lo = (int64_t)p * W, hi = lo + W < A->n ? lo + W : A->n;
    if (hi > lo) SLA_CUDA(c, cudaMemcpyAsync(dx + lo, x_host + lo, sizeof(double) * (size_t)(hi - lo), cudaMemcpyHostToDevice, c->copy_stream));
    SLA_CUDA(c, cudaEventRecord(c->ev_copy[p], c->copy_stream));
  }
  // 2. passes
  SpmvArgs a;
  a.hints = A->hints; a.x = dx; a.u0 = nullptr; a.fin = FIN_STORE; a.dst = S_TMP0; memset(&a.dx, 0, sizeof(a.dx));
  for (int p = 0; p < P; ++p) {
    const bool last = p + 1 == P;
    if (P > 1) {
      const sla_panel& pn = A->panels[p];
      a.row_ptr = pn.row_ptr; a.col = pn.col; a.val = pn.val; a.tile_row = pn.tile_row; a.ntiles = pn.ntiles; a.skew_a = pn.skew_a;
    } else {
      a.row_ptr = A->row_ptr; a.col = A->col; a.val = A->val; a.tile_row = A->tile_row; a.ntiles = A->ntiles; a.skew_a = A->skew_a;
    }
    a.yin = p == 0 ? nullptr : dy; a.y = dy; a.tile0 = 0;
    SLA_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copy[p], 0));
    if (!last) { SLA_TRY(launch_any(c, EPI_NONE, p > 0, 0, a)); continue; }
    for (int q = 0; q < SLA_HOST_CHUNKS; ++q) {
      const int t0 = A->chunk_tile[q], t1 = A->chunk_tile[q + 1];
      const int r0 = A->chunk_row[q], r1 = q + 1 == SLA_HOST_CHUNKS ? (int)A->m : A->chunk_row[q + 1];
      if (t1 > t0) { a.tile0 = t0; a.ntiles = t1 - t0; SLA_TRY(launch_any(c, EPI_NONE, p > 0, 0, a)); }
      cudaEvent_t ev = c->ev_copy[SLA_MAX_PANELS + 1 + q];
      SLA_CUDA(c, cudaEventRecord(ev, c->stream));
      SLA_CUDA(c, cudaStreamWaitEvent(c->copy_stream, ev, 0));
      if (r1 > r0) SLA_CUDA(c, cudaMemcpyAsync(y_host + r0, dy + r0, sizeof(double) * (size_t)(r1 - r0), cudaMemcpyDeviceToHost, c->copy_stream));
    }
  }
  SLA_CUDA(c, cudaStreamSynchronize(c->copy_stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  return SLA_OK;
}
