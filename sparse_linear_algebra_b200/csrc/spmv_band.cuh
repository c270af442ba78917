// spmv_band.cuh — (#>) for matrices with COLUMN LOCALITY (banded, stencil): TMA-staged tiles of the dense x vector in shared
// memory (included by spmv.cu; uses its row_epilogue / mbarrier / bulk-copy helpers).
//
// Why.  The tile kernel gathers x[col] through L1TEX, one request per 8-byte gather, and that port issues ~0.95 requests per
// SM-cycle (profiles/r02_gather_paths.jsonl): 46 % of the HBM roofline on the +-65536 banded family of cfg 2.  A gather
// from SHARED memory costs a quarter of that (4.0 per SM-cycle measured).  So when the columns a block of rows touches fit a
// modest window, x is staged through shared memory piece by piece and every gather becomes an LDS:
//
//   * rows are cut into blocks of R rows; block rb touches columns [win_lo, win_hi], cut into SUB-PANELS of W columns;
//   * the plan re-sorts the block's entries by (sub-panel, row, column) — the (rb, sub-panel) segments are contiguous, padded to
//     a multiple of 4 entries so that bulk copies stay 16-byte aligned — and stores each entry as (fp64 value, 32-bit
//     {local row : 16 | local column : 16}): the same 12 bytes per entry as CSR, no row_ptr;
//   * one persistent CTA per SM walks row blocks; per sub-panel it has the W doubles of x in shared memory (cp.async.bulk,
//     double-buffered: sub-panel s+1 arrives while s is consumed) and streams the segment through a 3-stage ring of
//     2048-entry tiles (cp.async.bulk + mbarrier, one elected thread), multiplies in place (a_ij * x_j, __dmul_rn) and adds
//     each row's products — contiguous in the tile — to that row's accumulator in shared memory, one thread per row
//     segment, strictly in ascending column order with __dadd_rn;
//   * after the last sub-panel the R accumulators are the rows' results (y, or the fused Krylov epilogues).
// A row's sum is therefore the reference's left fold over ascending columns, sub-panel after sub-panel, tile after tile:
// BIT-IDENTICAL to the tile kernel and to the Haskell result — for rows of ANY length.
//
// Cost model (what the plan checks): the stream is 12 B per entry as before; x is read R + band columns per row block instead
// of once (from L2: neighbouring row blocks share their windows), so the plan is taken only when those reads stay below half
// the stream.
#pragma once

#ifndef BAND_THREADS
#define BAND_THREADS 1024
#endif
#ifndef BAND_TILE
#define BAND_TILE 2048                      // entries per ring stage
#endif
#ifndef BAND_STAGES
#define BAND_STAGES 3
#endif
#define BAND_GROUPS (BAND_TILE / 32)        // 32-entry groups per tile, dealt round-robin to the warps
#define BAND_WARPS (BAND_THREADS / 32)
#define BAND_STAGE_BYTES (BAND_TILE * 12)
#define BAND_PAD_META 0xFFFFFFFFu           // padding entry (never a real one: local row 65535 would need R = 65536)
#define BAND_MAX_SP 1024                    // sub-panels per row block (shared-memory table)

struct sla_bsell_dev;                       // spmv_bandsell.cuh: the sliced-ELL form of the plan
struct sla_band_plan {
  sla_bsell_dev* sell;  // non-null: the plan is the sliced-ELL one and nothing below is used
  int R, W, nrb;
  int* win_lo;          // nrb      first column of the block's window (even)
  int* sp_base;         // nrb + 1  global sub-panel index of the block's first sub-panel
  int* seg_off;         // nsp + 1  entry offset of every (block, sub-panel) segment (multiples of 4)
  double* val;          // padded entries, (sub-panel, row, column) order
  unsigned* meta;       // (local row << 16) | local column ; BAND_PAD_META = padding
  int64_t nent;         // padded entry count
  int nsp;
};

#include "spmv_bandsell.cuh"

struct BandArgs {
  const int* win_lo; const int* sp_base; const int* seg_off; const double* val; const unsigned* meta;
  int R, W, nrb, m; long long n_even;
};

// One 32-entry group of a tile, one entry per lane.  The entries of a row are contiguous and ascending, so a row's sum is
// formed by the lane that holds the row's FIRST entry in this tile: it starts from the row's accumulator and adds the products
// of the following lanes one by one (warp shuffles), walks on through shared memory if the row continues past the group, and
// writes the accumulator back.  No two lanes of a tile ever own the same row, so the tile needs no barrier inside.
__device__ __forceinline__ void band_group(const double* __restrict__ pv, const unsigned* __restrict__ pm, const double* __restrict__ xs,
                                           double* __restrict__ acc, int k0, int cnt, int lane) {
  const int k = k0 + lane;
  const unsigned mt = k < cnt ? pm[k] : BAND_PAD_META;
  const unsigned row = mt >> 16;                                   // padding: 65535
  double p = 0.0;
  if (mt != BAND_PAD_META) p = __dmul_rn(pv[k], xs[mt & 0xffffu]);  // dotu: a_ij * x_j, matrix entry on the left
  unsigned prev = __shfl_up_sync(0xffffffffu, row, 1);
  if (lane == 0) prev = k0 > 0 ? pm[k0 - 1] >> 16 : 0xffffffffu;    // a row that began in the previous group belongs to that group's lane
  const bool head = mt != BAND_PAD_META && row != prev;
  double a = head ? __dadd_rn(acc[row], p) : 0.0;                   // the fold continues from the accumulator: ((acc + p0) + p1) + ...
  bool open = head;                                                 // this head's row may continue in the next lane
  for (int d = 1; d < 32; ++d) {
    const unsigned nr = __shfl_down_sync(0xffffffffu, row, d);
    const double np = __shfl_down_sync(0xffffffffu, p, d);
    open = open && lane + d < 32 && nr == row;
    if (open) a = __dadd_rn(a, np);
    if (!__any_sync(0xffffffffu, open && lane + d + 1 < 32)) break;
  }
  // tail: rows that are still open at lane 31 continue beyond the group
  const unsigned row31 = __shfl_sync(0xffffffffu, row, 31);
  if (head && row == row31) {
    for (int j = k0 + 32; j < cnt; ++j) {
      const unsigned m2 = pm[j];
      if ((m2 >> 16) != row) break;                                 // padding (65535) ends the walk too
      a = __dadd_rn(a, __dmul_rn(pv[j], xs[m2 & 0xffffu]));
    }
  }
  if (head) acc[row] = a;
}

template <int EPI>
__global__ void __launch_bounds__(BAND_THREADS, 1)
spmv_band_kernel(BandArgs P, const double* __restrict__ x, double* y, const double* __restrict__ u0, double* partials) {
  extern __shared__ __align__(128) unsigned char band_raw[];
  unsigned char* ring = band_raw;                                                   // BAND_STAGES x (val | meta)
  double* xbuf = reinterpret_cast<double*>(band_raw + BAND_STAGES * BAND_STAGE_BYTES);   // 2 x W doubles
  double* acc = xbuf + 2 * (size_t)P.W;                                             // R doubles
  __shared__ uint64_t full_bar[BAND_STAGES], x_bar[2];
  __shared__ int seg_s[BAND_MAX_SP + 1];
  __shared__ double red[2 * 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
  if (tid == 0) {
    for (int s = 0; s < BAND_STAGES; ++s) mbar_init(&full_bar[s], 1);
    mbar_init(&x_bar[0], 1); mbar_init(&x_bar[1], 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  uint32_t st_c = 0, ph_c = 0;            // consumer cursor of the tile ring (all threads, in lock step)
  uint32_t st_p = 0;                      // producer cursor (thread 0)
  uint32_t xph0 = 0, xph1 = 0;            // phases of the two x buffers
  double e0 = 0.0, e1 = 0.0;

  for (int rb = blockIdx.x; rb < P.nrb; rb += gridDim.x) {
    const int row0 = rb * P.R;
    const int nrows = min(P.R, P.m - row0);
    const int sp0 = P.sp_base[rb], nsp = P.sp_base[rb + 1] - sp0;
    const int wlo = P.win_lo[rb];
    for (int j = tid; j <= nsp; j += BAND_THREADS) seg_s[j] = P.seg_off[sp0 + j];
    for (int j = tid; j < P.R; j += BAND_THREADS) acc[j] = 0.0;             // sum = strict left fold from 0
    __syncthreads();                                                         // also: the previous block is done with every buffer
    // producer state: tile (ps, pt) = next tile to request; primed BAND_STAGES tiles ahead
    int ps = 0, pt = 0;
    auto issue_x = [&](int s) {
      const long long start = (long long)wlo + (long long)s * P.W;
      long long cnt = P.n_even - start; if (cnt > P.W) cnt = P.W;
      uint64_t* bar = &x_bar[s & 1];
      mbar_expect_tx(bar, (uint32_t)(cnt * 8));
      bulk_g2s(xbuf + (size_t)(s & 1) * P.W, x + start, (uint32_t)(cnt * 8), bar, pol_keep);
    };
    auto issue_tile = [&]() -> bool {       // requests tile (ps, pt) into stage st_p; false when the block has no tile left
      while (ps < nsp && pt >= seg_s[ps + 1] - seg_s[ps]) { ++ps; pt = 0; }
      if (ps >= nsp) return false;
      const int seg_len = seg_s[ps + 1] - seg_s[ps];
      const int cnt = min(BAND_TILE, seg_len - pt);
      const size_t off = (size_t)seg_s[ps] + pt;
      unsigned char* stg = ring + st_p * BAND_STAGE_BYTES;
      mbar_expect_tx(&full_bar[st_p], (uint32_t)cnt * 12u);
      bulk_g2s(stg, P.val + off, (uint32_t)cnt * 8u, &full_bar[st_p], pol_stream);
      bulk_g2s(stg + BAND_TILE * 8, P.meta + off, (uint32_t)cnt * 4u, &full_bar[st_p], pol_stream);
      pt += cnt;
      st_p = st_p + 1 == BAND_STAGES ? 0 : st_p + 1;
      return true;
    };
    if (tid == 0) {
      if (nsp > 0) issue_x(0);
      for (int q = 0; q < BAND_STAGES; ++q) if (!issue_tile()) break;
    }
    for (int s = 0; s < nsp; ++s) {
      if (tid == 0 && s + 1 < nsp) issue_x(s + 1);                            // buffer (s+1)&1 was last read in sub-panel s-1: finished (barrier below)
      mbar_wait_bounded(&x_bar[s & 1], (s & 1) ? xph1 : xph0);
      if (s & 1) xph1 ^= 1u; else xph0 ^= 1u;
      const double* xs = xbuf + (size_t)(s & 1) * P.W;
      const int seg_len = seg_s[s + 1] - seg_s[s];
      for (int t0 = 0; t0 < seg_len; t0 += BAND_TILE) {
        const int cnt = min(BAND_TILE, seg_len - t0);
        const double* pv = reinterpret_cast<const double*>(ring + st_c * BAND_STAGE_BYTES);
        const unsigned* pm = reinterpret_cast<const unsigned*>(ring + st_c * BAND_STAGE_BYTES + BAND_TILE * 8);
        mbar_wait_bounded(&full_bar[st_c], ph_c);
#pragma unroll
        for (int g = 0; g < BAND_GROUPS / BAND_WARPS; ++g) {
          const int k0 = (g * BAND_WARPS + warp) * 32;
          if (k0 < cnt) band_group(pv, pm, xs, acc, k0, cnt, lane);
        }
        __syncthreads();                                    // accumulators consistent for the next tile; the stage is free
        if (tid == 0) issue_tile();                         // the stage was only READ through the generic proxy: no proxy fence needed
        st_c = st_c + 1 == BAND_STAGES ? 0 : st_c + 1;
        if (st_c == 0) ph_c ^= 1u;
      }
      if (seg_len == 0) __syncthreads();                    // an empty sub-panel: keep the ranks of the x barrier's phases apart anyway
    }
    // the accumulators are the rows' results
    for (int j = tid; j < nrows; j += BAND_THREADS) row_epilogue<EPI>(row0 + j, acc[j], y, u0, e0, e1);
    __syncthreads();
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

// ---- plan -------------------------------------------------------------------------------------------------------------

// per row block: min / max stored column
__global__ void band_range_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int m, int R, int* __restrict__ cmin, int* __restrict__ cmax) {
  __shared__ int smin[32], smax[32];
  const int rb = blockIdx.x;
  const int r0 = rb * R, r1 = min(m, r0 + R);
  const int s = row_ptr[r0], e = row_ptr[r1];
  int mn = 0x7fffffff, mx = -1;
  // columns ascend inside a row, so the first and last entry of every row suffice
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    const int a = row_ptr[r], b = row_ptr[r + 1];
    if (b > a) { mn = min(mn, col[a]); mx = max(mx, col[b - 1]); }
  }
  (void)s; (void)e;
  for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = min(mn, smin[w]); mx = max(mx, smax[w]); }
    cmin[rb] = mn; cmax[rb] = mx;
  }
}

// key of entry q = (global sub-panel index << 32) | q : sorting the keys groups the entries by sub-panel and keeps the CSR
// order (row, then column) inside each group
__global__ void band_keys_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int m, int64_t nnz, int R, int W,
                                 const int* __restrict__ win_lo, const int* __restrict__ sp_base, unsigned long long* __restrict__ keys) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = m;                     // last row r with row_ptr[r] <= q
    while (lo < hi) {
      const int mid = lo + ((hi - lo + 1) >> 1);
      if (row_ptr[mid] <= q) lo = mid; else hi = mid - 1;
    }
    const int rb = lo / R;
    const int sp = sp_base[rb] + (col[q] - win_lo[rb]) / W;
    keys[q] = ((unsigned long long)sp << 32) | (unsigned long long)q;
  }
}

// seg_cnt[sp] = entries of sub-panel sp, from the sorted keys
__global__ void band_count_kernel(const unsigned long long* __restrict__ keys, int64_t nnz, int nsp, int* __restrict__ seg_cnt) {
  const int sp = blockIdx.x * blockDim.x + threadIdx.x;
  if (sp >= nsp) return;
  auto lower = [&](unsigned long long key) {
    int64_t lo = 0, hi = nnz;
    while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if (keys[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
  };
  const int64_t a = lower((unsigned long long)sp << 32), b = lower((unsigned long long)(sp + 1) << 32);
  seg_cnt[sp] = (int)(b - a);
}

__global__ void band_pad4_kernel(const int* __restrict__ cnt, int nsp, int* __restrict__ padded) {
  const int sp = blockIdx.x * blockDim.x + threadIdx.x;
  if (sp <= nsp) padded[sp] = sp < nsp ? (cnt[sp] + 3) & ~3 : 0;
}

// scatter the sorted entries into the padded layout
__global__ void band_fill_kernel(const unsigned long long* __restrict__ keys, int64_t nnz, const int* __restrict__ row_ptr, const int* __restrict__ col,
                                 const double* __restrict__ val, int m, int R, int W, const int* __restrict__ win_lo, const int* __restrict__ sp_base,
                                 const int* __restrict__ seg_raw /* exclusive scan of the unpadded counts */, const int* __restrict__ seg_off,
                                 double* __restrict__ bval, unsigned* __restrict__ bmeta) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = keys[i];
    const int sp = (int)(key >> 32);
    const int64_t q = (int64_t)(key & 0xffffffffull);
    int lo = 0, hi = m;
    while (lo < hi) {
      const int mid = lo + ((hi - lo + 1) >> 1);
      if (row_ptr[mid] <= q) lo = mid; else hi = mid - 1;
    }
    const int rb = lo / R;
    const int lcol = col[q] - win_lo[rb] - (sp - sp_base[rb]) * W;
    const int64_t dst = (int64_t)seg_off[sp] + (i - seg_raw[sp]);
    bval[dst] = val[q];
    bmeta[dst] = ((unsigned)(lo - rb * R) << 16) | (unsigned)lcol;
  }
}

__global__ void band_windows_kernel(const int* __restrict__ cmin, const int* __restrict__ cmax, int nrb, int W, int* __restrict__ win_lo, int* __restrict__ nsp) {
  const int rb = blockIdx.x * blockDim.x + threadIdx.x;
  if (rb > nrb) return;
  if (rb == nrb) { nsp[rb] = 0; return; }
  if (cmax[rb] < cmin[rb]) { win_lo[rb] = 0; nsp[rb] = 0; return; }
  const int lo = cmin[rb] & ~1;
  win_lo[rb] = lo;
  nsp[rb] = (cmax[rb] - lo) / W + 1;
}

void sla_csr_free_band(sla_csr* A) {
  sla_band_plan* B = (sla_band_plan*)A->band;
  if (!B) return;
  bsell_free_dev(B->sell);
  cudaFree(B->win_lo); cudaFree(B->sp_base); cudaFree(B->seg_off); cudaFree(B->val); cudaFree(B->meta);
  delete B;
  A->band = nullptr;
}

static size_t band_smem_bytes(int R, int W) { return (size_t)BAND_STAGES * BAND_STAGE_BYTES + 8 * (size_t)(2 * W + R); }

// Builds the band plan when SLA_SPMV_BAND=1 asks for it (OPT-IN), for single-GPU matrices whose row blocks stay within
// BAND_MAX_SP sub-panels.  SLA_SPMV_BAND=2 additionally applies the automatic test (>= 2^20 entries and x re-reads, 8 W bytes per
// sub-panel, below half of the 12 nnz byte stream).
// MEASURED (B200, profiles/r02_band_ab.txt, r02_spmv_band.summary.txt): bit-identical to the tile kernel on every test, DRAM traffic
// equal to the algorithmic bytes (3.92 GB for cfg 2 banded) — but 2.0-2.1 ms against 1.32 ms for the tile kernel on the +-65536
// family, 0.46 vs 0.25 ms on the 4096^2 stencil: with x in shared memory the kernel is bound by INSTRUCTION ISSUE, not by the
// gather port — ~120 thread-instructions per entry for the in-order row sums over 12-byte packed entries (ncu: issue slots 32 % busy
// at 16-32 warps per SM, top stalls short_scoreboard / wait / barrier), where the HBM roofline leaves ~65.  The shared-memory
// gather itself delivers what the micro-benchmark promised; the ordered segmented sum around it is what has to get cheaper
// (rank-in-row bits in the packed entry + per-warp rounds is the next design) before this can become the default.
static sla_status build_band_plan(sla_ctx* c, sla_csr* A) {
  sla_csr_free_band(A);
  int want = 0;
  if (const char* e = getenv("SLA_SPMV_BAND")) want = atoi(e);
  if (want <= 0 || A->dist || A->m == 0 || A->nnz == 0 || A->n >= (1LL << 31) - 2) return SLA_OK;
  if (want == 3 || want == 4) {
    // sliced-ELL form (spmv_bandsell.cuh), built on the host; 4 = with the automatic test
    if (want == 4 && A->nnz < (1 << 20)) return SLA_OK;
    int R = 8192, W = 8192;
    if (const char* e = getenv("SLA_BAND_R")) R = atoi(e);
    if (const char* e = getenv("SLA_BAND_W")) W = atoi(e);
    if (R < 32 || R > 32768 || W < 32 || W > 32768 || (W & 1) || bsell_smem_bytes(R, W) > 227u * 1024u - 1024u)
      return sla_fail(c, SLA_ERR_INVALID, "band plan: SLA_BAND_R / SLA_BAND_W out of range");
    // a cheap look at the column windows on the device before the matrix is copied to the host for the conversion
    {
      const int m = (int)A->m, nrb = (m + R - 1) / R;
      int *cmin = nullptr, *cmax = nullptr;
      if (cudaMalloc(&cmin, sizeof(int) * nrb) != cudaSuccess || cudaMalloc(&cmax, sizeof(int) * nrb) != cudaSuccess) {
        cudaGetLastError(); cudaFree(cmin); return SLA_OK;
      }
      band_range_kernel<<<nrb, 256, 0, c->stream>>>(A->row_ptr, A->col, m, R, cmin, cmax);
      c->launches++;
      std::vector<int> hmin(nrb), hmax(nrb);
      cudaMemcpyAsync(hmin.data(), cmin, sizeof(int) * nrb, cudaMemcpyDeviceToHost, c->stream);
      cudaMemcpyAsync(hmax.data(), cmax, sizeof(int) * nrb, cudaMemcpyDeviceToHost, c->stream);
      const cudaError_t e = cudaStreamSynchronize(c->stream);
      cudaFree(cmin); cudaFree(cmax);
      if (e != cudaSuccess) return sla_fail(c, SLA_ERR_CUDA, "band plan: CUDA error");
      long long cells = 0; int worst = 0;
      for (int rb = 0; rb < nrb; ++rb) {
        if (hmax[rb] < hmin[rb]) continue;
        const int nc = hmax[rb] / W - hmin[rb] / W + 1;
        cells += nc; worst = nc > worst ? nc : worst;
      }
      if (worst > BSELL_MAX_CELLS) return SLA_OK;                                        // not a banded matrix
      if (want == 4) {
        if (8.0 * W * (double)cells > 0.5 * 12.0 * (double)A->nnz) return SLA_OK;          // staging x would cost more than half the entry stream
        if (cells < 4LL * nrb) return SLA_OK;        // a narrow band: the tile kernel's gathers already hit L1 / L2 lines (cfg 3: 82 %)
      }
    }
    sla_bsell_dev* D = nullptr;
    SLA_TRY(bsell_build_dev(c, A, R, W, want == 4, &D));
    if (!D) return SLA_OK;
    sla_band_plan* B = new (std::nothrow) sla_band_plan();
    if (!B) { bsell_free_dev(D); return sla_fail(c, SLA_ERR_ALLOC, "band plan alloc"); }
    memset(B, 0, sizeof(*B));
    B->sell = D; B->R = R; B->W = W; B->nrb = D->nrb;
    A->band = B;
    return SLA_OK;
  }
  if (want == 2) { want = -1; if (A->nnz < (1 << 20)) return SLA_OK; }
  int R = 8192, W = 4096;
  if (const char* e = getenv("SLA_BAND_R")) R = atoi(e);
  if (const char* e = getenv("SLA_BAND_W")) W = atoi(e);
  if (R < 16 || R > 65535 || W < 16 || W > 65536 || (W & 1) || band_smem_bytes(R, W) > 227u * 1024u - 6144u)       // static shared memory of the kernel: ~4.2 KB
    return sla_fail(c, SLA_ERR_INVALID, "band plan: SLA_BAND_R / SLA_BAND_W out of range");
  const int m = (int)A->m;
  const int nrb = (m + R - 1) / R;
  int *cmin = nullptr, *cmax = nullptr, *nsp_d = nullptr, *seg_cnt = nullptr, *seg_pad = nullptr, *seg_raw = nullptr;
  unsigned long long *keys = nullptr, *keys2 = nullptr;
  void* tmp = nullptr;
  sla_band_plan* B = new (std::nothrow) sla_band_plan();
  if (!B) return sla_fail(c, SLA_ERR_ALLOC, "band plan alloc");
  memset(B, 0, sizeof(*B));
  B->R = R; B->W = W; B->nrb = nrb;
  sla_status s = SLA_OK;
  bool take = false;
  do {
    if (cudaMalloc(&cmin, sizeof(int) * nrb) != cudaSuccess || cudaMalloc(&cmax, sizeof(int) * nrb) != cudaSuccess ||
        cudaMalloc(&nsp_d, sizeof(int) * (nrb + 1)) != cudaSuccess || cudaMalloc(&B->win_lo, sizeof(int) * nrb) != cudaSuccess ||
        cudaMalloc(&B->sp_base, sizeof(int) * (nrb + 1)) != cudaSuccess) { s = sla_fail(c, SLA_ERR_ALLOC, "band plan: cudaMalloc failed"); break; }
    band_range_kernel<<<nrb, 256, 0, c->stream>>>(A->row_ptr, A->col, m, R, cmin, cmax);
    band_windows_kernel<<<(nrb + 1 + 255) / 256, 256, 0, c->stream>>>(cmin, cmax, nrb, W, B->win_lo, nsp_d);
    c->launches += 2;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, nsp_d, B->sp_base, nrb + 1, c->stream);
    if (cudaMalloc(&tmp, tb ? tb : 1) != cudaSuccess) { s = sla_fail(c, SLA_ERR_ALLOC, "band plan: cudaMalloc failed"); break; }
    cub::DeviceScan::ExclusiveSum(tmp, tb, nsp_d, B->sp_base, nrb + 1, c->stream);
    std::vector<int> h_nsp(nrb + 1);
    int h_total = 0;
    cudaMemcpyAsync(h_nsp.data(), nsp_d, sizeof(int) * (nrb + 1), cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(&h_total, B->sp_base + nrb, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { s = sla_fail(c, SLA_ERR_CUDA, "band plan: CUDA error"); break; }
    int worst = 0;
    for (int rb = 0; rb < nrb; ++rb) worst = h_nsp[rb] > worst ? h_nsp[rb] : worst;
    if (worst > BAND_MAX_SP) break;                                        // not a banded matrix
    const double x_bytes = 8.0 * W * (double)h_total, stream_bytes = 12.0 * (double)A->nnz;
    if (want < 0 && x_bytes > 0.5 * stream_bytes) break;
    B->nsp = h_total;
    // entries sorted by (sub-panel, CSR position)
    const int64_t nnz = A->nnz;
    if (cudaMalloc(&keys, sizeof(unsigned long long) * (size_t)nnz) != cudaSuccess || cudaMalloc(&keys2, sizeof(unsigned long long) * (size_t)nnz) != cudaSuccess) {
      cudaGetLastError(); break;                                           // not enough memory for the conversion: keep the tile kernel
    }
    int64_t blocks = (nnz + 255) / 256; if (blocks > SLA_NUM_SMS * 32) blocks = SLA_NUM_SMS * 32;
    band_keys_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(A->row_ptr, A->col, m, nnz, R, W, B->win_lo, B->sp_base, keys);
    c->launches++;
    int sp_bits = 1; while ((1LL << sp_bits) <= h_total) ++sp_bits;
    size_t sb = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, sb, keys, keys2, (int)nnz, 32, 32 + sp_bits, c->stream);
    cudaFree(tmp); tmp = nullptr;
    if (cudaMalloc(&tmp, sb ? sb : 1) != cudaSuccess) { cudaGetLastError(); break; }
    // the low 32 bits (CSR position) are already ascending and the sort is stable: only the sub-panel bits need sorting
    cub::DeviceRadixSort::SortKeys(tmp, sb, keys, keys2, (int)nnz, 32, 32 + sp_bits, c->stream);
    c->launches += 4;
    if (cudaMalloc(&seg_cnt, sizeof(int) * (h_total + 1)) != cudaSuccess || cudaMalloc(&seg_pad, sizeof(int) * (h_total + 1)) != cudaSuccess ||
        cudaMalloc(&seg_raw, sizeof(int) * (h_total + 1)) != cudaSuccess || cudaMalloc(&B->seg_off, sizeof(int) * (h_total + 1)) != cudaSuccess) {
      s = sla_fail(c, SLA_ERR_ALLOC, "band plan: cudaMalloc failed"); break;
    }
    cudaMemsetAsync(seg_cnt, 0, sizeof(int) * (h_total + 1), c->stream);
    band_count_kernel<<<(h_total + 255) / 256, 256, 0, c->stream>>>(keys2, nnz, h_total, seg_cnt);
    band_pad4_kernel<<<(h_total + 1 + 255) / 256, 256, 0, c->stream>>>(seg_cnt, h_total, seg_pad);
    c->launches += 2;
    size_t tb2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, seg_cnt, seg_raw, h_total + 1, c->stream);
    if (tb2 > sb) { cudaFree(tmp); tmp = nullptr; if (cudaMalloc(&tmp, tb2) != cudaSuccess) { s = sla_fail(c, SLA_ERR_ALLOC, "band plan: cudaMalloc failed"); break; } }
    cub::DeviceScan::ExclusiveSum(tmp, tb2, seg_cnt, seg_raw, h_total + 1, c->stream);
    cub::DeviceScan::ExclusiveSum(tmp, tb2, seg_pad, B->seg_off, h_total + 1, c->stream);
    int h_nent = 0;
    cudaMemcpyAsync(&h_nent, B->seg_off + h_total, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { s = sla_fail(c, SLA_ERR_CUDA, "band plan: CUDA error"); break; }
    B->nent = h_nent;
    if (cudaMalloc(&B->val, sizeof(double) * (size_t)(h_nent + 4)) != cudaSuccess || cudaMalloc(&B->meta, sizeof(unsigned) * (size_t)(h_nent + 4)) != cudaSuccess) {
      cudaGetLastError(); break;
    }
    cudaMemsetAsync(B->val, 0, sizeof(double) * (size_t)(h_nent + 4), c->stream);
    cudaMemsetAsync(B->meta, 0xff, sizeof(unsigned) * (size_t)(h_nent + 4), c->stream);          // padding = BAND_PAD_META
    band_fill_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(keys2, nnz, A->row_ptr, A->col, A->val, m, R, W, B->win_lo, B->sp_base, seg_raw, B->seg_off,
                                                             B->val, B->meta);
    c->launches++;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) { s = sla_fail(c, SLA_ERR_CUDA, "band plan: CUDA error while filling"); break; }
    take = true;
  } while (0);
  cudaFree(cmin); cudaFree(cmax); cudaFree(nsp_d); cudaFree(seg_cnt); cudaFree(seg_pad); cudaFree(seg_raw); cudaFree(keys); cudaFree(keys2); cudaFree(tmp);
  if (take && s == SLA_OK) { A->band = B; return SLA_OK; }
  A->band = B; sla_csr_free_band(A);
  cudaGetLastError();
  return s;
}

template <int EPI>
static sla_status band_launch_epi(sla_ctx* c, const sla_csr* A, const double* x, double* y, const double* u0, int fin, int dst) {
  const sla_band_plan* B = (const sla_band_plan*)A->band;
  const size_t smem = band_smem_bytes(B->R, B->W);
  static size_t attr_set[64] = {0};              // largest dynamic size registered per device for this instantiation
  if (attr_set[c->device & 63] < smem) {
    SLA_CUDA(c, cudaFuncSetAttribute(spmv_band_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[c->device & 63] = smem;
  }
  BandArgs P;
  P.win_lo = B->win_lo; P.sp_base = B->sp_base; P.seg_off = B->seg_off; P.val = B->val; P.meta = B->meta;
  P.R = B->R; P.W = B->W; P.nrb = B->nrb; P.m = (int)A->m; P.n_even = (long long)((A->n + 1) & ~(int64_t)1);
  const int grid = B->nrb < SLA_NUM_SMS ? B->nrb : SLA_NUM_SMS;
  spmv_band_kernel<EPI><<<grid, BAND_THREADS, smem, c->stream>>>(P, x, y, u0, c->partials);
  SLA_LAUNCH_CHECK(c);
  if (EPI != EPI_NONE) {
    const sla_red_plan rp = sla_red_begin(c, fin, 2);
    partials_reduce_kernel<<<1, PRED_THREADS, 0, c->stream>>>(c->partials, grid, c->partials + 2 * (size_t)SLA_MAX_PARTIALS, c->counter, c->scal, rp.fin, dst, rp.pa);
    SLA_LAUNCH_CHECK(c);
    SLA_TRY(sla_red_end(c, rp, 2, fin, dst));
  }
  return SLA_OK;
}

static sla_status band_launch(sla_ctx* c, const sla_csr* A, const double* x, double* y, int epi, const double* u0, int fin, int dst) {
  if (((const sla_band_plan*)A->band)->sell) return bsell_launch(c, A, ((const sla_band_plan*)A->band)->sell, x, y, epi, u0, fin, dst);
  switch (epi) {
    case EPI_NONE:    return band_launch_epi<EPI_NONE>(c, A, x, y, u0, fin, dst);
    case EPI_DOT1:    return band_launch_epi<EPI_DOT1>(c, A, x, y, u0, fin, dst);
    case EPI_DOT2_YY: return band_launch_epi<EPI_DOT2_YY>(c, A, x, y, u0, fin, dst);
    case EPI_RESNORM: return band_launch_epi<EPI_RESNORM>(c, A, x, y, u0, fin, dst);
  }
  return sla_fail(c, SLA_ERR_INVALID, "spmv: unknown epilogue");
}
