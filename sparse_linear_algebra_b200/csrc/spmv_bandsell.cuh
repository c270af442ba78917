// spmv_bandsell.cuh — second form of the band plan (spmv_band.cuh): x staged through shared memory, entries in a SLICED-ELL layout
// per (row block, column sub-panel) cell so that one THREAD walks one row's entries of the cell (included by spmv_band.cuh).
//
// Why.  The first band kernel keeps the 12-byte CSR-like stream and pays ~120 thread-instructions per entry for in-order segmented
// row sums across lanes (instruction-issue-bound, profiles/r02_band_ab.txt).  Here a row of a cell belongs to one lane:
//
//   * rows are cut into blocks of R rows, columns into a GLOBAL grid of sub-panels of W columns; cell (rb, p) = the entries of block rb
//     whose column lies in sub-panel p.  A block touches a contiguous range of sub-panels (its cells, ascending);
//   * inside a cell only the rows that HAVE entries there are listed, sorted by their entry count in the cell (descending, stable), in
//     slices of 32 rows: descriptor (local row : 16 | count : 16) per listed row, one entry offset per slice.  The entries of a
//     slice are stored column-major — entry j of lane l at slice_base + 32 j + l — as fp64 value + 16-bit LOCAL column: 10 bytes
//     per entry; a slice is as long as its first (longest) row, so padding only appears where two count classes meet (< 2 %);
//   * one persistent CTA per SM walks row blocks; per cell it has the W doubles of x in shared memory (cp.async, double-buffered:
//     sub-panel c+1 arrives while c is consumed); every warp takes BATCH slices at a time: descriptors, then up to
//     UNROLL x BATCH independent (value, column) loads per lane in flight, then each lane folds ITS row:
//     acc = acc[row]; acc = acc + a_ij * xs[j] for the row's entries in ascending column order (__dmul_rn / __dadd_rn); acc[row] = acc;
//   * after the last cell the R accumulators are the rows' results (y, or the fused Krylov epilogues).
// A row's sum is the reference's strict left fold over ascending columns — cell after cell, entry after entry — so the result is
// BIT-IDENTICAL to the tile kernel and to the Haskell result for rows of any length.
//
// Bytes: 10 per entry + 4 per (row, cell) pair that has entries + x from L2; for the +-65536 family of cfg 2 (R = W = 8192: 17 cells per
// block, ~1.9 entries per row and cell) ~388 B per row against the 404 B of the algorithmic count.
//
// The plan is built ON THE HOST from a copy of the CSR arrays (plain loops over row blocks, a thread per block): that code and
// bsell_emulate_host — the kernel's loop nest, statement for statement, on the CPU — are exercised WITHOUT a GPU by
// tests/test_bandsell_host.py through sla_debug_bsell_host, which pins layout and summation order against the CPU restatement of the reference.
#pragma once
#include <atomic>
#include <thread>
#include <vector>

// Launch shapes (template parameters of the kernel): threads per CTA, slices a warp has in flight, entries per lane and slice loaded
// before the fold starts.  One CTA per SM: the threads share the whole register file, so more warps mean fewer slices per warp.
#define BSELL_MAX_CELLS 1024                // cells per row block the plan accepts

struct bsell_cell {
  long long ent_base;                       // first entry of the cell (multiple of 32)
  unsigned desc_base;                       // first descriptor (multiple of 32)
  unsigned slice_base;                      // first slice offset
  int nslices;
  int x_start;                              // first column of the sub-panel (multiple of W, W even)
};

struct bsell_host_plan {
  int R, W, nrb, m;
  long long n;
  std::vector<int> cell_first;              // nrb + 1
  std::vector<bsell_cell> cells;
  std::vector<unsigned> desc;               // (local row << 16) | count ; 0 = padding
  std::vector<unsigned> slice_off;          // entry offset of the slice inside its cell
  std::vector<double> val;
  std::vector<unsigned short> col;
  int max_cells;                            // largest cell count of a block
};

// ---- host side: plan -------------------------------------------------------------------------------------------------------------
struct bsell_block_tmp {
  std::vector<bsell_cell> cells;            // bases relative to the block
  std::vector<unsigned> desc, slice_off;
  long long nent;
  bool bad;                                 // a cell's entries do not fit a 32-bit slice offset
};

// layout of one row block (no entries yet): descriptors, slices, sizes
static void bsell_layout_block(const int* row_ptr, const int* col, int m, int R, int W, int rb, bsell_block_tmp* out, std::vector<unsigned short>& cnt) {
  out->cells.clear(); out->desc.clear(); out->slice_off.clear(); out->nent = 0; out->bad = false;
  const int r0 = rb * R, r1 = (r0 + R < m) ? r0 + R : m;
  int cmin = 0x7fffffff, cmax = -1;
  for (int r = r0; r < r1; ++r) {
    const int a = row_ptr[r], b = row_ptr[r + 1];
    if (b > a) { if (col[a] < cmin) cmin = col[a]; if (col[b - 1] > cmax) cmax = col[b - 1]; }      // columns ascend inside a row
  }
  if (cmax < cmin) return;
  const int p0 = cmin / W, p1 = cmax / W, nc = p1 - p0 + 1;
  const int nr = r1 - r0;
  cnt.assign((size_t)nc * nr, 0);
  for (int r = r0; r < r1; ++r)
    for (int q = row_ptr[r]; q < row_ptr[r + 1]; ++q) cnt[(size_t)(col[q] / W - p0) * nr + (r - r0)]++;
  std::vector<int> hist, order;
  for (int c = 0; c < nc; ++c) {
    const unsigned short* cc = &cnt[(size_t)c * nr];
    int mx = 0, nk = 0;
    for (int i = 0; i < nr; ++i) { if (cc[i] > mx) mx = cc[i]; nk += cc[i] ? 1 : 0; }
    bsell_cell cl;
    cl.ent_base = out->nent; cl.desc_base = (unsigned)out->desc.size(); cl.slice_base = (unsigned)out->slice_off.size();
    cl.x_start = (p0 + c) * W; cl.nslices = (nk + 31) / 32;
    // rows with entries here, by count descending, ties by row ascending (counting sort)
    hist.assign((size_t)mx + 2, 0);
    for (int i = 0; i < nr; ++i) if (cc[i]) hist[mx - cc[i] + 1]++;
    for (int k = 1; k <= mx + 1; ++k) hist[k] += hist[k - 1];
    order.assign((size_t)cl.nslices * 32, -1);
    for (int i = 0; i < nr; ++i) if (cc[i]) order[hist[mx - cc[i]]++] = i;
    long long off = 0;
    for (int s = 0; s < cl.nslices; ++s) {
      out->slice_off.push_back((unsigned)off);
      const int len = cc[order[(size_t)s * 32]];                         // the slice's first row is its longest
      for (int l = 0; l < 32; ++l) {
        const int i = order[(size_t)s * 32 + l];
        out->desc.push_back(i < 0 ? 0u : ((unsigned)i << 16) | (unsigned)cc[i]);
      }
      off += 32LL * len;
      if (off > 0xfffffff0LL) out->bad = true;
    }
    out->nent += off;
    out->cells.push_back(cl);
  }
}

// entries of one row block into the global arrays (the block's descriptors are already in place)
static void bsell_fill_block(const int* row_ptr, const int* col, const double* val, int m, int R, int W, int rb, const bsell_host_plan& P,
                             double* pval, unsigned short* pcol, std::vector<unsigned>& slot, std::vector<unsigned short>& seen) {
  const int r0 = rb * R, r1 = (r0 + R < m) ? r0 + R : m, nr = r1 - r0;
  const int c0 = P.cell_first[rb], nc = P.cell_first[rb + 1] - c0;
  if (nc == 0) return;
  const int p0 = P.cells[c0].x_start / W;
  // slot[c][i] = offset of row i's first entry inside cell c (slice offset + lane)
  slot.assign((size_t)nc * nr, 0);
  seen.assign((size_t)nc * nr, 0);
  for (int c = 0; c < nc; ++c) {
    const bsell_cell& cl = P.cells[c0 + c];
    for (int s = 0; s < cl.nslices; ++s)
      for (int l = 0; l < 32; ++l) {
        const unsigned d = P.desc[(size_t)cl.desc_base + 32u * s + l];
        if ((d & 0xffffu) == 0) continue;
        slot[(size_t)c * nr + (d >> 16)] = P.slice_off[(size_t)cl.slice_base + s] + (unsigned)l;
      }
  }
  for (int r = r0; r < r1; ++r)
    for (int q = row_ptr[r]; q < row_ptr[r + 1]; ++q) {
      const int c = col[q] / W - p0;
      const size_t at = (size_t)c * nr + (r - r0);
      const long long dst = P.cells[c0 + c].ent_base + slot[at] + 32LL * seen[at];
      seen[at]++;
      pval[dst] = val[q];
      pcol[dst] = (unsigned short)(col[q] - P.cells[c0 + c].x_start);
    }
}

// false: the matrix does not fit the plan (a block spans more than max_cells sub-panels, or sizes overflow the 32-bit bases)
static bool bsell_build_host(const int* row_ptr, const int* col, const double* val, int m, long long n, int R, int W, int max_cells,
                             bsell_host_plan* P, int nthreads) {
  P->R = R; P->W = W; P->m = m; P->n = n; P->nrb = (m + R - 1) / R; P->max_cells = 0;
  const int nrb = P->nrb;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nrb) nthreads = nrb > 0 ? nrb : 1;
  std::vector<bsell_block_tmp> tmp((size_t)nrb);
  {
    std::atomic<int> next(0);
    auto work = [&]() {
      std::vector<unsigned short> cnt;
      for (int rb = next++; rb < nrb; rb = next++) bsell_layout_block(row_ptr, col, m, R, W, rb, &tmp[rb], cnt);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
  }
  P->cell_first.assign((size_t)nrb + 1, 0);
  long long nent = 0;
  unsigned long long ndesc = 0, nslice = 0;
  for (int rb = 0; rb < nrb; ++rb) {
    const int nc = (int)tmp[rb].cells.size();
    if (nc > max_cells || tmp[rb].bad) return false;
    if (nc > P->max_cells) P->max_cells = nc;
    P->cell_first[rb + 1] = P->cell_first[rb] + nc;
    ndesc += tmp[rb].desc.size(); nslice += tmp[rb].slice_off.size(); nent += tmp[rb].nent;
  }
  if (ndesc >= 0xffffffffull || nslice >= 0xffffffffull) return false;
  P->cells.resize((size_t)P->cell_first[nrb]);
  P->desc.resize((size_t)ndesc); P->slice_off.resize((size_t)nslice);
  {
    long long e = 0; unsigned long long d = 0, s = 0;
    for (int rb = 0; rb < nrb; ++rb) {
      bsell_block_tmp& t = tmp[rb];
      for (size_t c = 0; c < t.cells.size(); ++c) {
        bsell_cell cl = t.cells[c];
        cl.ent_base += e; cl.desc_base += (unsigned)d; cl.slice_base += (unsigned)s;
        P->cells[(size_t)P->cell_first[rb] + c] = cl;
      }
      if (!t.desc.empty()) memcpy(&P->desc[(size_t)d], t.desc.data(), t.desc.size() * sizeof(unsigned));
      if (!t.slice_off.empty()) memcpy(&P->slice_off[(size_t)s], t.slice_off.data(), t.slice_off.size() * sizeof(unsigned));
      e += t.nent; d += t.desc.size(); s += t.slice_off.size();
      std::vector<unsigned>().swap(t.desc); std::vector<unsigned>().swap(t.slice_off);
    }
  }
  P->val.assign((size_t)nent, 0.0);
  P->col.assign((size_t)nent, 0);
  {
    std::atomic<int> next(0);
    auto work = [&]() {
      std::vector<unsigned> slot; std::vector<unsigned short> seen;
      for (int rb = next++; rb < nrb; rb = next++) bsell_fill_block(row_ptr, col, val, m, R, W, rb, *P, P->val.data(), P->col.data(), slot, seen);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
  }
  return true;
}

// The kernel's loop nest on the CPU: same cells, slices, lanes and statement order (lanes of a slice are independent rows, so running
// them one after the other changes nothing).  x is staged exactly as the kernel stages it (clipped at n).
static void bsell_emulate_host(const bsell_host_plan& P, const double* x, double* y) {
  std::vector<double> acc((size_t)P.R), xs((size_t)P.W);
  for (int rb = 0; rb < P.nrb; ++rb) {
    const int row0 = rb * P.R, nrows = (P.R < P.m - row0) ? P.R : P.m - row0;
    for (int j = 0; j < P.R; ++j) acc[j] = 0.0;
    for (int c = P.cell_first[rb]; c < P.cell_first[rb + 1]; ++c) {
      const bsell_cell& cl = P.cells[c];
      long long cntx = P.n - cl.x_start; if (cntx > P.W) cntx = P.W;
      for (long long i = 0; i < cntx; ++i) xs[(size_t)i] = x[cl.x_start + i];
      for (int s = 0; s < cl.nslices; ++s)
        for (int lane = 0; lane < 32; ++lane) {
          const unsigned d = P.desc[(size_t)cl.desc_base + 32u * s + lane];
          const int cnt = (int)(d & 0xffffu), row = (int)(d >> 16);
          if (cnt == 0) continue;
          const long long base = cl.ent_base + P.slice_off[(size_t)cl.slice_base + s] + lane;
          double a = acc[row];
          for (int j = 0; j < cnt; ++j) {
            volatile double prod = P.val[(size_t)(base + 32LL * j)] * xs[P.col[(size_t)(base + 32LL * j)]];      // volatile: no contraction into an FMA
            a = a + prod;
          }
          acc[row] = a;
        }
    }
    for (int j = 0; j < nrows; ++j) y[row0 + j] = acc[j];
  }
}

// Host-only entry (no GPU, no context): plan + emulation for a host CSR.  stats[0..4] = cells, padded entries, descriptors, largest
// cell count of a block, slices.  Returns 0, or 1 when the matrix does not fit the plan, 2 on bad arguments.
extern "C" int sla_debug_bsell_host(int m, int64_t n, const int32_t* row_ptr, const int32_t* col, const double* val, int R, int W, int threads,
                                    const double* x, double* y, int64_t* stats) {
  if (m < 0 || n < 0 || !row_ptr || R < 1 || R > 32768 || W < 2 || W > 32768 || (W & 1) || (m > 0 && row_ptr[m] > 0 && (!col || !val)) || !x || !y) return 2;
  bsell_host_plan P;
  if (!bsell_build_host(row_ptr, col, val, m, (long long)n, R, W, BSELL_MAX_CELLS, &P, threads)) return 1;
  bsell_emulate_host(P, x, y);
  if (stats) {
    stats[0] = (int64_t)P.cells.size(); stats[1] = (int64_t)P.val.size(); stats[2] = (int64_t)P.desc.size(); stats[3] = P.max_cells;
    stats[4] = (int64_t)P.slice_off.size();
  }
  return 0;
}

// ---- device side ---------------------------------------------------------------------------------------------------------------
struct sla_bsell_dev {
  int R, W, nrb;
  int* cell_first; bsell_cell* cells; unsigned* desc; unsigned* slice_off; double* val; unsigned short* col;
  long long nent, ndesc;
};

struct BsellArgs {
  const int* cell_first; const bsell_cell* cells; const unsigned* desc; const unsigned* slice_off; const double* val; const unsigned short* col;
  int R, W, nrb, m; long long n;
};

__device__ __forceinline__ unsigned ld_stream_u16(const unsigned short* p, uint64_t pol) {
  unsigned short r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(r) : "l"(p), "l"(pol));
  return (unsigned)r;
}
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}

template <int EPI, int BSELL_THREADS, int BSELL_BATCH, int BSELL_UNROLL, int BSELL_PF>
__global__ void __launch_bounds__(BSELL_THREADS, 1)
spmv_bsell_kernel(BsellArgs P, const double* __restrict__ x, double* y, const double* __restrict__ u0, double* partials) {
  extern __shared__ __align__(128) unsigned char bsell_raw[];
  double* xbuf = reinterpret_cast<double*>(bsell_raw);                  // 2 x W doubles
  double* acc = xbuf + 2 * (size_t)P.W;                                 // R doubles
  __shared__ double red[2 * 32];
  __shared__ int next_slice[2];                                         // per cell parity: the next batch of slices nobody has claimed yet
  const int tid = threadIdx.x, lane = tid & 31;
  const uint64_t pol_stream = policy_evict_first();
  const bool x16 = (reinterpret_cast<uintptr_t>(x) & 15u) == 0;         // sub-panels start at even columns: 16-byte copies if x itself is aligned
  double e0 = 0.0, e1 = 0.0;

  for (int rb = blockIdx.x; rb < P.nrb; rb += gridDim.x) {
    const int row0 = rb * P.R;
    const int nrows = min(P.R, P.m - row0);
    const int c0 = P.cell_first[rb], nc = P.cell_first[rb + 1] - c0;
    for (int j = tid; j < P.R; j += BSELL_THREADS) acc[j] = 0.0;        // sum = strict left fold from 0
    if (tid == 0) next_slice[0] = 0;                                    // (the block's previous cells are behind the barrier that ended it)
    auto stage = [&](int ci) {                                          // x sub-panel of cell ci -> buffer ci & 1 (asynchronous)
      const long long start = P.cells[c0 + ci].x_start;
      long long cnt = P.n - start; if (cnt > P.W) cnt = P.W;
      double* dstb = xbuf + (size_t)(ci & 1) * P.W;
      if (x16) {
        const int pairs = (int)(cnt >> 1);
        for (int i = tid; i < pairs; i += BSELL_THREADS) cp_async_16(dstb + 2 * i, x + start + 2 * i);
        if ((cnt & 1) && tid == 0) cp_async_8(dstb + cnt - 1, x + start + cnt - 1);
      } else {
        for (int i = tid; i < (int)cnt; i += BSELL_THREADS) cp_async_8(dstb + i, x + start + i);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (nc > 0) stage(0);
    for (int ci = 0; ci < nc; ++ci) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");              // my part of sub-panel ci has landed
      __syncthreads();                                                  // everybody's has; cell ci-1 is finished (accumulators, other x buffer)
      if (ci + 1 < nc) stage(ci + 1);
      if (tid == 0) next_slice[(ci + 1) & 1] = 0;                       // cell ci-1 used it; everybody is past that cell
      const bsell_cell cl = P.cells[c0 + ci];
      const double* xs = xbuf + (size_t)(ci & 1) * P.W;
      const unsigned* dsc = P.desc + cl.desc_base;
      const unsigned* sof = P.slice_off + cl.slice_base;
      const double* cv = P.val + cl.ent_base;
      const unsigned short* cc = P.col + cl.ent_base;
      // Batches of BSELL_BATCH slices are handed out through a shared counter, longest slices first (they are sorted that way): the
      // warps stay balanced up to the barrier that ends the cell.  Descriptors are requested TWO batches ahead; with BSELL_PF the
      // entries of the NEXT batch are pulled into L2 (prefetch.global.L2) while the current batch is folded, so that its loads find
      // them there instead of waiting for DRAM.
      auto claim = [&]() {
        int s = 0;
        if (lane == 0) s = atomicAdd(&next_slice[ci & 1], BSELL_BATCH);
        return __shfl_sync(0xffffffffu, s, 0);
      };
      unsigned dn[BSELL_BATCH], son[BSELL_BATCH], d2[BSELL_BATCH], so2[BSELL_BATCH];
      int s0 = claim();
#pragma unroll
      for (int b = 0; b < BSELL_BATCH; ++b) {
        const bool ok = s0 + b < cl.nslices;
        dn[b] = ok ? __ldg(dsc + 32 * (s0 + b) + lane) : 0u;
        son[b] = ok ? __ldg(sof + s0 + b) : 0u;
      }
      int s1 = claim();
#pragma unroll
      for (int b = 0; b < BSELL_BATCH; ++b) {
        const bool ok = s1 + b < cl.nslices;
        d2[b] = ok ? __ldg(dsc + 32 * (s1 + b) + lane) : 0u;
        so2[b] = ok ? __ldg(sof + s1 + b) : 0u;
      }
#pragma unroll 1
      while (s0 < cl.nslices) {
        unsigned d[BSELL_BATCH];
        unsigned base[BSELL_BATCH];                                     // entry offsets inside a cell fit 32 bits (checked by the plan)
        double a[BSELL_BATCH];
        int len = 0;                                                    // longest row of the batch (warp-uniform)
#pragma unroll
        for (int b = 0; b < BSELL_BATCH; ++b) {
          d[b] = dn[b]; base[b] = son[b] + (unsigned)lane;
          dn[b] = d2[b]; son[b] = so2[b];                               // batch s1: requested one iteration ago
          len = max(len, (int)(__shfl_sync(0xffffffffu, d[b], 0) & 0xffffu));      // a slice's first row is its longest
          a[b] = (d[b] & 0xffffu) ? acc[d[b] >> 16] : 0.0;
        }
        const int s2 = claim();
#pragma unroll 1
        for (int j0 = 0; j0 < len; j0 += BSELL_UNROLL) {
          double v[BSELL_BATCH][BSELL_UNROLL];
          unsigned k[BSELL_BATCH][BSELL_UNROLL];
#pragma unroll
          for (int b = 0; b < BSELL_BATCH; ++b) {
            const int cnt = (int)(d[b] & 0xffffu);
            const double* pv = cv + (base[b] + 32u * (unsigned)j0);     // entry u of this round: a constant 32 u behind
            const unsigned short* pc = cc + (base[b] + 32u * (unsigned)j0);
#pragma unroll
            for (int u = 0; u < BSELL_UNROLL; ++u) {
              if (j0 + u < cnt) {                                       // (slots beyond a row's count stay unset and unused)
                v[b][u] = ld_stream_double(pv + 32 * u, pol_stream);
                k[b][u] = ld_stream_u16(pc + 32 * u, pol_stream);
              }
            }
          }
          if (j0 == 0) {
#pragma unroll
            for (int b = 0; b < BSELL_BATCH; ++b) {
              const bool ok = s2 + b < cl.nslices;
              d2[b] = ok ? __ldg(dsc + 32 * (s2 + b) + lane) : 0u;
              so2[b] = ok ? __ldg(sof + s2 + b) : 0u;
              if (BSELL_PF) {
                // batch s1 -> L2: a slice of length L is 2 L lines of values and L half-lines of columns, one per lane
                const int l1 = (int)(__shfl_sync(0xffffffffu, dn[b], 0) & 0xffffu);
                if (lane < 2 * l1) asm volatile("prefetch.global.L2 [%0];" ::"l"(cv + son[b] + 16 * lane));
                if (lane < l1) asm volatile("prefetch.global.L2 [%0];" ::"l"(cc + son[b] + 32 * lane));
              }
            }
          }
#pragma unroll
          for (int b = 0; b < BSELL_BATCH; ++b) {
            const int cnt = (int)(d[b] & 0xffffu);
#pragma unroll
            for (int u = 0; u < BSELL_UNROLL; ++u)
              if (j0 + u < cnt) a[b] = __dadd_rn(a[b], __dmul_rn(v[b][u], xs[k[b][u]]));      // dotu: a_ij * x_j, matrix entry on the left
          }
        }
#pragma unroll
        for (int b = 0; b < BSELL_BATCH; ++b)
          if (d[b] & 0xffffu) acc[d[b] >> 16] = a[b];
        s0 = s1; s1 = s2;
      }
    }
    __syncthreads();                                                    // the accumulators are the rows' results
    for (int j = tid; j < nrows; j += BSELL_THREADS) row_epilogue<EPI>(row0 + j, acc[j], y, u0, e0, e1);
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

static void bsell_free_dev(sla_bsell_dev* D) {
  if (!D) return;
  cudaFree(D->cell_first); cudaFree(D->cells); cudaFree(D->desc); cudaFree(D->slice_off); cudaFree(D->val); cudaFree(D->col);
  delete D;
}

static size_t bsell_smem_bytes(int R, int W) { return 8 * (size_t)(2 * W + R); }

// CSR (device) -> host copy -> plan -> device.  *out stays null when the matrix does not fit the plan (no error).
// auto_test: also require the x re-reads (8 W bytes per cell) to stay below half of the entry stream.
static sla_status bsell_build_dev(sla_ctx* c, const sla_csr* A, int R, int W, bool auto_test, sla_bsell_dev** out) {
  *out = nullptr;
  const int m = (int)A->m;
  const int64_t nnz = A->nnz;
  std::vector<int> h_rp((size_t)m + 1), h_col((size_t)nnz);
  std::vector<double> h_val((size_t)nnz);
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  SLA_CUDA(c, cudaMemcpy(h_rp.data(), A->row_ptr, sizeof(int) * ((size_t)m + 1), cudaMemcpyDeviceToHost));
  SLA_CUDA(c, cudaMemcpy(h_col.data(), A->col, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost));
  SLA_CUDA(c, cudaMemcpy(h_val.data(), A->val, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost));
  bsell_host_plan P;
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 4;
  if (hw > 32) hw = 32;
  if (!bsell_build_host(h_rp.data(), h_col.data(), h_val.data(), m, (long long)A->n, R, W, BSELL_MAX_CELLS, &P, (int)hw)) return SLA_OK;
  if (auto_test && 8.0 * W * (double)P.cells.size() > 0.5 * 12.0 * (double)nnz) return SLA_OK;
  if (auto_test && (double)P.val.size() > 1.25 * (double)nnz) return SLA_OK;      // long rows drag their whole slice along: too much padding
  std::vector<int>().swap(h_col); std::vector<double>().swap(h_val);
  sla_bsell_dev* D = new (std::nothrow) sla_bsell_dev();
  if (!D) return sla_fail(c, SLA_ERR_ALLOC, "band plan alloc");
  memset(D, 0, sizeof(*D));
  D->R = R; D->W = W; D->nrb = P.nrb; D->nent = (long long)P.val.size(); D->ndesc = (long long)P.desc.size();
  auto up = [&](void** dst, const void* src, size_t bytes) -> bool {
    if (cudaMalloc(dst, bytes ? bytes : 16) != cudaSuccess) return false;
    return bytes == 0 || cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
  };
  const bool ok = up((void**)&D->cell_first, P.cell_first.data(), sizeof(int) * P.cell_first.size()) &&
                  up((void**)&D->cells, P.cells.data(), sizeof(bsell_cell) * P.cells.size()) &&
                  up((void**)&D->desc, P.desc.data(), sizeof(unsigned) * P.desc.size()) &&
                  up((void**)&D->slice_off, P.slice_off.data(), sizeof(unsigned) * P.slice_off.size()) &&
                  up((void**)&D->val, P.val.data(), sizeof(double) * P.val.size()) &&
                  up((void**)&D->col, P.col.data(), sizeof(unsigned short) * P.col.size());
  if (!ok) { cudaGetLastError(); bsell_free_dev(D); return SLA_OK; }      // not enough memory for the second copy: keep the tile kernel
  *out = D;
  return SLA_OK;
}

template <int EPI, int THREADS, int BATCH, int UNROLL, int PF>
static sla_status bsell_launch_shape(sla_ctx* c, const sla_csr* A, const sla_bsell_dev* D, const double* x, double* y, const double* u0, int fin, int dst) {
  const size_t smem = bsell_smem_bytes(D->R, D->W);
  static size_t attr_set[64] = {0};              // largest dynamic size registered per device for this instantiation
  if (attr_set[c->device & 63] < smem) {
    SLA_CUDA(c, cudaFuncSetAttribute(spmv_bsell_kernel<EPI, THREADS, BATCH, UNROLL, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[c->device & 63] = smem;
  }
  BsellArgs P;
  P.cell_first = D->cell_first; P.cells = D->cells; P.desc = D->desc; P.slice_off = D->slice_off; P.val = D->val; P.col = D->col;
  P.R = D->R; P.W = D->W; P.nrb = D->nrb; P.m = (int)A->m; P.n = (long long)A->n;
  const int grid = D->nrb < SLA_NUM_SMS ? D->nrb : SLA_NUM_SMS;
  spmv_bsell_kernel<EPI, THREADS, BATCH, UNROLL, PF><<<grid, THREADS, smem, c->stream>>>(P, x, y, u0, c->partials);
  SLA_LAUNCH_CHECK(c);
  if (EPI != EPI_NONE) {
    const sla_red_plan rp = sla_red_begin(c, fin, 2);
    partials_reduce_kernel<<<1, PRED_THREADS, 0, c->stream>>>(c->partials, grid, c->partials + 2 * (size_t)SLA_MAX_PARTIALS, c->counter, c->scal, rp.fin, dst, rp.pa);
    SLA_LAUNCH_CHECK(c);
    SLA_TRY(sla_red_end(c, rp, 2, fin, dst));
  }
  return SLA_OK;
}

template <int EPI>
static sla_status bsell_launch_epi(sla_ctx* c, const sla_csr* A, const sla_bsell_dev* D, const double* x, double* y, const double* u0, int fin, int dst) {
  switch (c->bsell_variant) {                    // option "bsell_variant" / SLA_BSELL_VARIANT; measured on cfg 2 banded: profiles/r02_bandsell_ab.md
    case 1:  return bsell_launch_shape<EPI, 768, 3, 4, 0>(c, A, D, x, y, u0, fin, dst);
    case 2:  return bsell_launch_shape<EPI, 1024, 2, 4, 0>(c, A, D, x, y, u0, fin, dst);
    case 4:  return bsell_launch_shape<EPI, 512, 4, 4, 0>(c, A, D, x, y, u0, fin, dst);
    case 5:  return bsell_launch_shape<EPI, 1024, 2, 3, 1>(c, A, D, x, y, u0, fin, dst);
    case 6:  return bsell_launch_shape<EPI, 1024, 2, 2, 1>(c, A, D, x, y, u0, fin, dst);
    case 7:  return bsell_launch_shape<EPI, 768, 3, 3, 1>(c, A, D, x, y, u0, fin, dst);
    default: return bsell_launch_shape<EPI, 1024, 2, 3, 0>(c, A, D, x, y, u0, fin, dst);      // 3
  }
}

static sla_status bsell_launch(sla_ctx* c, const sla_csr* A, const sla_bsell_dev* D, const double* x, double* y, int epi, const double* u0, int fin, int dst) {
  switch (epi) {
    case EPI_NONE:    return bsell_launch_epi<EPI_NONE>(c, A, D, x, y, u0, fin, dst);
    case EPI_DOT1:    return bsell_launch_epi<EPI_DOT1>(c, A, D, x, y, u0, fin, dst);
    case EPI_DOT2_YY: return bsell_launch_epi<EPI_DOT2_YY>(c, A, D, x, y, u0, fin, dst);
    case EPI_RESNORM: return bsell_launch_epi<EPI_RESNORM>(c, A, D, x, y, u0, fin, dst);
  }
  return sla_fail(c, SLA_ERR_INVALID, "spmv: unknown epilogue");
}
