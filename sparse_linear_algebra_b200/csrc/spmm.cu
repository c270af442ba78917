// spmm.cu — (##) with a dense right operand: C = A ## B, A CSR (m x n), B dense row-major (n x k), C dense
// row-major (m x k).   Reference: matMat_ AB -> matMatUnsafeWith, src/Data/Sparse/SpMatrix.hs:768-811:
//     C_ic = sum over k ascending of  b_kc * a_ik      (dott x y = sum $ liftI2 (*) x y, x = column of B)
// for every stored row of A x every stored column of B (explicit zeros kept — the result is dense).
//
// Two element types:
//   f64  : products rounded once (__dmul_rn, column of B on the left as in `dott`), summed sequentially in
//          ascending k with __dadd_rn from 0.0 — bit-identical to the reference for rows of any length.
//   bf16 : the cfg-5 path of BASELINE.json — A values and B in bf16, fp32 accumulation (FMA) in ascending k,
//          C rounded to bf16 (round-to-nearest-even).
// Layout: one warp per row of A; lanes run along the columns of B, so every B-row gather is one coalesced
// 8 k-byte (f64) / 2 k-byte (bf16) read; A's (col, val) pairs of the row are read once and broadcast.
// The kernel is bound by the B-row gathers (k * elt bytes per stored entry), i.e. by L2 / HBM bandwidth —
// 2 flop per gathered element, far below the tensor-core ridge — so there is no MMA here; a tensor-core tile
// path only pays for block-structured A.
#include "common.cuh"

#include <cuda_bf16.h>
#include <cub/device/device_scan.cuh>
#include <new>
#include <stdlib.h>

sla_status sla_spmm_bsr_tc(sla_ctx* c, const sla_csr* A, const sla_dense* B, sla_dense* C, double min_fill);

#define SPMM_THREADS 256

// ---- f64: exact ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPMM_THREADS)
spmm_f64_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, const double* __restrict__ val,
                const double* __restrict__ B, double* __restrict__ C, int m, int k) {
  const int warp = (blockIdx.x * SPMM_THREADS + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * SPMM_THREADS) >> 5;
  for (int r = warp; r < m; r += nwarps) {
    const int s = row_ptr[r], e = row_ptr[r + 1];
    for (int c0 = 0; c0 < k; c0 += 32) {
      const int c = c0 + lane;
      double acc = 0.0;
      if (c < k)
        for (int p = s; p < e; ++p)
          acc = __dadd_rn(acc, __dmul_rn(B[(size_t)col[p] * k + c], val[p]));   // b_kc * a_ik, ascending k
      if (c < k) C[(size_t)r * k + c] = acc;
    }
  }
}

// ---- bf16 in, fp32 accumulate, bf16 out; k a multiple of 4, lanes take 4 columns each per 128-column slab -----
__global__ void __launch_bounds__(SPMM_THREADS)
spmm_bf16_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, const __nv_bfloat16* __restrict__ val,
                 const __nv_bfloat16* __restrict__ B, __nv_bfloat16* __restrict__ C, int m, int k) {
  const int warp = (blockIdx.x * SPMM_THREADS + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * SPMM_THREADS) >> 5;
  for (int r = warp; r < m; r += nwarps) {
    const int s = row_ptr[r], e = row_ptr[r + 1];
    for (int c0 = 0; c0 < k; c0 += 128) {
      const int c = c0 + 4 * lane;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      if (c < k) {
        int p = s;
        for (; p + 1 < e; p += 2) {          // two independent gathers in flight
          const uint2 u = *reinterpret_cast<const uint2*>(B + (size_t)col[p] * k + c);
          const uint2 w = *reinterpret_cast<const uint2*>(B + (size_t)col[p + 1] * k + c);
          const float v = __bfloat162float(val[p]), v2 = __bfloat162float(val[p + 1]);
          const __nv_bfloat162 b01 = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b23 = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
          const __nv_bfloat162 d01 = *reinterpret_cast<const __nv_bfloat162*>(&w.x), d23 = *reinterpret_cast<const __nv_bfloat162*>(&w.y);
          a0 = fmaf(__low2float(b01), v, a0);  a1 = fmaf(__high2float(b01), v, a1);
          a2 = fmaf(__low2float(b23), v, a2);  a3 = fmaf(__high2float(b23), v, a3);
          a0 = fmaf(__low2float(d01), v2, a0); a1 = fmaf(__high2float(d01), v2, a1);
          a2 = fmaf(__low2float(d23), v2, a2); a3 = fmaf(__high2float(d23), v2, a3);
        }
        if (p < e) {
          const uint2 u = *reinterpret_cast<const uint2*>(B + (size_t)col[p] * k + c);
          const float v = __bfloat162float(val[p]);
          const __nv_bfloat162 b01 = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b23 = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
          a0 = fmaf(__low2float(b01), v, a0); a1 = fmaf(__high2float(b01), v, a1);
          a2 = fmaf(__low2float(b23), v, a2); a3 = fmaf(__high2float(b23), v, a3);
        }
        __nv_bfloat162 o01 = __floats2bfloat162_rn(a0, a1), o23 = __floats2bfloat162_rn(a2, a3);
        uint2 o;
        o.x = *reinterpret_cast<unsigned*>(&o01); o.y = *reinterpret_cast<unsigned*>(&o23);
        *reinterpret_cast<uint2*>(C + (size_t)r * k + c) = o;
      }
    }
  }
}

__global__ void f64_to_bf16_kernel(const double* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __double2bfloat16(in[i]);
}
__global__ void bf16_to_f64_kernel(const __nv_bfloat16* __restrict__ in, double* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (double)__bfloat162float(in[i]);
}

static unsigned gsb(int64_t n) { int64_t b = (n + 255) / 256; if (b < 1) b = 1; if (b > SLA_NUM_SMS * 16) b = SLA_NUM_SMS * 16; return (unsigned)b; }

// ---- dense row-major blocks ------------------------------------------------------------------------------------
static size_t elt_bytes(int dtype) { return dtype == SLA_BF16 ? 2 : 8; }

extern "C" sla_status sla_dense_create(sla_ctx* c, int64_t rows, int64_t cols, int dtype, sla_dense** out) {
  if (!c || !out || rows < 0 || cols < 0 || (dtype != SLA_F64 && dtype != SLA_BF16)) return SLA_ERR_INVALID;
  if (dtype == SLA_BF16 && cols % 4 != 0) return sla_fail(c, SLA_ERR_INVALID, "dense bf16 blocks need a column count that is a multiple of 4");
  sla_dense* d = new (std::nothrow) sla_dense();
  if (!d) return sla_fail(c, SLA_ERR_ALLOC, "dense alloc");
  d->ctx = c; d->rows = rows; d->cols = cols; d->ld = cols; d->dtype = dtype; d->rowmajor = 1; d->d = nullptr;
  const size_t bytes = (size_t)rows * (size_t)cols * elt_bytes(dtype);
  d->bytes = bytes < 16 ? 16 : bytes;
  if (sla_pool_alloc(c, (void**)&d->d, d->bytes) != cudaSuccess) { cudaGetLastError(); delete d; return sla_fail(c, SLA_ERR_ALLOC, "cudaMalloc failed for a dense block"); }
  cudaMemsetAsync(d->d, 0, bytes, c->stream);
  *out = d;
  return SLA_OK;
}

extern "C" sla_status sla_dense_from_host(sla_ctx* c, int64_t rows, int64_t cols, const double* rowmajor, int dtype, sla_dense** out) {
  if (!rowmajor && rows * cols > 0) return SLA_ERR_INVALID;
  SLA_TRY(sla_dense_create(c, rows, cols, dtype, out));
  const int64_t n = rows * cols;
  if (n == 0) return SLA_OK;
  if (dtype == SLA_F64) {
    SLA_CUDA(c, cudaMemcpyAsync((*out)->d, rowmajor, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  } else {
    double* tmp = nullptr;
    SLA_CUDA(c, cudaMalloc(&tmp, sizeof(double) * (size_t)n));
    cudaMemcpyAsync(tmp, rowmajor, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
    f64_to_bf16_kernel<<<gsb(n), 256, 0, c->stream>>>(tmp, (__nv_bfloat16*)(*out)->d, n);
    c->launches++;
    cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(tmp);
    SLA_CUDA(c, e);
  }
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  return SLA_OK;
}

extern "C" sla_status sla_dense_to_host_f64(sla_ctx* c, const sla_dense* d, double* rowmajor_out) {
  if (!c || !d || !rowmajor_out) return SLA_ERR_INVALID;
  if (!d->rowmajor) return sla_fail(c, SLA_ERR_INVALID, "dense_to_host_f64: block is column-major (use sla_dense_to_host)");
  const int64_t n = d->rows * d->cols;
  if (n == 0) return SLA_OK;
  if (d->dtype == SLA_F64) {
    SLA_CUDA(c, cudaMemcpyAsync(rowmajor_out, d->d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    SLA_CUDA(c, cudaStreamSynchronize(c->stream));
    return SLA_OK;
  }
  double* tmp = nullptr;
  SLA_CUDA(c, cudaMalloc(&tmp, sizeof(double) * (size_t)n));
  bf16_to_f64_kernel<<<gsb(n), 256, 0, c->stream>>>((const __nv_bfloat16*)d->d, tmp, n);
  c->launches++;
  cudaMemcpyAsync(rowmajor_out, tmp, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(tmp);
  SLA_CUDA(c, e);
  return SLA_OK;
}

// bf16 copy of A's values, built on first use
static sla_status ensure_val_bf16(sla_ctx* c, const sla_csr* A) {
  if (A->val_bf16 || A->nnz == 0) return SLA_OK;
  void* p = nullptr;
  SLA_CUDA(c, cudaMalloc(&p, sizeof(__nv_bfloat16) * (size_t)A->nnz));
  f64_to_bf16_kernel<<<gsb(A->nnz), 256, 0, c->stream>>>(A->val, (__nv_bfloat16*)p, A->nnz);
  SLA_LAUNCH_CHECK(c);
  const_cast<sla_csr*>(A)->val_bf16 = p;
  return SLA_OK;
}

// the single-GPU product: A's column indices address rows of B directly
static sla_status spmm_local(sla_ctx* c, const sla_csr* A, const sla_dense* B, sla_dense* C) {
  if (A->m == 0 || B->cols == 0) return SLA_OK;
  const int m = (int)A->m, k = (int)B->cols;
  const int64_t warps = A->m;
  int64_t blocks = (warps * 32 + SPMM_THREADS - 1) / SPMM_THREADS;
  if (blocks > SLA_NUM_SMS * 32) blocks = SLA_NUM_SMS * 32;
  if (B->dtype == SLA_F64) {
    spmm_f64_kernel<<<(unsigned)blocks, SPMM_THREADS, 0, c->stream>>>(A->row_ptr, A->col, A->val, (const double*)B->d, (double*)C->d, m, k);
  } else {
    // block-structured A and a 128-column B: tensor-core tile path (SLA_SPMM_TC=0 disables, =1 forces)
    const char* tc = getenv("SLA_SPMM_TC");
    if (!(tc && atoi(tc) == 0)) {
      const sla_status ts = sla_spmm_bsr_tc(c, A, B, C, tc && atoi(tc) == 1 ? 0.0 : 0.5);
      if (ts == SLA_OK) return SLA_OK;
      if (ts != SLA_ERR_INVALID) return ts;
    }
    SLA_TRY(ensure_val_bf16(c, A));
    spmm_bf16_kernel<<<(unsigned)blocks, SPMM_THREADS, 0, c->stream>>>(A->row_ptr, A->col, (const __nv_bfloat16*)A->val_bf16,
                                                                      (const __nv_bfloat16*)B->d, (__nv_bfloat16*)C->d, m, k);
  }
  SLA_LAUNCH_CHECK(c);
  return SLA_OK;
}

extern "C" sla_status sla_spmm_dense(sla_ctx* c, const sla_csr* A, const sla_dense* B, sla_dense* C) {
  if (!c || !A || !B || !C) return SLA_ERR_INVALID;
  if (!B->rowmajor || !C->rowmajor) return sla_fail(c, SLA_ERR_INVALID, "## : dense operands must be row-major blocks (sla_dense_create)");
  // a row block of a distributed matrix multiplies the matching ROW SLICE of B: B holds rows [row0, row0 + m) of the n x k operand
  const int64_t b_rows_expected = A->dist ? A->m : A->n;
  if (b_rows_expected != B->rows) {    // matMatCheck | c1 == r2 ... | otherwise = error   SpMatrix.hs:790-797
    snprintf(c->err, sizeof(c->err), "matMat : incompatible matrix sizes((%lld,%lld),(%lld,%lld))", (long long)A->m, (long long)A->n,
             (long long)B->rows, (long long)B->cols);
    return SLA_ERR_SIZE_MISMATCH;
  }
  if (C->rows != A->m || C->cols != B->cols) return sla_fail(c, SLA_ERR_SIZE_MISMATCH, "## : output block has the wrong shape");
  if (B->dtype != C->dtype) return sla_fail(c, SLA_ERR_INVALID, "## : B and C must have the same element type");
  if (!A->dist || c->world <= 1) return spmm_local(c, A, B, C);
  // ---- row-partitioned (NOT YET RUN ON HARDWARE): gather the rows of B the block references, then the local product.
  // Every rank takes part in the gather even when it has no rows.
  const size_t esz = B->dtype == SLA_BF16 ? 2 : 8;
  const size_t need = (size_t)A->n * (size_t)B->cols * esz;
  if (c->bfull_bytes < need) {
    SLA_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->bfull); c->bfull = nullptr; c->bfull_bytes = 0;
    if (cudaMalloc(&c->bfull, need ? need : 1) != cudaSuccess) { cudaGetLastError(); return sla_fail(c, SLA_ERR_ALLOC, "## : cudaMalloc failed for the gathered right operand"); }
    c->bfull_bytes = need;
  }
  SLA_TRY(sla_dist_gather_rows(c, A, B->d, c->bfull, B->cols, B->dtype));
  sla_dense view = *B;
  view.rows = A->n; view.d = (double*)c->bfull;
  return spmm_local(c, A, &view, C);
}

// =================================================================================================================
// Tensor-core tile path (tcgen05 + TMEM) for block-structured A and a 128-column bf16 B.
//
// A is re-blocked once into 16 x 16 bf16 blocks (BSR).  For one block row the product is computed TRANSPOSED,
//     C^T[128 x 16] += Bslab^T[128 x 16] * Ablk^T[16 x 16]        (M = 128 columns of B, N = 16 rows, K = 16)
// so that one tcgen05.mma (kind::f16, M128 N16 K16, fp32 accumulator in TMEM) consumes a whole block: the 16 rows
// of B the block touches are staged ONCE in shared memory and reused by all 16 rows of the block — that reuse, not
// the flops, is what the gather kernel lacks (it re-reads a B row per stored entry and is L2-bound).
//   operand A of the MMA = Bslab^T, MN-major, no-swizzle canonical layout (8 x 16-byte core matrices,
//                          SBO = 128 B between cores along M, LBO = 2048 B between the two K halves)
//   operand B of the MMA = Ablk^T,  K-major,  no-swizzle canonical layout (SBO = 256 B, LBO = 128 B); the BSR
//                          values are stored in that order at conversion time, so staging a block is a 512 B copy.
// One CTA of 128 threads walks block rows; thread t owns TMEM lane t = column t of C for the epilogue
// (tcgen05.ld 32x32b.x16), fp32 -> bf16, stores 64 B per warp per row.
#define BSR_B 16

__global__ void bsr_count_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int m, int nbr, int* __restrict__ cnt) {
  const int br = blockIdx.x * blockDim.x + threadIdx.x;
  if (br >= nbr) return;
  int cur[BSR_B], end[BSR_B];
  for (int i = 0; i < BSR_B; ++i) {
    const int r = br * BSR_B + i;
    cur[i] = r < m ? row_ptr[r] : 0;
    end[i] = r < m ? row_ptr[r + 1] : 0;
  }
  int n = 0;
  for (;;) {
    int mn = 0x7fffffff;
    for (int i = 0; i < BSR_B; ++i) if (cur[i] < end[i]) mn = min(mn, col[cur[i]] >> 4);
    if (mn == 0x7fffffff) break;
    ++n;
    for (int i = 0; i < BSR_B; ++i) while (cur[i] < end[i] && (col[cur[i]] >> 4) == mn) ++cur[i];
  }
  cnt[br] = n;
}

__global__ void bsr_fill_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, const double* __restrict__ val, int m,
                                int nbr, const int* __restrict__ brow_ptr, int* __restrict__ bcol, __nv_bfloat16* __restrict__ bval) {
  const int br = blockIdx.x * blockDim.x + threadIdx.x;
  if (br >= nbr) return;
  int cur[BSR_B], end[BSR_B];
  for (int i = 0; i < BSR_B; ++i) {
    const int r = br * BSR_B + i;
    cur[i] = r < m ? row_ptr[r] : 0;
    end[i] = r < m ? row_ptr[r + 1] : 0;
  }
  int b = brow_ptr[br];
  for (;;) {
    int mn = 0x7fffffff;
    for (int i = 0; i < BSR_B; ++i) if (cur[i] < end[i]) mn = min(mn, col[cur[i]] >> 4);
    if (mn == 0x7fffffff) break;
    bcol[b] = mn;
    __nv_bfloat16* blk = bval + (size_t)b * 256;
    for (int i = 0; i < BSR_B; ++i)
      while (cur[i] < end[i] && (col[cur[i]] >> 4) == mn) {
        const int k = col[cur[i]] & 15;
        // canonical K-major core-matrix order of the MMA's B operand: element (n = i, k)
        blk[(i >> 3) * 128 + (k >> 3) * 64 + (i & 7) * 8 + (k & 7)] = __double2bfloat16(val[cur[i]]);
        ++cur[i];
      }
    ++b;
  }
}

__device__ __forceinline__ uint32_t spmm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SmemDescriptor: start [0,14) >>4 | LBO [16,30) >>4 | SBO [32,46) >>4 | version = 1 at [46,48) | layout_type = 0 (no swizzle)
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) |
         (1ULL << 46);
}

#ifndef TC_STAGE
#define TC_STAGE 2
#endif
#define TC_SBO 144              // bytes between core matrices adjacent along M (128 + 16 of padding)
#define TC_LBO (16 * TC_SBO)    // bytes between the two K halves
//      // blocks staged (and multiplied) per synchronisation

#ifndef TC_MINBLOCKS
#define TC_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(128, TC_MINBLOCKS)
spmm_bsr_tc_kernel(const int* __restrict__ brow_ptr, const int* __restrict__ bcol, const __nv_bfloat16* __restrict__ bval,
                   const __nv_bfloat16* __restrict__ B, __nv_bfloat16* __restrict__ C, int m, int nbr) {
  __shared__ __align__(128) __nv_bfloat16 sA[TC_STAGE][256];          // 16 x 16 blocks of A (operand B of the MMA)
  // 16 rows of B each (operand A), canonical MN-major.  The 32 core matrices (8 k-rows x 16 B) are spaced 144 B apart
  // along M instead of 128 B: the staging stores of 8 lanes (same k, consecutive 8-column groups) then fall into
  // distinct banks (ncu: with SBO = 128 B they were 8-way conflicts and L1TEX sat at 67 %).
  __shared__ __align__(128) __nv_bfloat16 sB[TC_STAGE][2 * TC_LBO / 2];
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(spmm_smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(spmm_smem_u32(&mma_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // kind::f16 instruction descriptor: D = F32, A = B = BF16, A MN-major, B K-major, N = 16, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (0u << 16) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t phase = 0;

  for (int br = blockIdx.x; br < nbr; br += gridDim.x) {
    const int s = brow_ptr[br], e = brow_ptr[br + 1];
    for (int b0 = s; b0 < e; b0 += TC_STAGE) {
      const int nb = e - b0 < TC_STAGE ? e - b0 : TC_STAGE;
      // stage nb blocks of A (already in canonical order) and the 16 B rows each of them multiplies; all the
      // global loads of the stage are issued before the first store so that they overlap
      uint4 av[TC_STAGE], bv[TC_STAGE][2];
#pragma unroll
      for (int j = 0; j < TC_STAGE; ++j) {
        if (j < nb) {
          if (tid < 32) av[j] = reinterpret_cast<const uint4*>(bval + (size_t)(b0 + j) * 256)[tid];
          const size_t brow0 = (size_t)bcol[b0 + j] * 16;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int q = tid + 128 * h, k = q >> 4, mc = q & 15;
            bv[j][h] = *reinterpret_cast<const uint4*>(B + (brow0 + k) * 128 + mc * 8);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < TC_STAGE; ++j) {
        if (j < nb) {
          if (tid < 32) reinterpret_cast<uint4*>(sA[j])[tid] = av[j];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int q = tid + 128 * h, k = q >> 4, mc = q & 15;
            *reinterpret_cast<uint4*>(sB[j] + mc * (TC_SBO / 2) + (k >> 3) * (TC_LBO / 2) + (k & 7) * 8) = bv[j][h];
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int j = 0; j < nb; ++j) {
          const uint64_t adesc = umma_desc(spmm_smem_u32(sB[j]), TC_LBO, TC_SBO);
          const uint64_t bdesc = umma_desc(spmm_smem_u32(sA[j]), 128, 256);
          const uint32_t acc = (b0 + j) > s ? 1u : 0u;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
              ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(spmm_smem_u32(&mma_bar)) : "memory");
      }
      // wait until the MMAs have consumed the staged operands / produced the accumulator
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "SPMM_WAIT:\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
          "@p bra SPMM_DONE;\n\t"
          "bra SPMM_WAIT;\n\t"
          "SPMM_DONE:\n\t}" ::"r"(spmm_smem_u32(&mma_bar)), "r"(phase) : "memory");
      phase ^= 1u;
    }
    // epilogue: thread t = TMEM lane t = column t of C; 16 accumulator columns = the 16 rows of this block row
    uint32_t r[16];
    if (e > s) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = 0u;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int row = br * 16 + i;
      if (row < m) C[(size_t)row * 128 + tid] = __float2bfloat16_rn(__uint_as_float(r[i]));
    }
    // every lane has drained its accumulator before the next block row's first MMA overwrites it
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }

  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

// =================================================================================================================
// Pipelined tile path (the default): the same product as spmm_bsr_tc_kernel, fed by TMA and warp-specialised.
//
//   warp 0      TMA producer   one lane issues, per 16 x 16 block of A, ONE cp.async.bulk.tensor (SASS UTMALDG) for the 16 x 128
//                              slab of B it multiplies — B is described to the TMA unit as a 3-D tensor (64 columns, rows, 2 column
//                              halves) with SWIZZLE_128B, so the 4 KB land directly in the MN-major 128-byte-swizzled canonical
//                              layout tcgen05.mma reads (two 64-column groups 2 KB apart = LBO, 8-row groups 1 KB apart = SBO) —
//                              and one 512-byte bulk copy (UBLKCP) for the block of A, into a ring of TC_RING stages guarded by
//                              full / empty mbarriers; the 32 lanes fetch the block-column indices of the next 32 blocks at once
//   warp 1      MMA issuer     waits for a stage, issues tcgen05.mma (M128 N16 K16, bf16 -> fp32 in TMEM) and commits the stage's
//                              `empty` barrier; the accumulator is double-buffered in TMEM (2 x 16 columns): after the last block
//                              of a block row it commits `tmem_full` and starts the next row in the other half
//   warps 2-5   epilogue       tcgen05.ld their TMEM lane quarter, release the accumulator, convert to bf16 into a double-buffered
//                              16 x 128 shared-memory tile and hand it to the TMA unit (cp.async.bulk.tensor store, SASS UTMASTG)
// so that loads, MMAs and stores of different block rows overlap inside one CTA; CTAs own CONTIGUOUS ranges of block rows
// (the producer and the issuer walk brow_ptr / bcol as plain sequential streams).  Every wait is bounded: a protocol error
// traps instead of hanging the GPU.
#include <cuda.h>

#define TC_RING 8
#define TCP_THREADS 192
#define TCP_SLAB_BYTES 4096
#define TCP_ABLK_BYTES 512
#define TCP_CT_BYTES 4096
#define TCP_SMEM_BYTES (TC_RING * (TCP_SLAB_BYTES + TCP_ABLK_BYTES) + 2 * TCP_CT_BYTES + 1024)
#define TCP_SPIN_LIMIT 4000000000LL     // ~2 s of SM clocks

__device__ __forceinline__ void tcp_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(spmm_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tcp_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = spmm_smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > TCP_SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void tcp_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(spmm_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcp_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(spmm_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // as umma_desc, with layout_type = SWIZZLE_128B (2) at [61,64); the operand is 1024-byte aligned, so base_offset = 0
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) |
         (1ULL << 46) | (2ULL << 61);
}

__global__ void __launch_bounds__(TCP_THREADS, 1)
spmm_bsr_tc_pipe_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC, const int* __restrict__ brow_ptr,
                        const int* __restrict__ bcol, const __nv_bfloat16* __restrict__ bval, int nbr, int rows_per_cta) {
  extern __shared__ unsigned char tcp_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tcp_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* slab = base;                                           // TC_RING x 4 KB, each 1024-byte aligned (swizzle atoms)
  unsigned char* ablk = base + TC_RING * TCP_SLAB_BYTES;                // TC_RING x 512 B
  unsigned char* ctile = ablk + TC_RING * TCP_ABLK_BYTES;               // 2 x 4 KB
  __shared__ __align__(8) uint64_t full_bar[TC_RING], empty_bar[TC_RING], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(nbr, r0 + rows_per_cta);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(spmm_smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < TC_RING; ++s) { tcp_mbar_init(&full_bar[s], 1); tcp_mbar_init(&empty_bar[s], 1); }
    tcp_mbar_init(&tfull_bar[0], 1); tcp_mbar_init(&tfull_bar[1], 1);
    tcp_mbar_init(&tempty_bar[0], 128); tcp_mbar_init(&tempty_bar[1], 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (r0 < r1) {
    if (warp == 0) {
      // ===== TMA producer =====
      const int b_begin = brow_ptr[r0], b_end = brow_ptr[r1];
      uint32_t stage = 0, phase = 0;
      for (int b0 = b_begin; b0 < b_end; b0 += 32) {
        const int mine = b0 + lane < b_end ? bcol[b0 + lane] : 0;
        const int cnt = min(32, b_end - b0);
        for (int j = 0; j < cnt; ++j) {
          const int bc = __shfl_sync(0xffffffffu, mine, j);
          if (lane == 0) {
            tcp_mbar_wait(&empty_bar[stage], phase ^ 1u);
            tcp_mbar_expect_tx(&full_bar[stage], TCP_SLAB_BYTES + TCP_ABLK_BYTES);
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(spmm_smem_u32(slab + stage * TCP_SLAB_BYTES)), "l"(&tmB), "r"(spmm_smem_u32(&full_bar[stage])),
                           "r"(0), "r"(bc * 16), "r"(0) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(spmm_smem_u32(ablk + stage * TCP_ABLK_BYTES)), "l"(bval + (size_t)(b0 + j) * 256), "r"(TCP_ABLK_BYTES),
                           "r"(spmm_smem_u32(&full_bar[stage])) : "memory");
          }
          if (++stage == TC_RING) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      // ===== MMA issuer =====
      // kind::f16 instruction descriptor: D = F32, A = B = BF16, A MN-major, B K-major, N = 16, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (0u << 16) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      for (int rb = r0; rb < r1; rb += 32) {
        const int rr = rb + lane;
        const int my_s = rr < r1 ? brow_ptr[rr] : 0, my_e = rr < r1 ? brow_ptr[rr + 1] : 0;
        const int cnt = min(32, r1 - rb);
        for (int q = 0; q < cnt; ++q) {
          const int nb = __shfl_sync(0xffffffffu, my_e, q) - __shfl_sync(0xffffffffu, my_s, q);
          if (nb == 0) continue;
          if (lane == 0) tcp_mbar_wait(&tempty_bar[as], aphase ^ 1u);        // the epilogue has drained this half of the accumulator
          __syncwarp();
          for (int j = 0; j < nb; ++j) {
            if (lane == 0) {
              tcp_mbar_wait(&full_bar[stage], phase);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint64_t adesc = umma_desc_sw128(spmm_smem_u32(slab + stage * TCP_SLAB_BYTES), 2048, 1024);
              const uint64_t bdesc = umma_desc(spmm_smem_u32(ablk + stage * TCP_ABLK_BYTES), 128, 256);
              const uint32_t acc = j > 0 ? 1u : 0u;
              asm volatile(
                  "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                  ::"r"(tmem + as * 16u), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
              // the stage may be refilled once this MMA (and everything before it) has read its operands
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(spmm_smem_u32(&empty_bar[stage])) : "memory");
              if (j + 1 == nb)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(spmm_smem_u32(&tfull_bar[as])) : "memory");
            }
            __syncwarp();
            if (++stage == TC_RING) { stage = 0; phase ^= 1u; }
          }
          as ^= 1u;
          if (as == 0) aphase ^= 1u;
        }
      }
    } else {
      // ===== epilogue: 128 threads, thread = column of C =====
      const int q4 = warp & 3;                                   // the TMEM lane quarter this warp may read
      const int ccol = q4 * 32 + lane;
      const bool leader = tid == 64;
      uint32_t as = 0, aphase = 0;
      int it = 0;
      for (int br = r0; br < r1; ++br, ++it) {
        const int nb = brow_ptr[br + 1] - brow_ptr[br];
        uint32_t r[16];
        if (nb > 0) {
          tcp_mbar_wait(&tfull_bar[as], aphase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + as * 16u;
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(taddr) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          tcp_mbar_arrive(&tempty_bar[as]);                      // this half of the accumulator may be overwritten
          as ^= 1u;
          if (as == 0) aphase ^= 1u;
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = 0u;
        }
        __nv_bfloat16* ct = reinterpret_cast<__nv_bfloat16*>(ctile + (it & 1) * TCP_CT_BYTES);
        // the TMA store that read this tile two rows ago must be done with it
        if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) ct[i * 128 + ccol] = __float2bfloat16_rn(__uint_as_float(r[i]));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (leader) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(&tmC), "r"(spmm_smem_u32(ct)), "r"(0), "r"(br * 16) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

typedef CUresult (*sla_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static sla_encode_tiled_fn get_encode_tiled() {
  static sla_encode_tiled_fn fn = nullptr;
  static int tried = 0;
  if (!tried) {
    tried = 1;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (sla_encode_tiled_fn)p;
    cudaGetLastError();
  }
  return fn;
}

// returns SLA_ERR_INVALID when the pipelined kernel cannot be set up (the caller falls back to spmm_bsr_tc_kernel)
static sla_status spmm_bsr_tc_pipe(sla_ctx* c, const sla_csr* A, const sla_dense* B, sla_dense* C) {
  sla_encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return SLA_ERR_INVALID;
  if (((uintptr_t)B->d & 15u) || ((uintptr_t)C->d & 15u)) return SLA_ERR_INVALID;
  CUtensorMap tmB, tmC;
  {
    // B (rows x 128 bf16, row-major) as (64 columns, rows, 2 column halves): one box = a 16-row slab, half by half
    const cuuint64_t dims[3] = {64, (cuuint64_t)B->rows, 2};
    const cuuint64_t strides[2] = {256, 128};                   // bytes: next row, next column half
    const cuuint32_t box[3] = {64, 16, 2}, es[3] = {1, 1, 1};
    if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)B->d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return SLA_ERR_INVALID;
  }
  {
    const cuuint64_t dims[2] = {128, (cuuint64_t)C->rows};
    const cuuint64_t strides[1] = {256};
    const cuuint32_t box[2] = {128, 16}, es[2] = {1, 1};
    if (enc(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)C->d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return SLA_ERR_INVALID;
  }
  static char attr_set[64] = {0};
  if (!attr_set[c->device & 63]) {
    if (cudaFuncSetAttribute(spmm_bsr_tc_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_BYTES) != cudaSuccess) { cudaGetLastError(); return SLA_ERR_INVALID; }
    attr_set[c->device & 63] = 1;
  }
  int per_sm = 4;                     // measured on cfg 5 K16: 2 -> 1.84 ms, 3 -> 1.362, 4 -> 1.356, 6 -> 1.52 (profiles/r02_spmm_pipe_ab.txt)
  if (const char* e = getenv("SLA_SPMM_TC_CTAS")) per_sm = atoi(e) > 0 ? atoi(e) : per_sm;
  int grid = SLA_NUM_SMS * per_sm;
  if (grid > A->bsr_nbr) grid = A->bsr_nbr;
  if (grid < 1) grid = 1;
  const int rows_per_cta = (A->bsr_nbr + grid - 1) / grid;
  grid = (A->bsr_nbr + rows_per_cta - 1) / rows_per_cta;
  spmm_bsr_tc_pipe_kernel<<<grid, TCP_THREADS, TCP_SMEM_BYTES, c->stream>>>(tmB, tmC, A->bsr_row_ptr, A->bsr_col, (const __nv_bfloat16*)A->bsr_val,
                                                                           A->bsr_nbr, rows_per_cta);
  SLA_LAUNCH_CHECK(c);
  return SLA_OK;
}

// Builds the BSR copy of A on first use; returns the fill ratio nnz / (16*16*nblocks) through *fill.
static sla_status ensure_bsr(sla_ctx* c, const sla_csr* A, double* fill) {
  sla_csr* Am = const_cast<sla_csr*>(A);
  if (!Am->bsr_ready) {
    const int m = (int)A->m, nbr = (m + BSR_B - 1) / BSR_B;
    int *cnt = nullptr; void* tmp = nullptr;
    SLA_CUDA(c, cudaMalloc(&cnt, sizeof(int) * (size_t)(nbr + 1)));
    SLA_CUDA(c, cudaMalloc(&Am->bsr_row_ptr, sizeof(int) * (size_t)(nbr + 1)));
    SLA_CUDA(c, cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)(nbr + 1), c->stream));
    bsr_count_kernel<<<(nbr + 127) / 128, 128, 0, c->stream>>>(A->row_ptr, A->col, m, nbr, cnt);
    SLA_LAUNCH_CHECK(c);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, Am->bsr_row_ptr, nbr + 1, c->stream);
    SLA_CUDA(c, cudaMalloc(&tmp, tb ? tb : 1));
    cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, Am->bsr_row_ptr, nbr + 1, c->stream);
    c->launches += 2;
    int nblk = 0;
    SLA_CUDA(c, cudaMemcpyAsync(&nblk, Am->bsr_row_ptr + nbr, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SLA_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(cnt); cudaFree(tmp);
    Am->bsr_nblk = nblk; Am->bsr_nbr = nbr;
    SLA_CUDA(c, cudaMalloc(&Am->bsr_col, sizeof(int) * (size_t)(nblk > 0 ? nblk : 1)));
    SLA_CUDA(c, cudaMalloc(&Am->bsr_val, 512 * (size_t)(nblk > 0 ? nblk : 1)));
    SLA_CUDA(c, cudaMemsetAsync(Am->bsr_val, 0, 512 * (size_t)(nblk > 0 ? nblk : 1), c->stream));
    bsr_fill_kernel<<<(nbr + 127) / 128, 128, 0, c->stream>>>(A->row_ptr, A->col, A->val, m, nbr, Am->bsr_row_ptr, Am->bsr_col,
                                                              (__nv_bfloat16*)Am->bsr_val);
    SLA_LAUNCH_CHECK(c);
    Am->bsr_ready = 1;
  }
  *fill = A->bsr_nblk > 0 ? (double)A->nnz / (256.0 * (double)A->bsr_nblk) : 0.0;
  return SLA_OK;
}

void sla_csr_free_bsr(sla_csr* A) {
  cudaFree(A->bsr_row_ptr); cudaFree(A->bsr_col); cudaFree(A->bsr_val);
  A->bsr_row_ptr = nullptr; A->bsr_col = nullptr; A->bsr_val = nullptr; A->bsr_ready = 0;
}

// C = A ## B on the tensor cores; returns SLA_ERR_INVALID (without touching C) when the path does not apply.
sla_status sla_spmm_bsr_tc(sla_ctx* c, const sla_csr* A, const sla_dense* B, sla_dense* C, double min_fill) {
  if (B->dtype != SLA_BF16 || B->cols != 128 || A->n % 16 != 0 || A->m == 0) return SLA_ERR_INVALID;   // whole 16-row slabs of B only
  double fill = 0;
  SLA_TRY(ensure_bsr(c, A, &fill));
  if (fill < min_fill) return SLA_ERR_INVALID;
  // the TMA-fed warp-specialised pipeline is the default; SLA_SPMM_TC_PIPE=0 selects the one-stage kernel it replaced
  const char* pipe = getenv("SLA_SPMM_TC_PIPE");
  if (!(pipe && atoi(pipe) == 0)) {
    const sla_status ps = spmm_bsr_tc_pipe(c, A, B, C);
    if (ps != SLA_ERR_INVALID) return ps;
  }
  int grid = A->bsr_nbr < SLA_NUM_SMS * 16 ? A->bsr_nbr : SLA_NUM_SMS * 16;
  if (grid < 1) grid = 1;
  spmm_bsr_tc_kernel<<<grid, 128, 0, c->stream>>>(A->bsr_row_ptr, A->bsr_col, (const __nv_bfloat16*)A->bsr_val,
                                                 (const __nv_bfloat16*)B->d, (__nv_bfloat16*)C->d, (int)A->m, A->bsr_nbr);
  SLA_LAUNCH_CHECK(c);
  return SLA_OK;
}

// synthetic dense block: entry (r, c) = sla_synth_vec(seed, r * cols + c), rounded to the block's element type
#include "../../include/sla_synth.h"
__global__ void dense_synth_kernel(uint64_t seed, int64_t n, double* __restrict__ o64, __nv_bfloat16* __restrict__ o16) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = sla_synth_vec(seed, i);
    if (o64) o64[i] = v; else o16[i] = __double2bfloat16(v);
  }
}

extern "C" sla_status sla_dense_generate(sla_ctx* c, int64_t rows, int64_t cols, uint64_t seed, int dtype, sla_dense** out) {
  SLA_TRY(sla_dense_create(c, rows, cols, dtype, out));
  const int64_t n = rows * cols;
  if (n > 0) {
    dense_synth_kernel<<<gsb(n), 256, 0, c->stream>>>(seed, n, dtype == SLA_F64 ? (double*)(*out)->d : nullptr,
                                                     dtype == SLA_BF16 ? (__nv_bfloat16*)(*out)->d : nullptr);
    SLA_LAUNCH_CHECK(c);
  }
  return SLA_OK;
}
