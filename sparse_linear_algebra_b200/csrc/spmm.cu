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
#include <new>

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
  if (cudaMalloc((void**)&d->d, bytes ? bytes : 16) != cudaSuccess) { cudaGetLastError(); delete d; return sla_fail(c, SLA_ERR_ALLOC, "cudaMalloc failed for a dense block"); }
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

extern "C" sla_status sla_spmm_dense(sla_ctx* c, const sla_csr* A, const sla_dense* B, sla_dense* C) {
  if (!c || !A || !B || !C) return SLA_ERR_INVALID;
  if (A->dist) return sla_fail(c, SLA_ERR_INVALID, "## : row-partitioned operands are not supported yet");
  if (!B->rowmajor || !C->rowmajor) return sla_fail(c, SLA_ERR_INVALID, "## : dense operands must be row-major blocks (sla_dense_create)");
  if (A->n != B->rows) {    // matMatCheck | c1 == r2 ... | otherwise = error   SpMatrix.hs:790-797
    snprintf(c->err, sizeof(c->err), "matMat : incompatible matrix sizes((%lld,%lld),(%lld,%lld))", (long long)A->m, (long long)A->n,
             (long long)B->rows, (long long)B->cols);
    return SLA_ERR_SIZE_MISMATCH;
  }
  if (C->rows != A->m || C->cols != B->cols) return sla_fail(c, SLA_ERR_SIZE_MISMATCH, "## : output block has the wrong shape");
  if (B->dtype != C->dtype) return sla_fail(c, SLA_ERR_INVALID, "## : B and C must have the same element type");
  if (A->m == 0 || B->cols == 0) return SLA_OK;
  const int m = (int)A->m, k = (int)B->cols;
  const int64_t warps = A->m;
  int64_t blocks = (warps * 32 + SPMM_THREADS - 1) / SPMM_THREADS;
  if (blocks > SLA_NUM_SMS * 32) blocks = SLA_NUM_SMS * 32;
  if (B->dtype == SLA_F64) {
    spmm_f64_kernel<<<(unsigned)blocks, SPMM_THREADS, 0, c->stream>>>(A->row_ptr, A->col, A->val, (const double*)B->d, (double*)C->d, m, k);
  } else {
    SLA_TRY(ensure_val_bf16(c, A));
    spmm_bf16_kernel<<<(unsigned)blocks, SPMM_THREADS, 0, c->stream>>>(A->row_ptr, A->col, (const __nv_bfloat16*)A->val_bf16,
                                                                      (const __nv_bfloat16*)B->d, (__nv_bfloat16*)C->d, m, k);
  }
  SLA_LAUNCH_CHECK(c);
  return SLA_OK;
}
