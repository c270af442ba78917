// matring.cu — the rest of the reference's MatrixRing surface for a DENSE right operand (src/Numeric/LinearAlgebra/Class.hs:195-207,
// instance src/Data/Sparse/SpMatrix.hs:751-773), as compositions of the kernels that already exist:
//   (##^)  a b = matMat_ ABt a b = a ## transpose b      -> dense transpose of the operand, then the (##) kernels
//   (#^#)  a b = transpose a ## b                          -> the cached transpose of a (built once), then the (##) kernels
//   normFrobenius m = sqrt (trace (m ##^ m))               -> per row the left fold of a_ij * a_ij over ascending j (what the
//                                                             diagonal of m ##^ m holds), then the sum over the rows
// plus a column accessor for the Arnoldi basis.
#include "blas1.cuh"

#include <cuda_bf16.h>
#include <math.h>

// out[c][r] = in[r][c] for a row-major rows x cols block (32 x 32 tiles through shared memory, both sides coalesced)
template <class T>
__global__ void dense_transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t rows, int64_t cols) {
  __shared__ T tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int64_t r = r0 + dy, cc = c0 + threadIdx.x;
    if (r < rows && cc < cols) tile[dy][threadIdx.x] = in[r * cols + cc];
  }
  __syncthreads();
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int64_t cc = c0 + dy, r = r0 + threadIdx.x;
    if (r < rows && cc < cols) out[cc * rows + r] = tile[threadIdx.x][dy];
  }
}

// (##^): C (m x k) = A (m x n) ## transpose Bt, Bt a k x n row-major block
extern "C" sla_status sla_spmm_dense_abt(sla_ctx* c, const sla_csr* A, const sla_dense* Bt, sla_dense* C) {
  if (!c || !A || !Bt || !C) return SLA_ERR_INVALID;
  if (A->dist) return sla_fail(c, SLA_ERR_INVALID, "##^ : not available on a row-partitioned matrix");
  if (!Bt->rowmajor || !C->rowmajor) return sla_fail(c, SLA_ERR_INVALID, "##^ : dense operands must be row-major blocks (sla_dense_create)");
  if (Bt->cols != A->n) {                     // matMatCheck on (trDim mm2)   SpMatrix.hs:787-797
    snprintf(c->err, sizeof(c->err), "matMat : incompatible matrix sizes((%lld,%lld),(%lld,%lld))", (long long)A->m, (long long)A->n,
             (long long)Bt->cols, (long long)Bt->rows);
    return SLA_ERR_SIZE_MISMATCH;
  }
  sla_dense* B = nullptr;
  SLA_TRY(sla_dense_create(c, Bt->cols, Bt->rows, Bt->dtype, &B));
  if (Bt->rows > 0 && Bt->cols > 0) {
    const dim3 grid((unsigned)((Bt->cols + 31) / 32), (unsigned)((Bt->rows + 31) / 32)), block(32, 8);
    if (Bt->dtype == SLA_F64) dense_transpose_kernel<double><<<grid, block, 0, c->stream>>>(Bt->d, B->d, Bt->rows, Bt->cols);
    else dense_transpose_kernel<__nv_bfloat16><<<grid, block, 0, c->stream>>>((const __nv_bfloat16*)Bt->d, (__nv_bfloat16*)B->d, Bt->rows, Bt->cols);
    c->launches++;
  }
  sla_status s = cudaGetLastError() == cudaSuccess ? sla_spmm_dense(c, A, B, C) : sla_fail(c, SLA_ERR_CUDA, "##^ : transpose kernel failed");
  sla_dense_free(B);                          // stream-ordered free: after the product above
  return s;
}

// (#^#): C (n x k) = transpose A ## B, B an m x k row-major block.  The transpose is cached in A (as for (<#)).
extern "C" sla_status sla_spmm_dense_atb(sla_ctx* c, const sla_csr* A, const sla_dense* B, sla_dense* C) {
  if (!c || !A || !B || !C) return SLA_ERR_INVALID;
  if (A->dist && !A->T) return sla_fail(c, SLA_ERR_INVALID, "#^# : attach the distributed transpose of the row-partitioned matrix first");
  if (!A->T) {
    sla_csr* t = nullptr;
    SLA_TRY(sla_csr_transpose(c, A, &t));
    const_cast<sla_csr*>(A)->T = t;
  }
  return sla_spmm_dense(c, A->T, B, C);
}

// one thread per row: the left fold of a_ij * a_ij over the row's stored entries (ascending column), then the grid sum
__global__ void __launch_bounds__(EW_THREADS)
frob_rows_kernel(const int* __restrict__ row_ptr, const double* __restrict__ val, int64_t m, double* scal, double* partials,
                 unsigned int* counter, int fin, int dst, sla_p2p_args pa) {
  __shared__ double red[32];
  double acc[1] = {0.0};
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int q = row_ptr[r]; q < row_ptr[r + 1]; ++q) s = __dadd_rn(s, __dmul_rn(val[q], val[q]));
    acc[0] += s;
  }
  block_sum<1>(acc, red);
  grid_reduce_finish<1>(acc, partials, counter, scal, fin, dst, red, pa);
}

extern "C" sla_status sla_csr_norm_frobenius(sla_ctx* c, const sla_csr* A, double* out) {
  if (!c || !A || !out) return SLA_ERR_INVALID;
  int64_t blocks = (A->m + EW_THREADS - 1) / EW_THREADS;
  if (blocks < 1) blocks = 1;
  if (blocks > EW_MAX_BLOCKS) blocks = EW_MAX_BLOCKS;
  const sla_red_plan rp = sla_red_begin(c, FIN_STORE, 1);
  frob_rows_kernel<<<(unsigned)blocks, EW_THREADS, 0, c->stream>>>(A->row_ptr, A->val, A->m, c->scal, c->partials, c->counter, rp.fin, S_TMP0, rp.pa);
  SLA_LAUNCH_CHECK(c);
  SLA_TRY(sla_red_end(c, rp, 1, FIN_STORE, S_TMP0));
  double s = 0;
  SLA_TRY(sla_read_scalars(c, S_TMP0, 1, &s));
  *out = sqrt(s);
  return SLA_OK;
}

// column j of a column-major block (the Arnoldi basis Q) as a new vector: extractCol of the reference's Q (SpMatrix.hs:329-337)
extern "C" sla_status sla_dense_column(sla_ctx* c, const sla_dense* Q, int64_t j, sla_vec** out) {
  if (!c || !Q || !out) return SLA_ERR_INVALID;
  *out = nullptr;
  if (Q->rowmajor || Q->dtype != SLA_F64) return sla_fail(c, SLA_ERR_INVALID, "dense_column: a column-major fp64 block (Krylov basis) is expected");
  if (j < 0 || j >= Q->cols) return sla_fail(c, SLA_ERR_OOB_INDEX, "dense_column: column index out of bounds");
  SLA_TRY(sla_vec_create(c, Q->rows, out));
  SLA_CUDA(c, cudaMemcpyAsync((*out)->d, Q->d + j * Q->ld, sizeof(double) * (size_t)Q->rows, cudaMemcpyDeviceToDevice, c->stream));
  return SLA_OK;
}
