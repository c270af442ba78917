// lu.cu — ilu0Pre of the reference (src/Numeric/LinearAlgebra/Sparse.hs:696-706) on the device.
//
// The reference's "ILU(0)" is NOT the incomplete recurrence of the literature: it runs the COMPLETE Doolittle factorisation
// `lu` (Sparse.hs:489-538, O(n^3), element at a time) and afterwards drops the entries of L and U that sit where aa stores
// nothing (sparsifyLU).  A drop-in has to return those numbers, so this file evaluates the same recurrences in the same
// order on a dense n x n work area —
//     luInit   U[0][:] = row 0 of aa ; L = eye n ; L[i][0] = recip u00 * a_i0 for the stored a_i0       (`./` = times the reciprocal)
//     step i   U[i][j] = a_ij - sum_{k<i, l_ik stored, ascending} l_ik u_kj      (j = i .. n-1), kept when isNz
//              L[k][i] = (a_ki - sum_{q<i, l_kq stored, ascending} l_kq u_qi) / u_ii   (k = i+1 .. n-1), kept when isNz;
//              a nearZero u_ii with rows left below it raises NeedsPivoting (solveForLij)
// with __dmul_rn / __dadd_rn / __ddiv_rn (no FMA), hence bit-identical to the reference's evaluation — and it is meant for the
// sizes the reference's own algorithm can handle (n <= SLA_LU_MAX_N).  The row sums of a step run in parallel over j (or k);
// the sum inside each is sequential, as the fold it restates (contractSub, SpMatrix.hs:857-864).
#include "common.cuh"

#include <cub/device/device_scan.cuh>
#include <math.h>

#define SLA_LU_MAX_N 4096
#define LU_THREADS 128

namespace {

struct LuWork {
  double *A, *L, *LT, *U;          // dense n x n, row-major (LT = L transposed: column sums read it with unit stride)
  unsigned char *Ap, *Lf, *Uf;     // stored-entry flags of aa, L, U
  int* err;                        // pivot index whose u_jj is nearZero, -1 otherwise
};

__global__ void lu_scatter_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, const double* __restrict__ val, int n, LuWork w) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  for (int q = row_ptr[r]; q < row_ptr[r + 1]; ++q) {
    w.A[(size_t)r * n + col[q]] = val[q];
    w.Ap[(size_t)r * n + col[q]] = 1;
  }
  w.L[(size_t)r * n + r] = 1.0; w.LT[(size_t)r * n + r] = 1.0; w.Lf[(size_t)r * n + r] = 1;      // eye n
}

// luInit   Sparse.hs:500-507
__global__ void lu_init_kernel(int n, LuWork w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (w.Ap[i]) { w.U[i] = w.A[i]; w.Uf[i] = 1; }                       // extractRow aa 0
  const double u00 = w.Ap[0] ? w.A[0] : 0.0;
  if (fabs(u00) <= 1e-12) { if (i == 0) *w.err = 0; return; }
  if (i >= 1 && w.Ap[(size_t)i * n]) {
    const double v = __dmul_rn(__ddiv_rn(1.0, u00), w.A[(size_t)i * n]);   // extractSubCol aa 0 (1, n-1) ./ u00
    w.L[(size_t)i * n] = v; w.LT[i] = v; w.Lf[(size_t)i * n] = 1;
  }
}

// uUpd: row ix of U   Sparse.hs:518-523
__global__ void lu_urow_kernel(int n, int ix, LuWork w) {
  if (*w.err >= 0) return;
  const int j = ix + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double acc = 0.0;
  for (int k = 0; k < ix; ++k)
    if (w.Lf[(size_t)ix * n + k]) acc = __dadd_rn(acc, __dmul_rn(w.L[(size_t)ix * n + k], w.U[(size_t)k * n + j]));
  const double v = __dsub_rn(w.A[(size_t)ix * n + j], acc);
  if (fabs(v) > 1e-12) { w.U[(size_t)ix * n + j] = v; w.Uf[(size_t)ix * n + j] = 1; }
}

// lUpd: column ix of L   Sparse.hs:524-535
__global__ void lu_lcol_kernel(int n, int ix, LuWork w) {
  if (*w.err >= 0) return;
  const int k = ix + 1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double ujj = w.U[(size_t)ix * n + ix];                          // 0 when not stored
  if (fabs(ujj) <= 1e-12) { atomicCAS(w.err, -1, ix); return; }
  double acc = 0.0;
  for (int q = 0; q < ix; ++q)
    if (w.Lf[(size_t)k * n + q]) acc = __dadd_rn(acc, __dmul_rn(w.LT[(size_t)q * n + k], w.U[(size_t)q * n + ix]));
  const double v = __ddiv_rn(__dsub_rn(w.A[(size_t)k * n + ix], acc), ujj);
  if (fabs(v) > 1e-12) { w.L[(size_t)k * n + ix] = v; w.LT[(size_t)ix * n + k] = v; w.Lf[(size_t)k * n + ix] = 1; }
}

// sparsifyLU: the entries of L (U) that are stored AND sit on a stored position of aa   Sparse.hs:702-705
__global__ void lu_count_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int n, LuWork w, int* cl, int* cu) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n) return;
  int a = 0, b = 0;
  if (r < n)
    for (int q = row_ptr[r]; q < row_ptr[r + 1]; ++q) {
      a += w.Lf[(size_t)r * n + col[q]];
      b += w.Uf[(size_t)r * n + col[q]];
    }
  cl[r] = a; cu[r] = b;
}

__global__ void lu_fill_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int n, LuWork w, const int* __restrict__ pl,
                               int* lcol, double* lval, const int* __restrict__ pu, int* ucol, double* uval) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int a = pl[r], b = pu[r];
  for (int q = row_ptr[r]; q < row_ptr[r + 1]; ++q) {
    const int j = col[q];
    if (w.Lf[(size_t)r * n + j]) { lcol[a] = j; lval[a] = w.L[(size_t)r * n + j]; ++a; }
    if (w.Uf[(size_t)r * n + j]) { ucol[b] = j; uval[b] = w.U[(size_t)r * n + j]; ++b; }
  }
}

struct Buf {
  void* p = nullptr;
  ~Buf() { if (p) cudaFree(p); }
  cudaError_t zeros(size_t bytes, cudaStream_t s) {
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    return e != cudaSuccess ? e : cudaMemsetAsync(p, 0, bytes ? bytes : 1, s);
  }
};

}  // namespace

extern "C" sla_status sla_ilu0_pre(sla_ctx* c, const sla_csr* A, sla_csr** Lout, sla_csr** Uout) {
  if (!c || !A || !Lout || !Uout) return SLA_ERR_INVALID;
  *Lout = *Uout = nullptr;
  if (A->dist) return sla_fail(c, SLA_ERR_INVALID, "ilu0Pre: not available on a row-partitioned matrix");
  if (A->m != A->n || A->m < 1) return sla_fail(c, SLA_ERR_SIZE_MISMATCH, "ilu0Pre: the matrix must be square");
  if (A->m > SLA_LU_MAX_N)
    return sla_fail(c, SLA_ERR_INVALID, "ilu0Pre: the reference's algorithm is a complete O(n^3) LU followed by a mask; supported up to n = 4096");
  const int n = (int)A->m;
  const size_t nn = (size_t)n * n;
  Buf bA, bL, bLT, bU, bAp, bLf, bUf, berr, bcnt, bptr, btmp;
  LuWork w;
  SLA_CUDA(c, bA.zeros(nn * 8, c->stream)); SLA_CUDA(c, bL.zeros(nn * 8, c->stream)); SLA_CUDA(c, bLT.zeros(nn * 8, c->stream));
  SLA_CUDA(c, bU.zeros(nn * 8, c->stream)); SLA_CUDA(c, bAp.zeros(nn, c->stream)); SLA_CUDA(c, bLf.zeros(nn, c->stream));
  SLA_CUDA(c, bUf.zeros(nn, c->stream)); SLA_CUDA(c, berr.zeros(sizeof(int), c->stream));
  w.A = (double*)bA.p; w.L = (double*)bL.p; w.LT = (double*)bLT.p; w.U = (double*)bU.p;
  w.Ap = (unsigned char*)bAp.p; w.Lf = (unsigned char*)bLf.p; w.Uf = (unsigned char*)bUf.p; w.err = (int*)berr.p;
  SLA_CUDA(c, cudaMemsetAsync(w.err, 0xff, sizeof(int), c->stream));                      // -1
  const unsigned gb = (unsigned)((n + LU_THREADS - 1) / LU_THREADS);
  lu_scatter_kernel<<<gb, LU_THREADS, 0, c->stream>>>(A->row_ptr, A->col, A->val, n, w);
  lu_init_kernel<<<gb, LU_THREADS, 0, c->stream>>>(n, w);
  c->launches += 2;
  for (int ix = 1; ix < n; ++ix) {
    const int nu = n - ix, nl = n - ix - 1;
    lu_urow_kernel<<<(unsigned)((nu + LU_THREADS - 1) / LU_THREADS), LU_THREADS, 0, c->stream>>>(n, ix, w);
    c->launches++;
    if (nl > 0) {
      lu_lcol_kernel<<<(unsigned)((nl + LU_THREADS - 1) / LU_THREADS), LU_THREADS, 0, c->stream>>>(n, ix, w);
      c->launches++;
    }
  }
  SLA_CUDA(c, cudaGetLastError());
  int h_err = -1;
  SLA_CUDA(c, cudaMemcpyAsync(&h_err, w.err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  if (h_err >= 0) {                         // NeedsPivoting "solveForLij" ("U" ++ show (j, j))   Sparse.hs:497
    snprintf(c->err, sizeof(c->err), "solveForLij : U(%d,%d) is close to 0. Permute the rows to obtain a nonzero diagonal", h_err, h_err);
    return SLA_ERR_NEEDS_PIVOTING;
  }
  // masked CSR factors: count -> scan -> fill
  SLA_CUDA(c, bcnt.zeros(sizeof(int) * 2 * (size_t)(n + 1), c->stream));
  SLA_CUDA(c, bptr.zeros(sizeof(int) * 2 * (size_t)(n + 1), c->stream));
  int *cl = (int*)bcnt.p, *cu = cl + (n + 1), *pl = (int*)bptr.p, *pu = pl + (n + 1);
  lu_count_kernel<<<(unsigned)((n + 1 + LU_THREADS - 1) / LU_THREADS), LU_THREADS, 0, c->stream>>>(A->row_ptr, A->col, n, w, cl, cu);
  c->launches++;
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, cl, pl, n + 1, c->stream);
  SLA_CUDA(c, btmp.zeros(tb, c->stream));
  cub::DeviceScan::ExclusiveSum(btmp.p, tb, cl, pl, n + 1, c->stream);
  cub::DeviceScan::ExclusiveSum(btmp.p, tb, cu, pu, n + 1, c->stream);
  c->launches += 2;
  int nzl = 0, nzu = 0;
  SLA_CUDA(c, cudaMemcpyAsync(&nzl, pl + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  SLA_CUDA(c, cudaMemcpyAsync(&nzu, pu + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  sla_csr *L = nullptr, *U = nullptr;
  SLA_TRY(sla_csr_alloc(c, n, n, nzl, &L));
  sla_status s = sla_csr_alloc(c, n, n, nzu, &U);
  if (s != SLA_OK) { sla_csr_free(L); return s; }
  cudaMemcpyAsync(L->row_ptr, pl, sizeof(int) * (size_t)(n + 1), cudaMemcpyDeviceToDevice, c->stream);
  cudaMemcpyAsync(U->row_ptr, pu, sizeof(int) * (size_t)(n + 1), cudaMemcpyDeviceToDevice, c->stream);
  lu_fill_kernel<<<gb, LU_THREADS, 0, c->stream>>>(A->row_ptr, A->col, n, w, pl, L->col, L->val, pu, U->col, U->val);
  c->launches++;
  if (cudaGetLastError() != cudaSuccess) s = sla_fail(c, SLA_ERR_CUDA, "ilu0Pre: kernel launch failed");
  if (s == SLA_OK) s = sla_csr_build_plan(c, L);
  if (s == SLA_OK) s = sla_csr_build_plan(c, U);
  if (s == SLA_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) s = sla_fail(c, SLA_ERR_CUDA, "ilu0Pre: CUDA error");
  if (s != SLA_OK) { sla_csr_free(L); sla_csr_free(U); return s; }
  *Lout = L; *Uout = U;
  return SLA_OK;
}
