/* abi_smoke.c — the drop-in boundary exercised from plain C (no Python, no C++): builds the reference's 2x2
 * fixture aa0 = [[1,2],[3,4]] (test/LibSpec.hs:1171-1183) from COO triples, checks (#>), (<#), (<.>), transpose,
 * linSolve0 BICGSTAB_ (x0 = 0.1, ||x - xhat|| <= 1e-12 as LibSpec.hs:286-300) and the error codes.
 * Build:  gcc -std=c99 -Iinclude tests/c/abi_smoke.c -Lsparse_linear_algebra_b200 -lsla_b200 -lm
 * Exit code 0 = all checks passed, 77 = no CUDA device (skipped), anything else = failure. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "sla_b200.h"

#define CHECK(cond)                                                                          \
  do {                                                                                       \
    if (!(cond)) { fprintf(stderr, "FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, sla_last_error(ctx)); return 1; } \
  } while (0)

int main(void) {
  sla_ctx* ctx = NULL;
  sla_status st = sla_init(0, &ctx);
  if (st == SLA_ERR_CUDA) { fprintf(stderr, "skip: %s\n", sla_last_error(NULL)); return 77; }
  CHECK(st == SLA_OK);

  /* fromListSM (2,2), with a duplicate that must be overwritten (last write wins, SpMatrix.hs:218-224) */
  const int64_t ci[] = {0, 1, 0, 1, 0}, cj[] = {0, 0, 1, 1, 1};
  const double cv[] = {1.0, 3.0, 9.0, 4.0, 2.0};
  sla_csr* A = NULL;
  CHECK(sla_csr_from_coo(ctx, 2, 2, 5, ci, cj, cv, &A) == SLA_OK);
  int64_t m, n, nnz;
  CHECK(sla_csr_dims(A, &m, &n, &nnz) == SLA_OK && m == 2 && n == 2 && nnz == 4);
  int32_t rp[3], col[4]; double val[4];
  CHECK(sla_csr_to_host(ctx, A, rp, col, val) == SLA_OK);
  CHECK(rp[0] == 0 && rp[1] == 2 && rp[2] == 4 && col[1] == 1 && val[1] == 2.0 && val[2] == 3.0);

  const double x0true[] = {2.0, 3.0}, b0[] = {8.0, 18.0};
  sla_vec *x = NULL, *y = NULL, *b = NULL, *x0 = NULL, *sol = NULL;
  CHECK(sla_vec_from_host(ctx, 2, x0true, &x) == SLA_OK);
  CHECK(sla_vec_create(ctx, 2, &y) == SLA_OK);
  double out[2], d;
  CHECK(sla_spmv(ctx, A, x, y) == SLA_OK && sla_vec_to_host(ctx, y, out) == SLA_OK);          /* aa0 #> x0true = [8,18] */
  CHECK(out[0] == 8.0 && out[1] == 18.0);
  CHECK(sla_spmvT(ctx, A, x, y) == SLA_OK && sla_vec_to_host(ctx, y, out) == SLA_OK);         /* x0true <# aa0 = [11,16] */
  CHECK(out[0] == 11.0 && out[1] == 16.0);
  CHECK(sla_dot(ctx, x, x, &d) == SLA_OK && d == 13.0);
  CHECK(sla_spmv_host(ctx, A, x0true, out) == SLA_OK && out[0] == 8.0 && out[1] == 18.0);

  sla_csr* At = NULL;
  CHECK(sla_csr_transpose(ctx, A, &At) == SLA_OK && sla_csr_to_host(ctx, At, rp, col, val) == SLA_OK);
  CHECK(val[0] == 1.0 && val[1] == 3.0 && val[2] == 2.0 && val[3] == 4.0);

  /* linSolve0 BICGSTAB_ aa0 b0 (0.1, 0.1) */
  const double tenth[] = {0.1, 0.1};
  CHECK(sla_vec_from_host(ctx, 2, b0, &b) == SLA_OK && sla_vec_from_host(ctx, 2, tenth, &x0) == SLA_OK);
  CHECK(sla_vec_create(ctx, 2, &sol) == SLA_OK);
  int iters = -1; double res = -1;
  CHECK(sla_linsolve0(ctx, SLA_BICGSTAB_, A, b, x0, NULL, sol, &iters, &res) == SLA_OK);
  CHECK(sla_vec_to_host(ctx, sol, out) == SLA_OK);
  CHECK(iters == 2 && sqrt((out[0] - 2) * (out[0] - 2) + (out[1] - 3) * (out[1] - 3)) <= 1e-12);
  CHECK(sla_linsolve0(ctx, SLA_GMRES_, A, b, x0, NULL, sol, &iters, &res) == SLA_ERR_UNSUPPORTED_METHOD);
  CHECK(strstr(sla_last_error(ctx), "Only BICGSTAB_, CGS_, and CGNE_ are implemented") != NULL);

  /* error codes */
  sla_vec* bad = NULL;
  CHECK(sla_vec_create(ctx, 3, &bad) == SLA_OK);
  CHECK(sla_spmv(ctx, A, bad, y) == SLA_ERR_SIZE_MISMATCH);
  const int64_t oi[] = {0}, oj[] = {2}; const double ov[] = {1.0};
  sla_csr* B = NULL;
  CHECK(sla_csr_from_coo(ctx, 2, 2, 1, oi, oj, ov, &B) == SLA_ERR_OOB_INDEX && B == NULL);

  sla_vec_free(x); sla_vec_free(y); sla_vec_free(b); sla_vec_free(x0); sla_vec_free(sol); sla_vec_free(bad);
  sla_csr_free(A); sla_csr_free(At);
  printf("abi_smoke ok (%s, %lld kernel launches)\n", sla_version(), (long long)sla_launch_count(ctx));
  sla_finalize(ctx);
  return 0;
}
