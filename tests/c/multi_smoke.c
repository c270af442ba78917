/* multi_smoke.c — a plain C99 program that drives TWO GPUs from ONE process through include/sla_b200.h only
 * (sla_init_multi): the 5-point Laplacian on a 64 x 64 grid, (#>) and linSolve0 BICGSTAB_ on both GPUs, compared with the
 * same calls on GPU 0 alone through the single-GPU entry points.  Exit code 0 and "MULTI_SMOKE OK" on success.
 * build: gcc -std=c99 -Iinclude tests/c/multi_smoke.c -Lsparse_linear_algebra_b200 -lsla_b200 -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sla_b200.h"

#define CHECK(call, ctx_err)                                                              \
  do {                                                                                    \
    sla_status s_ = (call);                                                               \
    if (s_ != SLA_OK) { printf("FAILED %s -> %d (%s)\n", #call, (int)s_, ctx_err); return 1; } \
  } while (0)

int main(void) {
  const int g = 64;
  const int64_t n = (int64_t)g * g;
  sla_mctx* m = NULL;
  sla_status s = sla_init_multi(2, NULL, &m);
  if (s != SLA_OK) { printf("sla_init_multi failed: %s\n", sla_multi_last_error(NULL)); return s == SLA_ERR_INVALID ? 77 : 1; }
  if (sla_multi_world(m) != 2) { printf("world != 2\n"); return 1; }

  /* two GPUs, one process */
  sla_mcsr* A = NULL;
  sla_mvec *x = NULL, *b = NULL, *x0 = NULL, *sol = NULL;
  CHECK(sla_multi_csr_generate(m, 2 /* GEN_LAPLACE2D */, n, 5, 0, g, &A), sla_multi_last_error(m));
  CHECK(sla_multi_vec_generate(m, n, 3, &x), sla_multi_last_error(m));
  CHECK(sla_multi_vec_create(m, n, &b), sla_multi_last_error(m));
  CHECK(sla_multi_vec_create(m, n, &x0), sla_multi_last_error(m));
  CHECK(sla_multi_vec_create(m, n, &sol), sla_multi_last_error(m));
  CHECK(sla_multi_spmv(m, A, x, b), sla_multi_last_error(m));
  double* bh = (double*)malloc(sizeof(double) * (size_t)n);
  double* xh = (double*)malloc(sizeof(double) * (size_t)n);
  double* sh = (double*)malloc(sizeof(double) * (size_t)n);
  CHECK(sla_multi_vec_to_host(m, b, bh), sla_multi_last_error(m));
  CHECK(sla_multi_vec_to_host(m, x, xh), sla_multi_last_error(m));
  int iters = 0; double res = 0.0, dot2 = 0.0;
  sla_solve_opts o; sla_solve_opts_default(&o);
  o.tol_abs = 1e-9; o.tol_rel = 1e-12; o.max_iters = 400;
  CHECK(sla_multi_linsolve0(m, SLA_BICGSTAB_, A, b, x0, &o, sol, &iters, &res), sla_multi_last_error(m));
  CHECK(sla_multi_vec_to_host(m, sol, sh), sla_multi_last_error(m));
  CHECK(sla_multi_dot(m, x, b, &dot2), sla_multi_last_error(m));

  /* the same on GPU 0 alone */
  sla_ctx* c = NULL;
  CHECK(sla_init(0, &c), sla_last_error(NULL));
  sla_csr* A1 = NULL; sla_vec *x1 = NULL, *b1 = NULL, *z1 = NULL, *s1 = NULL;
  CHECK(sla_csr_generate(c, 2, n, 5, 0, g, &A1), sla_last_error(c));
  CHECK(sla_vec_from_host(c, n, xh, &x1), sla_last_error(c));
  CHECK(sla_vec_create(c, n, &b1), sla_last_error(c));
  CHECK(sla_vec_create(c, n, &z1), sla_last_error(c));
  CHECK(sla_vec_create(c, n, &s1), sla_last_error(c));
  CHECK(sla_spmv(c, A1, x1, b1), sla_last_error(c));
  double* b1h = (double*)malloc(sizeof(double) * (size_t)n);
  double* s1h = (double*)malloc(sizeof(double) * (size_t)n);
  CHECK(sla_vec_to_host(c, b1, b1h), sla_last_error(c));
  int iters1 = 0; double res1 = 0.0, dot1 = 0.0;
  CHECK(sla_linsolve0(c, SLA_BICGSTAB_, A1, b1, z1, &o, s1, &iters1, &res1), sla_last_error(c));
  CHECK(sla_vec_to_host(c, s1, s1h), sla_last_error(c));
  CHECK(sla_dot(c, x1, b1, &dot1), sla_last_error(c));

  int bad = memcmp(bh, b1h, sizeof(double) * (size_t)n) != 0;     /* (#>): the same bits on 1 and 2 GPUs */
  double emax = 0.0, xmax = 0.0;
  for (int64_t i = 0; i < n; ++i) { double d = fabs(sh[i] - xh[i]); if (d > emax) emax = d; if (fabs(xh[i]) > xmax) xmax = fabs(xh[i]); }
  double dmax = 0.0;
  for (int64_t i = 0; i < n; ++i) { double d = fabs(sh[i] - s1h[i]); if (d > dmax) dmax = d; }
  printf("matvec identical: %s ; iters %d (2 GPUs) vs %d (1 GPU) ; |x - x_true| = %.3e ; |x_2gpu - x_1gpu| = %.3e ; dot %.17g vs %.17g\n",
         bad ? "NO" : "yes", iters, iters1, emax, dmax, dot2, dot1);
  if (bad || emax > 1e-6 * xmax || dmax > 1e-6 * xmax || fabs(dot2 - dot1) > 1e-12 * fabs(dot1) || iters <= 0) { printf("MULTI_SMOKE FAILED\n"); return 1; }

  sla_vec_free(x1); sla_vec_free(b1); sla_vec_free(z1); sla_vec_free(s1); sla_csr_free(A1); sla_finalize(c);
  sla_multi_vec_free(x); sla_multi_vec_free(b); sla_multi_vec_free(x0); sla_multi_vec_free(sol); sla_multi_csr_free(A);
  sla_finalize_multi(m);
  free(bh); free(xh); free(sh); free(b1h); free(s1h);
  printf("MULTI_SMOKE OK\n");
  return 0;
}
