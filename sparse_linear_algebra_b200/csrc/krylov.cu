// krylov.cu — the Krylov layer of the hot path: bicgsInit/bicgstabStep, cgsInit/cgsStep, cgneInit/cgneStep,
// linSolve0, arnoldi (src/Numeric/LinearAlgebra/Sparse.hs:630-667, 855-981, 1016-1072) and a restarted
// GMRES built on the same Arnoldi kernels.  Every step is a fixed sequence of SpMV launches (with fused dot
// epilogues) and fused elementwise kernels; alpha / omega / beta are computed on the device by the last CTA
// of each reduction and never visit the host.
#include "blas1.cuh"

#include <math.h>
#include <new>
#include <vector>

// ---- state records -----------------------------------------------------------------------------------

static sla_status krylov_alloc(sla_ctx* c, int kind, int64_t n, bool need_u, sla_krylov** out) {
  sla_krylov* st = new (std::nothrow) sla_krylov();
  if (!st) return sla_fail(c, SLA_ERR_ALLOC, "krylov alloc");
  memset(st, 0, sizeof(*st));
  st->ctx = c; st->kind = kind; st->n = n;
  sla_status s = SLA_OK;
  sla_vec** all[] = {&st->x, &st->r, &st->p, &st->t0, &st->t1, &st->t2};
  for (auto v : all) if (s == SLA_OK) s = sla_vec_alloc(c, n, v);
  if (s == SLA_OK && need_u) s = sla_vec_alloc(c, n, &st->u);
  if (s != SLA_OK) { sla_krylov_free(st); return s; }
  *out = st;
  return SLA_OK;
}

extern "C" void sla_krylov_free(sla_krylov* st) {
  if (!st) return;
  if (st->ctx && st->ctx->scal_owner == st) st->ctx->scal_owner = nullptr;
  sla_vec_free(st->x); sla_vec_free(st->r); sla_vec_free(st->p); sla_vec_free(st->u);
  sla_vec_free(st->t0); sla_vec_free(st->t1); sla_vec_free(st->t2);
  delete st;
}

// A deep copy of the record (state AND work vectors are fresh allocations).  The reference's steps are pure:
// `iterate (bicgstabStep aa r0hat) st0 !! 20` (README.md:208) keeps st0 alive, so a drop-in `step` is clone-then-advance
// (hs/.../B200.hs); callers that own the record exclusively keep using the in-place step.
extern "C" sla_status sla_krylov_clone(sla_ctx* c, const sla_krylov* st, sla_krylov** out) {
  if (!c || !st || !out) return SLA_ERR_INVALID;
  *out = nullptr;
  sla_krylov* cp = nullptr;
  SLA_TRY(krylov_alloc(c, st->kind, st->n, st->u != nullptr, &cp));
  sla_status s = sla_vec_copy(c, st->x, cp->x);
  if (s == SLA_OK) s = sla_vec_copy(c, st->r, cp->r);
  if (s == SLA_OK) s = sla_vec_copy(c, st->p, cp->p);
  if (s == SLA_OK && st->u) s = sla_vec_copy(c, st->u, cp->u);
  if (s != SLA_OK) { sla_krylov_free(cp); return s; }
  cp->rho_valid = false;              // rho = r <.> r0hat is recomputed by the clone's first step (the reference recomputes it every step)
  *out = cp;
  return SLA_OK;
}

static const sla_vec* krylov_field(const sla_krylov* st, int field) {
  switch (field) {
    case SLA_FIELD_X: return st->x;
    case SLA_FIELD_R: return st->r;
    case SLA_FIELD_P: return st->p;
    case SLA_FIELD_U: return st->u;
  }
  return nullptr;
}

extern "C" sla_status sla_krylov_view(sla_ctx* c, const sla_krylov* st, int field, const sla_vec** view) {
  if (!c || !st || !view) return SLA_ERR_INVALID;
  const sla_vec* v = krylov_field(st, field);
  if (!v) return sla_fail(c, SLA_ERR_INVALID, "krylov_view: this record has no such field");
  *view = v;
  return SLA_OK;
}

extern "C" sla_status sla_krylov_get(sla_ctx* c, const sla_krylov* st, int field, double* host_out) {
  const sla_vec* v = nullptr;
  SLA_TRY(sla_krylov_view(c, st, field, &v));
  return sla_vec_to_host(c, v, host_out);
}

static sla_status check_system(sla_ctx* c, const char* who, const sla_csr* A, const sla_vec* b, const sla_vec* x0) {
  if (!c || !A || !b || !x0) return SLA_ERR_INVALID;
  if (csr_xdim(A) != x0->n) {       // aa #> x0 : matVec dimension check   Common.hs:248-250
    snprintf(c->err, sizeof(c->err), "%s: matVec : mismatched dimensions (%lld,%lld)", who, (long long)A->n, (long long)x0->n);
    return SLA_ERR_SIZE_MISMATCH;
  }
  if (A->m != b->n || A->m != csr_xdim(A)) {
    snprintf(c->err, sizeof(c->err), "%s: Matrix-vector dimensions are incompatible: Matrix is (%lld,%lld), whereas vector is %lld",
             who, (long long)A->m, (long long)A->n, (long long)b->n);
    return SLA_ERR_SIZE_MISMATCH;
  }
  return SLA_OK;
}

// r0 = b ^-^ (aa #> x0), copied into up to three records   (bicgsInit / cgsInit / cgneInit)
static sla_status init_residual(sla_ctx* c, const sla_csr* A, const sla_vec* b, const sla_vec* x0, sla_krylov* st,
                                double* o0, double* o1, double* o2) {
  SLA_TRY(sla_vec_copy(c, x0, st->x));
  SLA_TRY(sla_spmv_launch(c, A, x0->d, st->t0->d, EPI_NONE, nullptr, nullptr, FIN_STORE, S_TMP0));
  Ptrs<2> in{{b->d, st->t0->d}};
  Ptrs<3> out{{o0, o1, o2}};
  SLA_TRY(ew_launch(c, OpResidInit{}, st->n, in, out, FIN_STORE, S_TMP1));
  return SLA_OK;
}

static void bump(sla_vec* v) { sla_touch(v); }

// rho = r <.> rhat is carried on the device from the previous step while nobody touched r, rhat or the
// scalar slots; otherwise it is recomputed (the reference recomputes it every step, Sparse.hs:974).
static sla_status ensure_rho(sla_ctx* c, sla_krylov* st, const sla_vec* rhat) {
  const bool ok = st->rho_valid && c->scal_owner == st && st->rho_r0hat == rhat &&
                  st->rho_r0hat_version == rhat->version && st->rho_r_version == st->r->version;
  if (ok) return SLA_OK;
  Ptrs<2> in{{st->r->d, rhat->d}}; Ptrs<0> o{};
  SLA_TRY(ew_launch(c, OpDot{}, st->n, in, o, FIN_STORE, S_RHO));
  c->scal_owner = st;
  return SLA_OK;
}

static void cache_rho(sla_krylov* st, const sla_vec* rhat) {
  st->rho_valid = true; st->rho_r0hat = rhat;
  st->rho_r0hat_version = rhat->version; st->rho_r_version = st->r->version;
}

// ---- BiCGSTAB ------------------------------------------------------------------------------------------

extern "C" sla_status sla_bicgstab_init(sla_ctx* c, const sla_csr* A, const sla_vec* b, const sla_vec* x0, sla_krylov** out) {
  if (!out) return SLA_ERR_INVALID;
  *out = nullptr;
  SLA_TRY(check_system(c, "bicgsInit", A, b, x0));
  sla_krylov* st = nullptr;
  SLA_TRY(krylov_alloc(c, SLA_BICGSTAB_, A->m, false, &st));
  sla_status s = init_residual(c, A, b, x0, st, st->r->d, st->p->d, st->p->d);   // BICGSTAB x0 r0 r0
  if (s != SLA_OK) { sla_krylov_free(st); return s; }
  *out = st;
  return SLA_OK;
}

extern "C" sla_status sla_bicgstab_step(sla_ctx* c, const sla_csr* A, const sla_vec* r0hat, sla_krylov* st) {
  SLA_GUARD(c);
  if (!c || !A || !r0hat || !st || st->kind != SLA_BICGSTAB_) return SLA_ERR_INVALID;
  if (A->m != st->n || csr_xdim(A) != st->n || r0hat->n != st->n) return sla_fail(c, SLA_ERR_SIZE_MISMATCH, "bicgstabStep: dimensions differ");
  const int64_t n = st->n;
  double *x = st->x->d, *r = st->r->d, *p = st->p->d, *aap = st->t0->d, *s = st->t1->d, *aas = st->t2->d;
  SLA_TRY(ensure_rho(c, st, r0hat));
  // aap = aa #> p ; alphaj = (r <.> r0hat) / (aap <.> r0hat)
  SLA_TRY(sla_spmv_launch(c, A, p, aap, EPI_DOT1, r0hat->d, nullptr, FIN_BICG_ALPHA, 0));
  // sj = r ^-^ (alphaj .* aap)
  { Ptrs<2> in{{r, aap}}; Ptrs<1> o{{s}}; OpBicgS op; op.alpha = 0; SLA_TRY(ew_launch(c, op, n, in, o)); }
  // aasj = aa #> sj ; omegaj = (aasj <.> sj) / (aasj <.> aasj)
  SLA_TRY(sla_spmv_launch(c, A, s, aas, EPI_DOT2_YY, s, nullptr, FIN_BICG_OMEGA, 0));
  // xj1, rj1 ; betaj = (rj1 <.> r0hat)/(r <.> r0hat) * alphaj / omegaj
  { Ptrs<5> in{{x, p, s, aas, r0hat->d}}; Ptrs<2> o{{x, r}}; OpBicgXR op; op.alpha = op.omega = 0;
    SLA_TRY(ew_launch(c, op, n, in, o, FIN_BICG_BETA, 0)); }
  // pj1 = rj1 ^+^ (betaj .* (p ^-^ (omegaj .* aap)))
  { Ptrs<3> in{{r, p, aap}}; Ptrs<1> o{{p}}; OpBicgP op; op.beta = op.omega = 0; SLA_TRY(ew_launch(c, op, n, in, o)); }
  bump(st->x); bump(st->r); bump(st->p);
  cache_rho(st, r0hat);
  return SLA_OK;
}

// ---- CGS -----------------------------------------------------------------------------------------------

extern "C" sla_status sla_cgs_init(sla_ctx* c, const sla_csr* A, const sla_vec* b, const sla_vec* x0, sla_krylov** out) {
  if (!out) return SLA_ERR_INVALID;
  *out = nullptr;
  SLA_TRY(check_system(c, "cgsInit", A, b, x0));
  sla_krylov* st = nullptr;
  SLA_TRY(krylov_alloc(c, SLA_CGS_, A->m, true, &st));
  sla_status s = init_residual(c, A, b, x0, st, st->r->d, st->p->d, st->u->d);   // CGS x0 r0 r0 r0
  if (s != SLA_OK) { sla_krylov_free(st); return s; }
  *out = st;
  return SLA_OK;
}

extern "C" sla_status sla_cgs_step(sla_ctx* c, const sla_csr* A, const sla_vec* rhat, sla_krylov* st) {
  SLA_GUARD(c);
  if (!c || !A || !rhat || !st || st->kind != SLA_CGS_) return SLA_ERR_INVALID;
  if (A->m != st->n || csr_xdim(A) != st->n || rhat->n != st->n) return sla_fail(c, SLA_ERR_SIZE_MISMATCH, "cgsStep: dimensions differ");
  const int64_t n = st->n;
  double *x = st->x->d, *r = st->r->d, *p = st->p->d, *u = st->u->d, *t0 = st->t0->d, *q = st->t1->d, *upq = st->t2->d;
  SLA_TRY(ensure_rho(c, st, rhat));
  // aap = aa #> p ; alphaj = (r `dot` rhat) / (aap `dot` rhat)
  SLA_TRY(sla_spmv_launch(c, A, p, t0, EPI_DOT1, rhat->d, nullptr, FIN_BICG_ALPHA, 0));
  // q = u ^-^ (alphaj .* aap) ; xj1 = x ^+^ (alphaj .* (u ^+^ q))
  { Ptrs<3> in{{u, t0, x}}; Ptrs<3> o{{q, upq, x}}; OpCgsQ op; op.alpha = 0; SLA_TRY(ew_launch(c, op, n, in, o)); }
  // aa #> (u ^+^ q)
  SLA_TRY(sla_spmv_launch(c, A, upq, t0, EPI_NONE, nullptr, nullptr, FIN_STORE, S_TMP0));
  // rj1 = r ^-^ (alphaj .* ...) ; betaj = (rj1 `dot` rhat) / (r `dot` rhat)
  { Ptrs<3> in{{r, t0, rhat->d}}; Ptrs<1> o{{r}}; OpCgsR op; op.alpha = 0; SLA_TRY(ew_launch(c, op, n, in, o, FIN_CGS_BETA, 0)); }
  // uj1 = rj1 ^+^ (betaj .* q) ; pj1 = uj1 ^+^ (betaj .* (q ^+^ (betaj .* p)))
  { Ptrs<3> in{{r, q, p}}; Ptrs<2> o{{u, p}}; OpCgsUP op; op.beta = 0; SLA_TRY(ew_launch(c, op, n, in, o)); }
  bump(st->x); bump(st->r); bump(st->p); bump(st->u);
  cache_rho(st, rhat);
  return SLA_OK;
}

// ---- CGNE ----------------------------------------------------------------------------------------------

static sla_status ensure_transpose(sla_ctx* c, const sla_csr* A) {
  if (A->T) return SLA_OK;
  sla_csr* t = nullptr;
  SLA_TRY(sla_csr_transpose(c, A, &t));   // the reference re-transposes every step (Sparse.hs:878); built once here
  const_cast<sla_csr*>(A)->T = t;
  return SLA_OK;
}

extern "C" sla_status sla_cgne_init(sla_ctx* c, const sla_csr* A, const sla_vec* b, const sla_vec* x0, sla_krylov** out) {
  if (!out) return SLA_ERR_INVALID;
  *out = nullptr;
  SLA_TRY(check_system(c, "cgneInit", A, b, x0));
  // a row-partitioned matrix needs its distributed transpose attached first (sla_csr_transpose_dist + sla_csr_attach_transpose)
  if (A->dist && !A->T) return sla_fail(c, SLA_ERR_INVALID, "cgneInit: attach the distributed transpose of the row-partitioned matrix first");
  if (!A->dist) SLA_TRY(ensure_transpose(c, A));
  sla_krylov* st = nullptr;
  SLA_TRY(krylov_alloc(c, SLA_CGNE_, A->m, false, &st));
  sla_status s = init_residual(c, A, b, x0, st, st->r->d, st->t1->d, st->t1->d);
  // p0 = transposeSM aa #> r0
  if (s == SLA_OK) s = sla_spmv_launch(c, A->T, st->r->d, st->p->d, EPI_NONE, nullptr, nullptr, FIN_STORE, S_TMP0);
  if (s != SLA_OK) { sla_krylov_free(st); return s; }
  *out = st;
  return SLA_OK;
}

extern "C" sla_status sla_cgne_step(sla_ctx* c, const sla_csr* A, sla_krylov* st) {
  SLA_GUARD(c);
  if (!c || !A || !st || st->kind != SLA_CGNE_) return SLA_ERR_INVALID;
  if (A->dist && !A->T) return sla_fail(c, SLA_ERR_INVALID, "cgneStep: attach the distributed transpose of the row-partitioned matrix first");
  if (A->m != st->n || csr_xdim(A) != st->n) return sla_fail(c, SLA_ERR_SIZE_MISMATCH, "cgneStep: dimensions differ");
  if (!A->dist) SLA_TRY(ensure_transpose(c, A));
  const int64_t n = st->n;
  double *x = st->x->d, *r = st->r->d, *p = st->p->d, *ap = st->t0->d, *atr = st->t1->d;
  c->scal_owner = st;
  // alphai = (r `dot` r) / (p `dot` p)
  { Ptrs<2> in{{r, p}}; Ptrs<0> o{}; SLA_TRY(ew_launch(c, OpSelfDot2{}, n, in, o, FIN_CGNE_ALPHA, 0)); }
  SLA_TRY(sla_spmv_launch(c, A, p, ap, EPI_NONE, nullptr, nullptr, FIN_STORE, S_TMP0));
  // x1 = x ^+^ (alphai .* p) ; r1 = r ^-^ (alphai .* (aa #> p)) ; beta = (r1 `dot` r1) / (r `dot` r)
  { Ptrs<4> in{{x, p, r, ap}}; Ptrs<2> o{{x, r}}; OpCgneXR op; op.alpha = 0; SLA_TRY(ew_launch(c, op, n, in, o, FIN_CGNE_BETA, 0)); }
  // p1 = transpose aa #> r1 ^+^ (beta .* p)
  SLA_TRY(sla_spmv_launch(c, A->T, r, atr, EPI_NONE, nullptr, nullptr, FIN_STORE, S_TMP0));
  { Ptrs<2> in{{atr, p}}; Ptrs<1> o{{p}}; OpCgneP op; op.beta = 0; SLA_TRY(ew_launch(c, op, n, in, o)); }
  bump(st->x); bump(st->r); bump(st->p);
  return SLA_OK;
}

// ---- linSolve0 -----------------------------------------------------------------------------------------

extern "C" void sla_solve_opts_default(sla_solve_opts* o) {
  if (!o) return;
  o->max_iters = 200; o->tol_abs = 1e-6; o->tol_rel = 1e-4; o->true_residual = 1; o->check_every = 1;
}

static sla_status residual_norm(sla_ctx* c, const sla_csr* A, const sla_vec* x, const sla_vec* b, double* out) {
  // norm2 ((aa #> x) ^-^ b) with the SpMV, the subtraction and the sum of squares in one kernel   Sparse.hs:1041
  SLA_TRY(sla_spmv_launch(c, A, x->d, nullptr, EPI_RESNORM, b->d, nullptr, FIN_STORE, S_RES2));
  double r2 = 0;
  SLA_TRY(sla_read_scalars(c, S_RES2, 1, &r2));
  *out = sqrt(r2);
  return SLA_OK;
}

extern "C" sla_status sla_linsolve0(sla_ctx* c, int method, const sla_csr* A, const sla_vec* b, const sla_vec* x0,
                                    const sla_solve_opts* opts_in, sla_vec* x, int* iters, double* resnorm) {
  if (!c || !A || !b || !x0 || !x) return SLA_ERR_INVALID;
  sla_solve_opts o;
  sla_solve_opts_default(&o);
  if (opts_in) o = *opts_in;
  if (o.max_iters <= 0) o.max_iters = 200;
  if (o.check_every <= 0) o.check_every = 1;
  if (iters) *iters = 0;
  if (resnorm) *resnorm = 0.0;
  if (A->m != b->n) {                         // | m /= nb = throwM (MatVecSizeMismatchException "linSolve0" dm nb)   :1022
    snprintf(c->err, sizeof(c->err), "linSolve0 : Matrix-vector dimensions are incompatible: Matrix is (%lld,%lld) , whereas vector is %lld",
             (long long)A->m, (long long)A->n, (long long)b->n);
    return SLA_ERR_SIZE_MISMATCH;
  }
  if (x->n != csr_xdim(A)) return sla_fail(c, SLA_ERR_SIZE_MISMATCH, "linSolve0 : output vector has the wrong dimension");
  int diag = 0;
  SLA_TRY(sla_csr_is_diagonal(c, A, &diag));
  if (diag) {                                 // isDiagonalSM aa' = return $ reciprocal aa' #> b'   :1024-1025
    Ptrs<2> in{{A->val, b->d}}; Ptrs<1> out{{x->d}};
    SLA_TRY(ew_launch(c, OpDiagSolve{}, A->m, in, out));
    sla_touch(x);
    return SLA_OK;
  }
  if (method != SLA_BICGSTAB_ && method != SLA_CGS_ && method != SLA_CGNE_) {   // IterE   :1031
    const char* nm = method == SLA_GMRES_ ? "GMRES_" : method == SLA_BCG_ ? "BCG_" : "?";
    snprintf(c->err, sizeof(c->err), "linSolve0 : Only BICGSTAB_, CGS_, and CGNE_ are implemented, got: %s", nm);
    return SLA_ERR_UNSUPPORTED_METHOD;
  }
  SLA_TRY(check_system(c, "linSolve0", A, b, x0));
  const int64_t n = A->m;
  // r0hat = b ^-^ (aa #> x0) ; tol = max tolAbs (tolRel * norm2 r0hat)   :1032-1037
  sla_vec* r0hat = nullptr;
  SLA_TRY(sla_vec_alloc(c, n, &r0hat));
  sla_krylov* st = nullptr;
  sla_status s = method == SLA_BICGSTAB_ ? sla_bicgstab_init(c, A, b, x0, &st)
               : method == SLA_CGS_      ? sla_cgs_init(c, A, b, x0, &st)
                                         : sla_cgne_init(c, A, b, x0, &st);
  double r0norm = 0, res = 0;
  if (s == SLA_OK) s = sla_vec_copy(c, st->r, r0hat);          // same arithmetic as bicgsInit's r0
  if (s == SLA_OK) s = sla_norm2(c, r0hat, &r0norm);
  const double tol = o.tol_abs > o.tol_rel * r0norm ? o.tol_abs : o.tol_rel * r0norm;
  int it = 0;
  while (s == SLA_OK && it < o.max_iters) {                    // runIter   :1043-1052
    s = method == SLA_BICGSTAB_ ? sla_bicgstab_step(c, A, r0hat, st)
      : method == SLA_CGS_      ? sla_cgs_step(c, A, r0hat, st)
                                : sla_cgne_step(c, A, st);
    if (s != SLA_OK) break;
    ++it;
    if (it % o.check_every == 0 || it == o.max_iters) {
      if (o.true_residual) s = residual_norm(c, A, st->x, b, &res);
      else { s = sla_norm2(c, st->r, &res); }
      if (s != SLA_OK) break;
      if (res <= tol) break;                                   // NaN <= tol is False: runs to max_iters like the reference
    }
  }
  if (s == SLA_OK) s = sla_vec_copy(c, st->x, x);
  if (s == SLA_OK) s = sla_sync(c);
  if (iters) *iters = it;
  if (resnorm) *resnorm = res;
  sla_krylov_free(st);
  sla_vec_free(r0hat);
  return s;
}

extern "C" sla_status sla_linsolve0_host(sla_ctx* c, int method, const sla_csr* A, const double* b_host, const double* x0_host,
                                         const sla_solve_opts* opts, double* x_host, int* iters, double* resnorm) {
  if (!c || !A || !b_host || !x0_host || !x_host) return SLA_ERR_INVALID;
  sla_vec *b = nullptr, *x0 = nullptr, *x = nullptr;
  sla_status s = sla_vec_from_host(c, A->m, b_host, &b);
  if (s == SLA_OK) s = sla_vec_from_host(c, A->n, x0_host, &x0);
  if (s == SLA_OK) s = sla_vec_create(c, A->n, &x);
  if (s == SLA_OK) s = sla_linsolve0(c, method, A, b, x0, opts, x, iters, resnorm);
  if (s == SLA_OK) s = sla_vec_to_host(c, x, x_host);
  sla_vec_free(b); sla_vec_free(x0); sla_vec_free(x);
  return s;
}

// ---- Arnoldi -------------------------------------------------------------------------------------------
// One Arnoldi step = (#>) + three streaming kernels, all asynchronous: the Hessenberg column, the breakdown test and (for
// GMRES) the Givens rotations stay on the device, so a whole cycle of kn steps is queued without a host round trip.
//   tsmv_t_kernel   h_k = q_k <.> w for ALL k <= j in one launch (w read once; NC <= 32 columns per launch)
//   lincomb_kernel  w <- w - sum_k h_k q_k (the reference's left-to-right association) + ||w||^2
//   arn_finish_kernel  q_{j+1} = recip(||w||) .* w ; thread 0 files the column of H (and rotates it for GMRES)
// Algorithmic bytes per step j: B_spmv + 16 n (j+1) + 32 n — Q is read twice, w is read by the dots,
// read and written by the projection, read again by the scaling that writes q_{j+1}.

#define TS_NC 32   // basis columns per launch of the tall-skinny dot kernel
#define TS_WARPS (EW_THREADS / 32)
#define TS_CPW (TS_NC / TS_WARPS)   // columns per warp
#define TS_UNROLL 2

// h[k0 + k] = q_{k0+k} <.> w   for k < nc <= TS_NC        (hhcoli = fmap (`dot` aqi) qv, Sparse.hs:655)
// A CTA walks 512-byte pieces of the vectors; its 8 warps split the COLUMNS (warp v owns columns v, v + 8, v + 16, v + 24), so a
// thread carries 4 accumulators instead of 32 — the first version (every thread, every column) needed 128 registers, ran at
// 25 % occupancy with 64 KB of loads in flight per SM and reached 4.6 TB/s (profiles/r02_arnoldi_kernels.txt); w is read by
// all 8 warps but only the first read of a piece leaves the SM.  No cross-warp reduction is needed: a column belongs to one warp.
__global__ void __launch_bounds__(EW_THREADS, 4)
tsmv_t_kernel(const double* __restrict__ Q, int64_t ld, int64_t n, int k0, int nc, const double* __restrict__ w,
              double* scal, double* partials, unsigned int* counter, int fin, int slot0, sla_p2p_args pa) {
  __shared__ double colsum[TS_NC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[TS_CPW];
#pragma unroll
  for (int q = 0; q < TS_CPW; ++q) acc[q] = 0.0;
  const int64_t n2 = n >> 1, ld2 = ld >> 1;                         // ld is a multiple of 16 doubles
  const double2* Q2 = reinterpret_cast<const double2*>(Q + (int64_t)k0 * ld);
  const double2* w2 = reinterpret_cast<const double2*>(w);
  const int64_t step = (int64_t)gridDim.x * 32;
  int64_t i = (int64_t)blockIdx.x * 32 + lane;
#pragma unroll 1
  for (; i + (TS_UNROLL - 1) * step < n2; i += TS_UNROLL * step) {     // TS_UNROLL pieces per trip: 8 + 2 loads in flight per thread
    double2 wv[TS_UNROLL], qv[TS_UNROLL][TS_CPW];
#pragma unroll
    for (int u = 0; u < TS_UNROLL; ++u) {
      wv[u] = w2[i + u * step];
#pragma unroll
      for (int q = 0; q < TS_CPW; ++q)
        if (warp + q * TS_WARPS < nc) qv[u][q] = Q2[(int64_t)(warp + q * TS_WARPS) * ld2 + i + u * step];
    }
#pragma unroll
    for (int u = 0; u < TS_UNROLL; ++u)
#pragma unroll
      for (int q = 0; q < TS_CPW; ++q)
        if (warp + q * TS_WARPS < nc) { acc[q] += qv[u][q].x * wv[u].x; acc[q] += qv[u][q].y * wv[u].y; }
  }
#pragma unroll 1
  for (; i < n2; i += step) {
    const double2 wv = w2[i];
#pragma unroll
    for (int q = 0; q < TS_CPW; ++q)
      if (warp + q * TS_WARPS < nc) {
        const double2 qv = Q2[(int64_t)(warp + q * TS_WARPS) * ld2 + i];
        acc[q] += qv.x * wv.x; acc[q] += qv.y * wv.y;
      }
  }
  if ((n & 1) && blockIdx.x == 0 && lane == 0) {
#pragma unroll
    for (int q = 0; q < TS_CPW; ++q)
      if (warp + q * TS_WARPS < nc) acc[q] += Q[(int64_t)(k0 + warp + q * TS_WARPS) * ld + n - 1] * w[n - 1];
  }
#pragma unroll
  for (int q = 0; q < TS_CPW; ++q) {
    acc[q] = warp_sum(acc[q]);
    if (lane == 0) colsum[warp + q * TS_WARPS] = acc[q];
  }
  __syncthreads();
  // grid reduction, one column per warp again: per-CTA sums -> ticket -> the last CTA adds them in a fixed order
  // (lane-strided partial sums + shuffle tree), completes the all-reduce over peer memory when there are ranks, and stores
  __shared__ bool is_last;
  const unsigned int nblk = gridDim.x;
  if (threadIdx.x < TS_NC) partials[(size_t)threadIdx.x * nblk + blockIdx.x] = colsum[threadIdx.x];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == nblk - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
#pragma unroll
  for (int q = 0; q < TS_CPW; ++q) {
    const int k = warp + q * TS_WARPS;
    const volatile double* p = partials + (size_t)k * nblk;
    double s = 0.0;
    for (unsigned int b = lane; b < nblk; b += 32) s += p[b];
    s = warp_sum(s);
    if (lane == 0) colsum[k] = s;
  }
  __syncthreads();
  if (pa.world > 1) p2p_allreduce_block(pa, colsum, TS_NC);
  if (threadIdx.x == 0) {
    finalize_scalars(FIN_STORE, slot0 + k0, scal, colsum, TS_NC);    // with or without FIN_DEFER: the sums go to their slots
    *counter = 0u;
  }
  (void)fin;
}

// out_i = base_i -/+ (((c_0 q_0i) + c_1 q_1i) + ... + c_{nc-1} q_{nc-1,i}), coefficients in scal[slot0..];
// SIGN = -1: qipnn = aqi ^-^ foldl' (^+^) zv (zipWith (.*) hhcoli qv), plus sum of squares (Sparse.hs:657-659)
// SIGN = +1: x = x ^+^ Q y (GMRES update); nc_dev (optional) caps the column count with a device-side value
template <int SIGN>
__global__ void __launch_bounds__(EW_THREADS)
lincomb_kernel(const double* __restrict__ Q, int64_t ld, int64_t n, int nc, const double* base, double* out,
               double* scal, double* partials, unsigned int* counter, int fin, int slot0, const int* nc_dev, sla_p2p_args pa) {
  __shared__ double coef[SLA_MAX_KRYLOV + 2];
  __shared__ double red[32];
  if (nc_dev) { const int cap = *nc_dev; if (cap < nc) nc = cap; }
  for (int k = threadIdx.x; k < nc; k += blockDim.x) coef[k] = scal[slot0 + k];
  __syncthreads();
  double acc[1] = {0.0};
  const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * blockDim.x, gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const double2* Q2 = reinterpret_cast<const double2*>(Q);
  const int64_t ld2 = ld >> 1;
  if (nc > 0) {
    for (int64_t i = gtid; i < n2; i += stride) {
      double2 s = make_double2(0.0, 0.0);
      for (int kb = 0; kb < nc; kb += 8) {              // loads in batches of 8 columns, sums strictly in column order
        double2 q[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (kb + k < nc) q[k] = Q2[(int64_t)(kb + k) * ld2 + i];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (kb + k < nc) {
            const double ck = coef[kb + k];
            if (kb + k == 0) { s.x = __dmul_rn(ck, q[k].x); s.y = __dmul_rn(ck, q[k].y); }   // zeroSV ^+^ (h0 .* q0) = h0 .* q0
            else { s.x = __dadd_rn(s.x, __dmul_rn(ck, q[k].x)); s.y = __dadd_rn(s.y, __dmul_rn(ck, q[k].y)); }
          }
        }
      }
      const double2 b = reinterpret_cast<const double2*>(base)[i];
      double2 o;
      o.x = SIGN < 0 ? __dsub_rn(b.x, s.x) : __dadd_rn(b.x, s.x);
      o.y = SIGN < 0 ? __dsub_rn(b.y, s.y) : __dadd_rn(b.y, s.y);
      reinterpret_cast<double2*>(out)[i] = o;
      acc[0] += o.x * o.x + o.y * o.y;
    }
    if ((n & 1) && gtid == 0) {
      const int64_t i = n - 1;
      double s = __dmul_rn(coef[0], Q[i]);
      for (int k = 1; k < nc; ++k) s = __dadd_rn(s, __dmul_rn(coef[k], Q[(int64_t)k * ld + i]));
      const double o = SIGN < 0 ? __dsub_rn(base[i], s) : __dadd_rn(base[i], s);
      out[i] = o;
      acc[0] += o * o;
    }
  }
  if (SIGN < 0) {
    block_sum<1>(acc, red);
    grid_reduce_finish<1>(acc, partials, counter, scal, fin, 0, red, pa);
  }
}

// Device-resident bookkeeping of one Arnoldi run / GMRES cycle (one allocation).
//   H   (kn+1) x kn column-major: the Hessenberg matrix (arnoldi) or the rotated upper-triangular factor (GMRES)
//   cs, sn, g, y: Givens rotations, rotated right-hand side, solution of the least-squares problem (GMRES)
//   meta[0] = first column count at which the run has to stop (-1: none), meta[1] = columns used by the last solve
//   res[0] = |g_{j+1}| at the stop column (GMRES residual estimate)
struct arn_dev {
  double *H, *cs, *sn, *g, *y, *res;
  int* meta;
  int ldh;
  void* block;
};

static sla_status arn_dev_alloc(sla_ctx* c, int kn, arn_dev* d) {
  const size_t nd = (size_t)(kn + 1) * kn + 4 * (size_t)(kn + 2) + 2;
  const size_t bytes = nd * sizeof(double) + 4 * sizeof(int);
  memset(d, 0, sizeof(*d));
  if (cudaMalloc(&d->block, bytes) != cudaSuccess) { cudaGetLastError(); return sla_fail(c, SLA_ERR_ALLOC, "cudaMalloc failed for the Hessenberg block"); }
  SLA_CUDA(c, cudaMemsetAsync(d->block, 0, bytes, c->stream));
  double* p = (double*)d->block;
  d->ldh = kn + 1;
  d->H = p; p += (size_t)(kn + 1) * kn;
  d->cs = p; p += kn + 2; d->sn = p; p += kn + 2; d->g = p; p += kn + 2; d->y = p; p += kn + 2; d->res = p; p += 2;
  d->meta = (int*)p;
  return SLA_OK;
}

__global__ void arn_reset_kernel(int* meta, double* res) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { meta[0] = -1; meta[1] = 0; res[0] = 0.0; }
}

// q_{j+1} = (recip h_{j+1,j}) .* w, and thread 0 of CTA 0 files column j:
//   GM = false (arnoldi): H[:, j] = (h_0 .. h_j, ||w||); breakdown = nearZero ||w|| for j > 0 (arnInit has no test)  Sparse.hs:643-666
//   GM = true  (GMRES):   h = first-pass + re-orthogonalisation dots, previous rotations applied, new rotation formed,
//                         g updated; the run stops at the first column whose residual estimate meets tol (or on breakdown)
template <bool GM>
__global__ void __launch_bounds__(EW_THREADS)
arn_finish_kernel(const double* __restrict__ w, double* __restrict__ qnext, int64_t n, const double* __restrict__ scal, int j,
                  int reorth, arn_dev d, double tol, double g0) {
  const double a = scal[S_INVN];
  const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * blockDim.x, gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = gtid; i < n2; i += stride) {
    const double2 v = reinterpret_cast<const double2*>(w)[i];
    reinterpret_cast<double2*>(qnext)[i] = make_double2(__dmul_rn(a, v.x), __dmul_rn(a, v.y));
  }
  if (gtid == 0) {
    if (n & 1) qnext[n - 1] = __dmul_rn(a, w[n - 1]);
    double* hc = d.H + (size_t)j * d.ldh;
    const double nrm = scal[S_NRM];
    if (!GM) {
      for (int k = 0; k <= j; ++k) hc[k] = scal[S_HCOL + k];
      hc[j + 1] = nrm;
      if (j > 0 && fabs(nrm) <= 1e-12 && d.meta[0] < 0) d.meta[0] = j + 1;
    } else {
      if (j == 0) d.g[0] = g0;
      double prev = scal[S_HCOL] + (reorth ? scal[S_HCOL2] : 0.0);
      for (int k = 0; k < j; ++k) {                      // apply the previous rotations to the new column
        const double hn = scal[S_HCOL + k + 1] + (reorth ? scal[S_HCOL2 + k + 1] : 0.0);
        hc[k] = d.cs[k] * prev + d.sn[k] * hn;
        prev = -d.sn[k] * prev + d.cs[k] * hn;
      }
      const double bb = nrm, dd = hypot(prev, bb);
      const double cj = dd == 0.0 ? 1.0 : prev / dd, sj = dd == 0.0 ? 0.0 : bb / dd;
      d.cs[j] = cj; d.sn[j] = sj;
      hc[j] = cj * prev + sj * bb;
      const double gj = d.g[j];
      d.g[j + 1] = -sj * gj; d.g[j] = cj * gj;
      if (d.meta[0] < 0) {
        const double r = fabs(d.g[j + 1]);
        d.res[0] = r;
        if (r <= tol || fabs(bb) <= 1e-12) d.meta[0] = j + 1;
      }
    }
  }
}

// GMRES: back-substitution R y = g over the first jn columns (jn = the stop column, or jmax when the cycle ran to its end);
// y goes to scal[S_HCOL ..] where lincomb_kernel<+1> reads it, jn to meta[1]
__global__ void gmres_solve_kernel(arn_dev d, int jmax, double* scal) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int jn = d.meta[0] >= 0 && d.meta[0] < jmax ? d.meta[0] : jmax;
  for (int k = jn - 1; k >= 0; --k) {
    double t = d.g[k];
    for (int l = k + 1; l < jn; ++l) t -= d.H[(size_t)l * d.ldh + k] * d.y[l];
    d.y[k] = t / d.H[(size_t)k * d.ldh + k];
  }
  for (int k = 0; k < jn; ++k) scal[S_HCOL + k] = d.y[k];
  d.meta[1] = jn;
}

static unsigned ts_blocks(int64_t n) {
  int64_t b = ((n >> 1) + EW_THREADS - 1) / EW_THREADS;
  if (b < 1) b = 1;
  if (b > EW_MAX_BLOCKS) b = EW_MAX_BLOCKS;
  return (unsigned)b;
}

// grid of the tall-skinny dot kernel: whole waves of resident CTAs
static unsigned tsmv_grid(int64_t n) {
  static int per_sm = 0;
  if (!per_sm) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tsmv_t_kernel, EW_THREADS, 0) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
  }
  int64_t want = ((n >> 1) + 32 * TS_UNROLL - 1) / (32 * TS_UNROLL);
  if (want < 1) want = 1;
  const int64_t cap = (int64_t)SLA_NUM_SMS * per_sm;
  return (unsigned)(want < cap ? want : cap);
}

static sla_status tsmv_launch(sla_ctx* c, const sla_dense* Q, int nq, const double* w, int slot0) {
  const int64_t n = Q->rows, ld = Q->ld;
  for (int k0 = 0; k0 < nq; k0 += TS_NC) {
    const int nc = nq - k0 < TS_NC ? nq - k0 : TS_NC;
    const sla_red_plan rp = sla_red_begin(c, FIN_STORE, TS_NC);      // the kernel reduces all TS_NC slots (the unused ones are 0)
    tsmv_t_kernel<<<tsmv_grid(n), EW_THREADS, 0, c->stream>>>(Q->d, ld, n, k0, nc, w, c->scal, c->partials, c->counter, rp.fin, slot0, rp.pa);
    SLA_LAUNCH_CHECK(c);
    SLA_TRY(sla_red_end(c, rp, TS_NC, FIN_STORE, slot0 + k0));
  }
  return SLA_OK;
}

static sla_status dense_alloc(sla_ctx* c, int64_t rows, int64_t cols, sla_dense** out) {
  sla_dense* d = new (std::nothrow) sla_dense();
  if (!d) return sla_fail(c, SLA_ERR_ALLOC, "dense alloc");
  d->ctx = c; d->rows = rows; d->cols = cols; d->ld = (rows + 15) & ~(int64_t)15; d->d = nullptr;
  d->dtype = SLA_F64; d->rowmajor = 0;
  if (d->ld == 0) d->ld = 16;
  d->bytes = sizeof(double) * (size_t)d->ld * (size_t)(cols > 0 ? cols : 1);
  if (sla_pool_alloc(c, (void**)&d->d, d->bytes) != cudaSuccess) {
    cudaGetLastError(); delete d;
    return sla_fail(c, SLA_ERR_ALLOC, "cudaMalloc failed for a dense block");
  }
  *out = d;
  return SLA_OK;
}

// One Arnoldi step, queued on the stream: given basis columns 0..j of Q, append column j+1 and file column j of H.
//   aqi = aa #> q_j ; h_k = q_k <.> aqi (k <= j) ; w = aqi - sum_k h_k q_k ; h_{j+1} = norm2 w ; q_{j+1} = w ./ h_{j+1}
// reorth = false is the reference's single classical Gram-Schmidt pass (Sparse.hs:655-659); reorth = true repeats the
// projection once (CGS2) — used by GMRES only, because single-pass CGS loses orthogonality on clustered spectra.
template <bool GM>
static sla_status arnoldi_step(sla_ctx* c, const sla_csr* A, sla_dense* Q, int j, double* w, bool reorth, const arn_dev& d, double tol, double g0) {
  SLA_GUARD(c);
  const int64_t n = Q->rows, ld = Q->ld;
  SLA_TRY(sla_spmv_launch(c, A, Q->d + (int64_t)j * ld, w, EPI_NONE, nullptr, nullptr, FIN_STORE, S_TMP0));
  const int nq = j + 1;
  for (int pass = 0; pass < (reorth ? 2 : 1); ++pass) {
    const int slot0 = pass == 0 ? S_HCOL : S_HCOL2;
    SLA_TRY(tsmv_launch(c, Q, nq, w, slot0));
    const sla_red_plan rp = sla_red_begin(c, FIN_NORM_INV, 1);
    lincomb_kernel<-1><<<ts_blocks(n), EW_THREADS, 0, c->stream>>>(Q->d, ld, n, nq, w, w, c->scal, c->partials, c->counter, rp.fin, slot0, nullptr, rp.pa);
    SLA_LAUNCH_CHECK(c);
    SLA_TRY(sla_red_end(c, rp, 1, FIN_NORM_INV, 0));
  }
  arn_finish_kernel<GM><<<ts_blocks(n), EW_THREADS, 0, c->stream>>>(w, Q->d + (int64_t)(j + 1) * ld, n, c->scal, j, reorth ? 1 : 0, d, tol, g0);
  SLA_LAUNCH_CHECK(c);
  return SLA_OK;
}

extern "C" sla_status sla_arnoldi(sla_ctx* c, const sla_csr* A, const sla_vec* b, int kn, sla_dense** Qout, double* h_host, int* nmax_out) {
  if (!c || !A || !b || !Qout || !h_host || !nmax_out) return SLA_ERR_INVALID;
  *Qout = nullptr; *nmax_out = 0;
  if (csr_xdim(A) != b->n) {               // | n == nb ... | otherwise = throwM (MatVecSizeMismatchException "arnoldi" (m,n) nb)   :636-637
    snprintf(c->err, sizeof(c->err), "arnoldi : Matrix-vector dimensions are incompatible: Matrix is (%lld,%lld) , whereas vector is %lld",
             (long long)A->m, (long long)A->n, (long long)b->n);
    return SLA_ERR_SIZE_MISMATCH;
  }
  if (A->m != csr_xdim(A)) return sla_fail(c, SLA_ERR_SIZE_MISMATCH, "arnoldi : matrix must be square");
  if (kn < 2 || kn > SLA_MAX_KRYLOV)       // kn <= 1 never meets `i == kn` and runs to breakdown in the reference; not supported here
    return sla_fail(c, SLA_ERR_INVALID, "arnoldi : kn must be in [2, 384]");
  const int64_t n = A->m;
  sla_dense* Q = nullptr;
  SLA_TRY(dense_alloc(c, n, kn + 1, &Q));
  sla_vec* w = nullptr;
  arn_dev d;
  sla_status s = sla_vec_alloc(c, n, &w);
  if (s == SLA_OK) s = arn_dev_alloc(c, kn, &d); else memset(&d, 0, sizeof(d));
  if (s != SLA_OK) { sla_vec_free(w); sla_dense_free(Q); return s; }
  arn_reset_kernel<<<1, 32, 0, c->stream>>>(d.meta, d.res);
  c->launches++;
  // q0 = normalize2 b   :643
  {
    Ptrs<1> in{{b->d}}; Ptrs<0> o0{};
    s = ew_launch(c, OpNorm2Sq{}, n, in, o0, FIN_NORM_INV, 0);
    Ptrs<1> o{{Q->d}}; OpScaleDev op; op.slot = S_INVN; op.a = 0;
    if (s == SLA_OK) s = ew_launch(c, op, n, in, o);
  }
  // arnInit is the j = 0 step; modifyUntil then applies the step and tests `i == kn || breakdown`   :639-651.
  // All kn steps are queued without a host round trip; the breakdown column (if any) is found afterwards and the columns
  // computed past it are discarded — the same (Q, H) the reference returns when it stops there.
  for (int j = 0; j < kn && s == SLA_OK; ++j) s = arnoldi_step<false>(c, A, Q, j, w->d, false, d, 0.0, 0.0);
  std::vector<double> hfull((size_t)(kn + 1) * kn, 0.0);
  int meta[4] = {-1, 0, 0, 0};
  if (s == SLA_OK && cudaMemcpyAsync(hfull.data(), d.H, sizeof(double) * hfull.size(), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) s = sla_fail(c, SLA_ERR_CUDA, "arnoldi: copy of H failed");
  if (s == SLA_OK && cudaMemcpyAsync(meta, d.meta, sizeof(meta), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) s = sla_fail(c, SLA_ERR_CUDA, "arnoldi: copy of the breakdown flag failed");
  if (s == SLA_OK) s = sla_sync(c);
  sla_vec_free(w);
  cudaFree(d.block);
  if (s != SLA_OK) { cudaGetLastError(); sla_dense_free(Q); return s; }
  const bool brk = meta[0] >= 0 && meta[0] <= kn;
  const int i = brk ? meta[0] : kn;          // columns of H produced before the run stops
  // H is (i+1) x i column-major; on breakdown the reference returns the leading block with nmax = i
  for (int col = 0; col < i; ++col)
    for (int row = 0; row <= i; ++row) h_host[(int64_t)col * (i + 1) + row] = row <= col + 1 ? hfull[(size_t)col * (kn + 1) + row] : 0.0;
  for (int64_t z = (int64_t)i * (i + 1); z < (int64_t)(kn + 1) * kn; ++z) h_host[z] = 0.0;
  if (i < kn) Q->cols = i + 1;
  *nmax_out = i;
  *Qout = Q;
  return brk ? SLA_ERR_BREAKDOWN : SLA_OK;
}

// ---- GMRES(restart) --------------------------------------------------------------------------------------
// Restarted GMRES on the Arnoldi kernels above: Gram-Schmidt basis with one re-orthogonalisation pass, Givens rotations
// and the back-substitution of the (restart+1) x restart least-squares problem on the device (arn_finish_kernel<true>,
// gmres_solve_kernel), x += Q y on the device; the host synchronises twice per CYCLE (restart residual, columns used),
// never per step.  A cycle is queued to its end; when the residual estimate meets the tolerance at column j the later
// columns of that (last) cycle are computed but not used.  The reference's own gmres (arnoldi -> qr -> triUpperSolve ->
// Q y) is commented out (Sparse.hs:837-848); tolerance and iteration-cap conventions follow linSolve0 (Sparse.hs:1034-1037).
// opts->check_every < 0: no stopping test (every cycle runs `restart` steps until max_iters; breakdown still stops a cycle)
// — the fixed-work form "GMRES(30), 10 restarts" of BASELINE.json config 4.
extern "C" sla_status sla_gmres(sla_ctx* c, const sla_csr* A, const sla_vec* b, const sla_vec* x0, int restart,
                                const sla_solve_opts* opts_in, sla_vec* x, int* iters, double* resnorm) {
  if (!c || !A || !b || !x0 || !x) return SLA_ERR_INVALID;
  SLA_TRY(check_system(c, "gmres", A, b, x0));
  if (restart < 1 || restart > SLA_MAX_KRYLOV) return sla_fail(c, SLA_ERR_INVALID, "gmres : restart must be in [1, 384]");
  sla_solve_opts o;
  sla_solve_opts_default(&o);
  if (opts_in) o = *opts_in;
  if (o.max_iters <= 0) o.max_iters = 200;
  const bool never_stop = o.check_every < 0;
  const int64_t n = A->m;
  const int m = restart;
  sla_dense* Q = nullptr;
  sla_vec* w = nullptr;
  arn_dev d;
  memset(&d, 0, sizeof(d));
  SLA_TRY(dense_alloc(c, n, m + 1, &Q));
  sla_status s = sla_vec_alloc(c, n, &w);
  if (s == SLA_OK) s = arn_dev_alloc(c, m, &d);
  if (s == SLA_OK && x != x0) s = sla_vec_copy(c, x0, x);
  int total = 0;
  double res = 0, tol = 0;
  bool first = true, done = false;
  while (s == SLA_OK && !done) {
    // r = b - A x ; beta = ||r|| ; q0 = r / beta
    s = sla_spmv_launch(c, A, x->d, w->d, EPI_NONE, nullptr, nullptr, FIN_STORE, S_TMP0);
    if (s != SLA_OK) break;
    { Ptrs<2> in{{b->d, w->d}}; Ptrs<3> out{{w->d, w->d, w->d}}; s = ew_launch(c, OpResidInit{}, n, in, out, FIN_NORM_INV, 0); }
    if (s != SLA_OK) break;
    { Ptrs<1> in{{w->d}}; Ptrs<1> out{{Q->d}}; OpScaleDev op; op.slot = S_INVN; op.a = 0; s = ew_launch(c, op, n, in, out); }
    if (s != SLA_OK) break;
    double beta = 0;
    s = sla_read_scalars(c, S_NRM, 1, &beta);
    if (s != SLA_OK) break;
    res = beta;
    if (first) { tol = never_stop ? -1.0 : (o.tol_abs > o.tol_rel * beta ? o.tol_abs : o.tol_rel * beta); first = false; }
    if (!(beta > tol) || total >= o.max_iters) break;     // converged on the true residual (or NaN / cap)
    const int jmax = m < o.max_iters - total ? m : o.max_iters - total;
    arn_reset_kernel<<<1, 32, 0, c->stream>>>(d.meta, d.res);
    c->launches++;
    for (int j = 0; j < jmax && s == SLA_OK; ++j) s = arnoldi_step<true>(c, A, Q, j, w->d, true, d, tol, beta);
    if (s != SLA_OK) break;
    // R y = g on the device, then x = x + Q[:, 0..jn-1] y
    gmres_solve_kernel<<<1, 32, 0, c->stream>>>(d, jmax, c->scal);
    lincomb_kernel<1><<<ts_blocks(n), EW_THREADS, 0, c->stream>>>(Q->d, Q->ld, n, jmax, x->d, x->d, c->scal, c->partials, c->counter, FIN_STORE, S_HCOL, d.meta + 1, sla_red_begin(c, FIN_STORE, P2P_MAX_NV + 1).pa);
    c->launches += 2;
    int meta[2] = {-1, 0};
    double rest = 0;
    if (cudaMemcpyAsync(meta, d.meta, sizeof(meta), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaMemcpyAsync(&rest, d.res, sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
      s = sla_fail(c, SLA_ERR_CUDA, "gmres: CUDA error in the solution update");
      break;
    }
    total += meta[1];
    res = rest;
    if (total >= o.max_iters) {
      // report the true residual of the returned iterate
      s = residual_norm(c, A, x, b, &res);
      done = true;
    }
  }
  sla_touch(x);
  if (iters) *iters = total;
  if (resnorm) *resnorm = res;
  sla_vec_free(w); sla_dense_free(Q); cudaFree(d.block);
  return s;
}
