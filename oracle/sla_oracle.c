/* sla_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See sla_oracle.h.
 *
 * Every function cites the reference lines (under /root/reference/) it restates.
 * Build with -ffp-contract=off: GHC does not contract a*b+c into an FMA on x86-64.
 */
#include "sla_oracle.h"
#include "../include/sla_synth.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ containers */

static ora_sv* sv_alloc(int64_t dim, int64_t cap) {
  ora_sv* v = (ora_sv*)malloc(sizeof(ora_sv));
  v->dim = dim; v->nnz = 0; v->cap = cap > 0 ? cap : 0;
  v->idx = v->cap ? (int64_t*)malloc(sizeof(int64_t) * (size_t)v->cap) : NULL;
  v->val = v->cap ? (double*)malloc(sizeof(double) * (size_t)v->cap) : NULL;
  return v;
}

static void sv_reserve(ora_sv* v, int64_t cap) {
  if (cap <= v->cap) return;
  int64_t nc = v->cap ? v->cap * 2 : 4;
  if (nc < cap) nc = cap;
  v->idx = (int64_t*)realloc(v->idx, sizeof(int64_t) * (size_t)nc);
  v->val = (double*)realloc(v->val, sizeof(double) * (size_t)nc);
  v->cap = nc;
}

/* position of key k in the sorted key array, or the insertion point; *found says which */
static int64_t sv_find(const ora_sv* v, int64_t k, int* found) {
  int64_t lo = 0, hi = v->nnz;
  while (lo < hi) {
    int64_t mid = lo + (hi - lo) / 2;
    if (v->idx[mid] < k) lo = mid + 1; else hi = mid;
  }
  *found = (lo < v->nnz && v->idx[lo] == k);
  return lo;
}

/* IntMap.insert: replaces the value if the key is present (IntM.hs insert) */
static void sv_insert(ora_sv* v, int64_t k, double x) {
  int found;
  /* fast path: appending in ascending order */
  if (v->nnz == 0 || v->idx[v->nnz - 1] < k) {
    sv_reserve(v, v->nnz + 1);
    v->idx[v->nnz] = k; v->val[v->nnz] = x; v->nnz++;
    return;
  }
  int64_t p = sv_find(v, k, &found);
  if (found) { v->val[p] = x; return; }
  sv_reserve(v, v->nnz + 1);
  memmove(v->idx + p + 1, v->idx + p, sizeof(int64_t) * (size_t)(v->nnz - p));
  memmove(v->val + p + 1, v->val + p, sizeof(double) * (size_t)(v->nnz - p));
  v->idx[p] = k; v->val[p] = x; v->nnz++;
}

ora_sv* ora_sv_zero(int64_t dim) { return sv_alloc(dim, 0); }          /* SpVector.hs:157-158 */

/* mkSpVR d ll = SV d (mkIm ll); fromListDenseSV d ll = SV d (fromList $ indexed (take d ll))
 * SpVector.hs:183-195.  Every listed entry is stored, zeros included. */
ora_sv* ora_sv_from_dense(int64_t dim, const double* x, int64_t len) {
  int64_t n = len < dim ? len : dim;
  ora_sv* v = sv_alloc(dim, n);
  for (int64_t i = 0; i < n; ++i) { v->idx[i] = i; v->val[i] = x[i]; }
  v->nnz = n;
  return v;
}

/* fromListSV d iix = SV d $ foldr insf empty iix, insf (i,x) acc | inBounds0 d i = insert i x acc
 *                                                                 | otherwise = acc
 * SpVector.hs:275-278.  foldr inserts the LAST list element first, so for duplicate
 * indices the FIRST occurrence in the list wins; out-of-bounds entries are dropped. */
ora_sv* ora_sv_from_list(int64_t dim, int64_t n, const int64_t* idx, const double* val) {
  ora_sv* v = sv_alloc(dim, n);
  for (int64_t q = n - 1; q >= 0; --q)
    if (idx[q] >= 0 && idx[q] < dim) sv_insert(v, idx[q], val[q]);   /* inBounds0, Utils.hs:103-107 */
  return v;
}

ora_sv* ora_sv_copy(const ora_sv* s) {
  ora_sv* v = sv_alloc(s->dim, s->nnz);
  if (s->nnz) {
    memcpy(v->idx, s->idx, sizeof(int64_t) * (size_t)s->nnz);
    memcpy(v->val, s->val, sizeof(double) * (size_t)s->nnz);
  }
  v->nnz = s->nnz;
  return v;
}

void ora_sv_free(ora_sv* v) { if (!v) return; free(v->idx); free(v->val); free(v); }
int64_t ora_sv_dim(const ora_sv* v) { return v->dim; }
int64_t ora_sv_nnz(const ora_sv* v) { return v->nnz; }

/* toDenseListSV: findWithDefault 0 i, i in [0 .. d-1]   SpVector.hs:300-301 */
void ora_sv_to_dense(const ora_sv* v, double* out) {
  for (int64_t i = 0; i < v->dim; ++i) out[i] = 0.0;
  for (int64_t q = 0; q < v->nnz; ++q)
    if (v->idx[q] >= 0 && v->idx[q] < v->dim) out[v->idx[q]] = v->val[q];
}

void ora_sv_to_list(const ora_sv* v, int64_t* idx, double* val) {
  for (int64_t q = 0; q < v->nnz; ++q) { idx[q] = v->idx[q]; val[q] = v->val[q]; }
}

/* (^+^) = liftU2 (^+^): IntMap.unionWith (+); result dim = max n1 n2 — no dimension check.
 * SpVector.hs:62-63, 107-109; IntM.hs:79. */
ora_sv* ora_sv_add(const ora_sv* a, const ora_sv* b) {
  ora_sv* v = sv_alloc(a->dim > b->dim ? a->dim : b->dim, a->nnz + b->nnz);
  int64_t i = 0, j = 0, o = 0;
  while (i < a->nnz && j < b->nnz) {
    if (a->idx[i] < b->idx[j])      { v->idx[o] = a->idx[i]; v->val[o] = a->val[i]; ++i; }
    else if (a->idx[i] > b->idx[j]) { v->idx[o] = b->idx[j]; v->val[o] = b->val[j]; ++j; }
    else { v->idx[o] = a->idx[i]; v->val[o] = a->val[i] + b->val[j]; ++i; ++j; }
    ++o;
  }
  for (; i < a->nnz; ++i, ++o) { v->idx[o] = a->idx[i]; v->val[o] = a->val[i]; }
  for (; j < b->nnz; ++j, ++o) { v->idx[o] = b->idx[j]; v->val[o] = b->val[j]; }
  v->nnz = o;
  return v;
}

/* negateV v = fmap negateV v   SpVector.hs:110 */
ora_sv* ora_sv_negate(const ora_sv* s) {
  ora_sv* v = ora_sv_copy(s);
  for (int64_t q = 0; q < v->nnz; ++q) v->val[q] = -v->val[q];
  return v;
}

/* x ^-^ y = x ^+^ negateV y   (class default)   Class.hs:68-69 */
ora_sv* ora_sv_sub(const ora_sv* a, const ora_sv* b) {
  ora_sv* nb = ora_sv_negate(b);
  ora_sv* v = ora_sv_add(a, nb);
  ora_sv_free(nb);
  return v;
}

/* n .* v = fmap (n .*) v, (.*) = (*) on Double: scalar on the LEFT.  SpVector.hs:112-114, Class.hs:377-380 */
ora_sv* ora_sv_scale(double a, const ora_sv* s) {
  ora_sv* v = ora_sv_copy(s);
  for (int64_t q = 0; q < v->nnz; ++q) v->val[q] = a * v->val[q];
  return v;
}

/* v ./ s = (recip s) .* v   Class.hs:94-95 */
ora_sv* ora_sv_divs(const ora_sv* v, double s) { return ora_sv_scale(1.0 / s, v); }

/* Ascending-key intersection visitor shared by <.>, dotu and dott: calls the products in key
 * order.  IntMap.intersectionWith (IntM.hs:80) followed by the default Foldable `sum` — a strict
 * left fold from 0 in ascending key order.  `first` is the operand whose value is the LEFT
 * factor of (*). */
static double sv_dot_ordered(const ora_sv* first, const ora_sv* second) {
  double acc = 0.0;
  /* lopsided sizes: binary-search the small operand's keys in the large one (same visit order) */
  if (first->nnz * 16 < second->nnz || second->nnz * 16 < first->nnz) {
    const ora_sv* s = first->nnz < second->nnz ? first : second;
    const ora_sv* l = first->nnz < second->nnz ? second : first;
    int dense_keys = (l->nnz > 0 && l->idx[0] == 0 && l->idx[l->nnz - 1] == l->nnz - 1);
    for (int64_t q = 0; q < s->nnz; ++q) {
      int64_t k = s->idx[q], p; int found;
      if (dense_keys) { p = k; found = (k >= 0 && k < l->nnz); }
      else p = sv_find(l, k, &found);
      if (!found) continue;
      double fa = (s == first) ? s->val[q] : l->val[p];
      double fb = (s == first) ? l->val[p] : s->val[q];
      acc = acc + fa * fb;
    }
    return acc;
  }
  int64_t i = 0, j = 0;
  while (i < first->nnz && j < second->nnz) {
    if (first->idx[i] < second->idx[j]) ++i;
    else if (first->idx[i] > second->idx[j]) ++j;
    else { acc = acc + first->val[i] * second->val[j]; ++i; ++j; }
  }
  return acc;
}

/* v <.> w = sum $ liftI2 (<.>) v w ; Double: (<.>) = (*)   SpVector.hs:116-117, Class.hs:400 */
double ora_sv_dot(const ora_sv* v, const ora_sv* w) { return sv_dot_ordered(v, w); }

/* norm2Sq = sum . fmap norm2Sq ; Double: norm2Sq = (**2) = C pow(x, 2)   SpVector.hs:122, Class.hs:405-407 */
double ora_sv_norm2sq(const ora_sv* v) {
  double acc = 0.0;
  for (int64_t q = 0; q < v->nnz; ++q) acc = acc + pow(v->val[q], 2.0);
  return acc;
}

/* norm2 c = sqrt (norm2Sq c) ; norm2' likewise   SpVector.hs:127-128 */
double ora_sv_norm2(const ora_sv* v) { return sqrt(ora_sv_norm2sq(v)); }

/* normalize2 v = v ./ norm2 v   SpVector.hs:125 */
ora_sv* ora_sv_normalize2(const ora_sv* v) { return ora_sv_divs(v, ora_sv_norm2(v)); }

/* nearZero a = abs a <= 1e-12   Eps.hs:41-42 */
int ora_near_zero(double a) { return fabs(a) <= 1e-12; }

/* ------------------------------------------------------------------ SpMatrix */

ora_sm* ora_sm_zero(int64_t m, int64_t n) {
  ora_sm* a = (ora_sm*)malloc(sizeof(ora_sm));
  a->nrows = m; a->ncols = n; a->nstored = 0; a->cap = 0; a->rkey = NULL; a->row = NULL;
  return a;
}

static int64_t sm_find_row(const ora_sm* a, int64_t k, int* found) {
  int64_t lo = 0, hi = a->nstored;
  while (lo < hi) {
    int64_t mid = lo + (hi - lo) / 2;
    if (a->rkey[mid] < k) lo = mid + 1; else hi = mid;
  }
  *found = (lo < a->nstored && a->rkey[lo] == k);
  return lo;
}

/* outer-map row lookup, creating the row (IntMap.alter, IntMap2.hs:24-29) */
static ora_sv* sm_row_for_insert(ora_sm* a, int64_t i) {
  int found; int64_t p;
  if (a->nstored == 0 || a->rkey[a->nstored - 1] < i) { p = a->nstored; found = 0; }
  else p = sm_find_row(a, i, &found);
  if (found) return a->row[p];
  if (a->nstored + 1 > a->cap) {
    int64_t nc = a->cap ? a->cap * 2 : 4;
    a->rkey = (int64_t*)realloc(a->rkey, sizeof(int64_t) * (size_t)nc);
    a->row = (ora_sv**)realloc(a->row, sizeof(ora_sv*) * (size_t)nc);
    a->cap = nc;
  }
  memmove(a->rkey + p + 1, a->rkey + p, sizeof(int64_t) * (size_t)(a->nstored - p));
  memmove(a->row + p + 1, a->row + p, sizeof(ora_sv*) * (size_t)(a->nstored - p));
  a->rkey[p] = i; a->row[p] = sv_alloc(a->ncols, 0); a->nstored++;
  return a->row[p];
}

/* insertIM2 i j x: alter the outer map at i, insert j x in the inner map (replacing).  IntMap2.hs:24-29 */
static void sm_insert(ora_sm* a, int64_t i, int64_t j, double x) { sv_insert(sm_row_for_insert(a, i), j, x); }

/* fromListSM (m,n) iix = foldl' ins (zeroSM m n) iix ; ins t (i,j,x) = insertSpMatrix i j x t ;
 * insertSpMatrix errors when not inBounds02.  Left fold => LATER duplicates overwrite.
 * SpMatrix.hs:205-224. */
ora_sm* ora_sm_from_list(int64_t m, int64_t n, int64_t nnz, const int64_t* i, const int64_t* j,
                         const double* v, int* err) {
  if (err) *err = ORA_OK;
  ora_sm* a = ora_sm_zero(m, n);
  for (int64_t q = 0; q < nnz; ++q) {
    if (!(i[q] >= 0 && i[q] < m && j[q] >= 0 && j[q] < n)) {       /* inBounds02, Utils.hs:109-110 */
      if (err) *err = ORA_ERR_OOB_INDEX;                           /* error "insertSpMatrix : index out of bounds" */
      ora_sm_free(a);
      return NULL;
    }
    sm_insert(a, i[q], j[q], v[q]);
  }
  return a;
}

/* fromListDenseSM m ll = fromListSM (m, n) $ indexed2 m ll, n = length ll `div` m : COLUMN-major;
 * zip3 truncates to n*m elements.   SpMatrix.hs:239-241, Utils.hs:85-90 */
ora_sm* ora_sm_from_dense_colmajor(int64_t m, const double* ll, int64_t len) {
  int64_t n = m > 0 ? len / m : 0;
  ora_sm* a = ora_sm_zero(m, n);
  for (int64_t q = 0; q < n * m; ++q) sm_insert(a, q % m, q / m, ll[q]);
  return a;
}

/* marshalling helper (not a reference function): rows with row_ptr[i] == row_ptr[i+1] are not stored */
ora_sm* ora_sm_from_csr(int64_t m, int64_t n, const int64_t* row_ptr, const int64_t* col, const double* val) {
  ora_sm* a = ora_sm_zero(m, n);
  a->cap = m; a->rkey = (int64_t*)malloc(sizeof(int64_t) * (size_t)(m > 0 ? m : 1));
  a->row = (ora_sv**)malloc(sizeof(ora_sv*) * (size_t)(m > 0 ? m : 1));
  for (int64_t i = 0; i < m; ++i) {
    int64_t len = row_ptr[i + 1] - row_ptr[i];
    if (len == 0) continue;
    ora_sv* r = sv_alloc(n, len);
    for (int64_t q = 0; q < len; ++q) sv_insert(r, col[row_ptr[i] + q], val[row_ptr[i] + q]);
    a->rkey[a->nstored] = i; a->row[a->nstored] = r; a->nstored++;
  }
  return a;
}

void ora_sm_free(ora_sm* a) {
  if (!a) return;
  for (int64_t r = 0; r < a->nstored; ++r) ora_sv_free(a->row[r]);
  free(a->rkey); free(a->row); free(a);
}

int64_t ora_sm_nrows(const ora_sm* a) { return a->nrows; }
int64_t ora_sm_ncols(const ora_sm* a) { return a->ncols; }
int64_t ora_sm_nstored_rows(const ora_sm* a) { return a->nstored; }
int64_t ora_sm_nnz(const ora_sm* a) {
  int64_t s = 0;
  for (int64_t r = 0; r < a->nstored; ++r) s += a->row[r]->nnz;
  return s;
}

/* ascending (row, col) traversal — the order of `toList <$> immSM` (NOT toListSM, which conses
 * and therefore returns descending order, SpMatrix.hs:251-253) */
void ora_sm_to_coo(const ora_sm* a, int64_t* i, int64_t* j, double* v) {
  int64_t o = 0;
  for (int64_t r = 0; r < a->nstored; ++r)
    for (int64_t q = 0; q < a->row[r]->nnz; ++q, ++o) {
      i[o] = a->rkey[r]; j[o] = a->row[r]->idx[q]; v[o] = a->row[r]->val[q];
    }
}

void ora_sm_to_csr(const ora_sm* a, int64_t* row_ptr, int64_t* col, double* val) {
  int64_t o = 0, r = 0;
  for (int64_t i = 0; i < a->nrows; ++i) {
    row_ptr[i] = o;
    if (r < a->nstored && a->rkey[r] == i) {
      for (int64_t q = 0; q < a->row[r]->nnz; ++q, ++o) { col[o] = a->row[r]->idx[q]; val[o] = a->row[r]->val[q]; }
      ++r;
    }
  }
  row_ptr[a->nrows] = o;
}

/* transposeIM2 = ifoldlIM2 (flip insertIM2): visit (i, j, x) ascending, insert at (j, i).
 * Explicit zeros are kept; an empty stored row disappears.  IntMap2.hs:71-75, 88-90 */
static ora_sm* sm_transpose_dims(const ora_sm* a, int64_t m, int64_t n) {
  ora_sm* t = ora_sm_zero(m, n);
  for (int64_t r = 0; r < a->nstored; ++r)
    for (int64_t q = 0; q < a->row[r]->nnz; ++q)
      sm_insert(t, a->row[r]->idx[q], a->rkey[r], a->row[r]->val[q]);
  for (int64_t r = 0; r < t->nstored; ++r) t->row[r]->dim = n;
  return t;
}

/* transposeSM (SM (m, n) im) = SM (n, m) (transposeIM2 im)   SpMatrix.hs:717-718 */
ora_sm* ora_sm_transpose(const ora_sm* a) { return sm_transpose_dims(a, a->ncols, a->nrows); }

/* isDiagonalSM m = size d == nrows m, d = rows having exactly one entry, on the diagonal.
 * SpMatrix.hs:411-415 */
int ora_sm_is_diagonal(const ora_sm* a) {
  int64_t cnt = 0;
  for (int64_t r = 0; r < a->nstored; ++r)
    if (a->row[r]->nnz == 1 && a->row[r]->idx[0] == a->rkey[r]) ++cnt;
  return cnt == a->nrows;
}

static ora_sm* sm_copy(const ora_sm* a) {
  ora_sm* c = ora_sm_zero(a->nrows, a->ncols);
  c->cap = a->nstored;
  c->rkey = (int64_t*)malloc(sizeof(int64_t) * (size_t)(a->nstored > 0 ? a->nstored : 1));
  c->row = (ora_sv**)malloc(sizeof(ora_sv*) * (size_t)(a->nstored > 0 ? a->nstored : 1));
  for (int64_t r = 0; r < a->nstored; ++r) { c->rkey[r] = a->rkey[r]; c->row[r] = ora_sv_copy(a->row[r]); }
  c->nstored = a->nstored;
  return c;
}

/* reciprocal = fmap recip   Class.hs:174-175 */
ora_sm* ora_sm_reciprocal(const ora_sm* a) {
  ora_sm* c = sm_copy(a);
  for (int64_t r = 0; r < c->nstored; ++r)
    for (int64_t q = 0; q < c->row[r]->nnz; ++q) c->row[r]->val[q] = 1.0 / c->row[r]->val[q];
  return c;
}

/* sparsifySM = ifilterIM2 (\_ _ x -> isNz x): inner maps are filtered, outer keys stay.
 * SpMatrix.hs:648-654, IntMap2.hs:108-111, Eps.hs:79 */
ora_sm* ora_sm_sparsify(const ora_sm* a) {
  ora_sm* c = sm_copy(a);
  for (int64_t r = 0; r < c->nstored; ++r) {
    ora_sv* row = c->row[r]; int64_t o = 0;
    for (int64_t q = 0; q < row->nnz; ++q)
      if (!ora_near_zero(row->val[q])) { row->idx[o] = row->idx[q]; row->val[o] = row->val[q]; ++o; }
    row->nnz = o;
  }
  return c;
}

/* matVecSD (SM (nr,nc) mdata) (SV n sv) | nc == n = SV nr $ fmap (`dotu` sv) mdata
 * dotu u v = sum $ liftI2 (*) u v : u = the matrix row (left factor), un-conjugated.
 * One output entry per STORED row (an empty stored row yields a stored 0).  Common.hs:242-260 */
ora_sv* ora_sm_matvec(const ora_sm* a, const ora_sv* x, int* err) {
  if (err) *err = ORA_OK;
  if (a->ncols != x->dim) { if (err) *err = ORA_ERR_SIZE_MISMATCH; return NULL; }  /* error "matVec : mismatched dimensions" */
  ora_sv* y = sv_alloc(a->nrows, a->nstored);
  for (int64_t r = 0; r < a->nstored; ++r) { y->idx[r] = a->rkey[r]; y->val[r] = sv_dot_ordered(a->row[r], x); }
  y->nnz = a->nstored;
  return y;
}

/* vecMatSD (SV n sv) (SM (nr,nc) mdata) | n == nr = SV nc $ fmap (`dotu` sv) (transposeIM2 mdata)
 * Common.hs:253-256 */
ora_sv* ora_sm_vecmat(const ora_sv* x, const ora_sm* a, int* err) {
  if (err) *err = ORA_OK;
  if (x->dim != a->nrows) { if (err) *err = ORA_ERR_SIZE_MISMATCH; return NULL; }
  ora_sm* t = ora_sm_transpose(a);
  ora_sv* y = sv_alloc(a->ncols, t->nstored);
  for (int64_t r = 0; r < t->nstored; ++r) { y->idx[r] = t->rkey[r]; y->val[r] = sv_dot_ordered(t->row[r], x); }
  y->nnz = t->nstored;
  ora_sm_free(t);
  return y;
}

/* (##) = matMat_ AB = matMatCheck (matMatUnsafeWith transposeIM2):
 *   SM (nrows m1, ncols m2) (overRows2 <$> immSM m1)
 *   overRows2 vm1 = (`dott` vm1) <$> transposeIM2 (immSM m2) ; dott x y = sum $ liftI2 (*) x y
 * x = a stored column of B (left factor), y = the row of A.  Every (stored row of A, stored
 * column of B) pair gets an entry, explicit zeros included.   SpMatrix.hs:768-811 */
ora_sm* ora_sm_matmat(const ora_sm* a, const ora_sm* b, int* err) {
  if (err) *err = ORA_OK;
  if (a->ncols != b->nrows) { if (err) *err = ORA_ERR_SIZE_MISMATCH; return NULL; } /* error "matMat : incompatible matrix sizes" */
  ora_sm* bt = ora_sm_transpose(b);
  ora_sm* c = ora_sm_zero(a->nrows, b->ncols);
  c->cap = a->nstored;
  c->rkey = (int64_t*)malloc(sizeof(int64_t) * (size_t)(a->nstored > 0 ? a->nstored : 1));
  c->row = (ora_sv**)malloc(sizeof(ora_sv*) * (size_t)(a->nstored > 0 ? a->nstored : 1));
  for (int64_t r = 0; r < a->nstored; ++r) {
    ora_sv* row = sv_alloc(b->ncols, bt->nstored);
    for (int64_t cc = 0; cc < bt->nstored; ++cc) {
      row->idx[cc] = bt->rkey[cc];
      row->val[cc] = sv_dot_ordered(bt->row[cc], a->row[r]);
    }
    row->nnz = bt->nstored;
    c->rkey[r] = a->rkey[r]; c->row[r] = row;
  }
  c->nstored = a->nstored;
  ora_sm_free(bt);
  return c;
}

/* derived Eq on SM dims (IntM (IntM a)): same dims, same keys, same values (0.0 == -0.0, NaN /= NaN) */
int ora_sm_equal(const ora_sm* a, const ora_sm* b) {
  if (a->nrows != b->nrows || a->ncols != b->ncols || a->nstored != b->nstored) return 0;
  for (int64_t r = 0; r < a->nstored; ++r) {
    if (a->rkey[r] != b->rkey[r] || a->row[r]->nnz != b->row[r]->nnz) return 0;
    for (int64_t q = 0; q < a->row[r]->nnz; ++q)
      if (a->row[r]->idx[q] != b->row[r]->idx[q] || !(a->row[r]->val[q] == b->row[r]->val[q])) return 0;
  }
  return 1;
}

/* ------------------------------------------------------------------ preconditioners, triangular solves */

/* filterSM f sm = SM (dim sm) (ifilterIM2 f (dat sm)) ; ifilterIM2 = mapWithKey (\i row -> filterWithKey (f i) row):
 * inner maps are filtered, every outer row key stays (possibly with an empty inner map).
 * extractSubDiag / extractDiag / extractSuperDiag keep i > j / i == j / i < j.
 * SpMatrix.hs:306-315, IntMap2.hs:108-111.   which: -1 sub, 0 diag, +1 super */
ora_sm* ora_sm_extract_tri(const ora_sm* a, int which) {
  ora_sm* c = sm_copy(a);
  for (int64_t r = 0; r < c->nstored; ++r) {
    ora_sv* row = c->row[r]; int64_t i = c->rkey[r], o = 0;
    for (int64_t q = 0; q < row->nnz; ++q) {
      int64_t j = row->idx[q];
      int keep = which < 0 ? (i > j) : which == 0 ? (i == j) : (i < j);
      if (keep) { row->idx[o] = j; row->val[o] = row->val[q]; ++o; }
    }
    row->nnz = o;
  }
  return c;
}

/* eye n = mkDiagonal n (replicate n 1) = fromListSM (n,n) (zip3 [0..] [0..] ones) ; eye 0 = zeroSM 0 0.  SpMatrix.hs:128-135, 184-191 */
ora_sm* ora_sm_eye(int64_t n) {
  ora_sm* a = ora_sm_zero(n, n);
  for (int64_t i = 0; i < n; ++i) sm_insert(a, i, i, 1.0);
  return a;
}

/* scale n = fmap (* n): the scalar is the RIGHT factor.   Class.hs:179-180 */
ora_sm* ora_sm_scale_right(const ora_sm* a, double n) {
  ora_sm* c = sm_copy(a);
  for (int64_t r = 0; r < c->nstored; ++r)
    for (int64_t q = 0; q < c->row[r]->nnz; ++q) c->row[r]->val[q] = c->row[r]->val[q] * n;
  return c;
}

/* negateV = fmap negateV   SpMatrix.hs:79 */
ora_sm* ora_sm_negate(const ora_sm* a) {
  ora_sm* c = sm_copy(a);
  for (int64_t r = 0; r < c->nstored; ++r)
    for (int64_t q = 0; q < c->row[r]->nnz; ++q) c->row[r]->val[q] = -c->row[r]->val[q];
  return c;
}

/* (^+^) = liftU2 (^+^) = SM (maxTup n1 n2) ((liftU2 . liftU2) (+) x1 x2): union of the outer maps,
 * rows present in both are united entry-wise.   SpMatrix.hs:71-78 */
ora_sm* ora_sm_add(const ora_sm* a, const ora_sm* b) {
  ora_sm* c = ora_sm_zero(a->nrows > b->nrows ? a->nrows : b->nrows, a->ncols > b->ncols ? a->ncols : b->ncols);
  int64_t cap = a->nstored + b->nstored;
  c->cap = cap;
  c->rkey = (int64_t*)malloc(sizeof(int64_t) * (size_t)(cap > 0 ? cap : 1));
  c->row = (ora_sv**)malloc(sizeof(ora_sv*) * (size_t)(cap > 0 ? cap : 1));
  int64_t i = 0, j = 0, o = 0;
  while (i < a->nstored || j < b->nstored) {
    if (j >= b->nstored || (i < a->nstored && a->rkey[i] < b->rkey[j])) { c->rkey[o] = a->rkey[i]; c->row[o] = ora_sv_copy(a->row[i]); ++i; }
    else if (i >= a->nstored || a->rkey[i] > b->rkey[j])                { c->rkey[o] = b->rkey[j]; c->row[o] = ora_sv_copy(b->row[j]); ++j; }
    else { c->rkey[o] = a->rkey[i]; c->row[o] = ora_sv_add(a->row[i], b->row[j]); ++i; ++j; }
    ++o;
  }
  c->nstored = o;
  return c;
}

/* x ^-^ y = x ^+^ negateV y   Class.hs:68-69 */
ora_sm* ora_sm_sub(const ora_sm* a, const ora_sm* b) {
  ora_sm* nb = ora_sm_negate(b);
  ora_sm* c = ora_sm_add(a, nb);
  ora_sm_free(nb);
  return c;
}

/* jacobiPre x = recip <$> extractDiag x   Sparse.hs:686-687 */
ora_sm* ora_jacobi_pre(const ora_sm* a) {
  ora_sm* d = ora_sm_extract_tri(a, 0);
  ora_sm* c = ora_sm_reciprocal(d);
  ora_sm_free(d);
  return c;
}

/* mSsorPre aa omega = (l, r) where (e, d, f) = diagPartitions aa ; n = nrows e
 *   l = (eye n ^-^ scale omega e) ## reciprocal d ;  r = d ^-^ scale omega f          Sparse.hs:713-721
 * NB the (##) stores an entry for EVERY (row of the left factor, stored column of reciprocal d) pair — l has n x n
 * stored entries, the zeros explicit.  Meant for small n. */
int ora_mssor_pre(const ora_sm* aa, double omega, ora_sm** l_out, ora_sm** r_out) {
  ora_sm *e = ora_sm_extract_tri(aa, -1), *d = ora_sm_extract_tri(aa, 0), *f = ora_sm_extract_tri(aa, 1);
  ora_sm* eye = ora_sm_eye(e->nrows);
  ora_sm* we = ora_sm_scale_right(e, omega);
  ora_sm* lhs = ora_sm_sub(eye, we);
  ora_sm* rd = ora_sm_reciprocal(d);
  int err = ORA_OK;
  ora_sm* l = ora_sm_matmat(lhs, rd, &err);
  ora_sm* wf = ora_sm_scale_right(f, omega);
  ora_sm* r = ora_sm_sub(d, wf);
  ora_sm_free(e); ora_sm_free(d); ora_sm_free(f); ora_sm_free(eye); ora_sm_free(we); ora_sm_free(lhs); ora_sm_free(rd); ora_sm_free(wf);
  if (err != ORA_OK) { ora_sm_free(r); return err; }
  *l_out = l; *r_out = r;
  return ORA_OK;
}

/* m @@! (i, j): lookup with default 0 (no bounds check)   SpMatrix.hs:280-287 */
static double sm_at(const ora_sm* a, int64_t i, int64_t j) {
  int found; int64_t p = sm_find_row(a, i, &found);
  if (!found) return 0.0;
  int f2; int64_t q = sv_find(a->row[p], j, &f2);
  return f2 ? a->row[p]->val[q] : 0.0;
}

static int sm_has(const ora_sm* a, int64_t i, int64_t j) {
  int found; int64_t p = sm_find_row(a, i, &found);
  if (!found) return 0;
  int f2; sv_find(a->row[p], j, &f2);
  return f2;
}

/* contractSub a b i j n = foldlWithKey' (\acc k x -> if k > n then acc else acc + x * b @@! (k, j)) 0 (row i of a)
 * — the STORED entries of row i of a, ascending k, strict left fold from 0   SpMatrix.hs:857-864 */
static double contract_sub(const ora_sm* a, const ora_sm* b, int64_t i, int64_t j, int64_t n) {
  int found; int64_t p = sm_find_row(a, i, &found);
  double acc = 0.0;
  if (!found) return acc;
  const ora_sv* row = a->row[p];
  for (int64_t q = 0; q < row->nnz; ++q) {
    const int64_t k = row->idx[q];
    if (k > n) continue;
    acc = acc + row->val[q] * sm_at(b, k, j);
  }
  return acc;
}

/* lu aa = (L, U), Doolittle (unit diagonal of L)   Sparse.hs:489-538.
 *   luInit : l0 = insertCol (eye n) (extractSubCol aa 0 (1, n-1) ./ u00) 0 ; u0 = insertRow (zeroSM n n) (extractRow aa 0) 0
 *            (column 0 of L is a_i0 * recip u00 — `./` multiplies by the reciprocal, Class.hs:94-95; stored entries only)
 *   step i = 1 .. n-1: U row i : u_ij = a_ij - contractSub L U i j (i-1), j = i .. n-1, kept when isNz (onRangeSparse)
 *                      L col i : l_ki = (a_ki - contractSub L U k i (k-1)) / u_ii, k = i+1 .. n-1, kept when isNz;
 *                                a nearZero u_ii throws NeedsPivoting as soon as the range is not empty
 * *bad = the pivot index on ORA_ERR_NEEDS_PIVOTING.  O(n^3) element at a time, like the reference: small n only. */
int ora_lu(const ora_sm* aa, ora_sm** l_out, ora_sm** u_out, int64_t* bad) {
  const int64_t n = aa->nrows;
  *l_out = *u_out = NULL;
  if (bad) *bad = -1;
  if (n < 1 || aa->ncols != n) return ORA_ERR_SIZE_MISMATCH;
  ora_sm* l = ora_sm_eye(n);
  ora_sm* u = ora_sm_zero(n, n);
  {
    int found; int64_t p = sm_find_row(aa, 0, &found);
    if (found) for (int64_t q = 0; q < aa->row[p]->nnz; ++q) sm_insert(u, 0, aa->row[p]->idx[q], aa->row[p]->val[q]);   /* extractRow aa 0 */
  }
  const double u00 = sm_at(u, 0, 0);
  if (ora_near_zero(u00)) { ora_sm_free(l); ora_sm_free(u); if (bad) *bad = 0; return ORA_ERR_NEEDS_PIVOTING; }
  {
    const double r00 = 1.0 / u00;                                  /* v ./ s = recip s .* v */
    for (int64_t i = 1; i < n; ++i) if (sm_has(aa, i, 0)) sm_insert(l, i, 0, r00 * sm_at(aa, i, 0));
  }
  for (int64_t ix = 1; ix < n; ++ix) {
    /* uUpd */
    for (int64_t j = ix; j < n; ++j) {
      const double v = sm_at(aa, ix, j) - contract_sub(l, u, ix, j, ix - 1);
      if (!ora_near_zero(v)) sm_insert(u, ix, j, v);
    }
    /* lUpd */
    if (ix + 1 < n) {
      const double ujj = sm_at(u, ix, ix);
      if (ora_near_zero(ujj)) { ora_sm_free(l); ora_sm_free(u); if (bad) *bad = ix; return ORA_ERR_NEEDS_PIVOTING; }
      /* every l_k,ix is computed against L as it was BEFORE this column is inserted (insertCol happens after the list is built) */
      double* col = (double*)malloc(sizeof(double) * (size_t)n);
      for (int64_t k = ix + 1; k < n; ++k) col[k] = (sm_at(aa, k, ix) - contract_sub(l, u, k, ix, k - 1)) / ujj;
      for (int64_t k = ix + 1; k < n; ++k) if (!ora_near_zero(col[k])) sm_insert(l, k, ix, col[k]);
      free(col);
    }
  }
  *l_out = l; *u_out = u;
  return ORA_OK;
}

/* ilu0Pre aa = (sparsifyLU l aa, sparsifyLU u aa) where (l, u) = lu aa and sparsifyLU m m2 keeps the entries of m at the
 * positions STORED in m2 (ifilterSM (\i j _ -> isJust (lookupSM m2 i j)))   Sparse.hs:696-706.
 * NB this is the COMPLETE factorisation with the fill-in dropped afterwards, not the classical ILU(0) recurrence. */
static ora_sm* sm_mask_by(const ora_sm* m, const ora_sm* pat) {
  ora_sm* out = ora_sm_zero(m->nrows, m->ncols);
  for (int64_t p = 0; p < m->nstored; ++p)
    for (int64_t q = 0; q < m->row[p]->nnz; ++q)
      if (sm_has(pat, m->rkey[p], m->row[p]->idx[q])) sm_insert(out, m->rkey[p], m->row[p]->idx[q], m->row[p]->val[q]);
  return out;
}

int ora_ilu0_pre(const ora_sm* aa, ora_sm** l_out, ora_sm** u_out, int64_t* bad) {
  ora_sm *l = NULL, *u = NULL;
  const int err = ora_lu(aa, &l, &u, bad);
  if (err != ORA_OK) return err;
  *l_out = sm_mask_by(l, aa);
  *u_out = sm_mask_by(u, aa);
  ora_sm_free(l); ora_sm_free(u);
  return ORA_OK;
}

/* m @@ (i, j): bounds-checked lookup with default 0; out of bounds is `error "@@ : incompatible indices"`.
 * SpMatrix.hs:108-109, 280-287 */
static int sm_lookup_checked(const ora_sm* a, int64_t i, int64_t j, double* out) {
  if (!(i >= 0 && i < a->nrows && j >= 0 && j < a->ncols)) return 0;
  int found; int64_t p = sm_find_row(a, i, &found);
  *out = 0.0;
  if (found) { int f2; int64_t q = sv_find(a->row[p], j, &f2); if (f2) *out = a->row[p]->val[q]; }
  return 1;
}

static double sv_lookup0(const ora_sv* v, int64_t i) {
  int found; int64_t q = sv_find(v, i, &found);
  return found ? v->val[q] : 0.0;
}

/* extractSubRow m i (j1, j2) `dot` (the matching part of ww): stored entries of row i with j1 <= j <= j2, ascending,
 * times the entries of the partial solution, which holds EVERY index solved so far (insertSpVector stores zeros too).
 * Left factor = the matrix entry.   Common.hs:208-216, SpVector.hs:116-117, 350-353 */
static double subrow_dot(const ora_sm* a, int64_t i, int64_t j1, int64_t j2, const double* ww, int64_t nww) {
  int found; int64_t p = sm_find_row(a, i, &found);
  double acc = 0.0;
  if (!found) return acc;
  const ora_sv* row = a->row[p];
  for (int64_t q = 0; q < row->nnz; ++q) {
    int64_t j = row->idx[q];
    if (j >= j1 && j <= j2 && j < nww) acc = acc + row->val[q] * ww[j];   /* insertSpVector drops out-of-bounds keys */
  }
  return acc;
}

/* sparsifySV = filterSV isNz   SpVector.hs:390-391 */
static ora_sv* sv_from_dense_sparsified(int64_t n, const double* w) {
  ora_sv* v = sv_alloc(n, n);
  int64_t o = 0;
  for (int64_t i = 0; i < n; ++i)
    if (!ora_near_zero(w[i])) { v->idx[o] = i; v->val[o] = w[i]; ++o; }
  v->nnz = o;
  return v;
}

/* triLowerSolve ll b (Sparse.hs:750-777):
 *   lInit: w0 = b0 / l00 (NeedsPivoting when l00 is nearZero) ; state (w, 1)
 *   lStep (ww, i): lii = ll @@ (i,i) ; wi = (b @@ i - r) / lii, r = extractSubRow ll i (0, i-1) `dot` takeSV i ww
 *   modifyUntilM' steps FIRST and tests q (_, i) = i == svDim b afterwards (Iterative.hs:272-282): with dim b = 1 the
 *   first step looks up ll @@ (1,1), which is `error "@@ : incompatible indices"` -> ORA_ERR_OOB_INDEX here.
 *   result = sparsifySV w.   Entries of ll on or above the diagonal other than lii are never read. */
ora_sv* ora_tri_lower_solve(const ora_sm* ll, const ora_sv* b, int* err, int64_t* bad_row) {
  const int64_t nb = b->dim;
  double lii;
  if (err) *err = ORA_OK;
  if (bad_row) *bad_row = -1;
  if (!sm_lookup_checked(ll, 0, 0, &lii)) { if (err) *err = ORA_ERR_OOB_INDEX; return NULL; }
  if (ora_near_zero(lii)) { if (err) *err = ORA_ERR_NEEDS_PIVOTING; if (bad_row) *bad_row = 0; return NULL; }
  double* w = (double*)calloc((size_t)(nb > 0 ? nb : 1), sizeof(double));
  w[0] = sv_lookup0(b, 0) / lii;
  int64_t i = 1;
  for (;;) {
    if (!sm_lookup_checked(ll, i, i, &lii)) { if (err) *err = ORA_ERR_OOB_INDEX; free(w); return NULL; }
    if (ora_near_zero(lii)) { if (err) *err = ORA_ERR_NEEDS_PIVOTING; if (bad_row) *bad_row = i; free(w); return NULL; }
    const double r = subrow_dot(ll, i, 0, i - 1, w, nb);
    if (i < nb) w[i] = (sv_lookup0(b, i) - r) / lii;
    ++i;
    if (i == nb) break;
  }
  ora_sv* v = sv_from_dense_sparsified(nb, w);
  free(w);
  return v;
}

/* triUpperSolve uu w (Sparse.hs:784-811): from the last row upwards;
 *   uInit: i = nw-1 ; x_i = w_i / (uu @@! (i,i))  (unchecked lookup; NeedsPivoting reports index 0 there, :802)
 *   uStep (xx, i): uii = uu @@ (i,i) ; xi = (w @@ i - r) / uii, r = extractSubRow_RK uu i (i+1, nw-1) `dot` dropSV (i+1) xx
 *   stops when i == -1 AFTER a step: nw = 1 looks up uu @@ (-1,-1) -> ORA_ERR_OOB_INDEX. */
ora_sv* ora_tri_upper_solve(const ora_sm* uu, const ora_sv* wv, int* err, int64_t* bad_row) {
  const int64_t nw = wv->dim;
  if (err) *err = ORA_OK;
  if (bad_row) *bad_row = -1;
  if (nw < 1) { if (err) *err = ORA_ERR_OOB_INDEX; return NULL; }
  double uii = 0.0;
  { int found; int64_t p = sm_find_row(uu, nw - 1, &found);          /* (@@!) = lookupWD_SM: default 0, no bounds check */
    if (found) { int f2; int64_t q = sv_find(uu->row[p], nw - 1, &f2); if (f2) uii = uu->row[p]->val[q]; } }
  if (ora_near_zero(uii)) { if (err) *err = ORA_ERR_NEEDS_PIVOTING; if (bad_row) *bad_row = nw - 1; return NULL; }
  double* x = (double*)calloc((size_t)nw, sizeof(double));
  x[nw - 1] = sv_lookup0(wv, nw - 1) / uii;
  int64_t i = nw - 2;
  for (;;) {
    if (!sm_lookup_checked(uu, i, i, &uii)) { if (err) *err = ORA_ERR_OOB_INDEX; free(x); return NULL; }
    if (ora_near_zero(uii)) { if (err) *err = ORA_ERR_NEEDS_PIVOTING; if (bad_row) *bad_row = i; free(x); return NULL; }
    const double r = subrow_dot(uu, i, i + 1, nw - 1, x, nw);
    if (i >= 0) x[i] = (sv_lookup0(wv, i) - r) / uii;
    --i;
    if (i == -1) break;
  }
  ora_sv* v = sv_from_dense_sparsified(nw, x);
  free(x);
  return v;
}

/* ------------------------------------------------------------------ Krylov */

static ora_krylov* kry_new(ora_sv* x, ora_sv* r, ora_sv* p, ora_sv* u) {
  ora_krylov* k = (ora_krylov*)malloc(sizeof(ora_krylov));
  k->x = x; k->r = r; k->p = p; k->u = u;
  return k;
}

void ora_krylov_free(ora_krylov* k) {
  if (!k) return;
  ora_sv_free(k->x); ora_sv_free(k->r); ora_sv_free(k->p); ora_sv_free(k->u); free(k);
}

/* v ^+^ (a .* w) / v ^-^ (a .* w) helpers, keeping the reference's two-step rounding */
static ora_sv* sv_add_scaled(const ora_sv* v, double a, const ora_sv* w) {
  ora_sv* t = ora_sv_scale(a, w); ora_sv* o = ora_sv_add(v, t); ora_sv_free(t); return o;
}
static ora_sv* sv_sub_scaled(const ora_sv* v, double a, const ora_sv* w) {
  ora_sv* t = ora_sv_scale(a, w); ora_sv* o = ora_sv_sub(v, t); ora_sv_free(t); return o;
}

/* bicgsInit aa b x0 = BICGSTAB x0 r0 r0, r0 = b ^-^ (aa #> x0)   Sparse.hs:965-968 */
ora_krylov* ora_bicgs_init(const ora_sm* a, const ora_sv* b, const ora_sv* x0) {
  int err; ora_sv* ax = ora_sm_matvec(a, x0, &err);
  if (!ax) return NULL;
  ora_sv* r0 = ora_sv_sub(b, ax); ora_sv_free(ax);
  return kry_new(ora_sv_copy(x0), r0, ora_sv_copy(r0), NULL);
}

/* bicgstabStep aa r0hat (BICGSTAB x r p)   Sparse.hs:970-981 */
ora_krylov* ora_bicgstab_step(const ora_sm* a, const ora_sv* r0hat, const ora_krylov* st) {
  int err;
  ora_sv* aap = ora_sm_matvec(a, st->p, &err);                                  /* aap = aa #> p */
  if (!aap) return NULL;
  double rr0 = ora_sv_dot(st->r, r0hat);
  double alphaj = rr0 / ora_sv_dot(aap, r0hat);                                 /* (r <.> r0hat) / (aap <.> r0hat) */
  ora_sv* sj = sv_sub_scaled(st->r, alphaj, aap);                               /* r ^-^ (alphaj .* aap) */
  ora_sv* aasj = ora_sm_matvec(a, sj, &err);                                    /* aa #> sj */
  double omegaj = ora_sv_dot(aasj, sj) / ora_sv_dot(aasj, aasj);                /* (aasj <.> sj) / (aasj <.> aasj) */
  ora_sv* t1 = sv_add_scaled(st->x, alphaj, st->p);                             /* x ^+^ (alphaj .* p)  (infixl 6) */
  ora_sv* xj1 = sv_add_scaled(t1, omegaj, sj);                                  /*   ^+^ (omegaj .* sj) */
  ora_sv* rj1 = sv_sub_scaled(sj, omegaj, aasj);                                /* sj ^-^ (omegaj .* aasj) */
  double betaj = ora_sv_dot(rj1, r0hat) / ora_sv_dot(st->r, r0hat) * alphaj / omegaj; /* infixl 7: ((d/rr0)*alpha)/omega */
  ora_sv* t2 = sv_sub_scaled(st->p, omegaj, aap);                               /* p ^-^ (omegaj .* aap) */
  ora_sv* pj1 = sv_add_scaled(rj1, betaj, t2);                                  /* rj1 ^+^ (betaj .* ...) */
  ora_sv_free(aap); ora_sv_free(sj); ora_sv_free(aasj); ora_sv_free(t1); ora_sv_free(t2);
  (void)rr0;
  return kry_new(xj1, rj1, pj1, NULL);
}

/* cgsInit aa b x0 = CGS x0 r0 r0 r0   Sparse.hs:923-926 */
ora_krylov* ora_cgs_init(const ora_sm* a, const ora_sv* b, const ora_sv* x0) {
  int err; ora_sv* ax = ora_sm_matvec(a, x0, &err);
  if (!ax) return NULL;
  ora_sv* r0 = ora_sv_sub(b, ax); ora_sv_free(ax);
  return kry_new(ora_sv_copy(x0), r0, ora_sv_copy(r0), ora_sv_copy(r0));
}

/* cgsStep aa rhat (CGS x r p u)   Sparse.hs:928-939 */
ora_krylov* ora_cgs_step(const ora_sm* a, const ora_sv* rhat, const ora_krylov* st) {
  int err;
  ora_sv* aap = ora_sm_matvec(a, st->p, &err);                                  /* aap = aa #> p */
  if (!aap) return NULL;
  double alphaj = ora_sv_dot(st->r, rhat) / ora_sv_dot(aap, rhat);
  ora_sv* q = sv_sub_scaled(st->u, alphaj, aap);                                /* q = u ^-^ (alphaj .* aap) */
  ora_sv* upq = ora_sv_add(st->u, q);                                           /* u ^+^ q */
  ora_sv* xj1 = sv_add_scaled(st->x, alphaj, upq);                              /* x ^+^ (alphaj .* (u ^+^ q)) */
  ora_sv* aupq = ora_sm_matvec(a, upq, &err);                                   /* aa #> (u ^+^ q) */
  ora_sv* rj1 = sv_sub_scaled(st->r, alphaj, aupq);                             /* r ^-^ (alphaj .* ...) */
  double betaj = ora_sv_dot(rj1, rhat) / ora_sv_dot(st->r, rhat);
  ora_sv* uj1 = sv_add_scaled(rj1, betaj, q);                                   /* rj1 ^+^ (betaj .* q) */
  ora_sv* t = sv_add_scaled(q, betaj, st->p);                                   /* q ^+^ (betaj .* p) */
  ora_sv* pj1 = sv_add_scaled(uj1, betaj, t);                                   /* uj1 ^+^ (betaj .* t) */
  ora_sv_free(aap); ora_sv_free(q); ora_sv_free(upq); ora_sv_free(aupq); ora_sv_free(t);
  return kry_new(xj1, rj1, pj1, uj1);
}

/* cgneInit aa b x0 = CGNE x0 r0 p0, r0 = b ^-^ (aa #> x0), p0 = transposeSM aa #> r0   Sparse.hs:862-866 */
ora_krylov* ora_cgne_init(const ora_sm* a, const ora_sv* b, const ora_sv* x0) {
  int err; ora_sv* ax = ora_sm_matvec(a, x0, &err);
  if (!ax) return NULL;
  ora_sv* r0 = ora_sv_sub(b, ax); ora_sv_free(ax);
  ora_sm* at = ora_sm_transpose(a);
  ora_sv* p0 = ora_sm_matvec(at, r0, &err);
  ora_sm_free(at);
  if (!p0) { ora_sv_free(r0); return NULL; }
  return kry_new(ora_sv_copy(x0), r0, p0, NULL);
}

/* cgneStep aa (CGNE x r p)   Sparse.hs:868-878 */
ora_krylov* ora_cgne_step(const ora_sm* a, const ora_krylov* st) {
  int err;
  double rr = ora_sv_dot(st->r, st->r);
  double alphai = rr / ora_sv_dot(st->p, st->p);                                /* (r.r) / (p.p) */
  ora_sv* x1 = sv_add_scaled(st->x, alphai, st->p);                             /* x ^+^ (alphai .* p) */
  ora_sv* ap = ora_sm_matvec(a, st->p, &err);
  if (!ap) { ora_sv_free(x1); return NULL; }
  ora_sv* r1 = sv_sub_scaled(st->r, alphai, ap);                                /* r ^-^ (alphai .* (aa #> p)) */
  double beta = ora_sv_dot(r1, r1) / ora_sv_dot(st->r, st->r);
  ora_sm* at = ora_sm_transpose(a);                                             /* transpose aa, every step */
  ora_sv* atr = ora_sm_matvec(at, r1, &err);
  ora_sm_free(at);
  /* p1 = transpose aa #> r1 ^+^ (beta .* p): no fixity is declared for (#>) (Class.hs:224-229),
   * so it defaults to infixl 9 and binds tighter than ^+^ (infixl 6): (Aᵀ r1) + beta p. */
  ora_sv* p1 = sv_add_scaled(atr, beta, st->p);
  ora_sv_free(ap); ora_sv_free(atr);
  return kry_new(x1, r1, p1, NULL);
}

/* linSolve0 method aa b x0   Sparse.hs:1016-1072 */
ora_sv* ora_linsolve0(int method, const ora_sm* a, const ora_sv* b, const ora_sv* x0,
                      int nits, double tol_abs, double tol_rel,
                      int* iters, double* res_hist, int* err) {
  int e = ORA_OK;
  if (iters) *iters = 0;
  if (err) *err = ORA_OK;
  if (nits <= 0) nits = 200;                /* nits = 200     :1034 */
  if (tol_abs <= 0) tol_abs = 1e-6;         /* tolAbs = 1e-6  :1035 */
  if (tol_rel <= 0) tol_rel = 1e-4;         /* tolRel = 1e-4  :1036 */
  if (a->nrows != b->dim) { if (err) *err = ORA_ERR_SIZE_MISMATCH; return NULL; }   /* m /= nb  :1022 */
  if (ora_sm_is_diagonal(a)) {                                                      /* :1024-1025 */
    ora_sm* ra = ora_sm_reciprocal(a);
    ora_sv* x = ora_sm_matvec(ra, b, &e);
    ora_sm_free(ra);
    if (err) *err = e;
    return x;
  }
  if (method != ORA_BICGSTAB && method != ORA_CGS && method != ORA_CGNE) {          /* IterE :1031 */
    if (err) *err = ORA_ERR_UNSUPPORTED_METHOD;
    return NULL;
  }
  ora_sv* ax0 = ora_sm_matvec(a, x0, &e);
  if (!ax0) { if (err) *err = e; return NULL; }
  ora_sv* r0hat = ora_sv_sub(b, ax0); ora_sv_free(ax0);                             /* r0hat = b ^-^ (aa #> x0) :1032 */
  double r0norm = ora_sv_norm2(r0hat);
  double tol = tol_abs > tol_rel * r0norm ? tol_abs : tol_rel * r0norm;             /* max tolAbs (tolRel * r0norm) :1037 */
  ora_krylov* st = method == ORA_BICGSTAB ? ora_bicgs_init(a, b, x0)
                 : method == ORA_CGS      ? ora_cgs_init(a, b, x0)
                                          : ora_cgne_init(a, b, x0);
  int n = 0;
  while (n < nits) {                                                                /* runIter :1043-1052 */
    ora_krylov* st1 = method == ORA_BICGSTAB ? ora_bicgstab_step(a, r0hat, st)
                    : method == ORA_CGS      ? ora_cgs_step(a, r0hat, st)
                                             : ora_cgne_step(a, st);
    ora_krylov_free(st); st = st1;
    ora_sv* ax = ora_sm_matvec(a, st->x, &e);
    ora_sv* d = ora_sv_sub(ax, b);                                                  /* norm2 ((aa #> x) ^-^ b) :1041 */
    double res = ora_sv_norm2(d);
    ora_sv_free(ax); ora_sv_free(d);
    if (res_hist) res_hist[n] = res;
    ++n;
    if (res <= tol) break;                                                          /* NaN <= tol is False */
  }
  if (iters) *iters = n;
  ora_sv* x = ora_sv_copy(st->x);
  ora_krylov_free(st); ora_sv_free(r0hat);
  return x;
}

/* arnoldi aa b kn   Sparse.hs:630-667.  modifyUntil applies the step, then tests (Iterative.hs:246-252):
 * state starts at i = 1 and stops when i == kn or on breakdown; kn <= 1 runs until breakdown, which
 * max_steps bounds here (the reference would loop). */
int ora_arnoldi(const ora_sm* a, const ora_sv* b, int kn, int max_steps,
                double* q_out, double* h_out, int* ncols_q, int* nmax_out) {
  int err;
  if (a->ncols != b->dim) return ORA_ERR_SIZE_MISMATCH;                /* n == nb else MatVecSizeMismatchException */
  int64_t m = a->nrows;
  int cap = max_steps + 2;
  ora_sv** qv = (ora_sv**)malloc(sizeof(ora_sv*) * (size_t)cap);
  /* H entries as (i, j, v) triples in insertion order, later fed to fromListSM (last write wins) */
  int64_t hcap = (int64_t)cap * (cap + 1) / 2 + 4, hn = 0;
  int64_t* hi = (int64_t*)malloc(sizeof(int64_t) * (size_t)hcap);
  int64_t* hj = (int64_t*)malloc(sizeof(int64_t) * (size_t)hcap);
  double* hv = (double*)malloc(sizeof(double) * (size_t)hcap);
  /* arnInit :642-651 */
  ora_sv* q0 = ora_sv_normalize2(b);
  ora_sv* aq0 = ora_sm_matvec(a, q0, &err);
  double h11 = ora_sv_dot(q0, aq0);
  ora_sv* q1nn = sv_sub_scaled(aq0, h11, q0);
  double h21 = ora_sv_norm2(q1nn);
  ora_sv* q1 = ora_sv_normalize2(q1nn);
  ora_sv_free(aq0); ora_sv_free(q1nn);
  qv[0] = q0; qv[1] = q1; int nq = 2;
  hi[hn] = 0; hj[hn] = 0; hv[hn] = h11; ++hn;
  hi[hn] = 1; hj[hn] = 0; hv[hn] = h21; ++hn;
  int i = 1, fbreak = 0;
  /* modifyUntil tf arnoldiStep: step first, then test */
  for (;;) {
    if (i - 1 >= max_steps) break;
    /* arnoldiStep :652-667 */
    ora_sv* aqi = ora_sm_matvec(a, qv[nq - 1], &err);                  /* aa #> last qv */
    double* hhcoli = (double*)malloc(sizeof(double) * (size_t)nq);
    for (int k = 0; k < nq; ++k) hhcoli[k] = ora_sv_dot(qv[k], aqi);   /* fmap (`dot` aqi) qv : q_k <.> aqi */
    ora_sv* acc = ora_sv_zero(m);                                      /* foldl' (^+^) zv (zipWith (.*) hhcoli qv) */
    for (int k = 0; k < nq; ++k) { ora_sv* t = sv_add_scaled(acc, hhcoli[k], qv[k]); ora_sv_free(acc); acc = t; }
    ora_sv* qipnn = ora_sv_sub(aqi, acc);
    double qipnorm = ora_sv_norm2(qipnn);
    ora_sv* qip = ora_sv_normalize2(qipnn);
    for (int k = 0; k < nq; ++k) { hi[hn] = k; hj[hn] = i; hv[hn] = hhcoli[k]; ++hn; }   /* zip3 [0..] (replicate i) */
    hi[hn] = nq; hj[hn] = i; hv[hn] = qipnorm; ++hn;
    qv[nq++] = qip;
    fbreak = ora_near_zero(qipnorm);                                   /* nearZero qipnorm */
    free(hhcoli); ora_sv_free(aqi); ora_sv_free(acc); ora_sv_free(qipnn);
    i = i + 1;
    if (i == kn || fbreak) break;                                      /* tf */
  }
  int nmax = i;
  /* fromColsV qvfin: Q is m x nq; fromListSM (nmax+1, nmax) hhfin */
  for (int c = 0; c < nq; ++c) ora_sv_to_dense(qv[c], q_out + (int64_t)c * m);
  for (int64_t z = 0; z < (int64_t)(nmax + 1) * nmax; ++z) h_out[z] = 0.0;
  int rc = ORA_OK;
  for (int64_t z = 0; z < hn; ++z) {
    if (hi[z] < 0 || hi[z] >= nmax + 1 || hj[z] < 0 || hj[z] >= nmax) { rc = ORA_ERR_OOB_INDEX; continue; }
    h_out[hj[z] * (nmax + 1) + hi[z]] = hv[z];
  }
  *ncols_q = nq; *nmax_out = nmax;
  for (int c = 0; c < nq; ++c) ora_sv_free(qv[c]);
  free(qv); free(hi); free(hj); free(hv);
  return rc;
}

/* ------------------------------------------------------------------ synthetic workloads */

ora_sm* ora_synth_matrix(int kind, int64_t n, int k, uint64_t seed, int64_t band) {
  ora_sm* a = ora_sm_zero(n, n);
  a->cap = n;
  a->rkey = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  a->row = (ora_sv**)malloc(sizeof(ora_sv*) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) {
    int64_t cols[SLA_SYNTH_MAX_K]; double vals[SLA_SYNTH_MAX_K];
    int c = sla_synth_row(kind, n, k, seed, band, i, cols, vals);
    ora_sv* r = sv_alloc(n, c);
    memcpy(r->idx, cols, sizeof(int64_t) * (size_t)c);
    memcpy(r->val, vals, sizeof(double) * (size_t)c);
    r->nnz = c;
    a->rkey[i] = i; a->row[i] = r;
  }
  a->nstored = n;
  return a;
}

ora_sv* ora_synth_vector(uint64_t seed, int64_t n) {
  ora_sv* v = sv_alloc(n, n);
  for (int64_t i = 0; i < n; ++i) { v->idx[i] = i; v->val[i] = sla_synth_vec(seed, i); }
  v->nnz = n;
  return v;
}

void ora_synth_row(int kind, int64_t n, int k, uint64_t seed, int64_t band, int64_t i,
                   int64_t* cols, double* vals, int* count) {
  *count = sla_synth_row(kind, n, k, seed, band, i, cols, vals);
}

/* ------------------------------------------------------------------ CPU-baseline timers */

static double now_s(void) {
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Times `reps` evaluations of aa #> x.  threads == 1 is the reference's own shape (single-threaded,
 * fresh result vector per call).  threads > 1 splits the stored rows over OpenMP threads; each
 * row's dotu is the same code, so the result is bit-identical. */
double ora_time_matvec(const ora_sm* a, const ora_sv* x, int reps, int threads, double* checksum) {
  double cs = 0.0;
  double t0 = now_s();
  for (int rep = 0; rep < reps; ++rep) {
    ora_sv* y = sv_alloc(a->nrows, a->nstored);
    y->nnz = a->nstored;
#ifdef _OPENMP
    if (threads > 1) {
#pragma omp parallel for num_threads(threads) schedule(static)
      for (int64_t r = 0; r < a->nstored; ++r) { y->idx[r] = a->rkey[r]; y->val[r] = sv_dot_ordered(a->row[r], x); }
    } else
#endif
    {
      for (int64_t r = 0; r < a->nstored; ++r) { y->idx[r] = a->rkey[r]; y->val[r] = sv_dot_ordered(a->row[r], x); }
    }
    cs += y->nnz ? y->val[(rep * 7919) % y->nnz] : 0.0;
    ora_sv_free(y);
  }
  double t1 = now_s();
  if (checksum) *checksum = cs;
  (void)threads;
  return (t1 - t0) / (double)(reps > 0 ? reps : 1);
}

/* Times `steps` bicgstabStep calls (single thread, the reference's shape). Returns seconds per step. */
double ora_time_bicgstab(const ora_sm* a, const ora_sv* b, const ora_sv* x0, int steps, double* checksum) {
  int err; ora_sv* ax0 = ora_sm_matvec(a, x0, &err);
  ora_sv* r0hat = ora_sv_sub(b, ax0); ora_sv_free(ax0);
  ora_krylov* st = ora_bicgs_init(a, b, x0);
  double t0 = now_s();
  for (int s = 0; s < steps; ++s) { ora_krylov* n = ora_bicgstab_step(a, r0hat, st); ora_krylov_free(st); st = n; }
  double t1 = now_s();
  if (checksum) *checksum = ora_sv_norm2(st->r);
  ora_krylov_free(st); ora_sv_free(r0hat);
  return (t1 - t0) / (double)(steps > 0 ? steps : 1);
}
