// sla_b200.hpp — header-only C++17 host mirror of Numeric.LinearAlgebra.Sparse's operator surface over the C ABI
// (include/sla_b200.h).  The reference's host language is Haskell and no GHC exists in the build image, so next to
// the Haskell shim source (hs/) and the Python mirror (sparse_linear_algebra_b200/sparse.py) this header gives
// compiled-language callers the reference's names, argument order and error behaviour:
//
//     reference                                    here
//     aa #> v          Common.hs:242-250           aa.matVec(v)        /  aa * v
//     v <# aa          Common.hs:253-256           aa.vecMat(v)
//     v <.> w          SpVector.hs:116-117         dot(v, w)
//     v ^+^ w, v ^-^ w SpVector.hs:107-110         v + w, v - w
//     a .* v, v ./ s   SpVector.hs:112-114         a * v, v / s
//     norm2, normalize2  SpVector.hs:119-129       norm2(v), normalize2(v)
//     transpose aa     SpMatrix.hs:717-718         aa.transpose()
//     bicgsInit / bicgstabStep  Sparse.hs:965-981  bicgsInit(aa, b, x0) / bicgstabStep(aa, r0hat, st)
//     cgsInit / cgsStep         Sparse.hs:923-939  cgsInit / cgsStep
//     linSolve0 m aa b x0       Sparse.hs:1016-1072 linSolve0(method, aa, b, x0)
//     arnoldi aa b kn           Sparse.hs:630-667  arnoldi(aa, b, kn)
//
// Errors are thrown as exceptions named after the reference's (Control/Exception/Common.hs:44-76).
// Everything here is marshalling: all arithmetic runs in libsla_b200.so; there is no CPU path.
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "sla_b200.h"

namespace sla {

struct Error : std::runtime_error {
  sla_status status;
  Error(sla_status s, const std::string& m) : std::runtime_error(m), status(s) {}
};
struct MatVecSizeMismatchException : Error { using Error::Error; };   // OperandSizeMismatch
struct OutOfBoundsIndexError : Error { using Error::Error; };         // error "insertSpMatrix : index out of bounds"
struct IterE : Error { using Error::Error; };                         // IterationException IterE
struct NeedsPivoting : Error { using Error::Error; };                 // MatrixException NeedsPivoting (triangular solves)

enum LinSolveMethod { GMRES_ = SLA_GMRES_, CGNE_ = SLA_CGNE_, BCG_ = SLA_BCG_, CGS_ = SLA_CGS_, BICGSTAB_ = SLA_BICGSTAB_ };

class Context {
 public:
  explicit Context(int device = 0) {
    sla_ctx* c = nullptr;
    const sla_status s = sla_init(device, &c);
    if (s != SLA_OK) throw Error(s, sla_last_error(nullptr));
    ctx_.reset(c, [](sla_ctx* p) { sla_finalize(p); });
  }
  sla_ctx* get() const { return ctx_.get(); }
  void check(sla_status s) const {
    if (s == SLA_OK) return;
    const std::string msg = sla_last_error(ctx_.get());
    switch (s) {
      case SLA_ERR_SIZE_MISMATCH: throw MatVecSizeMismatchException(s, msg);
      case SLA_ERR_OOB_INDEX: throw OutOfBoundsIndexError(s, msg);
      case SLA_ERR_UNSUPPORTED_METHOD: throw IterE(s, msg);
      case SLA_ERR_NEEDS_PIVOTING: throw NeedsPivoting(s, msg);
      default: throw Error(s, msg);
    }
  }
  std::int64_t launches() const { return sla_launch_count(ctx_.get()); }

 private:
  std::shared_ptr<sla_ctx> ctx_;
};

// SpVector Double: a dense double[n] on the device; absent keys of the reference's IntMap are 0.0
class SpVector {
 public:
  SpVector(const Context& c, std::int64_t n) : c_(c) {                                  // zeroSV
    sla_vec* v = nullptr;
    c_.check(sla_vec_create(c_.get(), n, &v));
    own(v);
  }
  SpVector(const Context& c, const std::vector<double>& dense) : c_(c) {                // mkSpVR / fromListDenseSV
    sla_vec* v = nullptr;
    c_.check(sla_vec_from_host(c_.get(), (std::int64_t)dense.size(), dense.data(), &v));
    own(v);
  }
  static SpVector constv(const Context& c, std::int64_t n, double a) {                   // constv
    SpVector v(c, n);
    c.check(sla_vec_fill(c.get(), v.get(), a));
    return v;
  }
  static SpVector borrowed(const Context& c, const sla_vec* v) { return SpVector(c, const_cast<sla_vec*>(v), false); }
  std::int64_t dim() const { return sla_vec_dim(v_.get()); }
  std::vector<double> toDenseListSV() const {
    std::vector<double> out((std::size_t)dim());
    c_.check(sla_vec_to_host(c_.get(), v_.get(), out.data()));
    return out;
  }
  SpVector copy() const {
    SpVector z(c_, dim());
    c_.check(sla_vec_copy(c_.get(), v_.get(), z.get()));
    return z;
  }
  sla_vec* get() const { return v_.get(); }
  const Context& ctx() const { return c_; }

 private:
  SpVector(const Context& c, sla_vec* v, bool owned) : c_(c) {
    if (owned) own(v); else v_.reset(v, [](sla_vec*) {});
  }
  void own(sla_vec* v) { v_.reset(v, [](sla_vec* p) { sla_vec_free(p); }); }
  Context c_;
  std::shared_ptr<sla_vec> v_;
};

inline SpVector operator+(const SpVector& x, const SpVector& y) {   // ^+^
  SpVector z(x.ctx(), x.dim());
  x.ctx().check(sla_vec_add(x.ctx().get(), x.get(), y.get(), z.get()));
  return z;
}
inline SpVector operator-(const SpVector& x, const SpVector& y) {   // ^-^
  SpVector z(x.ctx(), x.dim());
  x.ctx().check(sla_vec_sub(x.ctx().get(), x.get(), y.get(), z.get()));
  return z;
}
inline SpVector operator*(double a, const SpVector& x) {            // .*
  SpVector z(x.ctx(), x.dim());
  x.ctx().check(sla_vec_scale(x.ctx().get(), a, x.get(), z.get()));
  return z;
}
inline SpVector operator/(const SpVector& x, double s) { return (1.0 / s) * x; }   // ./ = recip s .* v
inline double dot(const SpVector& x, const SpVector& y) {           // <.>
  double d = 0;
  x.ctx().check(sla_dot(x.ctx().get(), x.get(), y.get(), &d));
  return d;
}
inline double norm2(const SpVector& x) {
  double d = 0;
  x.ctx().check(sla_norm2(x.ctx().get(), x.get(), &d));
  return d;
}
inline SpVector normalize2(const SpVector& x) {
  SpVector z(x.ctx(), x.dim());
  x.ctx().check(sla_vec_normalize2(x.ctx().get(), x.get(), z.get()));
  return z;
}

// SpMatrix Double: CSR on the device
class SpMatrix {
 public:
  struct Triple { std::int64_t i, j; double v; };
  // fromListSM (m, n) [(i, j, v)]: later duplicates overwrite, out-of-bounds throws
  SpMatrix(const Context& c, std::int64_t m, std::int64_t n, const std::vector<Triple>& iix) : c_(c) {
    std::vector<std::int64_t> i(iix.size()), j(iix.size());
    std::vector<double> v(iix.size());
    for (std::size_t q = 0; q < iix.size(); ++q) { i[q] = iix[q].i; j[q] = iix[q].j; v[q] = iix[q].v; }
    sla_csr* a = nullptr;
    c_.check(sla_csr_from_coo(c_.get(), m, n, (std::int64_t)iix.size(), i.data(), j.data(), v.data(), &a));
    a_.reset(a, [](sla_csr* p) { sla_csr_free(p); });
  }
  std::int64_t nrows() const { std::int64_t m, n, z; sla_csr_dims(a_.get(), &m, &n, &z); return m; }
  std::int64_t ncols() const { std::int64_t m, n, z; sla_csr_dims(a_.get(), &m, &n, &z); return n; }
  std::int64_t nnz() const { std::int64_t m, n, z; sla_csr_dims(a_.get(), &m, &n, &z); return z; }
  SpVector matVec(const SpVector& x) const {     // aa #> x
    SpVector y(c_, nrows());
    c_.check(sla_spmv(c_.get(), a_.get(), x.get(), y.get()));
    return y;
  }
  SpVector vecMat(const SpVector& x) const {     // x <# aa
    SpVector y(c_, ncols());
    c_.check(sla_spmvT(c_.get(), a_.get(), x.get(), y.get()));
    return y;
  }
  SpVector operator*(const SpVector& x) const { return matVec(x); }
  SpMatrix transpose() const {
    sla_csr* t = nullptr;
    c_.check(sla_csr_transpose(c_.get(), a_.get(), &t));
    return SpMatrix(c_, t);
  }
  bool isDiagonalSM() const {
    int d = 0;
    c_.check(sla_csr_is_diagonal(c_.get(), a_.get(), &d));
    return d != 0;
  }
  sla_csr* get() const { return a_.get(); }
  const Context& ctx() const { return c_; }
  // takes ownership of a handle the C ABI returned (preconditioners, partitions)
  static SpMatrix adopt(const Context& c, sla_csr* a) { return SpMatrix(c, a); }

 private:
  SpMatrix(const Context& c, sla_csr* a) : c_(c) { a_.reset(a, [](sla_csr* p) { sla_csr_free(p); }); }
  Context c_;
  std::shared_ptr<sla_csr> a_;
};

// BICGSTAB / CGS records: the state lives on the device and is advanced IN PLACE by the step functions
class KrylovState {
 public:
  KrylovState(const Context& c, sla_krylov* st) : c_(c) { st_.reset(st, [](sla_krylov* p) { sla_krylov_free(p); }); }
  SpVector field(int f) const {
    const sla_vec* v = nullptr;
    c_.check(sla_krylov_view(c_.get(), st_.get(), f, &v));
    return SpVector::borrowed(c_, v);
  }
  SpVector x() const { return field(SLA_FIELD_X); }   // _x / _xBicgstab
  SpVector r() const { return field(SLA_FIELD_R); }
  SpVector p() const { return field(SLA_FIELD_P); }
  SpVector u() const { return field(SLA_FIELD_U); }
  sla_krylov* get() const { return st_.get(); }
  // deep copy: what a PURE step needs (`iterate (bicgstabStep aa r0hat) st0 !! 20` keeps st0 usable, README.md:208)
  KrylovState clone() const {
    sla_krylov* out = nullptr;
    c_.check(sla_krylov_clone(c_.get(), st_.get(), &out));
    return KrylovState(c_, out);
  }

 private:
  Context c_;
  std::shared_ptr<sla_krylov> st_;
};

inline KrylovState bicgsInit(const SpMatrix& aa, const SpVector& b, const SpVector& x0) {
  sla_krylov* st = nullptr;
  aa.ctx().check(sla_bicgstab_init(aa.ctx().get(), aa.get(), b.get(), x0.get(), &st));
  return KrylovState(aa.ctx(), st);
}
inline KrylovState& bicgstabStep(const SpMatrix& aa, const SpVector& r0hat, KrylovState& st) {
  aa.ctx().check(sla_bicgstab_step(aa.ctx().get(), aa.get(), r0hat.get(), st.get()));
  return st;
}
// the reference's signature: a new record, the argument untouched (clone, then advance the clone)
inline KrylovState bicgstabStepPure(const SpMatrix& aa, const SpVector& r0hat, const KrylovState& st) {
  KrylovState next = st.clone();
  bicgstabStep(aa, r0hat, next);
  return next;
}
inline KrylovState cgsInit(const SpMatrix& aa, const SpVector& b, const SpVector& x0) {
  sla_krylov* st = nullptr;
  aa.ctx().check(sla_cgs_init(aa.ctx().get(), aa.get(), b.get(), x0.get(), &st));
  return KrylovState(aa.ctx(), st);
}
inline KrylovState& cgsStep(const SpMatrix& aa, const SpVector& rhat, KrylovState& st) {
  aa.ctx().check(sla_cgs_step(aa.ctx().get(), aa.get(), rhat.get(), st.get()));
  return st;
}

inline KrylovState cgneInit(const SpMatrix& aa, const SpVector& b, const SpVector& x0) {      // Sparse.hs:862-866
  sla_krylov* st = nullptr;
  aa.ctx().check(sla_cgne_init(aa.ctx().get(), aa.get(), b.get(), x0.get(), &st));
  return KrylovState(aa.ctx(), st);
}
inline KrylovState& cgneStep(const SpMatrix& aa, KrylovState& st) {                            // Sparse.hs:868-878
  aa.ctx().check(sla_cgne_step(aa.ctx().get(), aa.get(), st.get()));
  return st;
}

// ---- preconditioners and triangular solves (Sparse.hs:673-721, 750-811)
struct DiagPartitions { SpMatrix e, d, f; };     // strictly sub-diagonal, diagonal, strictly super-diagonal
inline DiagPartitions diagPartitions(const SpMatrix& aa) {
  sla_csr *e = nullptr, *d = nullptr, *f = nullptr;
  aa.ctx().check(sla_csr_diag_partitions(aa.ctx().get(), aa.get(), &e, &d, &f));
  return DiagPartitions{SpMatrix::adopt(aa.ctx(), e), SpMatrix::adopt(aa.ctx(), d), SpMatrix::adopt(aa.ctx(), f)};
}
inline SpMatrix jacobiPre(const SpMatrix& aa) {                       // recip <$> extractDiag x
  sla_csr* m = nullptr;
  aa.ctx().check(sla_jacobi_pre(aa.ctx().get(), aa.get(), &m));
  return SpMatrix::adopt(aa.ctx(), m);
}
inline std::pair<SpMatrix, SpMatrix> mSsorPre(const SpMatrix& aa, double omega) {   // (l, r)
  sla_csr *l = nullptr, *r = nullptr;
  aa.ctx().check(sla_mssor_pre(aa.ctx().get(), aa.get(), omega, &l, &r));
  return {SpMatrix::adopt(aa.ctx(), l), SpMatrix::adopt(aa.ctx(), r)};
}
inline SpVector triLowerSolve(const SpMatrix& ll, const SpVector& b) {    // forward substitution; NeedsPivoting on a nearZero diagonal
  SpVector w(ll.ctx(), b.dim());
  ll.ctx().check(sla_tri_lower_solve(ll.ctx().get(), ll.get(), b.get(), w.get()));
  return w;
}
inline SpVector triUpperSolve(const SpMatrix& uu, const SpVector& w) {    // backward substitution
  SpVector x(uu.ctx(), w.dim());
  uu.ctx().check(sla_tri_upper_solve(uu.ctx().get(), uu.get(), w.get(), x.get()));
  return x;
}

struct SolveInfo { int iters = 0; double resnorm = 0; };

// linSolve0 method aa b x0: nits = 200, tol = max 1e-6 (1e-4 * ||r0||), true residual every iteration
inline SpVector linSolve0(LinSolveMethod method, const SpMatrix& aa, const SpVector& b, const SpVector& x0, SolveInfo* info = nullptr) {
  SpVector x(aa.ctx(), x0.dim());
  SolveInfo local;
  SolveInfo* out = info ? info : &local;
  aa.ctx().check(sla_linsolve0(aa.ctx().get(), (int)method, aa.get(), b.get(), x0.get(), nullptr, x.get(), &out->iters, &out->resnorm));
  return x;
}

// restarted GMRES(m) — the algorithm behind (<\>) (Sparse.hs:837-848): Arnoldi with Givens rotations on the device, restart until the
// true residual meets max tol_abs (tol_rel * ||r0||) or max_iters Arnoldi steps are spent
inline SpVector gmres(const SpMatrix& aa, const SpVector& b, const SpVector& x0, int restart = 30, SolveInfo* info = nullptr,
                      const sla_solve_opts* opts = nullptr) {
  SpVector x(aa.ctx(), x0.dim());
  SolveInfo local;
  SolveInfo* out = info ? info : &local;
  aa.ctx().check(sla_gmres(aa.ctx().get(), aa.get(), b.get(), x0.get(), restart, opts, x.get(), &out->iters, &out->resnorm));
  return x;
}
inline SpVector backslash(const SpMatrix& aa, const SpVector& b) {          // aa <\> b, from x0 = 0
  return gmres(aa, b, SpVector(aa.ctx(), b.dim()));
}

inline std::pair<SpMatrix, SpMatrix> ilu0Pre(const SpMatrix& aa) {           // (l, u): lu masked by the pattern of aa, Sparse.hs:696-706
  sla_csr *l = nullptr, *u = nullptr;
  aa.ctx().check(sla_ilu0_pre(aa.ctx().get(), aa.get(), &l, &u));
  return {SpMatrix::adopt(aa.ctx(), l), SpMatrix::adopt(aa.ctx(), u)};
}
inline double normFrobenius(const SpMatrix& aa) {                            // sqrt (sum of squares), Class.hs:206-207
  double d = 0;
  aa.ctx().check(sla_csr_norm_frobenius(aa.ctx().get(), aa.get(), &d));
  return d;
}

struct ArnoldiResult {
  std::vector<double> Q;   // n x (nmax + 1), column-major
  std::vector<double> H;   // (nmax + 1) x nmax, column-major
  int nmax = 0;
  bool breakdown = false;
};

inline ArnoldiResult arnoldi(const SpMatrix& aa, const SpVector& b, int kn) {
  ArnoldiResult r;
  r.H.assign((std::size_t)(kn + 1) * kn, 0.0);
  sla_dense* q = nullptr;
  const sla_status s = sla_arnoldi(aa.ctx().get(), aa.get(), b.get(), kn, &q, r.H.data(), &r.nmax);
  if (s != SLA_OK && s != SLA_ERR_BREAKDOWN) aa.ctx().check(s);
  r.breakdown = s == SLA_ERR_BREAKDOWN;
  std::int64_t rows = 0, cols = 0;
  sla_dense_dims(q, &rows, &cols);
  r.Q.resize((std::size_t)rows * cols);
  const sla_status s2 = sla_dense_to_host(aa.ctx().get(), q, r.Q.data());
  sla_dense_free(q);
  aa.ctx().check(s2);
  r.H.resize((std::size_t)(r.nmax + 1) * r.nmax);
  return r;
}

}  // namespace sla
