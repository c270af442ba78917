/* sla_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the hot path of ocramz/sparse-linear-algebra
 * (reference commit b940b12), operation for operation, in the reference's
 * evaluation order.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (libsla_b200.so) never links, imports or calls it.
 *
 * PARITY PINNING: the reference is Haskell and no GHC exists in the build image,
 * so this restatement cannot be run against the reference binary.  It is pinned
 * against every known-answer test the reference holds for this path
 * (test/LibSpec.hs:45-65, 226-232, 252-321; README.md:183-189), ported in
 * tests/test_oracle_golden.py.  Those tests pin VALUES to 1e-12 (nearZero), not
 * the bit-level summation order; the order adopted here (strict left fold,
 * ascending key, seed 0) is the one base's default `sum` gives for the derived
 * Foldable instances (IntM.hs:17, SpVector.hs:43) — bit-level order is
 * "parity-unpinned", see DESIGN.md.
 *
 * Containers: the reference's IntMap (ascending-key Patricia trie) is restated
 * as a sorted array of (key, value) pairs; union / intersection / fold visit
 * keys in the same ascending order, so every result is identical.
 */
#ifndef SLA_ORACLE_H
#define SLA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* SpVector a = SV dim (IntM a)            src/Data/Sparse/SpVector.hs:42-43 */
typedef struct ora_sv { int64_t dim, nnz, cap; int64_t* idx; double* val; } ora_sv;
/* SpMatrix a = SM (rows, cols) (IntM (IntM a))   src/Data/Sparse/SpMatrix.hs:52-54 */
typedef struct ora_sm { int64_t nrows, ncols, nstored, cap; int64_t* rkey; ora_sv** row; } ora_sm;

enum { ORA_OK = 0, ORA_ERR_SIZE_MISMATCH = 1, ORA_ERR_OOB_INDEX = 2, ORA_ERR_UNSUPPORTED_METHOD = 3,
       ORA_ERR_NEEDS_PIVOTING = 4 /* MatrixException NeedsPivoting, Control/Exception/Common.hs:57-61 */ };
/* LinSolveMethod, Sparse.hs:1007-1012 */
enum { ORA_GMRES = 0, ORA_CGNE = 1, ORA_BCG = 2, ORA_CGS = 3, ORA_BICGSTAB = 4 };

/* ---- SpVector ---- */
ora_sv* ora_sv_zero(int64_t dim);                                         /* zeroSV            SpVector.hs:157-158 */
ora_sv* ora_sv_from_dense(int64_t dim, const double* x, int64_t len);     /* mkSpVR / fromListDenseSV  :183-195 */
ora_sv* ora_sv_from_list(int64_t dim, int64_t n, const int64_t* idx, const double* val); /* fromListSV :275-278 */
ora_sv* ora_sv_copy(const ora_sv*);
void    ora_sv_free(ora_sv*);
int64_t ora_sv_dim(const ora_sv*);
int64_t ora_sv_nnz(const ora_sv*);
void    ora_sv_to_dense(const ora_sv*, double* out);                      /* toDenseListSV     :300-301 */
void    ora_sv_to_list(const ora_sv*, int64_t* idx, double* val);         /* toListSV          :293-294 */
ora_sv* ora_sv_add(const ora_sv* v, const ora_sv* w);                     /* (^+^)             :107-110 */
ora_sv* ora_sv_negate(const ora_sv* v);                                   /* negateV           :110     */
ora_sv* ora_sv_sub(const ora_sv* v, const ora_sv* w);                     /* (^-^)   Class.hs:68-69     */
ora_sv* ora_sv_scale(double a, const ora_sv* v);                          /* (.*)              :112-114 */
ora_sv* ora_sv_divs(const ora_sv* v, double s);                           /* (./)    Class.hs:94-95     */
double  ora_sv_dot(const ora_sv* v, const ora_sv* w);                     /* (<.>)             :116-117 */
double  ora_sv_norm2sq(const ora_sv* v);                                  /* norm2Sq           :122     */
double  ora_sv_norm2(const ora_sv* v);                                    /* norm2 / norm2'    :127-128 */
ora_sv* ora_sv_normalize2(const ora_sv* v);                               /* normalize2        :125     */
int     ora_near_zero(double a);                                          /* nearZero  Eps.hs:41-42     */

/* ---- SpMatrix ---- */
ora_sm* ora_sm_zero(int64_t m, int64_t n);
ora_sm* ora_sm_from_list(int64_t m, int64_t n, int64_t nnz, const int64_t* i, const int64_t* j,
                         const double* v, int* err);                      /* fromListSM        SpMatrix.hs:205-224 */
ora_sm* ora_sm_from_dense_colmajor(int64_t m, const double* ll, int64_t len); /* fromListDenseSM :239-241 */
ora_sm* ora_sm_from_csr(int64_t m, int64_t n, const int64_t* row_ptr, const int64_t* col, const double* val);
void    ora_sm_free(ora_sm*);
int64_t ora_sm_nrows(const ora_sm*);
int64_t ora_sm_ncols(const ora_sm*);
int64_t ora_sm_nnz(const ora_sm*);
int64_t ora_sm_nstored_rows(const ora_sm*);
void    ora_sm_to_coo(const ora_sm*, int64_t* i, int64_t* j, double* v);  /* ascending (row, col) */
void    ora_sm_to_csr(const ora_sm*, int64_t* row_ptr, int64_t* col, double* val);
ora_sm* ora_sm_transpose(const ora_sm*);                                  /* transposeSM       :717-718 */
int     ora_sm_is_diagonal(const ora_sm*);                                /* isDiagonalSM      :411-415 */
ora_sm* ora_sm_reciprocal(const ora_sm*);                                 /* reciprocal  Class.hs:174-175 */
ora_sm* ora_sm_sparsify(const ora_sm*);                                   /* sparsifySM (drops |x|<=1e-12) */
ora_sv* ora_sm_matvec(const ora_sm* a, const ora_sv* x, int* err);        /* (#>) matVecSD  Common.hs:242-250 */
ora_sv* ora_sm_vecmat(const ora_sv* x, const ora_sm* a, int* err);        /* (<#) vecMatSD  Common.hs:253-256 */
ora_sm* ora_sm_matmat(const ora_sm* a, const ora_sm* b, int* err);        /* (##)  SpMatrix.hs:768-811 */
int     ora_sm_equal(const ora_sm* a, const ora_sm* b);                   /* derived Eq */

/* ---- preconditioners and triangular solves (Sparse.hs:670-811) ---- */
ora_sm* ora_sm_extract_tri(const ora_sm* a, int which);                   /* extractSubDiag (-1) / extractDiag (0) / extractSuperDiag (+1)  SpMatrix.hs:306-315 */
ora_sm* ora_sm_eye(int64_t n);                                            /* eye               SpMatrix.hs:128-135 */
ora_sm* ora_sm_scale_right(const ora_sm* a, double n);                    /* scale n = fmap (* n)  Class.hs:179-180 */
ora_sm* ora_sm_negate(const ora_sm* a);                                   /* negateV           SpMatrix.hs:79 */
ora_sm* ora_sm_add(const ora_sm* a, const ora_sm* b);                     /* (^+^)             SpMatrix.hs:71-78 */
ora_sm* ora_sm_sub(const ora_sm* a, const ora_sm* b);                     /* (^-^)     Class.hs:68-69 */
ora_sm* ora_jacobi_pre(const ora_sm* a);                                  /* jacobiPre         Sparse.hs:686-687 */
int     ora_mssor_pre(const ora_sm* aa, double omega, ora_sm** l, ora_sm** r); /* mSsorPre     Sparse.hs:713-721 */
/* lu (Doolittle, Sparse.hs:489-538) and ilu0Pre = lu followed by a mask on aa's stored positions (Sparse.hs:696-706);
 * ORA_ERR_NEEDS_PIVOTING with *bad = the pivot whose u_jj is nearZero.  O(n^3): small matrices only. */
int     ora_lu(const ora_sm* aa, ora_sm** l, ora_sm** u, int64_t* bad);
int     ora_ilu0_pre(const ora_sm* aa, ora_sm** l, ora_sm** u, int64_t* bad);
/* *err: ORA_ERR_NEEDS_PIVOTING (with *bad_row = the row whose diagonal is nearZero) or ORA_ERR_OOB_INDEX (the
 * `@@` lookup past the matrix that a system of dimension 1 runs into). */
ora_sv* ora_tri_lower_solve(const ora_sm* ll, const ora_sv* b, int* err, int64_t* bad_row);  /* triLowerSolve Sparse.hs:750-777 */
ora_sv* ora_tri_upper_solve(const ora_sm* uu, const ora_sv* w, int* err, int64_t* bad_row);  /* triUpperSolve Sparse.hs:784-811 */

/* ---- Krylov states (Sparse.hs:855-981) ---- */
typedef struct { ora_sv *x, *r, *p, *u; } ora_krylov;   /* BICGSTAB: x r p; CGS: x r p u; CGNE: x r p */
void ora_krylov_free(ora_krylov*);
ora_krylov* ora_bicgs_init(const ora_sm* a, const ora_sv* b, const ora_sv* x0);                  /* :965-968 */
ora_krylov* ora_bicgstab_step(const ora_sm* a, const ora_sv* r0hat, const ora_krylov* st);     /* :970-981 */
ora_krylov* ora_cgs_init(const ora_sm* a, const ora_sv* b, const ora_sv* x0);                    /* :923-926 */
ora_krylov* ora_cgs_step(const ora_sm* a, const ora_sv* rhat, const ora_krylov* st);           /* :928-939 */
ora_krylov* ora_cgne_init(const ora_sm* a, const ora_sv* b, const ora_sv* x0);                   /* :862-866 */
ora_krylov* ora_cgne_step(const ora_sm* a, const ora_krylov* st);                              /* :868-878 */

/* linSolve0 (Sparse.hs:1016-1072).  nits/tol_abs/tol_rel default to 200/1e-6/1e-4 when <= 0.
 * Returns x; *iters = number of steps taken; res_hist (nullable, >= nits doubles) = true residual per step. */
ora_sv* ora_linsolve0(int method, const ora_sm* a, const ora_sv* b, const ora_sv* x0,
                      int nits, double tol_abs, double tol_rel,
                      int* iters, double* res_hist, int* err);

/* arnoldi (Sparse.hs:630-667).  Q returned dense column-major n x (*ncols_q); H dense column-major
 * (nmax+1) x nmax with nmax = *nmax_out.  Caller provides q_out (n*(kmax+1)) and h_out ((kmax+1)*kmax)
 * where kmax = max(kn, 1) bound used by the caller (see ora_arnoldi_bound). */
int ora_arnoldi(const ora_sm* a, const ora_sv* b, int kn, int max_steps,
                double* q_out, double* h_out, int* ncols_q, int* nmax_out);

/* ---- synthetic workloads (include/sla_synth.h) ---- */
ora_sm* ora_synth_matrix(int kind, int64_t n, int k, uint64_t seed, int64_t band);
ora_sv* ora_synth_vector(uint64_t seed, int64_t n);
void    ora_synth_row(int kind, int64_t n, int k, uint64_t seed, int64_t band, int64_t i,
                      int64_t* cols, double* vals, int* count);

/* ---- timing helper for the CPU baseline: y = A #> x repeated, rows optionally split over threads
 * (each row's arithmetic identical to ora_sm_matvec). Returns seconds per matvec. ---- */
double ora_time_matvec(const ora_sm* a, const ora_sv* x, int reps, int threads, double* checksum);
double ora_time_bicgstab(const ora_sm* a, const ora_sv* b, const ora_sv* x0, int steps, double* checksum);

#ifdef __cplusplus
}
#endif
#endif
