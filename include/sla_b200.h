/* sla_b200.h — C ABI of libsla_b200.so, the B200 (sm_100a) sparse Krylov backend that drops in behind
 * Numeric.LinearAlgebra.Sparse's operator surface (ocramz/sparse-linear-algebra, commit b940b12).
 *
 * The reference has no FFI; the boundary it exposes is the typeclass surface of
 * src/Numeric/LinearAlgebra/Class.hs as instantiated for SpVector Double / SpMatrix Double.  Each
 * entry point below names the reference function it replaces (paths under the reference root).
 * A Haskell shim binds these with `foreign import ccall safe` (hs/).
 *
 * Conventions
 *  - plain C: opaque handles, pointers and sizes; no C++ or torch types cross the boundary.
 *  - host pointers are BORROWED for the duration of one call; device memory never crosses.
 *  - every call returns sla_status; sla_last_error(ctx) gives the message (it mirrors the Show
 *    instance of the reference exception, src/Control/Exception/Common.hs:44-76).
 *  - one sla_ctx = one GPU = one host thread at a time (the reference is single-threaded and pure).
 *    Multi-GPU is one process (one ctx) per GPU; ranks are joined with sla_init_dist.
 *  - vectors are dense double[n] on the device: an absent SpVector key marshals as 0.0.
 *  - matrices are CSR (int32 row_ptr[m+1], int32 col_idx[nnz] ascending per row, double val[nnz]);
 *    field meaning follows vector/src/Data/Sparse/Internal/CSR.hs:38-50.
 *  - there is NO CPU fallback: without a CUDA device every compute entry returns SLA_ERR_CUDA.
 */
#ifndef SLA_B200_H
#define SLA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sla_ctx sla_ctx;
typedef struct sla_csr sla_csr;       /* SpMatrix Double   src/Data/Sparse/SpMatrix.hs:52-54 */
typedef struct sla_vec sla_vec;       /* SpVector Double   src/Data/Sparse/SpVector.hs:42-43 */
typedef struct sla_dense sla_dense;   /* dense column-major block (Krylov basis Q; `##` right operand) */
typedef struct sla_krylov sla_krylov; /* BICGSTAB / CGS / CGNE records   src/Numeric/LinearAlgebra/Sparse.hs:855, 921, 962 */

typedef enum {
  SLA_OK = 0,
  SLA_ERR_SIZE_MISMATCH = 1,      /* MatVecSizeMismatchException / error "matVec : mismatched dimensions"  Common.hs:250, Sparse.hs:637,1022 */
  SLA_ERR_OOB_INDEX = 2,          /* error "insertSpMatrix : index out of bounds"  SpMatrix.hs:205-208 */
  SLA_ERR_UNSUPPORTED_METHOD = 3, /* IterE "linSolve0" "Only BICGSTAB_, CGS_, and CGNE_ ..."  Sparse.hs:1031 */
  SLA_ERR_NOT_CONVERGED = 4,      /* informational only: the reference returns x silently after nits  Sparse.hs:1045 */
  SLA_ERR_BREAKDOWN = 5,          /* informational: Arnoldi breakdown, nearZero h_{i+1,i}  Sparse.hs:666 */
  SLA_ERR_CUDA = 6,
  SLA_ERR_COMM = 7,
  SLA_ERR_ALLOC = 8,
  SLA_ERR_INVALID = 9,
  SLA_ERR_NEEDS_PIVOTING = 10     /* MatrixException NeedsPivoting: a nearZero diagonal in a triangular solve  Control/Exception/Common.hs:57-61, Sparse.hs:757,791 */
} sla_status;

/* LinSolveMethod  Sparse.hs:1007-1012 (constructor order) */
typedef enum { SLA_GMRES_ = 0, SLA_CGNE_ = 1, SLA_BCG_ = 2, SLA_CGS_ = 3, SLA_BICGSTAB_ = 4 } sla_method;

/* fields of a Krylov record for sla_krylov_get: _x,_r,_p,_u  Sparse.hs:921, 962-963 */
typedef enum { SLA_FIELD_X = 0, SLA_FIELD_R = 1, SLA_FIELD_P = 2, SLA_FIELD_U = 3 } sla_field;

/* ---- context ------------------------------------------------------------------------------ */
sla_status  sla_init(int device, sla_ctx** out);
/* one rank of a row-partitioned multi-GPU job; nccl_id = 128 bytes from sla_nccl_unique_id on rank 0 */
sla_status  sla_nccl_unique_id(void* out128);
sla_status  sla_init_dist(int device, int rank, int world, const void* nccl_id128, sla_ctx** out);
void        sla_finalize(sla_ctx*);
const char* sla_last_error(const sla_ctx*);
const char* sla_version(void);
sla_status  sla_sync(sla_ctx*);
void*       sla_stream(sla_ctx*);                  /* the cudaStream_t every kernel of this ctx is launched on */
int         sla_rank(const sla_ctx*);
int         sla_world(const sla_ctx*);
int64_t     sla_launch_count(const sla_ctx*);      /* kernels launched by this ctx so far */
/* named integer switches; "skip_exchange" = 1 is a DIAGNOSTIC for bench.py (row-partitioned (#>) without its x exchange) */
sla_status  sla_set_option(sla_ctx*, const char* name, int64_t value);
/* page-locked host buffers for the host-pointer entry points (Haskell: wrap in a ForeignPtr with sla_host_free) */
sla_status  sla_host_alloc(sla_ctx*, int64_t bytes, void** out);
void        sla_host_free(void*);
sla_status  sla_timer_start(sla_ctx*);             /* CUDA events on the ctx stream */
sla_status  sla_timer_stop(sla_ctx*, float* ms);

/* ---- matrices (construction is setup, not the timed path) ---------------------------------- */
/* fromListSM (m,n) [(i,j,v)]: later duplicates overwrite, out-of-bounds is an error.  SpMatrix.hs:205-224 */
sla_status sla_csr_from_coo(sla_ctx*, int64_t m, int64_t n, int64_t nnz, const int64_t* i, const int64_t* j,
                            const double* v, sla_csr** out);
/* already-CSR input (columns ascending and unique per row; validated) */
sla_status sla_csr_from_csr(sla_ctx*, int64_t m, int64_t n, int64_t nnz, const int32_t* row_ptr,
                            const int32_t* col_idx, const double* val, sla_csr** out);
/* synthetic workloads of SURVEY.md §8(d), generated on the device from include/sla_synth.h */
sla_status sla_csr_generate(sla_ctx*, int kind, int64_t n, int nnz_per_row, uint64_t seed, int64_t band,
                            sla_csr** out);
/* ---- row-partitioned (multi-GPU) matrices: this rank holds global rows [row_lo, row_hi) with GLOBAL column
 * indices; vectors hold the matching slice.  sla_csr_col_range reports the columns the block references;
 * the host plans which contiguous pieces of x travel (sparse_linear_algebra_b200/dist.py) and installs the
 * plan with sla_csr_set_dist: dir 0 = receive global entries [goff, goff+count) from peer, 1 = send them. */
sla_status sla_csr_generate_rows(sla_ctx*, int kind, int64_t n, int nnz_per_row, uint64_t seed, int64_t band,
                                 int64_t row_lo, int64_t row_hi, sla_csr** out);
sla_status sla_csr_col_range(sla_ctx*, const sla_csr*, int64_t* lo, int64_t* hi);   /* hi < lo: no entries */
sla_status sla_csr_set_dist(sla_ctx*, sla_csr*, int64_t row0, int nseg, const int* dir, const int* peer,
                            const int64_t* goff, const int64_t* count, int allgather /* same value on EVERY rank */);
/* Peer-memory collectives over NVLink / NVSwitch (csrc/p2p.cu): the all-reduce behind every dot and the x exchange
 * before a row-partitioned (#>) run as single kernels that store straight into the peers' windows instead of
 * calling NCCL.  Set-up is host plumbing: every rank exports a 64-byte cudaIpcMemHandle_t, the host gathers the
 * `world` handles in rank order, every rank attaches, and the switch is thrown with the SAME value on every rank
 * (on = every rank exported and attached).  Context first, then each distributed matrix after sla_csr_set_dist. */
sla_status sla_p2p_export(sla_ctx*, void* handle64);
sla_status sla_p2p_attach(sla_ctx*, const void* handles /* world x 64 bytes */);
sla_status sla_p2p_enable(sla_ctx*, int on);
int        sla_p2p_enabled(const sla_ctx*);
/* LL halo exchange (mode 3): base[s] belongs to segment s of the plan given to sla_csr_set_dist — the compact offset of the
 * segment in this rank's halo buffer (receive) or in the destination's (send); capacity = entries of the LARGEST halo of the
 * job (the same value on every rank: all halo buffers have one size).  Call before sla_csr_p2p_export. */
sla_status sla_csr_set_halo(sla_ctx*, sla_csr*, int nseg, const int64_t* base, int64_t capacity);
sla_status sla_csr_p2p_export(sla_ctx*, sla_csr*, void* handle64);
sla_status sla_csr_p2p_attach(sla_ctx*, sla_csr*, const void* handles /* world x 64 bytes */);
sla_status sla_csr_p2p_enable(sla_ctx*, sla_csr*, int on /* 0 off, 1 push kernel, 2 copy engines + arrival-order panels, 3 LL halo, 4 copy-engine all-gather,
                                                              5 phased TMA push under rotated column panels */);
int        sla_csr_p2p_mode(const sla_csr*);      /* the mode in force (2 and 5 fall back to 4 / 1 when the plan is not dense / equal-block or x is small) */
/* Mode 5's panel schedule, pure host arithmetic: sizes[p] (room for 8) = column blocks in panel p (panel 0 = the own block, panel
 * p >= 1 = the next sizes[p] predecessors, exchanged in phase p); spec = "1,1,2"-style override or NULL for 1, 1, 2, 4, ...;
 * returns the number of panels. */
int        sla_p2p_phase_schedule(int world, const char* spec, int* sizes);
/* Test hook, one GPU: cut an ordinary matrix into mode 5's rotated panels as if its columns were `world` equal blocks and this GPU
 * held block `rank`; (#>) then runs panel by panel (own block first; results within the fp64 bound, not bit-identical).  world <= 1
 * removes the panels (one pass over all columns).  sla_csr_npanels: column panels of the plan in force (0 = one pass). */
sla_status sla_csr_debug_rot_panels(sla_ctx*, sla_csr*, int world, int rank, const char* spec);
/* Diagnostic, one GPU (scripts/prof_push_contention.py): mode 5's push kernels copying one vector into another of the same dimension
 * inside this GPU's memory from a high-priority side stream — start() orders the copy after what is queued on the ctx stream and
 * returns at once, join() makes the ctx stream wait for it; kind 0 = TMA bulk kernel, 1 = LSU kernel, ctas = resident CTAs. */
/* Host-only twin of the sliced-ELL band plan (csrc/spmv_bandsell.cuh; SLA_SPMV_BAND=3): builds the plan for a HOST CSR (int32
 * indices, columns ascending inside a row) and runs the kernel's loop nest on the CPU — no GPU, no context.  Test infrastructure for the
 * layout and the summation order.  stats[5] = cells, padded entries, descriptors, largest cell count of a row block, slices.
 * Returns 0; 1 when the matrix does not fit the plan; 2 on bad arguments. */
int        sla_debug_bsell_host(int m, int64_t n, const int32_t* row_ptr, const int32_t* col, const double* val, int R, int W, int threads,
                                const double* x, double* y, int64_t* stats);
typedef struct sla_debug_push sla_debug_push;
sla_status sla_debug_push_create(sla_ctx*, sla_vec* dst, const sla_vec* src, sla_debug_push** out);
sla_status sla_debug_push_start(sla_ctx*, sla_debug_push*, int ctas, int kind);
sla_status sla_debug_push_join(sla_ctx*, sla_debug_push*);
void       sla_debug_push_free(sla_debug_push*);
int        sla_csr_npanels(const sla_csr*);
/* transposeSM of a row-partitioned square matrix (all-to-all of entries; starts = world + 1 global row offsets, the same on
 * every rank): *out is this rank's row block of the transpose; give it an exchange plan like any block, then hand it to A
 * with sla_csr_attach_transpose so that (<#) and CGNE on the distributed A use it (A owns it afterwards). */
sla_status sla_csr_transpose_dist(sla_ctx*, const sla_csr* A, const int64_t* starts, sla_csr** out);
sla_status sla_csr_attach_transpose(sla_ctx*, sla_csr* A, sla_csr* T);
sla_status sla_vec_generate_slice(sla_ctx*, int64_t i0, int64_t n, uint64_t seed, sla_vec** out);
sla_status sla_csr_dims(const sla_csr*, int64_t* m, int64_t* n, int64_t* nnz);
sla_status sla_csr_to_host(sla_ctx*, const sla_csr*, int32_t* row_ptr, int32_t* col_idx, double* val);
sla_status sla_csr_transpose(sla_ctx*, const sla_csr*, sla_csr** out);         /* transposeSM  SpMatrix.hs:717-718 (bit-exact) */
sla_status sla_csr_is_diagonal(sla_ctx*, const sla_csr*, int* out);            /* isDiagonalSM SpMatrix.hs:411-415 */
int64_t    sla_csr_spmv_bytes(const sla_csr*);                                 /* algorithmic bytes of one (#>): 12 nnz + 20 n + 4 */
void       sla_csr_free(sla_csr*);

/* ---- vectors ------------------------------------------------------------------------------- */
sla_status sla_vec_create(sla_ctx*, int64_t n, sla_vec** out);                           /* zeros */
sla_status sla_vec_from_host(sla_ctx*, int64_t n, const double* x, sla_vec** out);       /* mkSpVR / fromListDenseSV  SpVector.hs:183-195 */
sla_status sla_vec_generate(sla_ctx*, int64_t n, uint64_t seed, sla_vec** out);          /* sla_synth_vec */
sla_status sla_vec_upload(sla_ctx*, sla_vec*, const double* x);                          /* overwrite from host */
sla_status sla_vec_to_host(sla_ctx*, const sla_vec*, double* x);                         /* toDenseListSV  SpVector.hs:300-301 */
sla_status sla_vec_copy(sla_ctx*, const sla_vec* src, sla_vec* dst);
sla_status sla_vec_fill(sla_ctx*, sla_vec*, double a);                                   /* constv  SpVector.hs:232-233 */
int64_t    sla_vec_dim(const sla_vec*);
void       sla_vec_free(sla_vec*);

/* ---- operator surface  (Class.hs:57-99, 126-153, 224-229) ---------------------------------- */
sla_status sla_spmv (sla_ctx*, const sla_csr* A, const sla_vec* x, sla_vec* y);          /* (#>) = matVecSD  Common.hs:242-250 */
sla_status sla_spmvT(sla_ctx*, const sla_csr* A, const sla_vec* x, sla_vec* y);          /* (<#) = vecMatSD  Common.hs:253-256 */
sla_status sla_dot  (sla_ctx*, const sla_vec* x, const sla_vec* y, double* out);         /* (<.>)  SpVector.hs:116-117 */
sla_status sla_norm2sq(sla_ctx*, const sla_vec* x, double* out);                         /* norm2Sq SpVector.hs:122 */
sla_status sla_norm2(sla_ctx*, const sla_vec* x, double* out);                           /* norm2 / norm2'  SpVector.hs:127-128 */
sla_status sla_vec_add  (sla_ctx*, const sla_vec* x, const sla_vec* y, sla_vec* z);      /* z = x ^+^ y   SpVector.hs:107-109 */
sla_status sla_vec_sub  (sla_ctx*, const sla_vec* x, const sla_vec* y, sla_vec* z);      /* z = x ^-^ y   Class.hs:68-69 */
sla_status sla_vec_scale(sla_ctx*, double a, const sla_vec* x, sla_vec* z);              /* z = a .* x    SpVector.hs:112-114 */
sla_status sla_vec_axpy (sla_ctx*, double a, const sla_vec* x, const sla_vec* y, sla_vec* z); /* z = y ^+^ (a .* x), rounded as written */
sla_status sla_vec_normalize2(sla_ctx*, const sla_vec* x, sla_vec* z);                   /* z = x ./ norm2 x  SpVector.hs:125, Class.hs:94-95 */
/* host-buffer form of (#>): x and y live in host memory, copies are part of the call (bench e2e leg) */
sla_status sla_spmv_host(sla_ctx*, const sla_csr* A, const double* x_host, double* y_host);

/* ---- Krylov  (Sparse.hs:855-981) ----------------------------------------------------------- */
sla_status sla_bicgstab_init(sla_ctx*, const sla_csr* A, const sla_vec* b, const sla_vec* x0, sla_krylov** st); /* bicgsInit :965-968 */
sla_status sla_bicgstab_step(sla_ctx*, const sla_csr* A, const sla_vec* r0hat, sla_krylov* st);                 /* bicgstabStep :970-981 */
sla_status sla_cgs_init(sla_ctx*, const sla_csr* A, const sla_vec* b, const sla_vec* x0, sla_krylov** st);      /* cgsInit :923-926 */
sla_status sla_cgs_step(sla_ctx*, const sla_csr* A, const sla_vec* rhat, sla_krylov* st);                       /* cgsStep :928-939 */
sla_status sla_cgne_init(sla_ctx*, const sla_csr* A, const sla_vec* b, const sla_vec* x0, sla_krylov** st);     /* cgneInit :862-866 */
sla_status sla_cgne_step(sla_ctx*, const sla_csr* A, sla_krylov* st);                                           /* cgneStep :868-878 */
sla_status sla_krylov_get(sla_ctx*, const sla_krylov*, int field, double* host_out);                            /* _x / _r / _p / _u */
sla_status sla_krylov_view(sla_ctx*, const sla_krylov*, int field, const sla_vec** view);                       /* borrowed device view */
/* The steps above advance the record IN PLACE.  The reference's steps are pure functions (`iterate (bicgstabStep aa r0hat)
 * st0 !! 20`, README.md:208, keeps st0 usable): sla_krylov_clone makes the deep copy a pure `step` needs (clone, then advance
 * the clone — what hs/Numeric/LinearAlgebra/Sparse/B200.hs does). */
sla_status sla_krylov_clone(sla_ctx*, const sla_krylov* st, sla_krylov** out);
void       sla_krylov_free(sla_krylov*);

typedef struct {
  int    max_iters;      /* nits   = 200   Sparse.hs:1034 */
  double tol_abs;        /* tolAbs = 1e-6  Sparse.hs:1035 */
  double tol_rel;        /* tolRel = 1e-4  Sparse.hs:1036 */
  int    true_residual;  /* 1: ||A x - b|| recomputed (reference behaviour, Sparse.hs:1041); 0: recurrence residual ||r|| */
  int    check_every;    /* 1: test every iteration (reference); sla_gmres only: < 0 = no stopping test, run max_iters steps */
} sla_solve_opts;
void sla_solve_opts_default(sla_solve_opts*);

/* linSolve0 method aa b x0  (Sparse.hs:1016-1072).  x receives the solution; *iters the number of steps taken;
 * *resnorm the last residual norm tested.  Reaching max_iters returns SLA_OK (the reference returns x silently). */
sla_status sla_linsolve0(sla_ctx*, int method, const sla_csr* A, const sla_vec* b, const sla_vec* x0,
                         const sla_solve_opts*, sla_vec* x, int* iters, double* resnorm);
sla_status sla_linsolve0_host(sla_ctx*, int method, const sla_csr* A, const double* b_host, const double* x0_host,
                              const sla_solve_opts*, double* x_host, int* iters, double* resnorm);

/* arnoldi aa b kn  (Sparse.hs:630-667).  Q: n x (*nmax + 1) dense column-major on the device (sla_dense);
 * H: (nmax+1) x nmax column-major, written to the host buffer h_host (capacity (kn+1)*kn doubles, kn >= 2).
 * Returns SLA_ERR_BREAKDOWN (informational, outputs valid) if nearZero h_{i+1,i} stopped it early. */
sla_status sla_arnoldi(sla_ctx*, const sla_csr* A, const sla_vec* b, int kn, sla_dense** Q, double* h_host, int* nmax);
/* restarted GMRES(restart) built on the same Arnoldi kernels (the reference's gmres is commented out,
 * Sparse.hs:837-848; `<\>` was meant to call it, Sparse.hs:1082-1088). */
sla_status sla_gmres(sla_ctx*, const sla_csr* A, const sla_vec* b, const sla_vec* x0, int restart,
                     const sla_solve_opts*, sla_vec* x, int* iters, double* resnorm);

/* ---- (##) with a dense right operand  (matMat_ AB, SpMatrix.hs:768-811) ---------------------
 * B (n x k) and C (m x k) are ROW-major dense blocks.  SLA_F64: products rounded once and summed in ascending
 * column order — bit-identical to the reference.  SLA_BF16: A values and B in bf16, fp32 accumulation, C bf16
 * (BASELINE config 5). */
typedef enum { SLA_F64 = 0, SLA_BF16 = 1 } sla_dtype;
sla_status sla_dense_create(sla_ctx*, int64_t rows, int64_t cols, int dtype, sla_dense** out);
sla_status sla_dense_from_host(sla_ctx*, int64_t rows, int64_t cols, const double* rowmajor, int dtype, sla_dense** out);
sla_status sla_dense_generate(sla_ctx*, int64_t rows, int64_t cols, uint64_t seed, int dtype, sla_dense** out); /* synthetic, on-device */
sla_status sla_dense_to_host_f64(sla_ctx*, const sla_dense*, double* rowmajor_out);
sla_status sla_spmm_dense(sla_ctx*, const sla_csr* A, const sla_dense* B, sla_dense* C);
/* the rest of MatrixRing (Class.hs:195-207; instance SpMatrix.hs:751-773) for a dense right operand, single GPU:
 *   (##^)  a ## transpose b : Bt is the k x n row-major block whose TRANSPOSE is multiplied (matMat_ ABt, SpMatrix.hs:787-791)
 *   (#^#)  transpose a ## b : B is m x k, C is n x k; A's transpose is built once and cached (as for (<#))
 *   normFrobenius = sqrt (trace (m ##^ m)) : per row the left fold of a_ij * a_ij, then the sum over the rows (SpMatrix.hs:751-752) */
sla_status sla_spmm_dense_abt(sla_ctx*, const sla_csr* A, const sla_dense* Bt, sla_dense* C);
sla_status sla_spmm_dense_atb(sla_ctx*, const sla_csr* A, const sla_dense* B, sla_dense* C);
sla_status sla_csr_norm_frobenius(sla_ctx*, const sla_csr* A, double* out);

/* ---- preconditioners and triangular solves -------------------------------
 * Single GPU (a row-partitioned matrix returns SLA_ERR_INVALID).  New matrices are owned by the caller.
 *
 * sla_csr_diag_partitions : diagPartitions aa = (extractSubDiag, extractDiag, extractSuperDiag)      Sparse.hs:673-679
 * sla_jacobi_pre          : jacobiPre x = recip <$> extractDiag x                                   Sparse.hs:686-687
 * sla_mssor_pre           : mSsorPre aa omega = (l, r),  l = (eye n ^-^ scale omega e) ## reciprocal d,
 *                           r = d ^-^ scale omega f                                                  Sparse.hs:713-721
 *     Values are bit-identical to the reference's.  The reference's (##) also stores an explicit 0 for every
 *     (row, column) pair whose intersection is empty (l has n x n stored entries); those are not materialised.
 * sla_tri_lower_solve     : triLowerSolve ll b — forward substitution, w_i = (b_i - sum_{j<i asc} l_ij w_j) / l_ii,
 *                           result passed through sparsifySV (|w_i| <= 1e-12 becomes 0)              Sparse.hs:750-777
 * sla_tri_upper_solve     : triUpperSolve uu w — backward substitution, x_i = (w_i - sum_{j>i asc} u_ij x_j) / u_ii
 *                                                                                                   Sparse.hs:784-811
 *     Only the named triangle and the diagonal of the matrix are read (as in the reference), so a general matrix
 *     may be passed for a Gauss-Seidel sweep.  Results are bit-identical to the reference's evaluation order.
 *     A nearZero or missing diagonal entry returns SLA_ERR_NEEDS_PIVOTING (sla_last_error names the row the sweep
 *     meets first).  A system of dimension 1 returns SLA_ERR_OOB_INDEX: the reference's loop steps before it tests
 *     and looks up (1,1) resp. (-1,-1) with the bounds-checked `@@` (Iterative.hs:272-282, SpMatrix.hs:108-109).
 *     The first solve with a matrix builds (and caches in the matrix) a level schedule of its rows.
 * sla_tri_analysis        : builds that schedule now; reports the number of dependency levels and the stored
 *                           entries of the triangle including the diagonal (what one solve reads). */
sla_status sla_csr_diag_partitions(sla_ctx*, const sla_csr* A, sla_csr** E, sla_csr** D, sla_csr** F);
sla_status sla_jacobi_pre(sla_ctx*, const sla_csr* A, sla_csr** M);
sla_status sla_mssor_pre(sla_ctx*, const sla_csr* A, double omega, sla_csr** L, sla_csr** R);
/* ilu0Pre aa = (l, u) with holes (Sparse.hs:696-706): the reference runs the COMPLETE Doolittle `lu` (Sparse.hs:489-538) and then
 * drops the entries of L and U where aa stores nothing — not the incomplete recurrence of the literature.  The same recurrences,
 * in the same order, on a dense work area: bit-identical to the reference's evaluation, n <= 4096 (its algorithm is O(n^3)).
 * A nearZero pivot returns SLA_ERR_NEEDS_PIVOTING ("solveForLij : U(j,j) is close to 0 ..."). */
sla_status sla_ilu0_pre(sla_ctx*, const sla_csr* A, sla_csr** L, sla_csr** U);
sla_status sla_tri_lower_solve(sla_ctx*, const sla_csr* L, const sla_vec* b, sla_vec* w);
sla_status sla_tri_upper_solve(sla_ctx*, const sla_csr* U, const sla_vec* w, sla_vec* x);
sla_status sla_tri_analysis(sla_ctx*, const sla_csr* A, int upper, int* nlevels, int64_t* nnz_tri);

/* ---- dense blocks -------------------------------------------------------------------------- */
sla_status sla_dense_dims(const sla_dense*, int64_t* rows, int64_t* cols);
sla_status sla_dense_to_host(sla_ctx*, const sla_dense*, double* out_colmajor);
sla_status sla_dense_column(sla_ctx*, const sla_dense* Q, int64_t j, sla_vec** out);   /* column j of the Arnoldi basis as a new vector (extractCol, SpMatrix.hs:329-337) */
void       sla_dense_free(sla_dense*);

/* ---- one host process, several GPUs ------------------------------------------------------------------------
 * The reference is a single-threaded pure library; its natural caller is ONE process.  sla_init_multi(n_gpus, device_ids) owns
 * one context per GPU (a worker thread each) and exposes GLOBAL objects: a matrix is row-partitioned over the GPUs (contiguous
 * blocks, rows n*p/P), a vector is the concatenation of the ranks' slices; the caller never sees ranks.  Every call below issues
 * the matching single-rank call on all GPUs at once (those calls are collective: the x exchange and the all-reduced dots of
 * dist.cu / p2p.cu).  Semantics, error codes and messages are those of the single-GPU entry points above.
 * (sla_init_dist remains for jobs that already run one process per GPU, e.g. under torchrun.)                                  */
typedef struct sla_mctx sla_mctx;
typedef struct sla_mcsr sla_mcsr;
typedef struct sla_mvec sla_mvec;
typedef struct sla_mkrylov sla_mkrylov;
typedef struct sla_mdense sla_mdense;
sla_status  sla_init_multi(int n_gpus, const int* device_ids /* nullable: 0 .. n_gpus-1 */, sla_mctx** out);
void        sla_finalize_multi(sla_mctx*);
const char* sla_multi_last_error(const sla_mctx*);
int         sla_multi_world(const sla_mctx*);
sla_ctx*    sla_multi_ctx(sla_mctx*, int rank);            /* the per-GPU context (diagnostics: sla_launch_count, sla_p2p_enabled) */
sla_status sla_multi_csr_generate(sla_mctx*, int kind, int64_t n, int nnz_per_row, uint64_t seed, int64_t band, sla_mcsr** out);
sla_status sla_multi_csr_from_csr(sla_mctx*, int64_t m, int64_t n, int64_t nnz, const int32_t* row_ptr, const int32_t* col_idx,
                                  const double* val, sla_mcsr** out);                      /* global CSR in host memory */
sla_status sla_multi_csr_dims(const sla_mcsr*, int64_t* m, int64_t* n, int64_t* nnz);
void       sla_multi_csr_free(sla_mcsr*);
sla_status sla_multi_vec_create(sla_mctx*, int64_t n, sla_mvec** out);
sla_status sla_multi_vec_from_host(sla_mctx*, int64_t n, const double* x, sla_mvec** out);
sla_status sla_multi_vec_generate(sla_mctx*, int64_t n, uint64_t seed, sla_mvec** out);
sla_status sla_multi_vec_to_host(sla_mctx*, const sla_mvec*, double* x);
sla_status sla_multi_vec_copy(sla_mctx*, const sla_mvec* src, sla_mvec* dst);
int64_t    sla_multi_vec_dim(const sla_mvec*);
void       sla_multi_vec_free(sla_mvec*);
sla_status sla_multi_spmv(sla_mctx*, const sla_mcsr* A, const sla_mvec* x, sla_mvec* y);                       /* (#>) */
sla_status sla_multi_dot(sla_mctx*, const sla_mvec* x, const sla_mvec* y, double* out);                        /* (<.>) */
sla_status sla_multi_norm2(sla_mctx*, const sla_mvec* x, double* out);
sla_status sla_multi_vec_axpy(sla_mctx*, double a, const sla_mvec* x, const sla_mvec* y, sla_mvec* z);         /* z = y ^+^ (a .* x) */
sla_status sla_multi_vec_scale(sla_mctx*, double a, const sla_mvec* x, sla_mvec* z);                           /* z = a .* x */
sla_status sla_multi_bicgstab_init(sla_mctx*, const sla_mcsr* A, const sla_mvec* b, const sla_mvec* x0, sla_mkrylov** st);
sla_status sla_multi_bicgstab_step(sla_mctx*, const sla_mcsr* A, const sla_mvec* r0hat, sla_mkrylov* st);
sla_status sla_multi_cgs_init(sla_mctx*, const sla_mcsr* A, const sla_mvec* b, const sla_mvec* x0, sla_mkrylov** st);
sla_status sla_multi_cgs_step(sla_mctx*, const sla_mcsr* A, const sla_mvec* rhat, sla_mkrylov* st);
sla_status sla_multi_krylov_clone(sla_mctx*, const sla_mkrylov* st, sla_mkrylov** out);
sla_status sla_multi_krylov_get(sla_mctx*, const sla_mkrylov*, int field, double* host_out /* n doubles */);
void       sla_multi_krylov_free(sla_mkrylov*);
sla_status sla_multi_linsolve0(sla_mctx*, int method /* BICGSTAB_ or CGS_ (CGNE_ needs the distributed transpose: sla_init_dist path) */,
                               const sla_mcsr* A, const sla_mvec* b, const sla_mvec* x0, const sla_solve_opts*, sla_mvec* x, int* iters,
                               double* resnorm);
sla_status sla_multi_gmres(sla_mctx*, const sla_mcsr* A, const sla_mvec* b, const sla_mvec* x0, int restart, const sla_solve_opts*,
                           sla_mvec* x, int* iters, double* resnorm);
sla_status sla_multi_arnoldi(sla_mctx*, const sla_mcsr* A, const sla_mvec* b, int kn, sla_mdense** Q, double* h_host, int* nmax);
sla_status sla_multi_dense_to_host(sla_mctx*, const sla_mdense* Q, double* out_colmajor /* n x (nmax + 1) */);
void       sla_multi_dense_free(sla_mdense*);

#ifdef __cplusplus
}
#endif
#endif /* SLA_B200_H */
