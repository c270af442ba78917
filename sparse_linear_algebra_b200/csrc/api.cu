// api.cu — context, handles and the operator surface of the C ABI (include/sla_b200.h).
#include "blas1.cuh"

#include <math.h>
#include <stdlib.h>
#include <new>

#define SLA_VERSION_STR "sla-b200 0.1 (sm_100a)"

extern "C" const char* sla_version(void) { return SLA_VERSION_STR; }

static char g_init_err[512] = "";

extern "C" const char* sla_last_error(const sla_ctx* c) { return c ? c->err : g_init_err; }

static sla_status ctx_create(int device, int rank, int world, sla_ctx** out) {
  if (!out) return SLA_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    // no CPU fallback: without a device the library refuses to work
    snprintf(g_init_err, sizeof(g_init_err), "sla_init: no CUDA device available (%s); this library has no CPU path",
             e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return SLA_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    snprintf(g_init_err, sizeof(g_init_err), "sla_init: device %d out of range (have %d)", device, ndev);
    return SLA_ERR_INVALID;
  }
  sla_ctx* c = new (std::nothrow) sla_ctx();
  if (!c) return SLA_ERR_ALLOC;
  memset(c, 0, sizeof(*c));
  c->device = device; c->rank = rank; c->world = world;
  c->spmv_hints = 3;
  if (const char* h = getenv("SLA_SPMV_HINTS")) c->spmv_hints = atoi(h);
  if (const char* h = getenv("SLA_SPMV_TMA")) c->spmv_tma = atoi(h);
  if (const char* h = getenv("SLA_SPMV_BULK")) c->spmv_bulk = atoi(h);
  if (const char* h = getenv("SLA_BSELL_VARIANT")) c->bsell_variant = atoi(h);
  cudaError_t ce = cudaSetDevice(device);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaEventCreate(&c->ev0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&c->ev1);
  if (ce == cudaSuccess) ce = cudaMalloc(&c->scal, sizeof(double) * SLA_SCAL_SLOTS);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(c->scal, 0, sizeof(double) * SLA_SCAL_SLOTS, c->stream);
  if (ce == cudaSuccess) ce = cudaMalloc(&c->partials, sizeof(double) * 4 * (size_t)SLA_MAX_PARTIALS);
  if (ce == cudaSuccess) ce = cudaMalloc(&c->counter, sizeof(unsigned int) * 4);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(c->counter, 0, sizeof(unsigned int) * 4, c->stream);
  if (ce == cudaSuccess) ce = cudaMallocHost(&c->h_scal, sizeof(double) * SLA_SCAL_SLOTS);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);
  if (ce != cudaSuccess) {                        // nothing of a half-built context survives
    snprintf(g_init_err, sizeof(g_init_err), "sla_init: CUDA error %s while creating the context on device %d", cudaGetErrorString(ce), device);
    cudaGetLastError();
    cudaFree(c->scal); cudaFree(c->partials); cudaFree(c->counter); cudaFreeHost(c->h_scal);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return SLA_ERR_CUDA;
  }
  *out = c;
  return SLA_OK;
}

extern "C" sla_status sla_init(int device, sla_ctx** out) { return ctx_create(device, 0, 1, out); }

sla_status sla_dist_attach(sla_ctx* c, const void* nccl_id128);   // dist.cu
void sla_dist_detach(sla_ctx* c);

extern "C" sla_status sla_init_dist(int device, int rank, int world, const void* nccl_id128, sla_ctx** out) {
  if (world < 1 || rank < 0 || rank >= world) return SLA_ERR_INVALID;
  SLA_TRY(ctx_create(device, rank, world, out));
  if (world > 1) {
    sla_status s = sla_dist_attach(*out, nccl_id128);
    if (s != SLA_OK) { snprintf(g_init_err, sizeof(g_init_err), "%s", (*out)->err); sla_finalize(*out); *out = nullptr; return s; }
  }
  return SLA_OK;
}

extern "C" void sla_finalize(sla_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  sla_p2p_free(c);
  cudaFree(c->bfull); c->bfull = nullptr; c->bfull_bytes = 0;
  cudaFree(c->dense_cache); c->dense_cache = nullptr; c->dense_cache_bytes = 0;
  for (int k = 0; k < c->n_parked; ++k) cudaFree(c->parked[k]);
  c->n_parked = 0;
  if (c->nccl) sla_dist_detach(c);
  sla_vec_free(c->scratch_x); sla_vec_free(c->scratch_y); sla_vec_free(c->scratch_r);
  if (c->copy_stream) {
    cudaStreamDestroy(c->copy_stream);
    for (int k = 0; k < SLA_MAX_PANELS + 8; ++k) cudaEventDestroy(c->ev_copy[k]);
  }
  cudaFree(c->scal); cudaFree(c->partials); cudaFree(c->counter); cudaFreeHost(c->h_scal);
  cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
  cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" sla_status sla_sync(sla_ctx* c) {
  if (!c) return SLA_ERR_INVALID;
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  return sla_p2p_check(c);            // a timed-out peer wait surfaces here as SLA_ERR_COMM
}
extern "C" void* sla_stream(sla_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int sla_rank(const sla_ctx* c) { return c ? c->rank : 0; }
extern "C" int sla_world(const sla_ctx* c) { return c ? c->world : 1; }
extern "C" int64_t sla_launch_count(const sla_ctx* c) { return c ? c->launches : 0; }

// Named integer switches of a context.  "skip_exchange" = 1: the row-partitioned (#>) launches its kernels without exchanging x
// first — a DIAGNOSTIC that lets bench.py time the kernels alone (exposed exchange time = step time - this); results are invalid.
extern "C" sla_status sla_set_option(sla_ctx* c, const char* name, int64_t value) {
  if (!c || !name) return SLA_ERR_INVALID;
  if (strcmp(name, "skip_exchange") == 0) { c->skip_exchange = value != 0; return SLA_OK; }
  if (strcmp(name, "spmv_hints") == 0) { c->spmv_hints = (int)value; return SLA_OK; }
  if (strcmp(name, "spmv_bulk") == 0) { c->spmv_bulk = value != 0; return SLA_OK; }
  if (strcmp(name, "bsell_variant") == 0) { c->bsell_variant = value; return SLA_OK; }
  return sla_fail(c, SLA_ERR_INVALID, "sla_set_option: unknown option");
}

extern "C" sla_status sla_host_alloc(sla_ctx* c, int64_t bytes, void** out) {
  if (!c || !out || bytes < 0) return SLA_ERR_INVALID;
  *out = nullptr;
  if (cudaMallocHost(out, (size_t)(bytes > 0 ? bytes : 1)) != cudaSuccess) {
    cudaGetLastError();
    return sla_fail(c, SLA_ERR_ALLOC, "cudaMallocHost failed");
  }
  return SLA_OK;
}
extern "C" void sla_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" sla_status sla_timer_start(sla_ctx* c) {
  if (!c) return SLA_ERR_INVALID;
  SLA_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  return SLA_OK;
}
extern "C" sla_status sla_timer_stop(sla_ctx* c, float* ms) {
  if (!c || !ms) return SLA_ERR_INVALID;
  SLA_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  SLA_CUDA(c, cudaEventSynchronize(c->ev1));
  SLA_CUDA(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return SLA_OK;
}

sla_status sla_read_scalars(sla_ctx* c, int first, int count, double* host_out) {
  SLA_GUARD(c);
  SLA_CUDA(c, cudaMemcpyAsync(c->h_scal + first, c->scal + first, sizeof(double) * (size_t)count,
                              cudaMemcpyDeviceToHost, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < count; ++k) host_out[k] = c->h_scal[first + k];
  return c->world > 1 ? sla_p2p_check(c) : SLA_OK;
}

// ---- vectors --------------------------------------------------------------------------------------

sla_status sla_vec_alloc(sla_ctx* c, int64_t n, sla_vec** out) {
  if (!c || !out || n < 0) return SLA_ERR_INVALID;
  SLA_GUARD(c);
  sla_vec* v = new (std::nothrow) sla_vec();
  if (!v) return sla_fail(c, SLA_ERR_ALLOC, "vec alloc");
  v->ctx = c; v->n = n; v->version = ++c->stamp; v->owns = true; v->d = nullptr;
  // round up so that 128-bit accesses of the last pair stay inside the allocation
  cudaError_t e = cudaMalloc(&v->d, sizeof(double) * (size_t)((n + 2) & ~(int64_t)1));
  if (e != cudaSuccess) { delete v; return sla_fail(c, SLA_ERR_ALLOC, "cudaMalloc failed for a vector"); }
  *out = v;
  return SLA_OK;
}

extern "C" sla_status sla_vec_create(sla_ctx* c, int64_t n, sla_vec** out) {
  SLA_TRY(sla_vec_alloc(c, n, out));
  SLA_CUDA(c, cudaMemsetAsync((*out)->d, 0, sizeof(double) * (size_t)((n + 2) & ~(int64_t)1), c->stream));
  return SLA_OK;
}

extern "C" sla_status sla_vec_upload(sla_ctx* c, sla_vec* v, const double* x) {
  if (!c || !v || (!x && v->n > 0)) return SLA_ERR_INVALID;
  SLA_CUDA(c, cudaMemcpyAsync(v->d, x, sizeof(double) * (size_t)v->n, cudaMemcpyHostToDevice, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));   // the host buffer is only borrowed for the call
  sla_touch(v);
  return SLA_OK;
}

extern "C" sla_status sla_vec_from_host(sla_ctx* c, int64_t n, const double* x, sla_vec** out) {
  SLA_TRY(sla_vec_create(c, n, out));
  return sla_vec_upload(c, *out, x);
}

extern "C" sla_status sla_vec_to_host(sla_ctx* c, const sla_vec* v, double* x) {
  if (!c || !v || (!x && v->n > 0)) return SLA_ERR_INVALID;
  SLA_CUDA(c, cudaMemcpyAsync(x, v->d, sizeof(double) * (size_t)v->n, cudaMemcpyDeviceToHost, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  return SLA_OK;
}

extern "C" sla_status sla_vec_copy(sla_ctx* c, const sla_vec* src, sla_vec* dst) {
  if (!c || !src || !dst) return SLA_ERR_INVALID;
  if (src->n != dst->n) return sla_fail(c, SLA_ERR_SIZE_MISMATCH, "vec_copy: dimensions differ");
  SLA_CUDA(c, cudaMemcpyAsync(dst->d, src->d, sizeof(double) * (size_t)src->n, cudaMemcpyDeviceToDevice, c->stream));
  sla_touch(dst);
  return SLA_OK;
}

__global__ void fill_kernel(double* d, int64_t n, double a) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = a;
}

extern "C" sla_status sla_vec_fill(sla_ctx* c, sla_vec* v, double a) {
  if (!c || !v) return SLA_ERR_INVALID;
  fill_kernel<<<SLA_NUM_SMS * 4, 256, 0, c->stream>>>(v->d, v->n, a);
  SLA_LAUNCH_CHECK(c);
  sla_touch(v);
  return SLA_OK;
}

extern "C" int64_t sla_vec_dim(const sla_vec* v) { return v ? v->n : -1; }

extern "C" void sla_vec_free(sla_vec* v) {
  if (!v) return;
  if (v->owns && v->d) { cudaStreamSynchronize(v->ctx->stream); cudaFree(v->d); }
  delete v;
}

// ---- operator surface -----------------------------------------------------------------------------

static sla_status check3(sla_ctx* c, const sla_vec* x, const sla_vec* y, const sla_vec* z, const char* what) {
  if (!c || !x || !y || !z) return SLA_ERR_INVALID;
  // the reference's liftU2 has no dimension check (SpVector.hs:62-64, result dim = max); a dense backend
  // cannot represent that, so mismatched dims are an error here.
  if (x->n != y->n || x->n != z->n) return sla_fail(c, SLA_ERR_SIZE_MISMATCH, what);
  return SLA_OK;
}

extern "C" sla_status sla_spmv(sla_ctx* c, const sla_csr* A, const sla_vec* x, sla_vec* y) {
  if (!c || !A || !x || !y) return SLA_ERR_INVALID;
  SLA_GUARD(c);
  if (csr_xdim(A) != x->n) {   // matVecSD | nc == n ... | otherwise = error   Common.hs:248-250
    snprintf(c->err, sizeof(c->err), "matVec : mismatched dimensions (%lld,%lld)", (long long)csr_xdim(A), (long long)x->n);
    return SLA_ERR_SIZE_MISMATCH;
  }
  if (A->m != y->n) return sla_fail(c, SLA_ERR_SIZE_MISMATCH, "matVec : output vector has the wrong dimension");
  if (x->d == y->d) return sla_fail(c, SLA_ERR_INVALID, "matVec : x and y must be distinct vectors");
  SLA_TRY(sla_spmv_launch(c, A, x->d, y->d, EPI_NONE, nullptr, nullptr, FIN_STORE, S_TMP0));
  sla_touch(y);
  return SLA_OK;
}

extern "C" sla_status sla_spmvT(sla_ctx* c, const sla_csr* A, const sla_vec* x, sla_vec* y) {
  if (!c || !A || !x || !y) return SLA_ERR_INVALID;
  if (A->dist && !A->T) return sla_fail(c, SLA_ERR_INVALID, "vecMat : attach the distributed transpose of the row-partitioned matrix first");
  if (A->m != x->n) {   // vecMatSD | n == nr ... | otherwise = error   Common.hs:254-256
    snprintf(c->err, sizeof(c->err), "vecMat : mismatching dimensions (%lld,%lld)", (long long)x->n, (long long)A->m);
    return SLA_ERR_SIZE_MISMATCH;
  }
  if (!A->T) {
    // the reference transposes on EVERY call (Common.hs:255); the transpose is built once and cached here
    sla_csr* t = nullptr;
    SLA_TRY(sla_csr_transpose(c, A, &t));
    const_cast<sla_csr*>(A)->T = t;
  }
  return sla_spmv(c, A->T, x, y);
}

extern "C" sla_status sla_dot(sla_ctx* c, const sla_vec* x, const sla_vec* y, double* out) {
  if (!c || !x || !y || !out) return SLA_ERR_INVALID;
  if (x->n != y->n) return sla_fail(c, SLA_ERR_SIZE_MISMATCH, "<.> : Incompatible dimensions");
  Ptrs<2> in{{x->d, y->d}}; Ptrs<0> o{};
  SLA_TRY(ew_launch(c, OpDot{}, x->n, in, o, FIN_STORE, S_TMP0));
  return sla_read_scalars(c, S_TMP0, 1, out);
}

extern "C" sla_status sla_norm2sq(sla_ctx* c, const sla_vec* x, double* out) {
  if (!c || !x || !out) return SLA_ERR_INVALID;
  Ptrs<1> in{{x->d}}; Ptrs<0> o{};
  SLA_TRY(ew_launch(c, OpNorm2Sq{}, x->n, in, o, FIN_STORE, S_TMP0));
  return sla_read_scalars(c, S_TMP0, 1, out);
}

extern "C" sla_status sla_norm2(sla_ctx* c, const sla_vec* x, double* out) {
  double s = 0;
  SLA_TRY(sla_norm2sq(c, x, &s));
  *out = sqrt(s);
  return SLA_OK;
}

extern "C" sla_status sla_vec_add(sla_ctx* c, const sla_vec* x, const sla_vec* y, sla_vec* z) {
  SLA_TRY(check3(c, x, y, z, "^+^ : dimensions differ"));
  Ptrs<2> in{{x->d, y->d}}; Ptrs<1> o{{z->d}};
  SLA_TRY(ew_launch(c, OpAdd{}, x->n, in, o));
  sla_touch(z);
  return SLA_OK;
}

extern "C" sla_status sla_vec_sub(sla_ctx* c, const sla_vec* x, const sla_vec* y, sla_vec* z) {
  SLA_TRY(check3(c, x, y, z, "^-^ : dimensions differ"));
  Ptrs<2> in{{x->d, y->d}}; Ptrs<1> o{{z->d}};
  SLA_TRY(ew_launch(c, OpSub{}, x->n, in, o));
  sla_touch(z);
  return SLA_OK;
}

extern "C" sla_status sla_vec_scale(sla_ctx* c, double a, const sla_vec* x, sla_vec* z) {
  SLA_TRY(check3(c, x, x, z, ".* : dimensions differ"));
  Ptrs<1> in{{x->d}}; Ptrs<1> o{{z->d}};
  OpScale op; op.a = a;
  SLA_TRY(ew_launch(c, op, x->n, in, o));
  sla_touch(z);
  return SLA_OK;
}

extern "C" sla_status sla_vec_axpy(sla_ctx* c, double a, const sla_vec* x, const sla_vec* y, sla_vec* z) {
  SLA_TRY(check3(c, x, y, z, "axpy : dimensions differ"));
  Ptrs<2> in{{x->d, y->d}}; Ptrs<1> o{{z->d}};
  OpAxpy op; op.a = a;
  SLA_TRY(ew_launch(c, op, x->n, in, o));
  sla_touch(z);
  return SLA_OK;
}

// normalize2 v = v ./ norm2 v = (recip (norm2 v)) .* v   SpVector.hs:125, Class.hs:94-95
extern "C" sla_status sla_vec_normalize2(sla_ctx* c, const sla_vec* x, sla_vec* z) {
  SLA_TRY(check3(c, x, x, z, "normalize2 : dimensions differ"));
  Ptrs<1> in{{x->d}}; Ptrs<0> o0{};
  SLA_TRY(ew_launch(c, OpNorm2Sq{}, x->n, in, o0, FIN_NORM_INV, 0));
  Ptrs<1> o{{z->d}};
  OpScaleDev op; op.slot = S_INVN; op.a = 0;
  SLA_TRY(ew_launch(c, op, x->n, in, o));
  sla_touch(z);
  return SLA_OK;
}

static sla_status scratch_vec(sla_ctx* c, sla_vec** slot, int64_t n) {
  if (*slot && (*slot)->n == n) return SLA_OK;
  sla_vec_free(*slot);
  *slot = nullptr;
  return sla_vec_alloc(c, n, slot);
}

// (#>) on host buffers: H2D of x, the SpMV kernel, D2H of y — all inside this call, on the ctx stream.
// Device staging vectors are cached in the ctx, so a steady-state call does no allocation.
extern "C" sla_status sla_spmv_host(sla_ctx* c, const sla_csr* A, const double* x_host, double* y_host) {
  if (!c || !A || !x_host || !y_host) return SLA_ERR_INVALID;
  SLA_GUARD(c);
  SLA_TRY(scratch_vec(c, &c->scratch_x, csr_xdim(A)));
  SLA_TRY(scratch_vec(c, &c->scratch_y, A->m));
  const bool uniform_panels = A->npanels < 2 || A->panel_width > 0;     // rotated panels (test hook below) have no upload order
  if (!A->dist && !c->spmv_tma && A->m > 0 && uniform_panels && !getenv("SLA_HOST_NO_PIPELINE"))
    return sla_spmv_host_pipelined(c, A, x_host, y_host, c->scratch_x->d, c->scratch_y->d);
  SLA_CUDA(c, cudaMemcpyAsync(c->scratch_x->d, x_host, sizeof(double) * (size_t)csr_xdim(A), cudaMemcpyHostToDevice, c->stream));
  SLA_TRY(sla_spmv(c, A, c->scratch_x, c->scratch_y));
  SLA_CUDA(c, cudaMemcpyAsync(y_host, c->scratch_y->d, sizeof(double) * (size_t)A->m, cudaMemcpyDeviceToHost, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  return SLA_OK;
}

// ---- dense blocks ---------------------------------------------------------------------------------

extern "C" sla_status sla_dense_dims(const sla_dense* d, int64_t* rows, int64_t* cols) {
  if (!d) return SLA_ERR_INVALID;
  if (rows) *rows = d->rows;
  if (cols) *cols = d->cols;
  return SLA_OK;
}

extern "C" sla_status sla_dense_to_host(sla_ctx* c, const sla_dense* d, double* out) {
  if (!c || !d || !out) return SLA_ERR_INVALID;
  if (d->rowmajor) return sla_fail(c, SLA_ERR_INVALID, "dense_to_host: block is row-major (use sla_dense_to_host_f64)");
  SLA_CUDA(c, cudaMemcpy2DAsync(out, sizeof(double) * (size_t)d->rows, d->d, sizeof(double) * (size_t)d->ld,
                                sizeof(double) * (size_t)d->rows, (size_t)d->cols, cudaMemcpyDeviceToHost, c->stream));
  SLA_CUDA(c, cudaStreamSynchronize(c->stream));
  return SLA_OK;
}

extern "C" void sla_dense_free(sla_dense* d) {
  if (!d) return;
  sla_pool_free(d->ctx, d->d, d->bytes);
  delete d;
}

// Test hook (tests/test_gpu_parity.py, one GPU): the rotated panels of the phased exchange on an ordinary matrix, as if its n columns
// were `world` equal blocks and this GPU held block `rank` — (#>) then multiplies panel by panel, own block first.  world <= 1
// removes the panels (one pass over all columns).
extern "C" sla_status sla_csr_debug_rot_panels(sla_ctx* c, sla_csr* A, int world, int rank, const char* spec) {
  if (!c || !A) return SLA_ERR_INVALID;
  SLA_GUARD(c);
  if (A->dist) return sla_fail(c, SLA_ERR_INVALID, "debug_rot_panels: not for row-partitioned matrices");
  if (world <= 1) { sla_csr_free_panels(A); return SLA_OK; }
  if (rank < 0 || rank >= world || A->n % world != 0) return sla_fail(c, SLA_ERR_INVALID, "debug_rot_panels: n must be a multiple of world, 0 <= rank < world");
  sla_rot_spec rs;
  int sizes[SLA_ROT_MAX];
  rs.n = A->n; rs.m = A->n / world; rs.own_end = (long long)(rank + 1) * rs.m; rs.kb[0] = 0;
  rs.P = sla_p2p_phase_schedule(world, spec, sizes);
  for (int i = 0; i < rs.P; ++i) rs.kb[i + 1] = rs.kb[i] + sizes[i];
  if (A->band) return sla_fail(c, SLA_ERR_INVALID, "debug_rot_panels: the matrix runs the band plan");
  return sla_csr_force_rot_panels(c, A, &rs);
}
extern "C" int sla_csr_npanels(const sla_csr* A) { return A ? A->npanels : 0; }
