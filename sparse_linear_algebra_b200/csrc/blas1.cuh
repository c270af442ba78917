// blas1.cuh — the dense-vector half of the hot path: (^+^) (^-^) (.*) (<.>) norm2 of the reference
// (src/Data/Sparse/SpVector.hs:107-129) and the fused update kernels of the Krylov recurrences.
//
// One generic kernel: 128-bit (double2) grid-stride loads of NIN inputs, an elementwise functor applied
// with explicit __dmul_rn/__dadd_rn/__dsub_rn in the reference's association order (so every elementwise
// result is bit-identical to the Haskell expression given the same scalars), 128-bit stores of NOUT
// outputs, and up to NRED fused dot products reduced deterministically over the grid (last CTA sums the
// per-CTA partials in a fixed order and post-processes the Krylov scalars on the device, so alpha / omega /
// beta never visit the host).
#pragma once
#include "common.cuh"

#define EW_THREADS 256
#define EW_MAX_BLOCKS (SLA_NUM_SMS * 8)

template <int N> struct Ptrs { double* p[N > 0 ? N : 1]; };

template <class OP>
__global__ void __launch_bounds__(EW_THREADS)
ew_kernel(OP op, int64_t n, Ptrs<OP::NIN> in, Ptrs<OP::NOUT> out, double* scal, double* partials,
          unsigned int* counter, int fin, int dst, sla_p2p_args pa) {
  constexpr int NIN = OP::NIN, NOUT = OP::NOUT, NRED = OP::NRED;
  __shared__ double red[(NRED > 0 ? NRED : 1) * 32];
  op.prep(scal);
  double acc[NRED > 0 ? NRED : 1];
#pragma unroll
  for (int k = 0; k < (NRED > 0 ? NRED : 1); ++k) acc[k] = 0.0;
  const int64_t n2 = n >> 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = gtid; i < n2; i += stride) {
    double2 a[NIN > 0 ? NIN : 1];
#pragma unroll
    for (int k = 0; k < NIN; ++k) a[k] = reinterpret_cast<const double2*>(in.p[k])[i];
    double ia[NIN > 0 ? NIN : 1], oa[NOUT > 0 ? NOUT : 1], ob[NOUT > 0 ? NOUT : 1];
#pragma unroll
    for (int k = 0; k < NIN; ++k) ia[k] = a[k].x;
    op.f(ia, oa, acc);
#pragma unroll
    for (int k = 0; k < NIN; ++k) ia[k] = a[k].y;
    op.f(ia, ob, acc);
#pragma unroll
    for (int k = 0; k < NOUT; ++k) reinterpret_cast<double2*>(out.p[k])[i] = make_double2(oa[k], ob[k]);
  }
  if ((n & 1) && gtid == 0) {
    double ia[NIN > 0 ? NIN : 1], oa[NOUT > 0 ? NOUT : 1];
#pragma unroll
    for (int k = 0; k < NIN; ++k) ia[k] = in.p[k][n - 1];
    op.f(ia, oa, acc);
#pragma unroll
    for (int k = 0; k < NOUT; ++k) out.p[k][n - 1] = oa[k];
  }
  if (NRED > 0) {
    block_sum<(NRED > 0 ? NRED : 1)>(acc, red);
    grid_reduce_finish<(NRED > 0 ? NRED : 1)>(acc, partials, counter, scal, fin, dst, red, pa);
  }
}

template <class OP>
static inline sla_status ew_launch(sla_ctx* c, OP op, int64_t n, Ptrs<OP::NIN> in, Ptrs<OP::NOUT> out,
                                   int fin = FIN_STORE, int dst = S_TMP0) {
  SLA_GUARD(c);
  int64_t blocks = ((n >> 1) + EW_THREADS - 1) / EW_THREADS;
  if (blocks < 1) blocks = 1;
  if (blocks > EW_MAX_BLOCKS) blocks = EW_MAX_BLOCKS;
  sla_red_plan rp = sla_red_begin(c, fin, OP::NRED > 0 ? OP::NRED : P2P_MAX_NV + 1);   // kernels without a reduction never take a sequence number
  if (OP::NRED == 0) { rp.fin = fin; rp.host = false; }
  ew_kernel<OP><<<(unsigned)blocks, EW_THREADS, 0, c->stream>>>(op, n, in, out, c->scal, c->partials, c->counter, rp.fin, dst, rp.pa);
  SLA_LAUNCH_CHECK(c);
  if (OP::NRED > 0) SLA_TRY(sla_red_end(c, rp, OP::NRED, fin, dst));
  return SLA_OK;
}

// ---- elementwise functors.  f(in, out, red) handles ONE element. ------------------------------------
#define OP_HEAD(nin, nout, nred) static constexpr int NIN = nin, NOUT = nout, NRED = nred
#define DEV __device__ __forceinline__

struct OpAdd {        // z = x ^+^ y
  OP_HEAD(2, 1, 0);
  DEV void prep(const double*) {}
  DEV void f(const double* i, double* o, double*) const { o[0] = __dadd_rn(i[0], i[1]); }
};
struct OpSub {        // z = x ^-^ y = x ^+^ negateV y
  OP_HEAD(2, 1, 0);
  DEV void prep(const double*) {}
  DEV void f(const double* i, double* o, double*) const { o[0] = __dsub_rn(i[0], i[1]); }
};
struct OpScale {      // z = a .* x
  OP_HEAD(1, 1, 0);
  double a;
  DEV void prep(const double*) {}
  DEV void f(const double* i, double* o, double*) const { o[0] = __dmul_rn(a, i[0]); }
};
struct OpScaleDev {   // z = scal[slot] .* x     (normalize2: slot = S_INVN)
  OP_HEAD(1, 1, 0);
  int slot; double a;
  DEV void prep(const double* s) { a = s[slot]; }
  DEV void f(const double* i, double* o, double*) const { o[0] = __dmul_rn(a, i[0]); }
};
struct OpAxpy {       // z = y ^+^ (a .* x) ; in = {x, y}
  OP_HEAD(2, 1, 0);
  double a;
  DEV void prep(const double*) {}
  DEV void f(const double* i, double* o, double*) const { o[0] = __dadd_rn(i[1], __dmul_rn(a, i[0])); }
};
struct OpDot {        // red0 = x <.> y
  OP_HEAD(2, 0, 1);
  DEV void prep(const double*) {}
  DEV void f(const double* i, double*, double* r) const { r[0] += i[0] * i[1]; }
};
struct OpDot2 {       // red0 = a <.> b ; red1 = c <.> d
  OP_HEAD(4, 0, 2);
  DEV void prep(const double*) {}
  DEV void f(const double* i, double*, double* r) const { r[0] += i[0] * i[1]; r[1] += i[2] * i[3]; }
};
struct OpNorm2Sq {    // red0 = sum x_i ** 2
  OP_HEAD(1, 0, 1);
  DEV void prep(const double*) {}
  DEV void f(const double* i, double*, double* r) const { r[0] += i[0] * i[0]; }
};
struct OpSelfDot2 {   // red0 = x <.> x ; red1 = y <.> y     (CGNE alpha)
  OP_HEAD(2, 0, 2);
  DEV void prep(const double*) {}
  DEV void f(const double* i, double*, double* r) const { r[0] += i[0] * i[0]; r[1] += i[1] * i[1]; }
};
struct OpResidInit {  // r = b ^-^ ax ; p = r (; u = r) ; red0 = r <.> r      in = {b, ax}   bicgsInit/cgsInit
  OP_HEAD(2, 3, 1);
  DEV void prep(const double*) {}
  DEV void f(const double* i, double* o, double* r) const {
    const double v = __dsub_rn(i[0], i[1]);
    o[0] = v; o[1] = v; o[2] = v;
    r[0] += v * v;
  }
};

// ---- BiCGSTAB (bicgstabStep, Sparse.hs:970-981) ---------------------------------------------------------
struct OpBicgS {      // sj = r ^-^ (alphaj .* aap)          in = {r, aap}
  OP_HEAD(2, 1, 0);
  double alpha;
  DEV void prep(const double* s) { alpha = s[S_ALPHA]; }
  DEV void f(const double* i, double* o, double*) const { o[0] = __dsub_rn(i[0], __dmul_rn(alpha, i[1])); }
};
struct OpBicgXR {     // xj1 = x ^+^ (alphaj .* p) ^+^ (omegaj .* sj) ; rj1 = sj ^-^ (omegaj .* aasj) ; red0 = rj1 <.> r0hat
  OP_HEAD(5, 2, 1);   // in = {x, p, s, aas, r0hat} ; out = {x, r}
  double alpha, omega;
  DEV void prep(const double* s) { alpha = s[S_ALPHA]; omega = s[S_OMEGA]; }
  DEV void f(const double* i, double* o, double* r) const {
    o[0] = __dadd_rn(__dadd_rn(i[0], __dmul_rn(alpha, i[1])), __dmul_rn(omega, i[2]));
    const double rj = __dsub_rn(i[2], __dmul_rn(omega, i[3]));
    o[1] = rj;
    r[0] += rj * i[4];
  }
};
struct OpBicgP {      // pj1 = rj1 ^+^ (betaj .* (p ^-^ (omegaj .* aap)))      in = {r, p, aap} ; out = {p}
  OP_HEAD(3, 1, 0);
  double beta, omega;
  DEV void prep(const double* s) { beta = s[S_BETA]; omega = s[S_OMEGA]; }
  DEV void f(const double* i, double* o, double*) const {
    o[0] = __dadd_rn(i[0], __dmul_rn(beta, __dsub_rn(i[1], __dmul_rn(omega, i[2]))));
  }
};

// ---- CGS (cgsStep, Sparse.hs:928-939) ---------------------------------------------------------------
struct OpCgsQ {       // q = u ^-^ (alphaj .* aap) ; upq = u ^+^ q ; xj1 = x ^+^ (alphaj .* upq)
  OP_HEAD(3, 3, 0);   // in = {u, aap, x} ; out = {q, upq, x}
  double alpha;
  DEV void prep(const double* s) { alpha = s[S_ALPHA]; }
  DEV void f(const double* i, double* o, double*) const {
    const double q = __dsub_rn(i[0], __dmul_rn(alpha, i[1]));
    const double upq = __dadd_rn(i[0], q);
    o[0] = q; o[1] = upq;
    o[2] = __dadd_rn(i[2], __dmul_rn(alpha, upq));
  }
};
struct OpCgsR {       // rj1 = r ^-^ (alphaj .* aupq) ; red0 = rj1 <.> rhat        in = {r, aupq, rhat} ; out = {r}
  OP_HEAD(3, 1, 1);
  double alpha;
  DEV void prep(const double* s) { alpha = s[S_ALPHA]; }
  DEV void f(const double* i, double* o, double* r) const {
    const double rj = __dsub_rn(i[0], __dmul_rn(alpha, i[1]));
    o[0] = rj;
    r[0] += rj * i[2];
  }
};
struct OpCgsUP {      // uj1 = rj1 ^+^ (betaj .* q) ; pj1 = uj1 ^+^ (betaj .* (q ^+^ (betaj .* p)))
  OP_HEAD(3, 2, 0);   // in = {r, q, p} ; out = {u, p}
  double beta;
  DEV void prep(const double* s) { beta = s[S_BETA]; }
  DEV void f(const double* i, double* o, double*) const {
    const double u = __dadd_rn(i[0], __dmul_rn(beta, i[1]));
    o[0] = u;
    o[1] = __dadd_rn(u, __dmul_rn(beta, __dadd_rn(i[1], __dmul_rn(beta, i[2]))));
  }
};

// ---- CGNE (cgneStep, Sparse.hs:868-878) -------------------------------------------------------------
struct OpCgneXR {     // x1 = x ^+^ (alphai .* p) ; r1 = r ^-^ (alphai .* ap) ; red0 = r1 <.> r1
  OP_HEAD(4, 2, 1);   // in = {x, p, r, ap} ; out = {x, r}
  double alpha;
  DEV void prep(const double* s) { alpha = s[S_ALPHA]; }
  DEV void f(const double* i, double* o, double* r) const {
    o[0] = __dadd_rn(i[0], __dmul_rn(alpha, i[1]));
    const double r1 = __dsub_rn(i[2], __dmul_rn(alpha, i[3]));
    o[1] = r1;
    r[0] += r1 * r1;
  }
};
struct OpCgneP {      // p1 = (transpose aa #> r1) ^+^ (beta .* p)       in = {atr, p} ; out = {p}
  OP_HEAD(2, 1, 0);
  double beta;
  DEV void prep(const double* s) { beta = s[S_BETA]; }
  DEV void f(const double* i, double* o, double*) const { o[0] = __dadd_rn(i[0], __dmul_rn(beta, i[1])); }
};

// ---- diagonal solve (linSolve0 shortcut, Sparse.hs:1024-1025): x_i = recip a_ii * b_i -----------------
struct OpDiagSolve {  // in = {a_ii, b} ; out = {x}
  OP_HEAD(2, 1, 0);
  DEV void prep(const double*) {}
  DEV void f(const double* i, double* o, double*) const { o[0] = __dmul_rn(__ddiv_rn(1.0, i[0]), i[1]); }
};
