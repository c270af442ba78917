/* sla_synth.h — specification of the synthetic inputs named in SURVEY.md §8(d).
 *
 * A counter-based hash RNG (splitmix64) so that any row of any synthetic matrix,
 * and any entry of any synthetic vector, can be regenerated independently on the
 * device (CUDA), on the host (C) and in the CPU oracle, bit for bit.  This header
 * is the single definition; it is plain C99 and compiles under nvcc as
 * __host__ __device__ code.
 *
 * The reference (ocramz/sparse-linear-algebra) has no generator of this kind: its
 * QuickCheck generators (test/LibSpec.hs:720-730, 773-780) draw sqrt(m*n) random
 * triples.  These inputs are the benchmark workloads of BASELINE.json, not a
 * restatement of reference code.
 */
#ifndef SLA_SYNTH_H
#define SLA_SYNTH_H

#include <stdint.h>

#ifdef __CUDACC__
#define SLA_HD __host__ __device__ __forceinline__
#else
#define SLA_HD static inline
#endif

/* matrix families */
#define SLA_GEN_UNIFORM   0 /* diag + (k-1) distinct hashed columns uniform over [0,n)           */
#define SLA_GEN_BANDED    1 /* diag + (k-1) distinct hashed columns within +-band of the row      */
#define SLA_GEN_LAPLACE2D 2 /* 5-point Laplacian on a band x band grid (n = band*band), Dirichlet */

#define SLA_GEN_BLOCK16   3 /* block-structured: every 16-row group holds k/16 full 16 x 16 blocks at hashed block columns */

#define SLA_SYNTH_MAX_K 128 /* max stored entries per generated row */

#define SLA_SALT_VAL 0xA5A5F00DBAADC0DEULL
#define SLA_SALT_VEC 0x5EEDC0FFEE15600DULL

SLA_HD uint64_t sla_splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

SLA_HD uint64_t sla_hash3(uint64_t seed, uint64_t a, uint64_t b) {
  return sla_splitmix64(sla_splitmix64(seed ^ (a * 0x9E3779B97F4A7C15ULL)) + b);
}

/* uniform double in [-1, 1): 53 hash bits, every step exact in IEEE-754 */
SLA_HD double sla_u11(uint64_t h) {
  double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
  return u * 2.0 - 1.0;
}

/* entry i of the synthetic dense vector with the given seed */
SLA_HD double sla_synth_vec(uint64_t seed, int64_t i) {
  return sla_u11(sla_hash3(seed ^ SLA_SALT_VEC, (uint64_t)i, 0));
}

/* number of stored entries row i of the family will have (no generation) */
SLA_HD int sla_synth_row_len(int kind, int64_t n, int k, int64_t band, int64_t i) {
  if (kind == SLA_GEN_LAPLACE2D) {
    int64_t g = band, ix = i % g, iy = i / g;
    return 1 + (iy > 0) + (ix > 0) + (ix < g - 1) + (iy < g - 1);
  }
  if (kind == SLA_GEN_BLOCK16) {
    int nb = k / 16;
    if (nb < 1) nb = 1;
    if (nb > SLA_SYNTH_MAX_K / 16) nb = SLA_SYNTH_MAX_K / 16;
    const int64_t nbc = n / 16;
    return (int)(16 * (nb < nbc ? nb : nbc));
  }
  int64_t lo = 0, hi = n - 1;
  if (kind == SLA_GEN_BANDED) {
    lo = i - band < 0 ? 0 : i - band;
    hi = i + band > n - 1 ? n - 1 : i + band;
  }
  int64_t avail = hi - lo + 1;
  if (k > SLA_SYNTH_MAX_K) k = SLA_SYNTH_MAX_K;
  return (int)(k < avail ? k : avail);
}

/* Generate row i: column indices ascending in cols[], values in vals[]; returns the
 * count (<= SLA_SYNTH_MAX_K).  UNIFORM/BANDED rows are strictly row-diagonally
 * dominant: a_ii = 1 + sum_j |a_ij| (sum taken in ascending column order). */
SLA_HD int sla_synth_row(int kind, int64_t n, int k, uint64_t seed, int64_t band,
                         int64_t i, int64_t* cols, double* vals) {
  if (kind == SLA_GEN_LAPLACE2D) {
    int64_t g = band, ix = i % g, iy = i / g;
    int c = 0;
    if (iy > 0)     { cols[c] = i - g; vals[c] = -1.0; ++c; }
    if (ix > 0)     { cols[c] = i - 1; vals[c] = -1.0; ++c; }
    cols[c] = i; vals[c] = 4.0; ++c;
    if (ix < g - 1) { cols[c] = i + 1; vals[c] = -1.0; ++c; }
    if (iy < g - 1) { cols[c] = i + g; vals[c] = -1.0; ++c; }
    return c;
  }
  if (kind == SLA_GEN_BLOCK16) {
    /* the block columns depend only on the 16-row group, so the 16 rows of a group share their blocks */
    const int nb = sla_synth_row_len(kind, n, k, band, i) / 16;
    const int64_t grp = i / 16, nbc = n / 16;
    int64_t bc[SLA_SYNTH_MAX_K / 16];
    int cnt = 0;
    for (uint64_t a = 0; cnt < nb; ++a) {
      const int64_t c = (int64_t)(sla_hash3(seed, (uint64_t)grp, a) % (uint64_t)nbc);
      int dup = 0;
      for (int q = 0; q < cnt; ++q) dup |= (bc[q] == c);
      if (!dup) bc[cnt++] = c;
    }
    for (int a = 1; a < cnt; ++a) {
      const int64_t c = bc[a];
      int b = a - 1;
      while (b >= 0 && bc[b] > c) { bc[b + 1] = bc[b]; --b; }
      bc[b + 1] = c;
    }
    int o = 0;
    for (int a = 0; a < cnt; ++a)
      for (int q = 0; q < 16; ++q, ++o) {
        cols[o] = bc[a] * 16 + q;
        vals[o] = sla_u11(sla_hash3(seed ^ SLA_SALT_VAL, (uint64_t)i, (uint64_t)o));
      }
    return o;
  }
  int64_t lo = 0, hi = n - 1;
  if (kind == SLA_GEN_BANDED) {
    lo = i - band < 0 ? 0 : i - band;
    hi = i + band > n - 1 ? n - 1 : i + band;
  }
  uint64_t span = (uint64_t)(hi - lo + 1);
  int kk = sla_synth_row_len(kind, n, k, band, i);
  int cnt = 1;
  cols[0] = i; vals[0] = 0.0;
  for (uint64_t a = 0; cnt < kk; ++a) {
    int64_t c = lo + (int64_t)(sla_hash3(seed, (uint64_t)i, a) % span);
    int dup = 0;
    for (int q = 0; q < cnt; ++q) dup |= (cols[q] == c);
    if (dup) continue;
    cols[cnt] = c;
    vals[cnt] = sla_u11(sla_hash3(seed ^ SLA_SALT_VAL, (uint64_t)i, (uint64_t)cnt));
    ++cnt;
  }
  /* insertion sort by column, values carried */
  for (int a = 1; a < cnt; ++a) {
    int64_t c = cols[a]; double v = vals[a];
    int b = a - 1;
    while (b >= 0 && cols[b] > c) { cols[b + 1] = cols[b]; vals[b + 1] = vals[b]; --b; }
    cols[b + 1] = c; vals[b + 1] = v;
  }
  double s = 1.0;
  int di = 0;
  for (int a = 0; a < cnt; ++a) {
    if (cols[a] == i) di = a;
    else s = s + (vals[a] < 0 ? -vals[a] : vals[a]);
  }
  vals[di] = s;
  return cnt;
}

#endif /* SLA_SYNTH_H */
