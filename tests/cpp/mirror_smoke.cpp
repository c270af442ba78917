// mirror_smoke.cpp — the reference's own specs (test/LibSpec.hs) written against the C++ host mirror
// (include/sla_b200.hpp).  Exit 0 = all passed, 77 = no CUDA device, else failure.
#include <cmath>
#include <cstdio>

#include "sla_b200.hpp"

#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

static bool nearZero(double a) { return std::fabs(a) <= 1e-12; }   // Eps.hs:41-42

int main() {
  using namespace sla;
  std::unique_ptr<Context> cp;
  try { cp.reset(new Context(0)); } catch (const Error& e) {
    if (e.status == SLA_ERR_CUDA) { std::fprintf(stderr, "skip: %s\n", e.what()); return 77; }
    throw;
  }
  const Context& c = *cp;
  // aa0 = fromListDenseSM 2 [1,3,2,4] ; b0 = [8,18] ; x0true = [2,3]   (LibSpec.hs:1171-1183)
  SpMatrix aa0(c, 2, 2, {{0, 0, 1}, {1, 0, 3}, {0, 1, 2}, {1, 1, 4}});
  SpVector b0(c, std::vector<double>{8, 18}), x0true(c, std::vector<double>{2, 3}), tv0(c, std::vector<double>{5, 6});
  CHECK(dot(tv0, tv0) == 61);                                               // "<.> : inner product (Real)"   :45-46
  CHECK(nearZero(std::pow(norm2((aa0 * x0true) - b0), 2)));                 // "(#>)"                          :51-52
  const auto vm = aa0.vecMat(x0true).toDenseListSV();
  CHECK(vm[0] == 11 && vm[1] == 16);                                        // "(<#)"                          :53-54
  const auto t = aa0.transpose().matVec(x0true).toDenseListSV();
  CHECK(t[0] == 11 && t[1] == 16);                                          // transpose consistent with <#
  // fromListSM: duplicate overwrite and out-of-bounds                       (LibSpec.hs:1268, SpMatrix.hs:205-224)
  SpMatrix m1p(c, 2, 3, {{0, 0, 2}, {1, 0, 3}, {1, 2, 4}, {1, 2, 1}});
  CHECK(m1p.nnz() == 3);
  bool threw = false;
  try { SpMatrix bad(c, 2, 2, {{0, 2, 1}}); } catch (const OutOfBoundsIndexError&) { threw = true; }
  CHECK(threw);
  threw = false;
  try { aa0 * SpVector(c, std::vector<double>{1, 2, 3}); } catch (const MatVecSizeMismatchException&) { threw = true; }
  CHECK(threw);
  // bicgsInit / bicgstabStep                                               (LibSpec.hs:265-279)
  SpVector x0(c, std::vector<double>{0.3, 1.4});
  const auto r0 = (b0 - (aa0 * x0)).toDenseListSV();
  KrylovState st = bicgsInit(aa0, b0, x0);
  CHECK(st.r().toDenseListSV() == r0 && st.p().toDenseListSV() == r0);
  SpVector r0hat = st.r().copy();
  bicgstabStep(aa0, r0hat, st);
  CHECK(st.x().dim() == 2);
  // linSolve0 x {BICGSTAB_, CGS_, CGNE_}: ||x - xhat|| <= 1e-12 from x0 = 0.1   (LibSpec.hs:286-300)
  for (LinSolveMethod m : {BICGSTAB_, CGS_, CGNE_}) {
    SolveInfo info;
    SpVector xhat = linSolve0(m, aa0, b0, SpVector::constv(c, 2, 0.1), &info);
    CHECK(nearZero(norm2(x0true - xhat)));
  }
  threw = false;
  try { linSolve0(GMRES_, aa0, b0, SpVector::constv(c, 2, 0.1)); } catch (const IterE&) { threw = true; }   // Sparse.hs:1031
  CHECK(threw);
  // arnoldi tm7 4: || A Q' - Q H ||_F nearZero                              (LibSpec.hs:226-232, 638-653)
  std::vector<SpMatrix::Triple> tri;
  for (int i = 0; i < 5; ++i) { tri.push_back({i, i, 2.0}); if (i < 4) { tri.push_back({i, i + 1, -1.0}); tri.push_back({i + 1, i, -1.0}); } }
  SpMatrix tm7(c, 5, 5, tri);
  ArnoldiResult ar = arnoldi(tm7, SpVector::constv(c, 5, 1.0), 4);
  const int n = 5, k = ar.nmax;
  double fro = 0;
  for (int col = 0; col < k; ++col)
    for (int row = 0; row < n; ++row) {
      double aq = 0;                          // (A Q')[row][col]
      for (int j = 0; j < n; ++j) { const double a = (j == row) ? 2.0 : (std::abs(j - row) == 1 ? -1.0 : 0.0); aq += a * ar.Q[(size_t)col * n + j]; }
      double qh = 0;                          // (Q H)[row][col]
      for (int j = 0; j <= k; ++j) qh += ar.Q[(size_t)j * n + row] * ar.H[(size_t)col * (k + 1) + j];
      fro += (aq - qh) * (aq - qh);
    }
  CHECK(std::sqrt(fro) <= 1e-12);
  std::printf("mirror_smoke ok (%lld kernel launches)\n", (long long)c.launches());
  return 0;
}
