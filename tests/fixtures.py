"""Fixture data of the reference's own test-suite for the hot path, transcribed as plain lists.

Citations are /root/reference/test/LibSpec.hs line numbers (README.md where noted).  Dense matrix
lists are COLUMN-major, as `fromListDenseSM` reads them (src/Data/Sparse/SpMatrix.hs:239-241).
"""

# 2x2 system, LibSpec.hs:1171-1183
AA0 = (2, [1, 3, 2, 4])            # aa0 = fromListDenseSM 2 [1,3,2,4]   == [[1,2],[3,4]]
B0 = [8, 18]
X0 = [0.3, 1.4]
X0TRUE = [2, 3]
AA0TX0 = [11, 16]

# 3x3 tridiagonal SPD, LibSpec.hs:1204-1209 (sparsifySM drops the two zeros)
AA2 = (3, [2, -1, 0, -1, 2, -1, 0, -1, 2])
X2 = [3, 2, 3]
B2 = [4, -2, 4]

# matMat fixtures, LibSpec.hs:1263-1275
M1 = (2, [1, 3, 2, 4])
M2 = (2, [5, 7, 6, 8])
M1M2 = (2, [19, 43, 22, 50])
M1P = ((2, 3), [(0, 0, 2), (1, 0, 3), (1, 2, 4), (1, 2, 1)])   # duplicate (1,2): last write wins
M2P = ((3, 2), [(0, 0, 5), (0, 1, 3), (2, 1, 4)])
M1M2P = (2, [10, 15, 6, 13])
M2M1P = ((3, 3), [(0, 0, 19), (2, 0, 12), (0, 2, 3), (2, 2, 4)])
M1T = (2, [1, 2, 3, 4])

# dot, LibSpec.hs:1323-1324
TV0 = [5, 6]

# Arnoldi fixtures, LibSpec.hs:1301, 1334-1339
AA4 = (3, [3, 2, -2, 2, 2, -1, 6, 5, -4])
TM7_N = 5                          # tm7 = tridiag(-1, 2, -1), n = 5

# README.md:97, 183-189
AMAT = ((3, 3), [(0, 0, 2), (1, 0, 4), (1, 1, 3), (1, 2, 2), (2, 2, 5)])
AMAT_B = [3, 2, 5]
AMAT_X = [1.5, -2.0, 1.0]


def tm7_triples():
    n = TM7_N
    t = [(i, i + 1, -1.0) for i in range(n - 1)]
    t += [(i, i, 2.0) for i in range(n)]
    t += [(i + 1, i, -1.0) for i in range(n - 1)]
    return (n, n), t

# triangular-solve fixtures, LibSpec.hs:1409-1434 (the reference checks them through the residual, :436-459;
# the "#>" comments beside ltri1 there do not match its right-hand side, the residual test still passes)
LTRI0 = ((2, 2), [(0, 0, 2), (1, 0, 1), (1, 1, 3)])
B_LTRI0 = [4, 11]
UTRI0 = ((2, 2), [(0, 0, 2), (0, 1, 1), (1, 1, 3)])
B_UTRI0 = [7, 9]
LTRI1 = ((3, 3), [(0, 0, 2), (1, 0, 1), (1, 1, 4), (2, 1, 2), (2, 2, 3)])
B_LTRI1 = [4, 10, 17]
UTRI1 = ((3, 3), [(0, 0, 2), (0, 1, 1), (0, 2, 1), (1, 1, 4), (1, 2, 2), (2, 2, 3)])
B_UTRI1 = [9, 14, 9]
TRI_SPECS = [("lower", LTRI0, B_LTRI0), ("upper", UTRI0, B_UTRI0), ("lower", LTRI1, B_LTRI1), ("upper", UTRI1, B_UTRI1)]
