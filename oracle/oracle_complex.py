"""Pure-Python restatement of the reference's container semantics for ANY scalar (Python complex included).

TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rule as oracle.py).  The product path is `SpMatrix Double`; this small
generic restatement exists only to pin the conventions the reference's COMPLEX known-answer tests fix and the real
restatement inherits: `<.>` conjugates its SECOND argument (Class.hs:401-402), `#>` / `<#` do NOT conjugate
(`dotu`, Common.hs:258-260), sums are strict left folds in ascending key order from 0 (IntM.hs:17, 78-80).
Containers are dicts walked in sorted key order — the traversal order of Data.IntMap.Strict.
"""


def from_list_dense_sv(xs):                       # fromListDenseSV  SpVector.hs:194-195
    return len(xs), {i: x for i, x in enumerate(xs)}


def from_list_dense_sm(n, xs):                    # fromListDenseSM n: COLUMN-major  SpMatrix.hs:239-241, Utils.hs:85-90
    m = len(xs) // n
    rows = {}
    for k, x in enumerate(xs):
        j, i = divmod(k, m)
        rows.setdefault(i, {})[j] = x
    return (m, n), rows


def _conj(x):
    return x.conjugate() if isinstance(x, complex) else x


def dot(v, w):                                    # v <.> w = sum $ liftI2 (<.>) v w   SpVector.hs:116-117 ; x <.> y = x * conjugate y   Class.hs:401-402
    acc = 0
    for k in sorted(set(v[1]) & set(w[1])):       # IM.intersectionWith, ascending keys   IntM.hs:80
        acc = acc + v[1][k] * _conj(w[1][k])
    return acc


def dotu(u, v):                                   # dotu u v = sum $ liftI2 (*) u v   Common.hs:259-260 (UN-conjugated)
    acc = 0
    for k in sorted(set(u) & set(v)):
        acc = acc + u[k] * v[k]
    return acc


def mat_vec(a, v):                                # matVecSD: one entry per STORED row   Common.hs:247-250
    (nr, nc), rows = a
    assert nc == v[0], "matVec : mismatched dimensions"
    return nr, {i: dotu(rows[i], v[1]) for i in sorted(rows)}


def transpose(a):                                 # transposeSM / transposeIM2   SpMatrix.hs:717-718, IntMap2.hs:88-90
    (nr, nc), rows = a
    t = {}
    for i in sorted(rows):
        for j in sorted(rows[i]):
            t.setdefault(j, {})[i] = rows[i][j]
    return (nc, nr), t


def vec_mat(v, a):                                # vecMatSD v m = matVecSD (transpose m) v   Common.hs:253-256
    return mat_vec(transpose(a), v)


def sub(v, w):                                    # v ^-^ w = v ^+^ negateV w over the key UNION   SpVector.hs:107-110, Class.hs:69
    out = {}
    for k in sorted(set(v[1]) | set(w[1])):
        out[k] = v[1].get(k, 0) + (-w[1][k] if k in w[1] else 0) if k in v[1] else -w[1][k]
    return max(v[0], w[0]), out


def near_zero(x):                                 # Complex: magnitude <= 1e-12   Eps.hs:41-42, 58-61
    return abs(x) <= 1e-12
