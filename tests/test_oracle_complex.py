"""The reference's COMPLEX known-answer tests for the path (test/LibSpec.hs:47-48, 55-60; data :1220-1230, :1370-1371),
against the generic pure-Python restatement.  They pin two conventions the real (`Double`) oracle and the CUDA path
inherit: `<.>` conjugates its second argument (the identity on reals), `#>` / `<#` never conjugate."""
from oracle import oracle_complex as oc


def test_dot_complex_conjugates_second_argument():              # LibSpec.hs:47-48, :1370-1371
    tvc2 = oc.from_list_dense_sv([1 + 1j, 2 - 1j])
    tvc3 = oc.from_list_dense_sv([3 - 2j, 1 + 1j])
    assert oc.dot(tvc2, tvc3) == 2 + 2j
    assert oc.dot(tvc3, tvc2) == 2 - 2j                         # conjugate-symmetric, so the convention is visible


def test_matvec_vecmat_complex_unconjugated():                  # LibSpec.hs:55-60, :1220-1230
    aa0c = oc.from_list_dense_sm(2, [3 + 1j, -3 + 2j, -2 - 1j, 1 - 2j])
    b0c = oc.from_list_dense_sv([3 - 4j, -1 + 0.5j])
    c0c = oc.from_list_dense_sv([15.5 - 9j, -1 + 20.5j])
    c0c_ = oc.from_list_dense_sv([15 - 12.5j, -10 + 7.5j])
    v = oc.sub(oc.mat_vec(aa0c, b0c), c0c)
    assert oc.near_zero(oc.dot(v, v))
    w = oc.sub(oc.vec_mat(b0c, aa0c), c0c_)
    assert oc.near_zero(oc.dot(w, w))
    # the values themselves are exact in binary floating point
    assert oc.mat_vec(aa0c, b0c)[1] == c0c[1] and oc.vec_mat(b0c, aa0c)[1] == c0c_[1]


def test_real_specialisation_matches_the_c_oracle(ora):         # LibSpec.hs:45-46, 51-54 through both restatements
    tv0 = oc.from_list_dense_sv([5.0, 6.0])
    assert oc.dot(tv0, tv0) == 61.0 == ora.SpVector.fromListDenseSV(2, [5.0, 6.0]).dot(ora.SpVector.fromListDenseSV(2, [5.0, 6.0]))
    aa0 = oc.from_list_dense_sm(2, [1.0, 3.0, 2.0, 4.0])
    x0 = oc.from_list_dense_sv([2.0, 3.0])
    assert oc.mat_vec(aa0, x0)[1] == {0: 8.0, 1: 18.0} and oc.vec_mat(x0, aa0)[1] == {0: 11.0, 1: 16.0}
    A = ora.SpMatrix.fromListDenseSM(2, [1.0, 3.0, 2.0, 4.0])
    xs = ora.SpVector.fromListDenseSV(2, [2.0, 3.0])
    assert list(A.matVec(xs).toDenseListSV()) == [8.0, 18.0] and list(A.vecMat(xs).toDenseListSV()) == [11.0, 16.0]
