"""CPU model of csrc/dist_transpose.cu (the row-partitioned transposeSM): the same steps in numpy, rank by rank in one
process, against the oracle's transposeSM.  It checks the ALGORITHM the CUDA code follows — contiguous destination
ranges after the local transpose, per-row concatenation of the sources in rank order, column globalisation by
starts[s] — not the kernels themselves (those are checked on the GPU box by tests/dist_check.py)."""
import numpy as np
import pytest

from sparse_linear_algebra_b200.dist import row_partition


def _local_transpose(rp, col, val, m, n):
    """n x m CSR of the block's transpose, columns = LOCAL row numbers, sorted by (new row, old local row) — what
    sla_csr_transpose returns for the block."""
    rows = np.repeat(np.arange(m), np.diff(rp))
    order = np.lexsort((rows, col))                       # primary: col (new row), secondary: old row
    trp = np.zeros(n + 1, dtype=np.int64)
    np.add.at(trp, col + 1, 1)
    return np.cumsum(trp), rows[order], val[order]


@pytest.mark.parametrize("world,n,k,kind", [(2, 200, 6, "GEN_UNIFORM"), (3, 301, 5, "GEN_UNIFORM"), (4, 24 * 24, 5, "GEN_LAPLACE2D"), (8, 1000, 7, "GEN_BANDED")])
def test_distributed_transpose_model(ora, world, n, k, kind):
    gk = getattr(ora, kind)
    band = 24 if kind == "GEN_LAPLACE2D" else 40
    A = ora.SpMatrix.synth(gk, n, k, 0x5EED0009, band)
    rp, col, val = (np.asarray(a) for a in A.toCSR())
    starts = row_partition(n, world)
    # step 1 on every rank: local transpose of the row block
    T = []
    for r in range(world):
        r0, r1 = starts[r], starts[r + 1]
        lrp = rp[r0:r1 + 1] - rp[r0]
        T.append(_local_transpose(lrp, col[rp[r0]:rp[r1]], val[rp[r0]:rp[r1]], r1 - r0, n))
    # step 2: range of destination q inside T_s = [trp[starts[q]], trp[starts[q + 1]])  (contiguous by construction)
    h_off = [[int(T[s][0][starts[q]]) for q in range(world + 1)] for s in range(world)]
    counts = [[h_off[s][q + 1] - h_off[s][q] for q in range(world)] for s in range(world)]      # the all-gathered table
    rpo, cio, vao = (np.asarray(a) for a in A.transpose().toCSR())
    for me in range(world):
        m = starts[me + 1] - starts[me]
        # what rank `me` holds after the exchange: per source, the raw row_ptr slice for its rows and the (col, val) range
        src = []
        for s in range(world):
            trp, tcol, tval = T[s]
            sl = trp[starts[me]:starts[me + 1] + 1]
            lo, hi = h_off[s][me], h_off[s][me + 1]
            assert hi - lo == counts[s][me] == int(sl[-1] - sl[0])
            src.append((sl, tcol[lo:hi], tval[lo:hi]))
        # step 3: td_len_kernel + exclusive scan + td_fill_kernel
        length = np.zeros(m + 1, dtype=np.int64)
        for sl, _, _ in src:
            length[:m] += np.diff(sl)
        out_ptr = np.concatenate([[0], np.cumsum(length[:m])])
        out_col = np.zeros(out_ptr[-1], dtype=np.int64)
        out_val = np.zeros(out_ptr[-1])
        for j in range(m):
            o = out_ptr[j]
            for s, (sl, c_s, v_s) in enumerate(src):
                base = sl[0]
                for kk in range(sl[j], sl[j + 1]):
                    out_col[o] = c_s[kk - base] + starts[s]
                    out_val[o] = v_s[kk - base]
                    o += 1
            assert o == out_ptr[j + 1]
        r0, r1 = starts[me], starts[me + 1]
        assert np.array_equal(out_ptr, rpo[r0:r1 + 1] - rpo[r0])
        assert np.array_equal(out_col, cio[rpo[r0]:rpo[r1]])
        assert out_val.tobytes() == np.ascontiguousarray(vao[rpo[r0]:rpo[r1]], dtype=np.float64).tobytes()
