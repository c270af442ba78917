"""The arithmetic behind DESIGN.md's rooflines, in one place: algorithmic bytes per operation, the
time each takes at the measured HBM peak, and — for (#>) — the floor set by the L1TEX -> crossbar request port
(at most ONE request per SM-cycle; a request = one (warp instruction, 128-byte line) pair).
usage: python scripts/roofline_model.py [hbm_GBps] [sm_GHz]"""
import json
import os
import sys

SMS = 148
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6534.8


def b_spmv(n, nnz):                 # 12 nnz (col + val) + 4 (n + 1) (row_ptr) + 8 n (x) + 8 n (y)
    return 12 * nnz + 20 * n + 4


def b_bicgstab(n, nnz):             # 2 (#>) + the minimal-traffic schedule of Sparse.hs:973-981: 8n + 24n + 8n + 56n + 32n
    return 2 * b_spmv(n, nnz) + 128 * n - 8


def b_arnoldi_cycle(n, nnz, kn):    # step j: (#>) + 16 n (j + 1) + 32 n
    return kn * b_spmv(n, nnz) + 16 * n * kn * (kn + 1) // 2 + 32 * n * kn


def b_spmm(n, nnz, k):              # bf16 A values + int32 cols + row_ptr + B once + C once (bf16)
    return 6 * nnz + 4 * (n + 1) + 4 * n * k


def spmv_requests(n, nnz, panels=1, lanes_share_lines=False, nnz_per_row=32):
    """L1TEX -> XBAR requests of one tile-streamed (#>): the (col, val) stream in 128-byte lines, one request per
    gathered x entry unless the lanes of a warp share lines (stencil / narrow band), row_ptr / y per panel pass."""
    stream = 12 * nnz / 128
    gathers = nnz / (16 if lanes_share_lines else 1)
    per_pass = (4 * n + 8 * n) / 128 + (8 * n / 128 if panels > 1 else 0)
    return stream + gathers + panels * per_pass


def main():
    hbm = float(sys.argv[1]) if len(sys.argv) > 1 else peak_hbm()
    ghz = float(sys.argv[2]) if len(sys.argv) > 2 else 1.965
    rows = []
    n2, z2 = 10_000_000, 320_000_000
    g = 4096
    n3, z3 = g * g, 5 * g * g - 4 * g
    n4, z4 = 4_000_000, 256_000_000
    for name, by in (("cfg2 (#>) 10M x 10M, 32/row", b_spmv(n2, z2)), ("cfg3 (#>) Laplacian 4096^2", b_spmv(n3, z3)),
                     ("cfg3 bicgstabStep", b_bicgstab(n3, z3)), ("cfg4 (#>) 4M x 4M, 64/row", b_spmv(n4, z4)),
                     ("cfg4 arnoldi cycle kn=30", b_arnoldi_cycle(n4, z4, 30)), ("cfg5 (##) k=128 bf16", b_spmm(n2, z2, 128))):
        rows.append((name, by, by / hbm / 1e6))
    print(f"HBM peak {hbm:.1f} GB/s, SM clock {ghz:.3f} GHz, {SMS} SMs")
    print(f"{'operation':34s} {'algorithmic GB':>15s} {'ms at peak':>11s}")
    for name, by, ms in rows:
        print(f"{name:34s} {by / 1e9:15.3f} {ms:11.3f}")
    print()
    port = SMS * ghz * 1e9
    for name, n, z, panels, share in (("cfg2 uniform, 2 panels", n2, z2, 2, False), ("cfg4 uniform (x L2-resident)", n4, z4, 1, False),
                                      ("cfg3 stencil (lanes share lines)", n3, z3, 1, True)):
        req = spmv_requests(n, z, panels, share)
        t = req / port * 1e3
        print(f"{name:34s} {req / 1e6:8.1f} M requests -> port floor {t:6.3f} ms = {b_spmv(n, z) / t / 1e6 / hbm * 100:5.1f} % of the HBM peak")


if __name__ == "__main__":
    main()
