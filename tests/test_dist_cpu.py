"""CPU tests of the multi-GPU host logic (world_size 2, gloo): the row partition and the x-exchange plan that
libsla_b200.so executes with NCCL on the device are computed by sparse_linear_algebra_b200.dist; here the same
plan is executed with gloo send/recv on numpy slices and the row-partitioned product, computed per rank by the
oracle, must equal the single-process oracle product bit for bit."""
import os
import socket

import numpy as np
import pytest


def test_row_partition_and_plan_unit():
    from sparse_linear_algebra_b200.dist import densify_needs, exchange_bytes, plan_exchange, row_partition

    assert row_partition(10, 3) == [0, 3, 6, 10]
    assert row_partition(7, 1) == [0, 7]
    starts = row_partition(100, 4)                       # [0, 25, 50, 75, 100]
    # banded: every rank needs its own block +- 5 columns
    needs = [(max(0, starts[p] - 5), min(99, starts[p + 1] - 1 + 5)) for p in range(4)]
    plans = [plan_exchange(p, starts, needs) for p in range(4)]
    assert sorted(plans[0]) == [(0, 1, 25, 5), (1, 1, 20, 5)]
    assert sorted(plans[1]) == [(0, 0, 20, 5), (0, 2, 50, 5), (1, 0, 25, 5), (1, 2, 45, 5)]
    # every send has its matching receive
    for p in range(4):
        for d, q, off, cnt in plans[p]:
            assert (1 - d, p, off, cnt) in plans[q]
    assert exchange_bytes(plans[1]) == (80, 80)
    # dense support: all-gather shape
    needs = [(0, 99)] * 4
    pl = plan_exchange(2, starts, needs)
    assert sorted(s for s in pl if s[0] == 0) == [(0, 0, 0, 25), (0, 1, 25, 25), (0, 3, 75, 25)]
    assert sorted(s for s in pl if s[0] == 1) == [(1, 0, 50, 25), (1, 1, 50, 25), (1, 3, 50, 25)]
    # the all-gather decision is collective: taken from the global table, so it is the same on every rank even
    # when one rank happens not to touch the first / last column
    starts8 = row_partition(800, 8)
    needs8 = [(0, 799)] * 7 + [(3, 790)]
    nd, ag = densify_needs(starts8, needs8)
    assert ag and nd == [(0, 799)] * 8
    assert densify_needs(row_partition(801, 8), [(0, 800)] * 8) == ([(0, 800)] * 8, False)      # ragged: p2p, full blocks
    assert densify_needs(starts8, [(max(0, s - 5), s + 104) for s in starts8[:-1]])[1] is False  # banded stays a halo plan
    # an empty block neither receives nor is sent anything on its behalf
    needs = [(0, 99), (0, -1), (0, 99), (0, 99)]
    assert all(s[0] == 1 for s in plan_exchange(1, starts, needs))
    assert not [s for s in plan_exchange(0, starts, needs) if s[0] == 1 and s[1] == 1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind_name, n, k, band, ret):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as ora
        from sparse_linear_algebra_b200.dist import densify_needs, plan_exchange, row_partition

        kind = getattr(ora, kind_name)
        seed = 0x5EED0005
        starts = row_partition(n, world)
        r0, r1 = starts[rank], starts[rank + 1]
        # local row block with GLOBAL column indices
        rows = [ora.synth_row(kind, n, k, seed, band, i) for i in range(r0, r1)]
        row_ptr = np.concatenate([[0], np.cumsum([len(c) for c, _ in rows])]).astype(np.int64)
        col = np.concatenate([c for c, _ in rows]) if rows else np.zeros(0, np.int64)
        val = np.concatenate([v for _, v in rows]) if rows else np.zeros(0)
        need = (int(col.min()), int(col.max())) if col.size else (0, -1)
        needs = [None] * world
        dist.all_gather_object(needs, need)
        needs, _ = densify_needs(starts, needs)
        segs = plan_exchange(rank, starts, needs)
        # execute the plan with gloo point-to-point copies
        x_local = ora.SpVector.synth(seed + 1, n).toDenseListSV()[r0:r1].copy()
        x_full = np.full(n, np.nan)                       # NaN everywhere: touching an un-exchanged entry poisons y
        x_full[r0:r1] = x_local
        reqs, bufs = [], []
        for d, q, off, cnt in segs:
            if d == 1:
                reqs.append(dist.isend(torch.from_numpy(x_local[off - r0: off - r0 + cnt].copy()), q))
        for d, q, off, cnt in segs:
            if d == 0:
                t = torch.empty(cnt, dtype=torch.float64)
                bufs.append((off, cnt, t))
                reqs.append(dist.irecv(t, q))
        for r in reqs:
            r.wait()
        for off, cnt, t in bufs:
            x_full[off: off + cnt] = t.numpy()
        # local product by the oracle on the exchanged vector (absent = never referenced)
        Aloc = ora.SpMatrix.fromCSR(r1 - r0, n, row_ptr, col, val)
        xs = ora.SpVector.mkSpVR(n, np.where(np.isnan(x_full), 0.0, x_full))
        touched = np.zeros(n, bool); touched[col] = True
        assert not np.isnan(x_full[touched]).any(), "the plan left a referenced x entry un-exchanged"
        y_local = Aloc.matVec(xs).toDenseListSV()
        ys = [None] * world
        dist.all_gather_object(ys, y_local)
        # the same plan applied to k-wide rows is the gather of the dense right operand of a row-partitioned (##)
        # (sla_dist_gather_rows): every B row the block references must arrive, and the local product must equal the
        # rows of the single-process product bit for bit
        kk = 3
        B_full_true = np.random.default_rng(99).standard_normal((n, kk))
        B_local = B_full_true[r0:r1].copy()
        B_full = np.full((n, kk), np.nan)
        B_full[r0:r1] = B_local
        reqs, bufs = [], []
        for d, q, off, cnt in segs:
            if d == 1:
                reqs.append(dist.isend(torch.from_numpy(B_local[off - r0: off - r0 + cnt].copy()), q))
        for d, q, off, cnt in segs:
            if d == 0:
                t = torch.empty((cnt, kk), dtype=torch.float64)
                bufs.append((off, cnt, t))
                reqs.append(dist.irecv(t, q))
        for r in reqs:
            r.wait()
        for off, cnt, t in bufs:
            B_full[off: off + cnt] = t.numpy()
        assert not np.isnan(B_full[touched]).any(), "the plan left a referenced row of B un-gathered"
        assert np.array_equal(B_full[touched], B_full_true[touched])
        Bo = ora.SpMatrix.fromListDenseSM(n, np.where(np.isnan(B_full), 0.0, B_full).T.reshape(-1))
        c_local = Aloc.matMat(Bo).toDense()
        cs = [None] * world
        dist.all_gather_object(cs, c_local)
        plans = [None] * world
        dist.all_gather_object(plans, segs)
        if rank == 0:
            Afull = ora.SpMatrix.synth(kind, n, k, seed, band)
            yref = Afull.matVec(ora.SpVector.synth(seed + 1, n)).toDenseListSV()
            ok = np.concatenate(ys).tobytes() == yref.tobytes()
            cref = Afull.matMat(ora.SpMatrix.fromListDenseSM(n, B_full_true.T.reshape(-1))).toDense()
            ok = ok and np.ascontiguousarray(np.concatenate(cs, axis=0)).tobytes() == np.ascontiguousarray(cref).tobytes()
            for p in range(world):
                for d, q, off, cnt in plans[p]:
                    ok = ok and (1 - d, p, off, cnt) in plans[q]
            recv_bytes = [8 * sum(c for d, _, _, c in pl if d == 0) for pl in plans]
            ret.put((ok, recv_bytes))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,n,k,band", [("GEN_UNIFORM", 600, 8, 0), ("GEN_BANDED", 1000, 8, 12), ("GEN_LAPLACE2D", 24 * 24, 5, 24)])
def test_row_partitioned_spmv_gloo(ora, kind, n, k, band):
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, n, k, band, ret)) for r in range(world)]
    for p in procs:
        p.start()
    ok, recv_bytes = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    if kind == "GEN_UNIFORM":
        assert recv_bytes == [8 * (n // 2)] * 2          # dense support: the whole remote block
    if kind == "GEN_BANDED":
        assert max(recv_bytes) <= 8 * band               # halo only
    if kind == "GEN_LAPLACE2D":
        assert recv_bytes == [8 * band] * 2              # one grid line from the neighbour


# ---- property test of the planner (hypothesis): for ANY partition and ANY per-rank column ranges the plans of the ranks
# are mutually consistent (every send has its receive), never self-directed, stay inside the sender's block, and deliver
# every remote column a rank references.
def test_plan_exchange_properties():
    from hypothesis import given, settings, strategies as st

    from sparse_linear_algebra_b200.dist import densify_needs, plan_exchange, row_partition

    @st.composite
    def cases(draw):
        world = draw(st.integers(2, 8))
        n = draw(st.integers(world, 400))
        starts = row_partition(n, world)
        needs = []
        for _ in range(world):
            if draw(st.booleans()) and draw(st.integers(0, 9)) == 0:
                needs.append((0, -1))                       # a block without entries
            else:
                lo = draw(st.integers(0, n - 1))
                needs.append((lo, draw(st.integers(lo, n - 1))))
        return world, n, starts, needs

    @settings(max_examples=300, deadline=None)
    @given(cases())
    def check(case):
        world, n, starts, needs0 = case
        for needs in (needs0, densify_needs(starts, needs0)[0]):
            plans = [plan_exchange(r, starts, needs) for r in range(world)]
            for r, pl in enumerate(plans):
                got = np.zeros(n, bool)
                got[starts[r]:starts[r + 1]] = True          # the own slice is read in place
                for d, q, off, cnt in pl:
                    assert q != r and cnt > 0 and 0 <= off and off + cnt <= n
                    assert (1 - d, r, off, cnt) in plans[q]
                    if d == 1:
                        assert starts[r] <= off and off + cnt <= starts[r + 1]
                    else:
                        assert starts[q] <= off and off + cnt <= starts[q + 1]
                        got[off:off + cnt] = True
                lo, hi = needs[r]
                if hi >= lo:
                    assert got[lo:hi + 1].all(), "a referenced column is neither local nor received"

    check()


# ---- LL halo plan (p2p mode 3): compact indices agree between sender and receiver --------------------------------------
from sparse_linear_algebra_b200 import dist as sd  # noqa: E402

def _banded_needs(starts, h, n):
    return [(max(0, starts[p] - h), min(n - 1, starts[p + 1] - 1 + h)) for p in range(len(starts) - 1)]


@pytest.mark.parametrize("world,n,h", [(2, 1000, 7), (4, 4096, 64), (8, 96 * 96, 96), (8, 1003, 200), (3, 100, 40)])
def test_halo_bases_agree_between_sender_and_receiver(world, n, h):
    starts = sd.row_partition(n, world)
    needs = _banded_needs(starts, h, n)
    assert sd.halo_eligible(starts, needs)
    plans = [sd.plan_exchange(q, starts, needs) for q in range(world)]
    bases = [sd.halo_bases(q, starts, needs) for q in range(world)]
    for q in range(world):
        b, total = bases[q]
        recv = [(b[i], plans[q][i][3]) for i in range(len(plans[q])) if plans[q][i][0] == 0]
        # receive segments tile [0, total) without gaps or overlap
        recv.sort()
        pos = 0
        for off, cnt in recv:
            assert off == pos
            pos += cnt
        assert pos == total == sum(c for d, _, _, c in plans[q] if d == 0)
        # every send lands exactly where the destination expects the matching receive
        for i, (d, peer, goff, cnt) in enumerate(plans[q]):
            if d != 1:
                continue
            j = next(k for k, (d2, p2, g2, c2) in enumerate(plans[peer]) if d2 == 0 and p2 == q)
            assert plans[peer][j][2:] == (goff, cnt)
            assert bases[peer][0][j] == b[i]


def test_halo_not_eligible_for_one_sided_or_empty_plans():
    starts = sd.row_partition(100, 2)
    # rank 0 needs a piece of rank 1's block, rank 1 needs nothing from rank 0: one-sided -> no LL halo (no implicit barrier)
    assert not sd.halo_eligible(starts, [(0, 60), (50, 99)])
    # block diagonal: nothing travels
    assert not sd.halo_eligible(starts, [(0, 49), (50, 99)])


def test_exchange_mode_env(monkeypatch):
    monkeypatch.delenv("SLA_P2P", raising=False)
    monkeypatch.delenv("SLA_P2P_X", raising=False)
    assert sd.p2p_exchange_mode() == -1
    for v, want in (("auto", -1), ("0", 0), ("2", 2), ("3", 3), ("5", 5), ("9", 5), ("x", -1)):
        monkeypatch.setenv("SLA_P2P_X", v)
        assert sd.p2p_exchange_mode() == want
    monkeypatch.setenv("SLA_P2P", "0")
    assert sd.p2p_exchange_mode() == 0


# ---- phased x exchange (SLA_P2P_X=5): the panel schedule and who talks to whom in which phase --------------------------
def test_phase_schedule_default():
    assert sd.phase_schedule(2) == [1, 1]
    assert sd.phase_schedule(3) == [1, 1, 1]
    assert sd.phase_schedule(4) == [1, 1, 2]
    assert sd.phase_schedule(8) == [1, 1, 2, 4]
    assert sd.phase_schedule(16) == [1, 1, 2, 4, 8]
    for w in range(2, 33):
        sizes = sd.phase_schedule(w)
        assert sum(sizes) == w and sizes[0] == 1 and len(sizes) <= 8
        # no phase is larger than everything multiplied before it (the last of a full schedule may take the remainder)
        if len(sizes) < 8:
            assert all(sizes[p] <= sum(sizes[:p]) for p in range(1, len(sizes)))


def test_phase_schedule_override_and_rejects():
    assert sd.phase_schedule(4, "1,3") == [1, 3]
    assert sd.phase_schedule(8, "1,1,1,1,2,2") == [1, 1, 1, 1, 2, 2]
    for bad in ("2,2", "1,1,1", "1,0,3", "4", "1,1,1,1,1,1,1,1,1", "x"):       # not starting at 1 / wrong sum / one panel / 9 entries
        assert sd.phase_schedule(4 if bad != "1,1,1,1,1,1,1,1,1" else 9, bad) == sd.phase_schedule(4 if bad != "1,1,1,1,1,1,1,1,1" else 9)


def test_phase_schedule_matches_the_library():
    import ctypes as C
    from sparse_linear_algebra_b200 import _lib
    L = _lib.load()
    buf = (C.c_int * 8)()
    for w in range(2, 17):
        for spec in (None, "1,1,2", "1,3", "1,1,1,1,2,2", "1,7", "2,2", "1,1,1,1,1,1,1,1"):
            n = L.sla_p2p_phase_schedule(w, spec.encode() if spec else None, buf)
            assert list(buf[:n]) == sd.phase_schedule(w, spec), (w, spec)


def test_phase_peers_are_consistent():
    for w in (2, 3, 4, 5, 8, 16):
        sizes = sd.phase_schedule(w)
        table = [sd.phase_peers(r, w, sizes) for r in range(w)]
        for r in range(w):
            assert table[r][0] == ([], [])                     # panel 0 is the own block: nothing travels
            got = []
            for p, (send, recv) in enumerate(table[r]):
                assert len(send) == len(recv) == (sizes[p] if p else 0)          # balanced: every rank sends what it receives
                for q in send:
                    assert r in table[q][p][1]                  # whoever I push to in phase p waits for me in phase p
                got += recv
            assert sorted(got) == [q for q in range(w) if q != r]               # every other block arrives exactly once


def test_auto_exchange_mode_policy():
    n = 10_000_000
    eq = lambda w, n_=n: [q * (n_ // w) for q in range(w + 1)]
    assert sd.auto_exchange_mode(True, 2, eq(2)) == 2                      # two ranks: copy engines in arrival order
    assert sd.auto_exchange_mode(True, 4, eq(4)) == 5 and sd.auto_exchange_mode(True, 8, eq(8)) == 5      # cfg 2 at 4 and 8 ranks
    assert sd.auto_exchange_mode(False, 8, eq(8)) == 0                     # halo plans (cfg 3, banded): NCCL send/recv
    assert sd.auto_exchange_mode(True, 8, eq(8, 4_000_000)) == 0           # cfg 4: x (32 MB) stays in L2, panels would cost more than they hide
    assert sd.auto_exchange_mode(True, 3, [0, 3_333_334, 6_666_667, 10_000_000]) == 0      # unequal blocks
    assert sd.auto_exchange_mode(True, 4, [0, 2_500_000, 5_000_000, 7_500_016, 10_000_000]) == 0
    assert sd.auto_exchange_mode(True, 5, eq(5, 10_000_040)) == 0          # blocks of 2 000 008 rows: a multiple of 8, not of 16
    assert sd.auto_exchange_mode(True, 16, eq(16, 16_000_000)) == 5
