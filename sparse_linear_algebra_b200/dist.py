"""Multi-GPU host logic: one process per GPU (torchrun), row-partitioned matrices, NCCL over NVLink.

torch.distributed is the plumbing (rendezvous, broadcasting the NCCL id, gathering the per-rank column
ranges); the data path — the x exchange before each (#>) and the all-reduce behind every dot — runs inside
libsla_b200.so on the library's own communicator and stream.

The planning functions (`row_partition`, `plan_exchange`) are pure Python and are exercised on CPU with the
gloo backend (tests/test_dist_cpu.py).
"""
import ctypes as C
import os

import numpy as np

from . import _lib as L
from .sparse import Context, SpMatrix, SpVector, set_default_context


def row_partition(n, world):
    """Contiguous row blocks, balanced by rows: rank p owns [starts[p], starts[p+1])."""
    return [(n * p) // world for p in range(world + 1)]


def _overlap(lo, hi, a, b):
    """Intersection of the inclusive range [lo, hi] with the half-open block [a, b) as (offset, count)."""
    s, e = max(lo, a), min(hi + 1, b)
    return (s, e - s) if e > s else None


def plan_exchange(rank, starts, needs):
    """Which contiguous pieces of x travel before a (#>).

    starts : the row partition (world + 1 entries), identical on every rank.
    needs  : per rank, the inclusive range (lo, hi) of GLOBAL columns its row block references
             (hi < lo when the block is empty).
    Returns [(dir, peer, goff, count)]: dir 0 = this rank receives global entries [goff, goff+count) from
    peer, dir 1 = this rank sends them.  A rank never exchanges with itself; its own slice is read in place.
    Every send on rank a to rank b has the matching receive on rank b from a, because both are derived from
    the same `needs` table.
    """
    world = len(starts) - 1
    segs = []
    for q in range(world):
        if q == rank:
            continue
        lo, hi = needs[rank]
        if hi >= lo:
            ov = _overlap(lo, hi, starts[q], starts[q + 1])
            if ov:
                segs.append((0, q, ov[0], ov[1]))
        lo, hi = needs[q]
        if hi >= lo:
            ov = _overlap(lo, hi, starts[rank], starts[rank + 1])
            if ov:
                segs.append((1, q, ov[0], ov[1]))
    return segs


def densify_needs(starts, needs):
    """Collective decision, identical on every rank because it only looks at the global tables: when the exact
    (halo) plan would already move more than half of every rank's remote entries, treat the column support as
    dense (each rank receives every other block whole).  Returns (needs, use_allgather); all-gather additionally
    needs equal blocks, because every rank must then issue the same NCCL collective.  (Deciding this per rank
    deadlocks: a rank whose rows happen not to touch column 0 would issue send/recv against its peers'
    all-gather.)"""
    world, n = len(starts) - 1, starts[-1]

    def halo_volume(q):      # entries rank q would receive under the exact (halo) plan
        lo, hi = needs[q]
        if hi < lo:
            return 0
        return sum((_overlap(lo, hi, starts[p], starts[p + 1]) or (0, 0))[1] for p in range(world) if p != q)

    live = [q for q in range(world) if needs[q][1] >= needs[q][0]]
    dense = bool(live) and all(2 * halo_volume(q) > n - (starts[q + 1] - starts[q]) for q in live)
    densify_needs.last_dense = dense            # read by distribute(): a dense plan never takes the LL halo path
    if not dense:
        return list(needs), False
    equal = n % world == 0 and all(starts[p] == p * (n // world) for p in range(world + 1))
    return [(0, n - 1)] * world, equal


def exchange_bytes(segs):
    """(bytes received, bytes sent) per (#>) under this plan."""
    return (8 * sum(c for d, _, _, c in segs if d == 0), 8 * sum(c for d, _, _, c in segs if d == 1))


def phase_schedule(world, spec=None):
    """Panel schedule of the phased x exchange (SLA_P2P_X=5; the C twin is sla_p2p_phase_schedule in csrc/p2p.cu): how many
    column blocks each rotated panel holds — panel 0 the own block, panel p >= 1 the next predecessors, whose blocks travel in
    phase p.  `spec` ("1,1,2") must start with 1, add up to `world` and have 2..8 entries; otherwise 1, 1, 2, 4, ...: no phase
    is larger than everything multiplied before it."""
    max_panels = 8
    if spec:
        try:
            sizes = [int(t) for t in spec.split(",")]
        except ValueError:
            sizes = []
        if 2 <= len(sizes) <= max_panels and all(v > 0 for v in sizes) and sum(sizes) == world and sizes[0] == 1:
            return sizes
    sizes, left, nxt = [1], world - 1, 1
    while left > 0:
        take = nxt if nxt < left and len(sizes) < max_panels - 1 else left
        sizes.append(take)
        left -= take
        nxt = 2 if len(sizes) == 2 else nxt * 2
    return sizes


def phase_peers(rank, world, sizes):
    """For each phase p of `sizes`: (ranks this rank pushes its block to, ranks whose blocks it receives).  Rank r is
    predecessor k of rank r + k; phase p carries the predecessors sum(sizes[:p]) .. sum(sizes[:p + 1]) - 1."""
    out, k0 = [], 0
    for p, sz in enumerate(sizes):
        ks = range(max(k0, 1), k0 + sz) if p else range(0)
        out.append(([(rank + k) % world for k in ks], [(rank - k) % world for k in ks]))
        k0 += sz
    return out


def auto_exchange_mode(allgather, world, starts):
    """The automatic transport of the x exchange (SLA_P2P_X unset), from global quantities only so that every rank decides alike:
    2 (copy engines, arrival order) for dense equal-block plans on two ranks; 5 (phased TMA push under rotated column panels) from three
    ranks on when the plan is dense, the blocks are equal, in rank order and a multiple of 16 rows, and x (8 n bytes) exceeds what stays
    L2-resident (56 MB: the matrices the single-GPU plan column-panelises anyway); NCCL (0) otherwise.  Measured: DESIGN.md section 5."""
    if not allgather:
        return 0
    if world == 2:
        return 2
    n_cols = starts[-1]
    block = n_cols // world if world else 0
    equal = world >= 3 and n_cols % world == 0 and block % 16 == 0 and all(starts[q] == q * block for q in range(world + 1))
    return 5 if equal and 8 * n_cols > (56 << 20) else 0


def p2p_wanted():
    """Peer-memory collectives (csrc/p2p.cu) are on unless SLA_P2P=0; the variable must agree on every rank."""
    return os.environ.get("SLA_P2P", "1") != "0"


def p2p_exchange_mode():
    """Transport of the x exchange (SLA_P2P_X), -1 = automatic (the default when the variable is unset or "auto"):
        0  NCCL (all-gather for dense equal-block plans, grouped send/recv otherwise)
        1  peer-memory push kernel + flag round trip (any plan)
        2  copy-engine all-gather (dense equal-block plans): every rank's block travels on the copy engines in a staggered
           order.  When x is too large to stay L2-resident (8 n > 56 MB: such matrices are column-panelised on one GPU too) the
           (#>) consumes the blocks in ARRIVAL order, one column panel per source rank, the transfer of block k+1 hidden
           behind the kernel of panel k — rows are then folded in rotated column order, within the fp64 bound of SURVEY.md
           section 8(d) instead of bit-identical; smaller x: the blocks are waited for as a whole (reported as mode 4, bit-exact)
        3  LL halo (symmetric halo plans): ONE kernel stores 16-byte tagged entries straight into the neighbours' halo
           buffers and unpacks the entries arriving from them — no fence, no flag round trip, no rendezvous
        4  copy-engine all-gather waited for as a whole (what 2 degrades to)
        5  phased push (dense equal-block plans, x too large for L2 as in 2): the (#>) runs ROTATED column panels — own block,
           then the blocks of the predecessors in groups of 1, 2, 4, ... (SLA_P2P_PANELS="1,1,2" overrides) — and group p travels
           as TMA bulk copies issued by SLA_P2P_PUSH_CTAS one-thread CTAs on a high-priority side stream under the kernels of the
           panels before it; rows fold panel by panel, within the same fp64 bound as 2
    Automatic (measured on B200, profiles/r02_*): 2 for dense equal-block plans on TWO ranks (cfg 2: 0.824 vs 0.933 ms, the
    exchange fully hidden), NCCL otherwise — at 8 ranks the seven copy-engine transfers per rank took 0.43 ms against 0.16 ms
    for ncclAllGather (cfg 2: 0.654 vs 0.388 ms), and the LL halo kernel only equals NCCL's grouped send/recv (11 us exposed
    per (#>) on the 4096^2 Laplacian either way), so 3 and 4 stay opt-in.  Must agree on every rank."""
    if not p2p_wanted():
        return 0
    v = os.environ.get("SLA_P2P_X", "auto")
    if v in ("", "auto"):
        return -1
    try:
        return max(0, min(5, int(v)))
    except ValueError:
        return -1


HALO_MAX_SEGS = 15         # neighbours per rank (SLA_MAX_WORLD - 1)


def halo_eligible(starts, needs):
    """Collective decision (global tables only): the LL halo exchange needs every rank's plan to be SYMMETRIC — it sends to
    exactly the ranks it receives from, which is what makes the double-buffered halo buffers safe without a barrier (a
    sender can only be one exchange ahead of a neighbour it also waits for)."""
    world = len(starts) - 1
    any_seg = False
    for q in range(world):
        segs = plan_exchange(q, starts, needs)
        recv = sorted(p for d, p, _, _ in segs if d == 0)
        send = sorted(p for d, p, _, _ in segs if d == 1)
        if recv != send or len(recv) > HALO_MAX_SEGS:
            return False
        any_seg = any_seg or bool(segs)
    return any_seg


def halo_bases(rank, starts, needs):
    """Compact halo indices of this rank's plan: for a receive segment the offset of its first entry in this rank's halo
    buffer (segments packed in plan order), for a send segment its offset in the DESTINATION's buffer.  Both sides derive
    the same numbers because both walk plan_exchange(q, ...) of the same global table.  Returns (bases, entries received)."""
    world = len(starts) - 1
    cache = {}

    def recv_offsets(q):
        if q not in cache:
            off, table = 0, {}
            for d, peer, _, cnt in plan_exchange(q, starts, needs):
                if d == 0:
                    table[peer] = off
                    off += cnt
            cache[q] = (table, off)
        return cache[q]

    mine, total = recv_offsets(rank)
    bases = [mine[peer] if d == 0 else recv_offsets(peer)[0][rank] for d, peer, _, _ in plan_exchange(rank, starts, needs)]
    return bases, total


def _p2p_handshake(export, attach, enable):
    """Collective set-up of one peer-memory window: every rank exports a 64-byte IPC handle, the handles are
    gathered in rank order, every rank attaches, and the switch is thrown with the same value everywhere —
    on only if EVERY rank exported and attached (otherwise the NCCL path keeps running).  Returns the decision."""
    import torch.distributed as dist

    world = dist.get_world_size()
    buf = C.create_string_buffer(64)
    st = export(buf) if p2p_wanted() else L.SLA_ERR_INVALID
    got = [None] * world
    dist.all_gather_object(got, (int(st), bytes(buf.raw)))
    ok = all(s == L.SLA_OK for s, _ in got)
    st2 = L.SLA_ERR_INVALID
    if ok:
        blob = C.create_string_buffer(b"".join(h for _, h in got), 64 * world)
        st2 = attach(blob)
    sts = [None] * world
    dist.all_gather_object(sts, int(st2))
    on = ok and all(s == L.SLA_OK for s in sts)
    enable(1 if on else 0)
    return on


def init_context(device=None):
    """Create this rank's Context inside an initialised torch.distributed job (any backend).

    Rank 0 draws the NCCL unique id from the library, torch.distributed broadcasts it, every rank joins."""
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        ctx = Context(device)
        set_default_context(ctx)
        return ctx
    lib = L.load()
    box = [None]
    if rank == 0:
        buf = C.create_string_buffer(128)
        st = lib.sla_nccl_unique_id(buf)
        if st != L.SLA_OK:
            raise RuntimeError("sla_nccl_unique_id failed: libnccl.so.2 not loadable")
        box[0] = bytes(buf.raw)
    dist.broadcast_object_list(box, src=0)
    ctx = Context(device, rank=rank, world=world, nccl_id=box[0])
    set_default_context(ctx)
    # the all-reduce behind every dot: peer-memory window over NVLink (falls back to NCCL if any rank cannot map it)
    ctx.p2p = _p2p_handshake(lambda buf: lib.sla_p2p_export(ctx.h, buf), lambda blob: lib.sla_p2p_attach(ctx.h, blob),
                             lambda on: ctx.check(lib.sla_p2p_enable(ctx.h, on)))
    return ctx


def _install_plan(ctx, A, row0, segs, allgather=False):
    n = len(segs)
    dirs = (C.c_int * max(n, 1))(*[s[0] for s in segs])
    peers = (C.c_int * max(n, 1))(*[s[1] for s in segs])
    goff = (C.c_int64 * max(n, 1))(*[s[2] for s in segs])
    cnt = (C.c_int64 * max(n, 1))(*[s[3] for s in segs])
    ctx.check(ctx.lib.sla_csr_set_dist(ctx.h, A.h, row0, n, dirs, peers, goff, cnt, 1 if allgather else 0))


def distribute(ctx, A, starts):
    """Turn the local row block A (global column indices) into a distributed matrix: gather every rank's
    column range, plan the exchange, install it.  Collective over torch.distributed."""
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = C.c_int64(0), C.c_int64(-1)
    ctx.check(ctx.lib.sla_csr_col_range(ctx.h, A.h, C.byref(lo), C.byref(hi)))
    needs = [None] * world
    dist.all_gather_object(needs, (lo.value, hi.value))
    needs, allgather = densify_needs(starts, needs)
    dense = densify_needs.last_dense
    segs = plan_exchange(rank, starts, needs)
    _install_plan(ctx, A, starts[rank], segs, allgather)
    # the x exchange: peers store their pieces straight into this rank's window (csrc/p2p.cu)
    A.dist_p2p = False
    A.dist_p2p_mode = 0
    mode = p2p_exchange_mode()
    halo_ok = not dense and halo_eligible(starts, needs)
    if mode == -1:
        mode = auto_exchange_mode(allgather, world, starts)
    elif mode == 3 and not halo_ok:
        mode = 0
    elif mode in (2, 4, 5) and not allgather:
        mode = 1
    if getattr(ctx, "p2p", False) and mode:
        lib = ctx.lib
        if mode == 3:
            bases, _ = halo_bases(rank, starts, needs)
            # every rank's halo buffer gets the SAME capacity (the largest halo of the job): a sender addresses the second
            # buffer of a peer's window with its own stride
            cap = max(halo_bases(q, starts, needs)[1] for q in range(world))
            arr = (C.c_int64 * max(len(bases), 1))(*bases)
            ctx.check(lib.sla_csr_set_halo(ctx.h, A.h, len(bases), arr, cap))
        A.dist_p2p = _p2p_handshake(lambda buf: lib.sla_csr_p2p_export(ctx.h, A.h, buf),
                                    lambda blob: lib.sla_csr_p2p_attach(ctx.h, A.h, blob),
                                    lambda on: ctx.check(lib.sla_csr_p2p_enable(ctx.h, A.h, mode if on else 0)))
        A.dist_p2p_mode = lib.sla_csr_p2p_mode(A.h)
    A.dist_plan = segs
    A.dist_allgather = allgather
    A.row_starts = starts
    return A


def transpose_distributed(ctx, A):
    """transposeSM (SpMatrix.hs:717-718) of a row-partitioned square matrix: an all-to-all of entries inside the library
    (csrc/dist_transpose.cu), then the usual exchange plan for the new block.  The result is attached to A as the cached
    transpose, so `x <# A` and cgneInit / cgneStep work on the distributed A; A owns it (the returned wrapper borrows).
    Collective.  NOT YET RUN ON HARDWARE."""
    starts = list(A.row_starts)
    arr = (C.c_int64 * len(starts))(*starts)
    h = C.c_void_p()
    ctx.check(ctx.lib.sla_csr_transpose_dist(ctx.h, A.h, arr, C.byref(h)))
    T = SpMatrix(ctx, h, owns=False)
    distribute(ctx, T, starts)
    ctx.check(ctx.lib.sla_csr_attach_transpose(ctx.h, A.h, T.h))
    T._owner = A                     # the handle lives as long as A does
    return T


def generate_distributed(ctx, kind, n, nnz_per_row, seed, band=0):
    """Rank-local block of the n x n synthetic family (include/sla_synth.h), row-partitioned over the job."""
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    starts = row_partition(n, world)
    h = C.c_void_p()
    ctx.check(ctx.lib.sla_csr_generate_rows(ctx.h, kind, n, nnz_per_row, seed, band, starts[rank], starts[rank + 1], C.byref(h)))
    A = SpMatrix(ctx, h)
    if world > 1:
        distribute(ctx, A, starts)
    else:
        A.row_starts = starts
    return A


def generate_vector_slice(ctx, n, seed, starts, rank):
    h = C.c_void_p()
    ctx.check(ctx.lib.sla_vec_generate_slice(ctx.h, starts[rank], starts[rank + 1] - starts[rank], seed, C.byref(h)))
    return SpVector(ctx, h)
