"""CPU tests of the peer-memory collectives' host logic (csrc/p2p.cu, dist.py):

* the collective ON/OFF decision of `_p2p_handshake` under gloo, world size 2 — the switch must end up with the same
  value on every rank whichever rank fails to export / attach;
* an interleaving model of the two device protocols (double-buffered by sequence parity, monotone flags, "wait
  until flag >= seq"): a randomised scheduler steps `world` simulated ranks one memory operation at a time and the
  test asserts that no rank ever reads a contribution that is not the one of its current collective — the
  write-after-read argument in the header of p2p.cu, executed — for the all-reduce, the push exchange, the copy-engine exchange
  consumed in arrival order (mode 2) and the phased push under rotated panels (mode 5).
"""
import os
import random
import socket

import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _handshake_worker(rank, world, port, scenario, ret):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if scenario == "env_off":
        os.environ["SLA_P2P"] = "0"
    else:
        os.environ.pop("SLA_P2P", None)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sparse_linear_algebra_b200 import _lib as L
        from sparse_linear_algebra_b200.dist import _p2p_handshake

        log = {"export": 0, "attach": None, "enable": None}

        def export(buf):
            log["export"] += 1
            if scenario == "export_fails_on_1" and rank == 1:
                return L.SLA_ERR_CUDA
            buf.raw = bytes([rank + 1]) * 64
            return L.SLA_OK

        def attach(blob):
            log["attach"] = bytes(blob.raw)
            if scenario == "attach_fails_on_0" and rank == 0:
                return L.SLA_ERR_COMM
            return L.SLA_OK

        def enable(on):
            log["enable"] = on

        on = _p2p_handshake(export, attach, enable)
        ret.put((rank, on, log))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("scenario,expect_on", [("ok", True), ("export_fails_on_1", False), ("attach_fails_on_0", False),
                                                ("env_off", False)])
def test_p2p_handshake_is_collective(scenario, expect_on):
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_handshake_worker, args=(r, world, port, scenario, ret)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(ret.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, on, log in got:
        assert on is expect_on and log["enable"] == (1 if expect_on else 0)      # the same decision on every rank
        if scenario == "env_off":
            assert log["export"] == 0 and log["attach"] is None
        if scenario == "export_fails_on_1":
            assert log["attach"] is None                                          # nobody maps a window that was not exported
        if scenario in ("ok", "attach_fails_on_0"):
            assert log["attach"] == bytes([1]) * 64 + bytes([2]) * 64            # handles in rank order


# ---------------------------------------------------------------------------------------------------------------
# interleaving model

def _allreduce_rank(rank, world, nred, win, out):
    """Generator: one simulated rank running `nred` all-reduces of p2p_allreduce_kernel; yields after every
    memory operation so the scheduler can interleave the ranks arbitrarily."""
    for seq in range(1, nred + 1):
        b = seq & 1
        mine = (rank + 1) * 1000 + seq                   # this rank's contribution to collective `seq`
        for t in range(world):                           # lane t stores into peer t's window, then raises the flag
            win[t]["val"][b][rank] = mine
            yield
            win[t]["flag"][b][rank] = seq
            yield
        for t in range(world):                           # wait for every peer's flag of this collective
            while win[rank]["flag"][b][t] < seq:
                yield
        got = []
        for r in range(world):                           # sum in rank order
            got.append(win[rank]["val"][b][r])
            yield
        out.append((seq, got))


def _exchange_rank(rank, world, nex, win, out):
    """Generator: p2p_push_kernel followed by the kernel that reads the gathered buffer."""
    for seq in range(1, nex + 1):
        b = seq & 1
        for t in range(world):
            if t != rank:
                win[t]["buf"][b][rank] = (rank + 1) * 1000 + seq          # data first ...
                yield
        for t in range(world):
            if t != rank:
                win[t]["flag"][rank] = seq                                 # ... then the flag
                yield
        for t in range(world):
            if t != rank:
                while win[rank]["flag"][t] < seq:
                    yield
        got = {}
        for t in range(world):                                             # the SpMV kernel reading xfull
            if t != rank:
                got[t] = win[rank]["buf"][b][t]
                yield
        out.append((seq, got))


def _run(gens, rng):
    live = list(range(len(gens)))
    steps = 0
    while live:
        # bursty scheduler: pick a rank and let it run a random number of operations (lets ranks race far ahead)
        i = rng.choice(live)
        for _ in range(rng.choice((1, 1, 2, 5, 40))):
            try:
                next(gens[i])
            except StopIteration:
                live.remove(i)
                break
        steps += 1
        assert steps < 2_000_000, "protocol model dead-locked"


@pytest.mark.parametrize("world", [2, 3, 8])
def test_allreduce_protocol_model(world):
    for seed in range(40):
        rng = random.Random(seed * 131 + world)
        nred = 12
        win = [{"val": [[None] * world for _ in range(2)], "flag": [[0] * world for _ in range(2)]} for _ in range(world)]
        outs = [[] for _ in range(world)]
        _run([_allreduce_rank(r, world, nred, win, outs[r]) for r in range(world)], rng)
        for r in range(world):
            assert [s for s, _ in outs[r]] == list(range(1, nred + 1))
            for seq, got in outs[r]:
                assert got == [(q + 1) * 1000 + seq for q in range(world)], f"rank {r} read a stale or early contribution at {seq}"


@pytest.mark.parametrize("world", [2, 4, 8])
def test_exchange_protocol_model(world):
    for seed in range(40):
        rng = random.Random(seed * 977 + world)
        nex = 10
        win = [{"buf": [[None] * world for _ in range(2)], "flag": [0] * world} for _ in range(world)]
        outs = [[] for _ in range(world)]
        _run([_exchange_rank(r, world, nex, win, outs[r]) for r in range(world)], rng)
        for r in range(world):
            for seq, got in outs[r]:
                assert got == {t: (t + 1) * 1000 + seq for t in range(world) if t != r}, f"rank {r}, exchange {seq}"


# ---- mode 2 (copy-engine all-gather consumed in arrival order): each rank is TWO actors — the compute stream, which
# records "x is final" at the start of (#>) number seq and then walks the panels waiting for one source at a time, and
# the comm stream, which after that event copies the local block to r+1, r+2, ... each followed by a flag.

def _arrival_compute(rank, world, nex, win, ev, out):
    for seq in range(1, nex + 1):
        ev[rank] = seq                                   # cudaEventRecord(ev_x0): earlier panel kernels are done
        yield
        b = seq & 1
        got = {}
        for k in range(1, world):
            src = (rank - k) % world
            while win[rank]["flag"][src] < seq:          # p2p_wait_kernel
                yield
            got[src] = win[rank]["buf"][b][src]          # the panel kernel reading xfull
            yield
        out.append((seq, got))


def _arrival_comm(rank, world, nex, win, ev):
    for seq in range(1, nex + 1):
        while ev[rank] < seq:                            # cudaStreamWaitEvent(comm_stream, ev_x0)
            yield
        b = seq & 1
        for k in range(1, world):
            q = (rank + k) % world
            win[q]["buf"][b][rank] = (rank + 1) * 1000 + seq      # cudaMemcpyAsync to the peer window
            yield
            win[q]["flag"][rank] = seq                            # p2p_flag_kernel
            yield


@pytest.mark.parametrize("world", [2, 4, 8])
def test_arrival_order_protocol_model(world):
    for seed in range(40):
        rng = random.Random(seed * 613 + world)
        nex = 10
        win = [{"buf": [[None] * world for _ in range(2)], "flag": [0] * world} for _ in range(world)]
        ev = [0] * world
        outs = [[] for _ in range(world)]
        gens = [_arrival_compute(r, world, nex, win, ev, outs[r]) for r in range(world)]
        gens += [_arrival_comm(r, world, nex, win, ev) for r in range(world)]
        _run(gens, rng)
        for r in range(world):
            assert [s for s, _ in outs[r]] == list(range(1, nex + 1))
            for seq, got in outs[r]:
                assert got == {t: (t + 1) * 1000 + seq for t in range(world) if t != r}, f"rank {r}, exchange {seq}"


# ---- mode 5 (phased push under rotated column panels): compute stream = record "x is final", then per panel p a wait on the flags of
# phase p's sources followed by the panel kernel reading exactly those blocks, then a wait for the own pushes (x_local may be overwritten
# afterwards); comm stream = after the event, per phase: the blocks to the phase's destinations, then (last CTA) the flags on them.

def _phased_compute(rank, world, nex, win, ev, done, peers, out):
    for seq in range(1, nex + 1):
        ev[rank] = seq                                   # cudaEventRecord(ev_x0): the panel kernels of seq - 1 are done
        yield
        b = seq & 1
        got = {}
        for _send, recv in peers:                        # panel p (panel 0: own block, nothing to wait for)
            for src in recv:
                while win[rank]["flag"][src] < seq:      # p2p_wait_kernel on the phase's source flags
                    yield
            for src in recv:
                got[src] = win[rank]["buf"][b][src]      # the panel kernel gathering from those blocks
                yield
        while done[rank] < seq:                          # sla_p2p_arrival_end: my pushes have read x_local
            yield
        out.append((seq, got))


def _phased_comm(rank, world, nex, win, ev, done, peers):
    for seq in range(1, nex + 1):
        while ev[rank] < seq:                            # cudaStreamWaitEvent(comm_stream, ev_x0)
            yield
        b = seq & 1
        for send, _recv in peers:                        # one push kernel per phase
            for q in send:
                win[q]["buf"][b][rank] = (rank + 1) * 1000 + seq
                yield
            for q in send:
                win[q]["flag"][rank] = seq               # raised by the last CTA, after every piece of the phase
                yield
        done[rank] = seq


@pytest.mark.parametrize("world,spec", [(2, None), (3, None), (4, None), (4, "1,3"), (8, None), (8, "1,1,1,1,2,2"), (8, "1,1,1,1,1,1,1,1")])
def test_phased_push_protocol_model(world, spec):
    from sparse_linear_algebra_b200 import dist as sd

    sizes = sd.phase_schedule(world, spec)
    for seed in range(30):
        rng = random.Random(seed * 389 + world)
        nex = 8
        win = [{"buf": [[None] * world for _ in range(2)], "flag": [0] * world} for _ in range(world)]
        ev, done = [0] * world, [0] * world
        outs = [[] for _ in range(world)]
        tables = [sd.phase_peers(r, world, sizes) for r in range(world)]
        gens = [_phased_compute(r, world, nex, win, ev, done, tables[r], outs[r]) for r in range(world)]
        gens += [_phased_comm(r, world, nex, win, ev, done, tables[r]) for r in range(world)]
        _run(gens, rng)
        for r in range(world):
            assert [s for s, _ in outs[r]] == list(range(1, nex + 1))
            for seq, got in outs[r]:
                assert got == {t: (t + 1) * 1000 + seq for t in range(world) if t != r}, f"rank {r}, exchange {seq}, schedule {sizes}"
