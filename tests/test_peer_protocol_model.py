"""CPU suite: an executable model of the peer-memory exchange protocol of ess_dist_bfs / ess_dist_sssp
(essentials_b200/csrc/capi_dist.cu), run under random interleavings.

The GPU tests show the exchange works on the schedules the hardware happened to produce; this model checks the
hazard argument itself. Every rank executes its per-level stream program in order; a random scheduler picks which
rank advances next, so a rank may run arbitrarily far ahead of the others (only the flag waits and the NCCL
collectives hold it back, exactly as on the device). The windows are modelled per (receiver, area, sender) cell:

    CLEAR --sender stores its non-zero words, then raises flag = epoch--> FULL(epoch)
    FULL(epoch) --receiver's consuming kernel reads and zeroes it--> CLEAR            (candidate inbox)
    FULL(epoch) --read in place as the next level's frontier, then memset--> CLEAR    (next-frontier areas)

The next-frontier areas are not unpacked: the area of epoch e IS the replicated frontier bitmap that level e+1's
kernels read; the receiver clears it after those kernels and BEFORE it raises its own level-(e+1) flag, which is what
every peer waits for before it writes that parity again (level e+2).

Because senders skip zero words, a store into a cell that is not CLEAR would leave stale bits behind, and a consumer
that finds anything but FULL(its own epoch) would read missing or future data: both are assertion failures here.
The model also shows that the two design choices the protocol rests on are necessary: with a single gather buffer
(no epoch parity) or without the per-round collective in SSSP, some interleaving breaks.
"""
import random

import pytest

CLEAR = None


class Rank:
    def __init__(self, rank, world, gather_buffers):
        self.rank, self.world = rank, world
        self.inbox = [CLEAR] * world                                   # candidate slices, one cell per sender
        self.gather = [[CLEAR] * world for _ in range(gather_buffers)]  # next-frontier rows, per buffer and sender
        self.flag_a = [0] * world
        self.flag_b = [0] * world
        self.pc = 0
        self.program = []


def bfs_program(levels, first_epoch):
    """Stream order of one rank for one BFS (capi_dist.cu, ess_dist_bfs, peer-memory driver): `levels` is a list of
    'push' / 'pull'. The last level of a run finds nothing, so its slices are empty (nothing is stored)."""
    ops = []
    for i, kind in enumerate(levels):
        epoch = first_epoch + i
        if i > 0:
            ops.append(("read_frontier", epoch - 1))   # this level's kernels read the area the previous level filled
        if kind == "push":
            ops += [("scatter", epoch), ("wait_a", epoch), ("absorb", epoch)]
        if i > 0:
            ops.append(("clear_frontier", epoch - 1))  # memset of that area, enqueued before the publish kernel
        ops += [("broadcast_empty" if i == len(levels) - 1 else "broadcast", epoch), ("wait_b", epoch)]
    ops.append(("collective", first_epoch + len(levels)))  # the seed all_reduce of the next run / end of the run
    return ops


def run_bfs_model(world, runs, rng, gather_buffers=2):
    ranks = [Rank(r, world, gather_buffers) for r in range(world)]
    epoch = 1
    for levels in runs:
        for rk in ranks:
            rk.program += bfs_program(levels, epoch)
        epoch += len(levels)
    arrived = {}  # collective id -> ranks that reached it
    while any(rk.pc < len(rk.program) for rk in ranks):
        runnable = []
        for rk in ranks:
            if rk.pc == len(rk.program):
                continue
            op, e = rk.program[rk.pc]
            if op == "wait_a" and min(rk.flag_a) < e:
                continue
            if op == "wait_b" and min(rk.flag_b) < e:
                continue
            if op == "collective":
                arrived.setdefault(e, set()).add(rk.rank)
                if len(arrived[e]) < world:
                    continue
            runnable.append(rk)
        assert runnable, "deadlock: every unfinished rank is blocked"
        rk = rng.choice(runnable)
        op, e = rk.program[rk.pc]
        buf = e % gather_buffers
        if op == "scatter":
            for peer in ranks:
                assert peer.inbox[rk.rank] is CLEAR, f"rank {rk.rank} overwrites an unconsumed inbox cell of {peer.rank}"
                peer.inbox[rk.rank] = e
                peer.flag_a[rk.rank] = e
        elif op == "absorb":
            for s in range(world):
                assert rk.inbox[s] == e, f"rank {rk.rank} absorbs epoch {rk.inbox[s]} from {s}, wants {e}"
                rk.inbox[s] = CLEAR
        elif op in ("broadcast", "broadcast_empty"):
            for peer in ranks:
                assert peer.gather[buf][rk.rank] is CLEAR, \
                    f"rank {rk.rank} (epoch {e}) stores into area[{buf}] of {peer.rank} before it was cleared"
                if op == "broadcast":
                    peer.gather[buf][rk.rank] = e
                peer.flag_b[rk.rank] = e
        elif op == "read_frontier":
            for s in range(world):
                assert rk.gather[buf][s] == e, f"rank {rk.rank} reads frontier epoch {rk.gather[buf][s]} from {s}, wants {e}"
        elif op == "clear_frontier":
            rk.gather[buf] = [CLEAR] * world
        rk.pc += 1
    for rk in ranks:  # the windows are all-clear again after a completed run (what the next run relies on)
        assert all(c is CLEAR for c in rk.inbox) and all(c is CLEAR for g in rk.gather for c in g)


def random_levels(rng):
    n = rng.randint(1, 9)
    return [rng.choice(["push", "pull"]) for _ in range(n)]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_bfs_exchange_has_no_hazard_under_random_interleavings(world):
    rng = random.Random(1000 + world)
    for _ in range(300):
        runs = [random_levels(rng) for _ in range(rng.randint(1, 3))]
        run_bfs_model(world, runs, rng)
    # the schedules that stress the double buffer most: only bottom-up levels, odd and even run lengths back to back
    for _ in range(100):
        run_bfs_model(world, [["pull"] * rng.randint(1, 7), ["pull"] * rng.randint(1, 7)], rng)


def test_single_gather_buffer_would_be_a_hazard():
    """Without the epoch-parity double buffer a fast peer delivers level L+1 into the area this rank still reads."""
    rng = random.Random(7)
    with pytest.raises(AssertionError):
        for _ in range(500):
            run_bfs_model(4, [["pull"] * 6], rng, gather_buffers=1)


# ---------------------------------------------------------------------------------------------------- SSSP
def run_sssp_model(world, rounds, rng, collective_per_round=True):
    """ess_dist_sssp, peer-memory path: per round  clear own dirty words -> relax (lowers entries of the OWN replica,
    raises dirty words) -> signal -> wait for all signals -> owners read the peers' dirty words and replicas ->
    all_reduce of the counts (also the barrier that lets the next round clear the dirty words)."""
    state = [{"dirty_epoch": 0, "being_read_by": set()} for _ in range(world)]
    flags = [[0] * world for _ in range(world)]
    programs = []
    for r in range(world):
        ops = []
        for e in range(1, rounds + 1):
            ops += [("clear_and_relax", e), ("signal", e), ("wait", e), ("reduce_begin", e), ("reduce_end", e)]
            if collective_per_round:
                ops.append(("collective", e))
        programs.append(ops)
    pc = [0] * world
    arrived = {}
    while any(pc[r] < len(programs[r]) for r in range(world)):
        runnable = []
        for r in range(world):
            if pc[r] == len(programs[r]):
                continue
            op, e = programs[r][pc[r]]
            if op == "wait" and min(flags[r]) < e:
                continue
            if op == "collective":
                arrived.setdefault(e, set()).add(r)
                if len(arrived[e]) < world:
                    continue
            runnable.append(r)
        assert runnable, "deadlock"
        r = rng.choice(runnable)
        op, e = programs[r][pc[r]]
        if op == "clear_and_relax":
            assert not state[r]["being_read_by"], \
                f"rank {r} clears its dirty words for round {e} while {state[r]['being_read_by']} still read round {e - 1}"
            state[r]["dirty_epoch"] = e
        elif op == "signal":
            for peer in range(world):
                flags[peer][r] = e
        elif op == "reduce_begin":
            for peer in range(world):
                if peer != r:
                    assert state[peer]["dirty_epoch"] >= e, "owner reads dirty words its peer has not produced yet"
                    state[peer]["being_read_by"].add(r)
        elif op == "reduce_end":
            for peer in range(world):
                state[peer]["being_read_by"].discard(r)
        pc[r] += 1


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sssp_exchange_has_no_hazard_under_random_interleavings(world):
    rng = random.Random(2000 + world)
    for _ in range(200):
        run_sssp_model(world, rng.randint(1, 12), rng)


def test_sssp_needs_its_per_round_collective():
    """The 2-word all_reduce is not only the termination test: without it a fast rank would clear the dirty words a
    slow owner is still reading."""
    rng = random.Random(3)
    with pytest.raises(AssertionError):
        for _ in range(500):
            run_sssp_model(4, 6, rng, collective_per_round=False)
