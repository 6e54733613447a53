"""Model check of the push exchanges inside the hybrid PCG loop (csrc/dist.cu push_kernel).

push_kernel stores a rank's boundary planes straight into its neighbours' ghost planes and raises a flag; it waits for the
neighbours' flags but NOT for the neighbours to have finished reading the previous content of those ghost planes.  The claim
that makes this safe: between two uses of a ghost plane the loop always passes an all-rank reduction or gather.  This test
replays the loop's exchange schedule (mg.cu cycle(), pcg.cu enqueue_iteration) on R asynchronous ranks under random
interleavings and asserts, for every remote store, that the destination's previous content has been consumed by all of its
readers, and for every read, that the ghost plane holds exactly the neighbour's data of the matching exchange.
"""
import random

import pytest

# one PCG iteration of a rank, in stream order.  ("push", a): push my boundary planes of array a into the neighbours' ghosts
# and wait for theirs; ("read", a): a kernel that reads my ghost planes of a; ("ar",): all-rank reduction; ("gpush",):
# all-rank push of the coarse right-hand side (waits for everybody's flags); ("write", a): a kernel that rewrites my owned
# planes of a (it never touches ghost planes: float4 kernels only store active groups)
CYCLE_REST = [                                # the multigrid cycle after its first sweep
    ("push", "xa"), ("read", "xa"), ("write", "xb"),   # pre-sweep 2
    ("push", "xb"), ("read", "xb"),           # restriction
    ("gpush",),                               # coarse right-hand side to every rank; coarse levels replicated
    ("read", "xb"), ("write", "xa"),          # prolongation + sweep (reads the same halo of xb)
    ("push", "xa"), ("read", "xa"), ("write", "xb"),   # last sweep (fused with z.r)
    ("ar",),                                  # z.r
]
ITERATION = [
    ("push", "s"), ("read", "s"),            # halo of the search direction, SpMV
    ("ar",),                                  # s.As
    ("write", "xa"),                          # CG update fused with the cycle's first sweep (no neighbour reads)
    ("ar",),                                  # ||r||_inf (decision)
] + CYCLE_REST + [
    ("write", "s"),                           # direction update
]
FIRST_CYCLE = [("write", "xa")] + CYCLE_REST  # the cycle before the loop (mg_apply ... start_kernel, AR_START)
PROGRAM_HEAD = FIRST_CYCLE + [("write", "s")]


def simulate(nranks, iterations, seed, drop_barriers=False, push_waits=True):
    rng = random.Random(seed)
    prog = PROGRAM_HEAD + ITERATION * iterations
    if drop_barriers:  # negative control: without the reductions the schedule must be able to race
        prog = [op for op in prog if op[0] not in ("ar", "gpush")]
    pc = [0] * nranks
    # ghost[(rank, array, side)] = {"ver": exchange index of the content, "readers_left": reads of this content still to come}
    ghost = {}
    pushed = {}    # (rank, array) -> number of pushes done (exchange index)
    flags = {}     # (rank, array, side) -> exchange index signalled by that neighbour
    barrier_count = [0] * nranks
    # how many reads consume one push of each array before the next push of the same array
    reads_after_push = {}
    for i, op in enumerate(prog):
        if op[0] == "push":
            n = 0
            for nxt in prog[i + 1:]:
                if nxt == op:
                    break
                if nxt == ("read", op[1]):
                    n += 1
            reads_after_push.setdefault(op[1], n)

    def neighbours(r):
        return [(r - 1, 1), (r + 1, 0)]  # (neighbour rank, which ghost side of the neighbour I write: I am its upper -> side 1)

    phase = ["start"] * nranks  # push ops have two halves: store + signal, then wait
    steps = 0
    while any(pc[r] < len(prog) for r in range(nranks)):
        steps += 1
        assert steps < 10_000_000
        r = rng.randrange(nranks)
        if pc[r] >= len(prog):
            continue
        op = prog[pc[r]]
        if op[0] == "push":
            a = op[1]
            if phase[r] == "start":
                k = pushed.get((r, a), 0) + 1
                for nb, side in neighbours(r):
                    if 0 <= nb < nranks:
                        g = ghost.setdefault((nb, a, side), {"ver": 0, "readers_left": 0})
                        assert g["readers_left"] == 0, f"rank {r} overwrites ghost {a} of rank {nb} before it was read (exchange {k})"
                        g["ver"], g["readers_left"] = k, reads_after_push[a]
                        flags[(nb, a, side)] = k
                pushed[(r, a)] = k
                phase[r] = "wait"
            else:
                k = pushed[(r, a)]
                need = [(r, a, 0 if nb < r else 1) for nb, _ in neighbours(r) if 0 <= nb < nranks]
                if not push_waits or all(flags.get(key, 0) >= k for key in need):
                    phase[r] = "start"
                    pc[r] += 1
        elif op[0] == "read":
            a = op[1]
            k = pushed[(r, a)]
            for nb, _ in neighbours(r):
                if 0 <= nb < nranks:
                    g = ghost.setdefault((r, a, 0 if nb < r else 1), {"ver": 0, "readers_left": 0})
                    assert g["ver"] == k, f"rank {r} reads exchange {g['ver']} of {a}, expected {k}"
                    g["readers_left"] -= 1
            pc[r] += 1
        elif op[0] in ("ar", "gpush"):
            if phase[r] == "start":
                barrier_count[r] += 1
                phase[r] = "wait"
            elif all(barrier_count[q] >= barrier_count[r] for q in range(nranks)):
                phase[r] = "start"
                pc[r] += 1
        else:  # local write of owned planes
            pc[r] += 1
    return steps


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_push_schedule_never_overwrites_unread_ghosts(nranks):
    for seed in range(60):
        simulate(nranks, iterations=6, seed=seed)


def test_pairwise_flag_waits_alone_are_sufficient_too():
    """Stronger than the documented argument: even without the reductions the schedule is safe, because a push also waits for
    the neighbour's push of the same exchange, and the neighbour's stream puts its reads of the previous array before that."""
    for seed in range(60):
        simulate(3, iterations=6, seed=seed, drop_barriers=True)


def test_model_detects_a_race_when_nothing_orders_the_ranks():
    """Negative control: a fire-and-forget push (no flag wait) in a schedule without reductions must be caught by the same
    assertions under some interleaving -- the checks above are not vacuous."""
    found = False
    for seed in range(200):
        try:
            simulate(3, iterations=6, seed=seed, drop_barriers=True, push_waits=False)
        except AssertionError:
            found = True
            break
    assert found


# ---- fused exchanges (csrc/fexch.cuh): the producer kernel stores its boundary planes into the neighbours' ghost planes and
# adds to their arrival counters without waiting for anything; the consumer kernel's boundary CTAs wait until the arrivals
# match the number of producer launches this rank has done itself for that array ------------------------------------------------
# ("wpush", a): producer of a (local write + remote stores + signal); ("read", a): consumer (waits, then reads the ghosts);
# ("push", "s"): the one stand-alone push per solve (start_kernel has no fused store)
FUSED_CYCLE_REST = [
    ("read", "xa"), ("wpush", "xb"),          # pre-sweep 2 (mg_jacobi4_kernel: waits for xa, pushes xb)
    ("read", "xb"),                           # restriction ...
    ("gstore",),                              # ... whose CTAs store the coarse right-hand side into every rank's array and count there
    ("gwait",),                               # first level-1 kernel: waits until every other rank's restriction has counted
    ("gdone",),                               # last level-1 kernel that reads the coarse right-hand side (replicated coarse levels)
    ("read", "xb"), ("wpush", "xa"),          # prolongation + sweep
    ("read", "xa"), ("write", "xb"),          # last sweep (fused with z.r inside the loop)
    ("ar",),                                  # z.r / AR_START
]
FUSED_HEAD = [("wpush", "xa")] + FUSED_CYCLE_REST + [("write", "s"), ("push", "s")]
FUSED_ITERATION = [
    ("read", "s"), ("ar",),                   # SpMV (waits for s), s.As
    ("wpush", "xa"), ("ar",),                 # CG update + first sweep (pushes xa), ||r||_inf
] + FUSED_CYCLE_REST + [
    ("wpush", "s"),                           # direction update (pushes s)
]


def simulate_fused(nranks, iterations, seed, drop_barriers=False, reads_wait=True):
    rng = random.Random(seed)
    prog = FUSED_HEAD + FUSED_ITERATION * iterations
    if drop_barriers:
        prog = [op for op in prog if op[0] not in ("ar", "gstore", "gwait", "gdone")]
    reads_per_push = {}
    for i, op in enumerate(prog):
        if op[0] in ("push", "wpush"):
            n = 0
            for nxt in prog[i + 1:]:
                if nxt[0] in ("push", "wpush") and nxt[1] == op[1]:
                    break
                if nxt == ("read", op[1]):
                    n += 1
            assert reads_per_push.setdefault(op[1], n) == n or i + 40 > len(prog)  # (the tail of the program may be cut short)
    pc = [0] * nranks
    ghost, pushed, arrived = {}, {}, {}
    barrier_count = [0] * nranks
    gcount, gread = [0] * nranks, [0] * nranks  # all-rank push: exchanges stored by each rank / consumed by each rank
    phase = ["start"] * nranks

    def sides(r):  # (neighbour, the ghost side of the neighbour that I write, my ghost side that it writes)
        return [(nb, s_theirs, s_mine) for nb, s_theirs, s_mine in ((r - 1, 1, 0), (r + 1, 0, 1)) if 0 <= nb < nranks]

    steps = 0
    while any(pc[r] < len(prog) for r in range(nranks)):
        steps += 1
        assert steps < 10_000_000
        r = rng.randrange(nranks)
        if pc[r] >= len(prog):
            continue
        op = prog[pc[r]]
        if op[0] in ("wpush", "push"):
            a = op[1]
            if phase[r] == "start":
                k = pushed.get((r, a), 0) + 1
                for nb, s_theirs, _ in sides(r):
                    g = ghost.setdefault((nb, a, s_theirs), {"ver": 0, "readers_left": 0})
                    assert g["readers_left"] == 0, f"rank {r} overwrites ghost {a} of rank {nb} before it was read (exchange {k})"
                    g["ver"], g["readers_left"] = k, reads_per_push[a]
                    arrived[(nb, a, s_theirs)] = k
                pushed[(r, a)] = k
                if op[0] == "wpush":
                    pc[r] += 1          # fire and forget
                else:
                    phase[r] = "wait"   # the stand-alone push_kernel also waits for the neighbours' flags
            elif all(arrived.get((r, a, s_mine), 0) >= pushed[(r, a)] for _, _, s_mine in sides(r)):
                phase[r] = "start"
                pc[r] += 1
        elif op[0] == "read":
            a = op[1]
            k = pushed[(r, a)]  # expectation = my own number of producer launches for this array
            if reads_wait and not all(arrived.get((r, a, s_mine), 0) >= k for _, _, s_mine in sides(r)):
                continue        # the boundary CTAs spin
            for _, _, s_mine in sides(r):
                g = ghost.setdefault((r, a, s_mine), {"ver": 0, "readers_left": 0})
                assert g["ver"] == k, f"rank {r} reads exchange {g['ver']} of {a}, expected {k}"
                g["readers_left"] -= 1
            pc[r] += 1
        elif op[0] == "ar":
            if phase[r] == "start":
                barrier_count[r] += 1
                phase[r] = "wait"
            elif all(barrier_count[q] >= barrier_count[r] for q in range(nranks)):
                phase[r] = "start"
                pc[r] += 1
        elif op[0] == "gstore":   # fire and forget: remote stores into every rank's coarse array + one count per rank
            gcount[r] += 1
            for q in range(nranks):
                if q != r:
                    assert gread[q] >= gcount[r] - 1, f"rank {r} overwrites the coarse right-hand side of rank {q} before it was read"
            pc[r] += 1
        elif op[0] == "gwait":    # consumer of the coarse right-hand side: needs every other rank's stores of this exchange
            if all(gcount[q] >= gcount[r] for q in range(nranks)):
                pc[r] += 1
        elif op[0] == "gdone":
            gread[r] = gcount[r]
            pc[r] += 1
        else:
            pc[r] += 1
    return steps


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_fused_schedule_never_overwrites_unread_ghosts(nranks):
    for seed in range(60):
        simulate_fused(nranks, iterations=6, seed=seed)


def test_fused_schedule_is_ordered_by_its_data_dependencies_alone():
    """Every producer of an array sits behind a consumer wait whose matching neighbour store comes after the neighbour's last
    read of the old contents, so the schedule stays race-free even with the reductions removed."""
    for seed in range(60):
        simulate_fused(3, iterations=6, seed=seed, drop_barriers=True)


def test_fused_model_detects_a_consumer_that_does_not_wait():
    found = False
    for seed in range(200):
        try:
            simulate_fused(3, iterations=6, seed=seed, reads_wait=False)
        except AssertionError:
            found = True
            break
    assert found


# ---- allreduce_kernel: slot[parity][source rank] on every rank, double-buffered by the parity of the epoch -----------------
def simulate_allreduce(nranks, rounds, seed, nbuf=2):
    """Every rank: for ep = 1..rounds: store (value(rank, ep), ep) into slot[ep % nbuf][rank] of EVERY rank (one store at a
    time, any interleaving), then wait until its own slots of that buffer all carry epoch >= ep, then read them.  Checks that
    what is read is the value of exactly that epoch (i.e. nobody was able to overwrite a slot that was still to be read)."""
    rng = random.Random(seed)
    slots = [[[(None, 0)] * nranks for _ in range(nbuf)] for _ in range(nranks)]   # slots[owner][buffer][source] = (value, epoch)
    ep = [1] * nranks
    stage = [0] * nranks          # 0..nranks-1: next peer to store to; nranks: waiting
    steps = 0
    while any(e <= rounds for e in ep):
        steps += 1
        assert steps < 5_000_000
        r = rng.randrange(nranks)
        if ep[r] > rounds:
            continue
        e, b = ep[r], ep[r] % nbuf
        if stage[r] < nranks:
            slots[stage[r]][b][r] = ((r, e), e)
            stage[r] += 1
        elif all(slots[r][b][q][1] >= e for q in range(nranks)):
            for q in range(nranks):
                assert slots[r][b][q][0] == (q, e), f"rank {r} epoch {e}: slot of rank {q} holds {slots[r][b][q]}"
            ep[r] += 1
            stage[r] = 0
    return steps


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_allreduce_double_buffering_is_enough(nranks):
    for seed in range(40):
        simulate_allreduce(nranks, rounds=12, seed=seed)


def test_allreduce_model_catches_a_single_buffer():
    found = False
    for seed in range(200):
        try:
            simulate_allreduce(3, rounds=12, seed=seed, nbuf=1)
        except AssertionError:
            found = True
            break
    assert found
