"""Model check of the mbarrier protocol of the experimental INT8 ZGEMM
(csrc/kernels_zgemm_ozaki.cu): one MMA-issuing thread, 16 worker warps, barriers
``planes`` (count 16), ``done[g]`` (count 1, arrived by tcgen05.commit when the MMAs of
accumulator group g complete, asynchronously but in order) and ``freed[g]`` (count 16).

The kernel cannot run in this container, so the wait / arrive sequence of both roles is
restated here with the kernel's own parity arithmetic and executed under many random
interleavings.  Checked: no deadlock; the tensor core never reads planes that are being
rewritten; a worker never reads an accumulator group before its MMAs completed or after the
next (tile, half) started overwriting it; the MMA thread never overwrites a group a warp has
still to read.  (A wrong parity or a missing wait fails these within a few seeds.)
"""
import random

import pytest

NW = 16


class MBar:
    """mbarrier with the PTX phase semantics: try_wait.parity(P) succeeds when the phase of
    parity P has completed, i.e. when the current (incomplete) phase has the other parity."""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def test(self, parity):
        return (self.phase & 1) != (parity & 1)


def mma_thread(st, tiles, NH, G):
    """lane 0 of the MMA warp (kernel lines: wait planes; per group: wait freed, issue, commit)"""
    it = 0
    for tile_no in range(tiles):
        yield ("wait", st["planes"], tile_no & 1)
        for h in range(NH):
            for g in range(G):
                if it > 0:
                    yield ("wait", st["freed"][g], (it - 1) & 1)
                yield ("issue", tile_no, it, g)
            it += 1


def worker(st, w, tiles, NH, G):
    it = 0
    for tile_no in range(tiles):
        yield ("bar",)                       # workers_barrier (row exponents)
        yield ("slice", tile_no)             # writes this warp's part of the A planes
        yield ("arrive", st["planes"])
        for h in range(NH):
            for g in range(G):
                yield ("wait", st["done"][g], it & 1)
                yield ("read", it, g)
                yield ("arrive", st["freed"][g])
            it += 1


def run(seed, tiles, NH, G):
    rng = random.Random(seed)
    st = {"planes": MBar(NW), "done": [MBar(1) for _ in range(G)], "freed": [MBar(NW) for _ in range(G)]}
    agents = {"mma": mma_thread(st, tiles, NH, G)}
    agents.update({w: worker(st, w, tiles, NH, G) for w in range(NW)})
    pending = {k: None for k in agents}          # the action an agent is blocked on
    inflight = []                                # issued, not yet completed MMA groups (FIFO)
    completed = {}                               # (it, g) -> True once the tensor core finished
    issued = set()
    reads = {}                                   # (it, g) -> set of warps that have read it
    sliced = {t: set() for t in range(tiles + 1)}
    at_bar = set()
    bar_gen = 0
    live = set(agents)
    steps = 0
    while live or inflight:
        steps += 1
        assert steps < 10_000_000
        choices = list(live) + (["tc"] if inflight else [])
        rng.shuffle(choices)
        progressed = False
        for who in choices:
            if who == "tc":                      # the tensor core completes the oldest group
                tile_no, it, g = inflight.pop(0)
                completed[(it, g)] = True
                st["done"][g].arrive()
                progressed = True
                break
            if who in at_bar:
                continue
            act = pending[who]
            if act is None:
                try:
                    act = next(agents[who])
                except StopIteration:
                    live.discard(who)
                    progressed = True
                    break
            kind = act[0]
            if kind == "wait":
                if not act[1].test(act[2]):
                    pending[who] = act
                    continue
            elif kind == "arrive":
                act[1].arrive()
            elif kind == "bar":
                at_bar.add(who)
                if len(at_bar) == NW:
                    at_bar.clear()
                    bar_gen += 1
            elif kind == "slice":
                t = act[1]
                # no MMA of an earlier tile may still be reading the planes
                assert not inflight, "planes rewritten while MMAs are in flight"
                sliced[t].add(who)
            elif kind == "issue":
                _, tile_no, it, g = act
                assert len(sliced[tile_no]) == NW, "MMA issued before all planes were written"
                assert not sliced[tile_no + 1], "MMA issued while the next tile is being sliced"
                if it > 0:
                    assert len(reads.get((it - 1, g), ())) == NW, "accumulator overwritten before all warps read it"
                issued.add((it, g))
                inflight.append((tile_no, it, g))
            elif kind == "read":
                _, it, g = act
                assert completed.get((it, g)), "accumulator read before its MMAs completed"
                assert (it + 1, g) not in issued, "accumulator read after the next MMAs were issued"
                reads.setdefault((it, g), set()).add(who)
            pending[who] = None
            progressed = True
            break
        assert progressed, "deadlock: %r" % {k: v for k, v in pending.items() if v is not None}
    total = tiles * NH * G
    assert len(completed) == total and all(len(v) == NW for v in reads.values()) and len(reads) == total


@pytest.mark.parametrize("tiles,NH,G", [(1, 1, 6), (1, 2, 6), (3, 2, 6), (4, 1, 7), (5, 2, 7)])
def test_protocol_random_interleavings(tiles, NH, G):
    for seed in range(40):
        run(seed, tiles, NH, G)


def test_model_catches_a_missing_wait():
    """The same harness with the `freed` wait removed must fail (the check has teeth)."""
    def broken_mma(st, tiles, NH, G):
        it = 0
        for tile_no in range(tiles):
            yield ("wait", st["planes"], tile_no & 1)
            for h in range(NH):
                for g in range(G):
                    yield ("issue", tile_no, it, g)
                it += 1
    global mma_thread
    good = mma_thread
    mma_thread = broken_mma
    try:
        with pytest.raises(AssertionError):
            for seed in range(40):
                run(seed, 3, 2, 6)
    finally:
        mma_thread = good


# ---------------------------------------------------------------------------
# k_zgemm_ozaki_kloop: K walked in chunks; planes rewritten per chunk behind `consumed`
# ---------------------------------------------------------------------------
def kloop_mma_thread(st, tiles, chunks, G):
    it = cc = 0
    for _ in range(tiles):
        for c in range(chunks):
            yield ("wait", st["planes"], cc & 1)
            for g in range(G):
                if c == 0 and it > 0:
                    yield ("wait", st["freed"][g], (it - 1) & 1)
                yield ("kissue", it, cc, c, g, c == chunks - 1)
            yield ("kcommit_consumed", cc)
            cc += 1
        it += 1


def kloop_worker(st, w, tiles, chunks, G):
    it = cc = 0
    for _ in range(tiles):
        for c in range(chunks):
            if cc > 0:
                yield ("wait", st["consumed"], (cc - 1) & 1)
            yield ("kslice", cc)
            yield ("arrive", st["planes"])
            cc += 1
        for g in range(G):
            yield ("wait", st["done"][g], it & 1)
            yield ("read", it, g)
            yield ("arrive", st["freed"][g])
        it += 1


def run_kloop(seed, tiles, chunks, G):
    rng = random.Random(seed)
    st = {"planes": MBar(NW), "consumed": MBar(1), "done": [MBar(1) for _ in range(G)],
          "freed": [MBar(NW) for _ in range(G)]}
    agents = {"mma": kloop_mma_thread(st, tiles, chunks, G)}
    agents.update({w: kloop_worker(st, w, tiles, chunks, G) for w in range(NW)})
    pending = {k: None for k in agents}
    inflight = []            # FIFO of async tensor-core events: ("mma", it, cc, g, last) / ("consumed", cc)
    mma_done_chunk = set()   # chunks whose MMAs have all completed
    completed, issued, reads = {}, set(), {}
    sliced = {}
    live = set(agents)
    steps = 0
    while live or inflight:
        steps += 1
        assert steps < 10_000_000
        choices = list(live) + (["tc"] if inflight else [])
        rng.shuffle(choices)
        progressed = False
        for who in choices:
            if who == "tc":
                ev = inflight.pop(0)
                if ev[0] == "mma":
                    _, it, cc, g, last = ev
                    if last:
                        completed[(it, g)] = True
                        st["done"][g].arrive()
                else:
                    mma_done_chunk.add(ev[1])
                    st["consumed"].arrive()
                progressed = True
                break
            act = pending[who]
            if act is None:
                try:
                    act = next(agents[who])
                except StopIteration:
                    live.discard(who)
                    progressed = True
                    break
            kind = act[0]
            if kind == "wait":
                if not act[1].test(act[2]):
                    pending[who] = act
                    continue
            elif kind == "arrive":
                act[1].arrive()
            elif kind == "kslice":
                cc = act[1]
                assert cc == 0 or (cc - 1) in mma_done_chunk, "planes rewritten while the previous chunk is in use"
                sliced.setdefault(cc, set()).add(who)
            elif kind == "kissue":
                _, it, cc, c, g, last = act
                assert len(sliced.get(cc, ())) == NW, "MMA issued before all planes of the chunk were written"
                assert not sliced.get(cc + 1), "MMA issued while the next chunk is being written"
                if c == 0 and it > 0:
                    assert len(reads.get((it - 1, g), ())) == NW, "accumulator overwritten before all warps read it"
                if c == 0:
                    issued.add((it, g))
                inflight.append(("mma", it, cc, g, last))
            elif kind == "kcommit_consumed":
                inflight.append(("consumed", act[1]))
            elif kind == "read":
                _, it, g = act
                assert completed.get((it, g)), "accumulator read before its MMAs completed"
                assert (it + 1, g) not in issued, "accumulator read after the next tile started on it"
                reads.setdefault((it, g), set()).add(who)
            pending[who] = None
            progressed = True
            break
        assert progressed, "deadlock: %r" % {k: v for k, v in pending.items() if v is not None}
    assert len(completed) == tiles * G and all(len(v) == NW for v in reads.values())


@pytest.mark.parametrize("tiles,chunks,G", [(1, 1, 6), (1, 4, 6), (3, 2, 7), (4, 3, 4), (2, 5, 6)])
def test_kloop_protocol_random_interleavings(tiles, chunks, G):
    for seed in range(30):
        run_kloop(seed, tiles, chunks, G)
