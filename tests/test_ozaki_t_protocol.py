"""Model check of the mbarrier protocol of the second-generation INT8 kernel
(csrc/kernels_zgemm_ozaki2.cu, k_ozaki_t): 8 producer warps, one MMA-issuing thread, 8 epilogue
warps, the tensor core as an in-order asynchronous agent; barriers ``full[s]`` (count 8),
``empty[s]`` (tcgen05.commit), ``done[b][g]`` (tcgen05.commit), ``freed[b][g]`` (count 8),
``wready`` (count 8: the epilogue warps have built the resident operand W), two X stages.  Accumulators: ComplexF32 (4 groups) all double-buffered; ComplexF64 (6 groups) groups
4 and 5 double-buffered and issued first, groups 0..3 single-buffered (``ot_dbuf`` /
``ot_issue_order`` in the kernel); the variant with W planes in tensor memory (``WT = 4``)
double-buffers nothing and issues the groups in natural order.

The wait / arrive sequence of every role is restated with the kernel's own parity arithmetic
and run under random interleavings.  Checked: no deadlock; a producer never rewrites a stage
the tensor core may still read; no MMA is issued before its tile's planes are complete or into
an accumulator an epilogue warp has still to read; an epilogue warp only reads completed
accumulators of ITS tile.  A wrong parity or a missing wait fails within a few seeds (see
``test_broken_protocols_are_caught``).
"""
import random

import pytest

NP_ = 8   # producer warps
NE = 8    # epilogue warps
NST = 2   # X stages


class MBar:
    """mbarrier with the PTX phase semantics: try_wait.parity(P) succeeds once the phase with
    parity P has completed."""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def test(self, parity):
        return (self.phase & 1) != (parity & 1)


WT = 0   # digit planes of W held in tensor memory (ComplexF64 variant k_ozaki_t<double, false, 4>):
         # the spare columns are taken, no group is double-buffered, groups in natural order


def dbuf(G, g):
    return G == 4 or (WT == 0 and g >= 4)


def buf_use(t, G, g):
    return (t & 1, t >> 1) if dbuf(G, g) else (0, t)


def issue_order(G):
    return list(range(G)) if (G == 4 or WT > 0) else [5, 4, 0, 1, 2, 3]


NSUB = 8   # sub-chunks of 4 tile rows per epilogue warp


def producer(st, tiles, broken):
    for t in range(tiles):
        stage, use = t & 1, t >> 1
        yield ("load", t)
        if use > 0 and broken != "no_empty_wait":
            yield ("wait", st["empty"][stage], (use - 1) & 1)
        yield ("write", t, stage)
        yield ("arrive", st["full"][stage])


def mma_thread(st, tiles, G, broken):
    if broken != "no_wready_wait":
        yield ("wait", st["wready"], 0)
    for t in range(tiles):
        stage = t & 1
        yield ("wait", st["full"][stage], ((t >> 1) & 1) if broken != "full_parity" else (t & 1))
        for g in issue_order(G):
            buf, u = buf_use(t, G, g)
            if u > 0 and broken != "no_freed_wait":
                yield ("wait", st["freed"][buf][g], (u - 1) & 1)
            yield ("issue", t, g, stage, buf)
            yield ("commit", st["done"][buf][g])
        yield ("commit", st["empty"][stage])


def epilogue(st, tiles, G, broken):
    # the epilogue warps build W (the resident operand) while the producers fetch the first tile
    for _ in range(12):          # gather, slice, store: longer than a producer's first tile
        yield ("setup_step",)
    yield ("setup",)
    yield ("arrive", st["wready"])
    for t in range(tiles):
        for sc in range(NSUB):
            for g in range(G):
                buf, u = buf_use(t, G, g)
                if sc == 0:
                    yield ("wait", st["done"][buf][g], (u & 1) if broken != "done_parity" else ((u + 1) & 1))
                yield ("read", t, g, buf)
            if sc == NSUB - 1:
                for g in range(G):
                    yield ("arrive", st["freed"][buf_use(t, G, g)[0]][g])


def run(seed, tiles, G, broken=None):
    NB = 2
    rng = random.Random(seed)
    st = {"full": [MBar(NP_) for _ in range(NST)], "empty": [MBar(1) for _ in range(NST)],
          "done": [[MBar(1) for _ in range(G)] for _ in range(NB)],
          "freed": [[MBar(NE) for _ in range(G)] for _ in range(NB)],
          "wready": MBar(NE)}
    agents = {("p", w): producer(st, tiles, broken) for w in range(NP_)}
    agents["mma"] = mma_thread(st, tiles, G, broken)
    agents.update({("e", w): epilogue(st, tiles, G, broken) for w in range(NE)})
    blocked = {k: None for k in agents}
    queue = []                    # tensor-core FIFO: ("mma", t, g, stage, buf) / ("commit", bar)
    completed = set()             # (t, g) whose MMAs have finished
    issued = set()                # (t, g) issued
    written = {}                  # tile -> producer warps that have written it
    reads = {}                    # (t, g) -> number of (warp, cb) reads done
    w_rows = [0]                  # epilogue warps that have stored their rows of W
    live = set(agents)
    while live or queue:
        choices = [k for k in live]
        if queue:
            choices.append("tc")
        rng.shuffle(choices)
        progressed = False
        for k in choices:
            if k == "tc":
                item = queue.pop(0)
                if item[0] == "mma":
                    completed.add((item[1], item[2]))
                else:
                    item[1].arrive()
                progressed = True
                break
            act = blocked[k]
            if act is None:
                try:
                    act = next(agents[k])
                except StopIteration:
                    live.discard(k)
                    progressed = True
                    break
            if act[0] == "wait":
                if not act[1].test(act[2]):
                    blocked[k] = act
                    continue
                blocked[k] = None
            elif act[0] == "arrive":
                act[1].arrive()
            elif act[0] == "write":
                _, t, stage = act
                # the tensor core must be done with the tile that used this stage before
                if t >= NST:
                    for g in range(G):
                        assert (t - NST, g) in completed, ("stage rewritten while in use", t, g)
                written[t] = written.get(t, 0) + 1
            elif act[0] == "setup_step":
                pass
            elif act[0] == "setup":
                w_rows[0] += 1
            elif act[0] == "issue":
                _, t, g, stage, buf = act
                assert w_rows[0] == NE, ("MMA issued before W is complete", t)
                assert written.get(t, 0) == NP_, ("MMA issued before the planes are complete", t)
                prev = t - (2 if dbuf(G, g) else 1)
                if prev >= 0:
                    assert reads.get((prev, g), 0) == NE * NSUB, ("accumulator overwritten before it was read", t, g)
                issued.add((t, g))
                queue.append(("mma", t, g, stage, buf))
            elif act[0] == "commit":
                queue.append(("commit", act[1]))
            elif act[0] == "read":
                _, t, g, buf = act
                assert (t, g) in completed, ("accumulator read before its MMAs completed", t, g)
                nxt = t + (2 if dbuf(G, g) else 1)
                assert (nxt, g) not in issued, ("accumulator read after the next tile started on it", t, g)
                reads[(t, g)] = reads.get((t, g), 0) + 1
            progressed = True
            break
        assert progressed, ("deadlock", {k: v for k, v in blocked.items() if v is not None and k in live})
    for t in range(tiles):
        for g in range(G):
            assert reads.get((t, g), 0) == NE * NSUB


@pytest.mark.parametrize("G", [6, 4])
def test_protocol_random_interleavings(G):
    for seed in range(60):
        run(seed, tiles=1 + seed % 7, G=G)


def test_protocol_w_planes_in_tensor_memory(monkeypatch):
    """The variant without double-buffered accumulators (option ozaki_tsw = 2)."""
    import sys
    monkeypatch.setattr(sys.modules[__name__], "WT", 4)
    for seed in range(60):
        run(seed, tiles=1 + seed % 7, G=6)


@pytest.mark.parametrize("broken", ["no_empty_wait", "no_freed_wait", "full_parity", "done_parity", "no_wready_wait"])
def test_broken_protocols_are_caught(broken):
    caught = 0
    for seed in range(40):
        try:
            run(seed, tiles=6, G=6, broken=broken)
        except AssertionError:
            caught += 1
    assert caught > 0, broken
