"""Executable model of the DYNAMICALLY SCHEDULED pipeline of csrc/glm_tc.cu (round 2): chunk ring with one
mbarrier per slot, two epilogue groups that alternate tiles, one or two residual buffers, gradient
accumulators flushed per chunk.

The model checks, over random interleavings of the roles and random completion delays of the asynchronous
engines (TMA, tensor core), what the barrier protocol has to guarantee:

* no dead-lock for any chunk list (including a CTA that gets no chunk at all);
* every role sees the same chunks in the same order, and never reads a ring slot that has been overwritten;
* a smem stage, an eta buffer or an R buffer is never overwritten before its consumer is done — in
  particular with ONE R buffer shared by both epilogue groups (the K = 16 configuration);
* every tile is accumulated exactly once into the gradient accumulator of ITS chunk, every chunk is read out
  exactly once, by the group that owns the chunk's last tile.

It also reproduces the bug that the first single-R-buffer version had on hardware (mbarrier parity waits alias
once a waiter can run two phases ahead): with one `r_empty` barrier for both groups the model finds a schedule
that corrupts R; with one barrier per tile parity (what the kernel does) it does not.
"""
import random

import pytest

from test_pipeline_model import MBarrier

K_RING = 16   # csrc/glm_tc.cu: kRing


class ChunkPipeline:
    def __init__(self, chunks, S, RB, rng, threads_per_group=3, one_r_empty_barrier=False):
        self.chunks = list(chunks)           # this CTA's chunks: numbers of tiles (all even)
        self.S, self.RB, self.rng, self.T = S, RB, rng, threads_per_group
        self.broken = one_r_empty_barrier
        self.full = [MBarrier(1) for _ in range(S)]
        self.empty = [MBarrier(1) for _ in range(S)]
        self.eta_full = [MBarrier(1) for _ in range(2)]
        self.eta_empty = [MBarrier(self.T) for _ in range(2)]
        self.r_full = [MBarrier(self.T) for _ in range(2)]
        self.r_empty = [MBarrier(1) for _ in range(2)]
        self.g_full = [MBarrier(1) for _ in range(2)]
        self.g_empty = [MBarrier(self.T) for _ in range(2)]
        self.ring_bar = [MBarrier(1) for _ in range(K_RING)]
        self.ring = [None] * K_RING          # (ordinal, n_tiles) or (ordinal, -1) = no more work
        self.stage = [None] * S
        self.eta = [None, None]
        self.r = [None] * RB
        self.r_writers = [0] * RB            # threads that have written the current content
        self.g = [[], []]
        self.flushed = []                    # (chunk ordinal, tiles)
        self.flush_arrivals = {}
        self.pending_async = []
        self.step = 0

    def later(self, fn, kind):
        last = max([t for t, _, k in self.pending_async if k == kind], default=self.step)
        self.pending_async.append((max(last, self.step + self.rng.randint(1, 6)), fn, kind))

    # -- consumers' view of the ring --------------------------------------------------------------------------
    def next_chunk(self, j):
        slot, parity = j % K_RING, (j // K_RING) & 1
        yield lambda: self.ring_bar[slot].passed(parity)
        entry = self.ring[slot]
        assert entry is not None and entry[0] == j, "a ring slot was overwritten before every role had read it"
        return entry[1]

    def producer(self):
        it = 0
        for j, n in enumerate(self.chunks + [-1]):
            yield lambda: True                      # claim latency: anything may happen in between
            self.ring[j % K_RING] = (j, n)
            self.ring_bar[j % K_RING].arrive()
            if n < 0:
                return
            for _ in range(n):
                st, ph = it % self.S, (it // self.S) & 1
                yield lambda: self.empty[st].passed(ph ^ 1)

                def land(it=it, st=st):
                    assert self.stage[st] is None, "TMA overwrote a stage MMA #2 had not released"
                    self.stage[st] = it
                    self.full[st].arrive()

                self.later(land, "tma")
                it += 1

    def mma1(self):
        it, j = 0, 0
        while True:
            n = yield from self.next_chunk(j)
            if n < 0:
                return
            for _ in range(n):
                st, ph = it % self.S, (it // self.S) & 1
                b, bph = it & 1, (it >> 1) & 1
                yield lambda: self.eta_empty[b].passed(bph ^ 1)
                yield lambda: self.full[st].passed(ph)
                assert self.stage[st] == it

                def done(it=it, b=b):
                    self.eta[b] = it
                    self.eta_full[b].arrive()

                self.later(done, "mma")
                it += 1
            j += 1

    def mma2(self):
        it, j = 0, 0
        while True:
            n = yield from self.next_chunk(j)
            if n < 0:
                return
            gb, gph = j & 1, (j >> 1) & 1
            yield lambda: self.g_empty[gb].passed(gph ^ 1)
            for t in range(n):
                st = it % self.S
                rb = (it & 1) if self.RB == 2 else 0
                rph = ((it >> 1) & 1) if self.RB == 2 else (it & 1)
                yield lambda: self.r_full[rb].passed(rph)
                assert self.r[rb] == it and self.r_writers[rb] == self.T, "MMA #2 read an R buffer that is not its tile's"
                assert self.stage[st] == it

                def done(it=it, st=st, rb=rb, gb=gb, j=j, t=t, n=n):
                    if t == 0:
                        assert self.g[gb] == [], "gradient accumulator reused before it was read out"
                    self.g[gb].append(it)
                    self.stage[st] = None
                    self.empty[st].arrive()
                    # R may be rewritten: one barrier per buffer, or (one buffer) one per tile parity
                    if self.RB == 2:
                        self.r_empty[rb].arrive()
                    else:
                        self.r_empty[0 if self.broken else (t & 1)].arrive()
                    self.r[rb], self.r_writers[rb] = None, 0
                    if t == n - 1:
                        self.g_full[gb].arrive()

                self.later(done, "mma")
                it += 1
            j += 1

    def epilogue_thread(self, tp):
        it, j, own = 0, 0, 0                 # own: tiles this group has processed (bph = own & 1)
        while True:
            n = yield from self.next_chunk(j)
            if n < 0:
                return
            for t in range(tp, n, 2):
                tile = it + t
                b, bph = tp, own & 1
                yield lambda: self.eta_full[b].passed(bph)
                assert self.eta[b] == tile, "epilogue group read eta of a tile it does not own"
                self.eta_empty[b].arrive()
                yield lambda: True                  # the link / likelihood maths
                if self.RB == 2:
                    yield lambda: self.r_empty[b].passed(bph ^ 1)
                    rb = b
                elif self.broken:                   # first hardware version: one barrier, phase = tile parity
                    yield lambda: self.r_empty[0].passed((tile & 1) ^ 1)
                    rb = 0
                else:                               # the other group's tile must have been consumed
                    yield lambda: self.r_empty[tp ^ 1].passed((bph ^ 1) if tp == 0 else bph)
                    rb = 0
                assert self.r[rb] in (None, tile), "R buffer overwritten while MMA #2 still needs it"
                self.r[rb] = tile
                self.r_writers[rb] += 1
                self.r_full[rb].arrive()
                own += 1
            if tp == 1:                             # owns the chunk's last tile: reads the accumulator out
                gb, gph = j & 1, (j >> 1) & 1
                yield lambda: self.g_full[gb].passed(gph)
                tiles = sorted(self.g[gb])
                yield lambda: True
                k = self.flush_arrivals.get(j, 0) + 1
                self.flush_arrivals[j] = k
                if k == self.T:
                    self.flushed.append((j, tiles))
                    self.g[gb] = []
                self.g_empty[gb].arrive()
            it += n
            j += 1

    def run(self):
        actors = [self.producer(), self.mma1(), self.mma2()]
        actors += [self.epilogue_thread(tp) for tp in (0, 1) for _ in range(self.T)]
        waiting, alive = [None] * len(actors), [True] * len(actors)
        for i, a in enumerate(actors):
            try:
                waiting[i] = next(a)
            except StopIteration:
                alive[i] = False
        limit = 600 * (sum(self.chunks) + 8) * len(actors)
        while any(alive) or self.pending_async:
            self.step += 1
            assert self.step < limit, "live-lock"
            for p in sorted([p for p in self.pending_async if p[0] <= self.step], key=lambda p: p[0]):
                self.pending_async.remove(p)
                p[1]()
            runnable = [i for i in range(len(actors)) if alive[i] and waiting[i]()]
            if not runnable:
                assert self.pending_async, f"dead-lock at step {self.step}"
                self.step = min(p[0] for p in self.pending_async) - 1
                continue
            i = self.rng.choice(runnable)
            try:
                waiting[i] = next(actors[i])
            except StopIteration:
                alive[i] = False
        return self


def _random_chunks(rng):
    return [rng.choice([2, 2, 4, 4, 6, 8, 32]) for _ in range(rng.choice([0, 1, 2, 3, 5, 9, 20, 40]))]


@pytest.mark.parametrize("RB", [2, 1])
def test_chunked_pipeline_protocol_holds_for_every_schedule(RB):
    rng = random.Random(99 + RB)
    for _ in range(120):
        chunks = _random_chunks(rng)
        S = rng.randint(2, 4)
        p = ChunkPipeline(chunks, S, RB, random.Random(rng.random())).run()
        assert [j for j, _ in sorted(p.flushed)] == list(range(len(chunks)))
        first = 0
        for (j, tiles), n in zip(sorted(p.flushed), chunks):
            assert tiles == list(range(first, first + n)), "a chunk's accumulator must hold exactly its own tiles"
            first += n


def test_model_reproduces_the_single_barrier_parity_aliasing_bug():
    """One `r_empty` barrier shared by both groups (first K = 16 version, dead-locked on hardware): a group can run
    two phases ahead of it, the parity wait aliases, R is overwritten early or arrivals mix — the model sees it."""
    rng = random.Random(5)
    with pytest.raises(AssertionError):
        for _ in range(400):
            ChunkPipeline([8, 4, 6, 32, 2, 4], rng.randint(3, 4), 1, random.Random(rng.random()), one_r_empty_barrier=True).run()
