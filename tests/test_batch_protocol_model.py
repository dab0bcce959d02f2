"""Executable model of the batch kernel's hand-off protocol (x-maps_b200/csrc/xm_batch_kernel.cuh).

The CUDA kernel lets event warps and epilogue groups of many CTAs run freely and synchronises them only
through per-frame counters: chunks are handed out in frame order by one global counter; a CTA's consumers
publish how many chunks of frame f they scattered when they leave the frame; an epilogue group processes
tiles of frame f only after the published counts add up to the frame's chunk count; a consumer may start
scattering frame f only after every tile of frame f - 3 has been read (three scatter maps in rotation).  The
consumers' software pipeline runs the FRONT half of the next chunk before the BACK half of the current one,
also across frame boundaries, with a drain rule that keeps this from deadlocking.

This test re-states exactly that control flow as Python coroutines, runs them under random interleavings for
many batch shapes (tiny frames, empty frames, more CTAs than chunks) and checks the three properties the
kernel relies on:

  1. no tile of frame f is processed before every chunk of frame f has been scattered,
  2. no chunk of frame f is scattered before every tile of frame f - 3 has been processed,
  3. every schedule terminates (no deadlock), with every chunk scattered and every tile processed once.

It is a model of the design, not of the CUDA code's arithmetic; the GPU tests (tests/test_gpu_batch.py) check
the real kernel's results.
"""
import random

import pytest

MAPS = 3  # kBatchMaps


class World:
    def __init__(self, chunks_per_frame, tiles_per_frame, hard_frames):
        self.C = list(chunks_per_frame)
        self.B = len(self.C)
        self.P = tiles_per_frame
        self.hard_frames = hard_frames
        self.first = [0]
        for c in self.C:
            self.first.append(self.first[-1] + c)
        self.total = self.first[-1]
        self.next_chunk = 0               # the global chunk counter
        self.blocks_done = [0] * self.B   # published chunk counts
        self.tiles_done = [0] * self.B
        self.ticket = [0] * self.B
        self.scattered = [0] * self.B     # ground truth: chunks whose BACK half ran
        self.tiles_processed = [0] * self.B
        self.violations = []

    def decode(self, it):
        f = 0
        while it >= self.first[f + 1]:
            f += 1
        return f, it - self.first[f]


def consumer(w: World, stages: int):
    """One CTA's consumer warps (they all see the same items in the same order, so one coroutine stands for
    the eight).  Yields ("wait", predicate) when it would spin, ("step",) between actions."""
    ring = []  # items the producer has already taken for this CTA (it runs `stages` ahead)

    def refill():
        while len(ring) < stages and not (ring and ring[-1] is None):
            it = w.next_chunk
            w.next_chunk += 1
            ring.append(None if it >= w.total else w.decode(it))  # None: past the end (every producer sees it)

    cur_f, front_f, my_chunks = -1, -1, 0

    def leave():
        nonlocal cur_f, my_chunks
        if cur_f >= 0:
            w.blocks_done[cur_f] += my_chunks
        my_chunks, cur_f = 0, -1

    def front(item):
        nonlocal front_f
        f = item[0]
        if f != front_f:
            front_f = f
            if f >= MAPS:
                yield ("wait", lambda f=f: w.tiles_done[f - MAPS] >= w.P)
        yield ("step",)

    def back(item):
        nonlocal cur_f, my_chunks
        f = item[0]
        if f != cur_f:
            leave()
            cur_f = f
        if f >= MAPS and w.tiles_done[f - MAPS] < w.P:
            w.violations.append(("scatter before the map was free", f))
        w.scattered[f] += 1
        my_chunks += 1
        yield ("step",)

    refill()
    item = ring.pop(0)
    refill()
    if item is not None:
        yield from front(item)
        cur = item
        while True:
            nxt = ring.pop(0) if ring else None
            refill()
            more = nxt is not None
            drain = more and w.hard_frames and nxt[0] != cur[0]
            if not drain and more and nxt[0] != front_f and nxt[0] >= MAPS:
                must_be_out = nxt[0] - MAPS
                drain = cur[0] <= must_be_out or (cur_f >= 0 and cur_f <= must_be_out)
            if more and not drain:
                yield from front(nxt)
            yield from back(cur)
            if not more:
                break
            if drain:
                leave()
                yield from front(nxt)
            cur = nxt
    leave()


def tile_group(w: World):
    for f in range(w.B):
        yield ("wait", lambda f=f: w.blocks_done[f] >= w.C[f])
        while True:
            t = w.ticket[f]
            w.ticket[f] += 1
            if t >= w.P:
                break
            if w.scattered[f] < w.C[f]:
                w.violations.append(("tile before the frame was complete", f))
            yield ("step",)
            w.tiles_processed[f] += 1
            w.tiles_done[f] += 1


def run(w: World, n_ctas: int, groups_per_cta: int, stages: int, rng: random.Random, max_steps=2_000_000):
    actors = [consumer(w, stages) for _ in range(n_ctas)] + [tile_group(w) for _ in range(n_ctas * groups_per_cta)]
    waiting = {}  # actor index -> predicate
    alive = set(range(len(actors)))
    steps = 0
    while alive:
        runnable = [i for i in alive if i not in waiting or waiting[i]()]
        if not runnable:
            return "deadlock"
        i = rng.choice(runnable)
        waiting.pop(i, None)
        try:
            msg = next(actors[i])
            if msg[0] == "wait" and not msg[1]():
                waiting[i] = msg[1]
        except StopIteration:
            alive.discard(i)
        steps += 1
        if steps > max_steps:
            return "runaway"
    return "done"


SHAPES = [
    # (chunks per frame, tiles per frame, CTAs)
    ([20, 0, 1, 1, 1, 2, 69, 1, 0, 33, 5, 1], 6, 8),      # the mixed-size GPU test: tiny and empty frames
    ([3] * 12, 4, 16),                                    # far more CTAs than chunks per frame
    ([1] * 20, 2, 5),
    ([40, 40, 40, 40, 40, 40], 10, 4),                    # many chunks per CTA and frame (the bench's regime)
    ([0, 0, 0, 0, 7], 3, 3),
    ([5], 3, 2),
    ([9, 0, 0, 0, 0, 0, 9, 1, 0, 0, 2], 5, 6),
]


@pytest.mark.parametrize("shape", range(len(SHAPES)))
@pytest.mark.parametrize("hard_frames", [0, 1])
@pytest.mark.parametrize("stages", [1, 2, 4])
def test_protocol_is_safe_and_live(shape, hard_frames, stages):
    chunks, tiles, ctas = SHAPES[shape]
    for seed in range(40):
        rng = random.Random(seed * 7919 + shape)
        w = World(chunks, tiles, hard_frames)
        assert run(w, ctas, 2, stages, rng) == "done", (shape, hard_frames, stages, seed)
        assert not w.violations, w.violations[:3]
        assert w.scattered == list(chunks) and w.blocks_done == list(chunks)
        assert w.tiles_processed == [tiles] * len(chunks)


def test_the_model_notices_a_missing_drain_rule():
    """Sanity of the model itself: without the drain rule the pipelined front half deadlocks on sparse frames
    (the situation the rule exists for), so the checker is able to see deadlocks."""
    global MAPS
    chunks, tiles, ctas = [1] * 20, 2, 5

    def consumer_without_drain(w, stages):
        ring = []

        def refill():
            while len(ring) < stages and not (ring and ring[-1] is None):
                it = w.next_chunk
                w.next_chunk += 1
                ring.append(None if it >= w.total else w.decode(it))

        cur_f, front_f, my = -1, -1, 0
        refill()
        cur = ring.pop(0)
        refill()
        if cur is None:
            return
        if cur[0] >= MAPS:
            yield ("wait", lambda f=cur[0]: w.tiles_done[f - MAPS] >= w.P)
        front_f = cur[0]
        while True:
            nxt = ring.pop(0) if ring else None
            refill()
            if nxt is not None and nxt[0] != front_f:
                front_f = nxt[0]
                if nxt[0] >= MAPS:
                    yield ("wait", lambda f=nxt[0]: w.tiles_done[f - MAPS] >= w.P)
            if cur[0] != cur_f:
                if cur_f >= 0:
                    w.blocks_done[cur_f] += my
                my, cur_f = 0, cur[0]
            w.scattered[cur[0]] += 1
            my += 1
            yield ("step",)
            if nxt is None:
                break
            cur = nxt
        w.blocks_done[cur_f] += my

    outcomes = set()
    for seed in range(30):
        rng = random.Random(seed)
        w = World(chunks, tiles, 0)
        actors = [consumer_without_drain(w, 2) for _ in range(ctas)] + [tile_group(w) for _ in range(ctas * 2)]
        waiting, alive = {}, set(range(len(actors)))
        result = "done"
        while alive:
            runnable = [i for i in alive if i not in waiting or waiting[i]()]
            if not runnable:
                result = "deadlock"
                break
            i = rng.choice(runnable)
            waiting.pop(i, None)
            try:
                msg = next(actors[i])
                if msg[0] == "wait" and not msg[1]():
                    waiting[i] = msg[1]
            except StopIteration:
                alive.discard(i)
        outcomes.add(result)
    assert "deadlock" in outcomes
