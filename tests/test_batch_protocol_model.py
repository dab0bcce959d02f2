"""Executable model of the batch kernel's hand-off protocol (x-maps_b200/csrc/xm_batch_kernel.cuh).

The CUDA kernel lets event warps and epilogue groups of many CTAs run freely and synchronises them only
through per-frame counters: chunks are handed out in frame order by one global counter; a CTA's consumers
publish how many chunks of frame f they scattered when they leave the frame; an epilogue group processes
tiles of frame f only after the published counts add up to the frame's chunk count; a consumer may start
scattering frame f only after every tile of frame f - 3 has been read (three scatter maps in rotation; a launch may
use up to eight).  The
consumers' software pipeline runs the FRONT half of the next chunk before the BACK half of the current one,
also across frame boundaries, with a drain rule that keeps this from deadlocking.

This test re-states exactly that control flow as Python coroutines, runs them under random interleavings for
many batch shapes (tiny frames, empty frames, more CTAs than chunks) and checks the three properties the
kernel relies on:

  1. no tile of frame f is processed before every chunk of frame f has been scattered,
  2. no chunk of frame f is scattered before every tile of frame f - 3 has been processed,
  3. every schedule terminates (no deadlock), with every chunk scattered and every tile processed once.

The strip epilogue (the projector view's default: two passes per frame, one warp per item, every item of the batch in
ONE ordered list handed out by a global counter, pass 2 of a frame `lag` blocks behind its pass 1) is modelled the same
way: pass 1 of frame f waits for the frame's chunks and for pass 2 of frame f - 4 (four dilated maps in rotation), pass
2 for the frame's finished pass-1 count; the finished pass-1 count is what frees the scatter map.

It is a model of the design, not of the CUDA code's arithmetic; the GPU tests (tests/test_gpu_batch.py) check
the real kernel's results.
"""
import random

import pytest

MAPS = 3  # kBatchMaps (default; a launch may rotate through up to 8)
DIL_MAPS = 4  # kBatchDilMaps


class World:
    def __init__(self, chunks_per_frame, tiles_per_frame, hard_frames, maps=MAPS, p2_items=0):
        self.maps = maps
        self.P2 = p2_items                # strip epilogue: pass-2 items per frame (pass-1 items = tiles_per_frame)
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
        self.strip_ticket = 0             # strip epilogue: the global item counter
        self.p2_done = [0] * self.B
        self.p2_processed = [0] * self.B
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
            if f >= w.maps:
                yield ("wait", lambda f=f: w.tiles_done[f - w.maps] >= w.P)
        yield ("step",)

    def back(item):
        nonlocal cur_f, my_chunks
        f = item[0]
        if f != cur_f:
            leave()
            cur_f = f
        if f >= w.maps and w.tiles_done[f - w.maps] < w.P:
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
            if not drain and more and nxt[0] != front_f and nxt[0] >= w.maps:
                must_be_out = nxt[0] - w.maps
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


STRIP_LAG = 1  # option "strip_lag": 1 (default) ... DIL_MAPS - 1


def strip_decode(w: World, t: int, lag=None):
    """Position t of the strip epilogue's item list -> (frame, pass, item).  Block b of the list = pass 1 of frame b
    (b < B) followed by pass 2 of frame b - lag (b >= lag) (batch_strip_warps)."""
    lag = STRIP_LAG if lag is None else lag
    per_frame = w.P + w.P2
    head = min(lag, w.B) * w.P
    mixed = max(w.B - lag, 0)
    if t < head:
        return t // w.P, 1, t % w.P
    if t - head < mixed * per_frame:
        b, r = (t - head) // per_frame, (t - head) % per_frame
        return (b + lag, 1, r) if r < w.P else (b, 2, r - w.P)
    k = t - head - mixed * per_frame
    return mixed + k // w.P2, 2, k % w.P2


def strip_warp(w: World, decode=strip_decode):
    """One epilogue warp of the strip epilogue: takes the next item of the global list when it is free (no look-ahead),
    waits for the item's dependencies, runs it, publishes it."""
    total = (w.P + w.P2) * w.B
    while True:
        t = w.strip_ticket
        w.strip_ticket += 1
        if t >= total:
            return
        f, ps, _ = decode(w, t)
        if ps == 1:
            yield ("wait", lambda f=f: w.blocks_done[f] >= w.C[f])
            if f >= DIL_MAPS:
                yield ("wait", lambda f=f: w.p2_done[f - DIL_MAPS] >= w.P2)
            if w.scattered[f] < w.C[f]:
                w.violations.append(("pass 1 before the frame was complete", f))
            if f >= DIL_MAPS and w.p2_processed[f - DIL_MAPS] < w.P2:
                w.violations.append(("dilated map overwritten while pass 2 of an earlier frame still reads it", f))
            yield ("step",)
            w.tiles_processed[f] += 1
            w.tiles_done[f] += 1  # `next_tile`: what the event warps of frame f + maps wait for
        else:
            yield ("wait", lambda f=f: w.tiles_done[f] >= w.P)
            if w.tiles_processed[f] < w.P:
                w.violations.append(("pass 2 before pass 1 was complete", f))
            yield ("step",)
            w.p2_processed[f] += 1
            w.p2_done[f] += 1


def run(w: World, n_ctas: int, groups_per_cta: int, stages: int, rng: random.Random, max_steps=2_000_000, epilogue=None):
    epilogue = epilogue or tile_group
    actors = [consumer(w, stages) for _ in range(n_ctas)] + [epilogue(w) for _ in range(n_ctas * groups_per_cta)]
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


STRIP_SHAPES = [
    # (chunks per frame, pass-1 items, pass-2 items, CTAs, epilogue warps per CTA)
    ([20, 0, 1, 1, 1, 2, 69, 1, 0, 33, 5, 1], 5, 7, 8, 4),
    ([3] * 12, 4, 6, 16, 2),
    ([1] * 20, 2, 3, 5, 4),
    ([40, 40, 40, 40, 40, 40], 9, 11, 4, 2),
    ([2] * 14, 7, 9, 1, 1),   # ONE epilogue warp runs the whole list in order
    ([0, 0, 0, 0, 7], 3, 1, 3, 4),
    ([5], 3, 4, 2, 2),
]


@pytest.mark.parametrize("shape", range(len(STRIP_SHAPES)))
@pytest.mark.parametrize("hard_frames", [0, 1])
@pytest.mark.parametrize("maps", [2, 3, 6])
@pytest.mark.parametrize("lag", [1, 2, 3])
def test_strip_epilogue_protocol_is_safe_and_live(shape, hard_frames, maps, lag):
    chunks, p1, p2, ctas, warps = STRIP_SHAPES[shape]
    for seed in range(12):
        rng = random.Random(seed * 104729 + shape)
        w = World(chunks, p1, hard_frames, maps=maps, p2_items=p2)
        assert run(w, ctas, warps, 2, rng, epilogue=lambda w: strip_warp(w, lambda w, t: strip_decode(w, t, lag))) == "done", (shape, hard_frames, maps, lag, seed)
        assert not w.violations, w.violations[:3]
        assert w.scattered == list(chunks) and w.blocks_done == list(chunks)
        assert w.tiles_processed == [p1] * len(chunks) and w.p2_processed == [p2] * len(chunks)


@pytest.mark.parametrize("lag", [1, 2, 3])
@pytest.mark.parametrize("frames", [1, 2, 3, 5, 9])
def test_strip_decode_enumerates_every_item_once(lag, frames):
    w = World([1] * frames, 3, 0, p2_items=4)
    seen = [strip_decode(w, t, lag) for t in range((3 + 4) * frames)]
    assert sorted(seen) == sorted([(f, 1, j) for f in range(frames) for j in range(3)] + [(f, 2, j) for f in range(frames) for j in range(4)])
    # an item only ever waits for items that come earlier in the list
    pos = {it: t for t, it in enumerate(seen)}
    for f in range(frames):
        assert max(pos[(f, 1, j)] for j in range(3)) < min(pos[(f, 2, j)] for j in range(4))
        if f >= DIL_MAPS:
            assert max(pos[(f - DIL_MAPS, 2, j)] for j in range(4)) < min(pos[(f, 1, j)] for j in range(3))


def test_the_model_notices_a_wrong_item_order():
    """Negative control for the strip epilogue: a list that hands out pass 2 of a frame BEFORE its pass 1 makes items
    wait for later ones; with few warps every warp ends up holding such an item."""

    def bad_decode(w, t):
        per_frame = w.P + w.P2
        f, r = t // per_frame, t % per_frame
        return (f, 2, r) if r < w.P2 else (f, 1, r - w.P2)

    outcomes = set()
    for seed in range(20):
        rng = random.Random(seed)
        w = World([2] * 6, 3, 0, p2_items=4)
        outcomes.add(run(w, 1, 2, 2, rng, epilogue=lambda w: strip_warp(w, bad_decode)))
    assert outcomes == {"deadlock"}


def test_the_model_notices_a_missing_drain_rule():
    """Sanity of the model itself: without the drain rule the pipelined front half deadlocks on sparse frames
    (the situation the rule exists for), so the checker is able to see deadlocks."""
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
        if cur[0] >= w.maps:
            yield ("wait", lambda f=cur[0]: w.tiles_done[f - w.maps] >= w.P)
        front_f = cur[0]
        while True:
            nxt = ring.pop(0) if ring else None
            refill()
            if nxt is not None and nxt[0] != front_f:
                front_f = nxt[0]
                if nxt[0] >= w.maps:
                    yield ("wait", lambda f=nxt[0]: w.tiles_done[f - w.maps] >= w.P)
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
