"""Protocol model of Mark's queue flush (csrc/stvl_mark.cu, `flush`): the warps of a CTA stream their units independently,
append new voxels to one CTA-wide queue, and meet at a named barrier either when a warp has seen the queue above the flush
threshold (it raises a flag every warp polls between two units) or when they have run out of units; the flush that finds every
warp out of units is the last one.  The model runs the protocol under random interleavings and checks:

  * no deadlock: every warp reaches the final flush, however the units and the flush requests fall;
  * every warp executes the same sequence of barriers (a named barrier counts arrivals, not call sites);
  * the queue never overflows as long as a warp's unit adds at most 128 voxels (flush threshold + warps * 128 <= capacity),
    and every voxel that was appended is written exactly once.

Pure Python, no GPU."""
import random

import pytest

QUEUE, FLUSH_AT, UNIT = 8192, 4096, 128


class Barrier:
    def __init__(self, n):
        self.n, self.count, self.gen = n, 0, 0


def wait(bar):
    """Arrive at the named barrier and yield until everyone has."""
    g = bar.gen
    bar.count += 1
    if bar.count == bar.n:
        bar.count = 0
        bar.gen += 1
        return
    while bar.gen == g:
        yield


class Cta:
    def __init__(self, n_warps):
        self.n_warps = n_warps
        self.bar = Barrier(n_warps)
        self.queue = []          # s_queue[0 : min(s_qn, QUEUE)]
        self.qn = 0              # s_qn
        self.req = False         # s_flush_req
        self.done = 0            # s_done
        self.written = []        # voxels that reached the store
        self.direct = 0
        self.trace = [[] for _ in range(n_warps)]


def flush(cta, w):
    yield from wait(cta.bar)
    cta.trace[w].append("a")
    final = cta.done == cta.n_warps
    if w == 0:                                   # (the drain is shared by all threads; one model warp does it)
        n = min(cta.qn, QUEUE)
        assert len(cta.queue) == n
        cta.written += cta.queue
    if final:
        return True
    yield from wait(cta.bar)
    cta.trace[w].append("b")
    if w == 0:
        cta.queue, cta.qn, cta.req = [], 0, False
    yield from wait(cta.bar)
    cta.trace[w].append("c")
    return False


def warp(cta, w, units, rng, next_id):
    for pushes in units:
        for _ in range(rng.randint(0, 3)):
            yield                                # processing
        if pushes:
            q0 = cta.qn                          # atomicAdd(&s_qn, n_push)
            cta.qn += pushes
            if q0 < FLUSH_AT <= q0 + pushes:
                cta.req = True
            yield
            for k in range(pushes):
                v = next_id[0]
                next_id[0] += 1
                if q0 + k < QUEUE:
                    cta.queue.append(v)
                else:
                    cta.direct += 1
                    cta.written.append(v)
        if cta.req:
            r = yield from flush(cta, w)
            assert r is False
    cta.done += 1
    while True:
        r = yield from flush(cta, w)
        if r:
            return


@pytest.mark.parametrize("n_warps,n_units,push_rate", [(32, 4, 0.2), (32, 40, 1.0), (16, 60, 1.0), (32, 25, 0.6), (32, 1, 1.0), (8, 300, 0.9)])
def test_flush_protocol(n_warps, n_units, push_rate):
    rng = random.Random(n_warps * 1000 + n_units)
    cta = Cta(n_warps)
    next_id = [0]
    gens = []
    for w in range(n_warps):
        k = rng.randint(max(0, n_units - 3), n_units + 3)          # warps run out of units at different times
        units = [int(rng.random() < push_rate) * rng.randint(1, UNIT) for _ in range(k)]
        gens.append(warp(cta, w, units, rng, next_id))
    live = list(range(n_warps))
    steps = 0
    while live:
        steps += 1
        assert steps < 20_000_000, "deadlock: the warps never all reach the final flush"
        i = rng.choice(live)
        try:
            next(gens[i])
        except StopIteration:
            live.remove(i)
    assert all(t == cta.trace[0] for t in cta.trace), "the warps executed different barrier sequences"
    if n_units * push_rate * UNIT / 2 * n_warps > 2 * FLUSH_AT:
        assert cta.trace[0].count("b") >= 1, "the case was meant to flush early at least once"
    assert cta.direct == 0, "the queue overflowed although the flush threshold leaves room for every warp's running unit"
    assert sorted(cta.written) == list(range(next_id[0])), "a queued voxel was lost or written twice"
