"""Protocol model of Mark's dynamic unit hand-out (csrc/stvl_mark.cu, `next_unit` / `claim_group`): every CTA starts with
one static group, claims further groups of 32 units from a never-reset global cursor one group ahead of need, and stops at
its first out-of-range claim.  The model runs the protocol's atomic steps under random interleavings of the warps of all
CTAs (CTAs start at random times, as they do behind the sweep) and checks what the kernel and the host rely on:

  * every unit of the batch is handed out exactly once, whatever the interleaving (no unit lost, none processed twice);
  * no warp waits forever (the claim a refilling warp waits for is always under way);
  * a launch moves the cursor by exactly n_groups (so the host can tell the next launch where the cursor stands), also
    over several launches back to back with different sizes.

Pure Python, no GPU: this pins the protocol, the GPU parity tests pin its CUDA transcription."""
import random

import pytest

GROUP = 32


class Cta:
    def __init__(self, c, n_warps, late_warps, n_units, n_groups, grid, cursor, base):
        self.c, self.n_units, self.n_groups, self.grid, self.cursor, self.base = c, n_units, n_groups, grid, cursor, base
        hi = min((c + 1) * GROUP, n_units)
        lo = min(c * GROUP + (n_warps - late_warps), hi)
        self.cur, self.end = lo, hi            # s_units
        self.claimed = None                    # s_claimed (None = 0: not there yet)
        self.exhausted = False
        self.static = [c * GROUP + (w - late_warps) if w >= late_warps and c * GROUP + (w - late_warps) < n_units else None
                       for w in range(n_warps)]

    def claim(self):                           # claim_group(): one global atomic
        v = self.cursor[0]
        self.cursor[0] += 1
        return self.grid + (v - self.base)


def warp(cta, w, out, late):
    """One warp as a generator: every `yield` is a point where other warps may run."""
    if w == 0:                                 # thread 0's first claim, installed before the set-up barrier
        cta.claimed = cta.claim()
    yield "barrier"
    if cta.static[w] is not None:
        out.append(cta.static[w])
    for _ in range(late):                      # costmap warps join late
        yield
    while True:
        # next_unit(), lane 0
        u = None
        while True:
            cur, end = cta.cur, cta.end        # atomicAdd(&s_units, 1)
            cta.cur += 1
            if cur < end:
                if cur % GROUP == 0:
                    yield                      # the global atomic takes a while
                    cta.claimed = cta.claim()
                u = cur
                break
            if cta.exhausted:
                break
            if cur == end:
                spins = 0
                while cta.claimed is None:     # claimed one group ahead: must always arrive
                    spins += 1
                    assert spins < 100000, "refilling warp waits forever"
                    yield
                grp, cta.claimed = cta.claimed, None
                if grp >= cta.n_groups:
                    cta.exhausted = True
                    break
                lo = grp * GROUP
                cta.cur, cta.end = lo, min(lo + GROUP, cta.n_units)   # atomicExch(&s_units, ...)
            else:
                yield                          # another warp is installing the next group
        if u is None:
            return
        out.append(u)
        yield                                  # processing the unit


def run_launch(rng, n_units, sm_count, n_warps, late_warps, cursor):
    n_groups = (n_units + GROUP - 1) // GROUP
    grid = max(1, min(sm_count, n_groups))
    base = cursor[0]
    out = []
    ctas = [Cta(c, n_warps, late_warps, n_units, n_groups, grid, cursor, base) for c in range(grid)]
    start = [rng.randint(0, 200) for _ in ctas]                        # CTAs become resident one by one
    gens, waiting = [], {}
    for ci, cta in enumerate(ctas):
        for w in range(n_warps):
            gens.append([start[ci], ci, warp(cta, w, out, late=rng.randint(0, 30) if w < late_warps else 0), False])
    barrier_count = [0] * grid
    t = 0
    live = list(gens)
    while live:
        t += 1
        ready = [g for g in live if g[0] <= t and not (g[3] and barrier_count[g[1]] < n_warps)]
        if not ready:
            continue
        g = rng.choice(ready)
        try:
            r = next(g[2])
            if r == "barrier":
                g[3] = True
                barrier_count[g[1]] += 1
        except StopIteration:
            live.remove(g)
        assert t < 5_000_000, "the launch does not finish"
    return out, n_groups, grid, base


@pytest.mark.parametrize("n_units,sm_count,n_warps,late", [
    (1, 148, 32, 8), (31, 148, 32, 8), (32, 148, 32, 0), (33, 4, 32, 8), (150, 148, 32, 8), (1000, 8, 32, 8),
    (1000, 8, 16, 8), (525 * 32, 148, 32, 8), (525 * 32 - 7, 37, 32, 0), (4097, 3, 16, 0)])
def test_every_unit_once_and_cursor_advance(n_units, sm_count, n_warps, late):
    rng = random.Random(n_units * 7919 + sm_count)
    cursor = [rng.randint(0, 1 << 20)]
    for _ in range(2 if n_units > 2000 else 4):                        # several launches back to back on one cursor
        before = cursor[0]
        out, n_groups, grid, base = run_launch(rng, n_units, sm_count, n_warps, late, cursor)
        assert sorted(out) == list(range(n_units))
        assert cursor[0] - before == max(n_groups - grid, 0) + grid == n_groups
