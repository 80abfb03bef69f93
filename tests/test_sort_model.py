"""CPU model of the one-sweep radix sort's tile protocol (j3d_b200/csrc/sort.cuh, onesweep_kernel).

A pass is one kernel: every tile publishes its digit counts in a status word (2 flag bits + 30 value bits: AGGREGATE = "my
own count", INCLUSIVE = "count of everything up to and including me") and finds the number of equal digits before it by
walking back over its predecessors, four per round trip, until it meets an INCLUSIVE word.  The GPU tests pin the sorted
result; this file replays the protocol in plain Python under RANDOM interleavings of the tiles' steps (tiles only ever
start in ticket order, as on the device) and checks what the kernel relies on:

  * every tile ends with exactly the exclusive prefix of the counts before it, whatever the interleaving,
  * a walk never reads past tile 0 and always terminates once its predecessors have published,
  * the keys-only packing of the builder (code bits | position) sorts like (code, position) pairs, stably,
  * the pass / bit bookkeeping of j3dg_build_bvh (sorted bits, first bit, packed or pairs).

Nothing here is imported by the product.
"""
import numpy as np
import pytest

AGG, INC, VAL = 1 << 30, 2 << 30, (1 << 30) - 1


class Tile:
    """One block of onesweep_kernel for ONE digit column, as a little state machine (each call of step() is one round trip)."""

    def __init__(self, tile, count, status):
        self.tile, self.count, self.status = tile, count, status
        self.state = "ranked"       # the keys are ranked, the count is known
        self.before = 0
        self.t = tile               # tiles [0, t) are still to be accounted for
        self.reads = 0

    def step(self):
        if self.state == "ranked":
            self.status[self.tile] = (INC if self.tile == 0 else AGG) | self.count
            self.state = "walking" if self.tile else "done"
            return
        assert self.state == "walking"
        t = self.t
        assert t >= 1
        vs = [self.status[t - 1 - u] if t - 1 - u >= 0 else INC for u in range(4)]   # four loads issued together
        self.reads += 1
        for v in vs:
            if (v >> 30) == 0:
                break               # not published yet: poll again from here
            self.before += v & VAL
            if v & INC:
                self.status[self.tile] = INC | (self.before + self.count)
                self.state = "done"
                return
            self.t -= 1


@pytest.mark.parametrize("ntiles,resident,seed", [(1, 1, 0), (7, 3, 1), (64, 8, 2), (300, 24, 3), (300, 300, 4), (97, 2, 5)])
def test_lookback_gives_the_exclusive_prefix_under_any_interleaving(ntiles, resident, seed):
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, 4097, ntiles)
    status = [0] * ntiles
    next_ticket = 0
    running, done = [], {}
    steps = 0
    while len(done) < ntiles:
        while len(running) < resident and next_ticket < ntiles:   # a block that starts draws the next ticket
            running.append(Tile(next_ticket, int(counts[next_ticket]), status))
            next_ticket += 1
        tl = running[rng.integers(len(running))]                  # any resident block may make the next step
        tl.step()
        steps += 1
        if tl.state == "done":
            done[tl.tile] = tl
            running.remove(tl)
        assert steps < 200 * ntiles + 1000, "the walk does not terminate"
    excl = np.concatenate([[0], np.cumsum(counts)[:-1]])
    for t in range(ntiles):
        assert done[t].before == excl[t]
        assert status[t] == INC | int(excl[t] + counts[t])


def lsd_sort(keys, first_shift, passes, vals=None):
    """What a sequence of stable 8-bit passes does (the kernel's ranking is stable inside a tile, tiles are ordered)."""
    order = np.arange(len(keys))
    for p in range(passes):
        d = (keys[order] >> np.uint64(first_shift + 8 * p)) & np.uint64(0xff)
        order = order[np.argsort(d, kind="stable")]
    return order


@pytest.mark.parametrize("n,axis_bits", [(1000, 13), (5000, 11), (70000, 12), (300, 16)])
def test_packed_keys_sort_like_pairs(n, axis_bits):
    """j3dg_build_bvh: passes / first bit / packed-or-pairs decision, and packed == pairs order."""
    MORTON_BITS = 48
    rng = np.random.default_rng(n)
    codes = rng.integers(0, 1 << 48, n, dtype=np.uint64)
    codes[rng.integers(0, n, n // 3)] = codes[rng.integers(0, n, n // 3)]      # duplicates: ties keep their input order
    passes = (3 * axis_bits + 7) // 8
    first_bit = MORTON_BITS - 8 * passes
    idx_bits = 1
    while idx_bits < 32 and (1 << idx_bits) < n:
        idx_bits += 1
    sorted_bits = 8 * passes
    if sorted_bits + idx_bits > 64 and 3 * axis_bits + idx_bits <= 64:
        sorted_bits = 64 - idx_bits
    packed_ok = sorted_bits <= 8 * passes and sorted_bits + idx_bits <= 64
    assert sorted_bits >= min(3 * axis_bits, 8 * passes) or not packed_ok
    pairs_order = lsd_sort(codes, first_bit, passes)
    # pairs order == stable order by the code bits [first_bit, 48)
    want = np.argsort(codes >> np.uint64(first_bit), kind="stable")
    assert (pairs_order == want).all()
    if packed_ok:
        fb = MORTON_BITS - sorted_bits
        packed = ((codes >> np.uint64(fb)) << np.uint64(idx_bits)) | np.arange(n, dtype=np.uint64)
        order = lsd_sort(packed, idx_bits, passes)
        got_tri = (packed[order] & np.uint64((1 << idx_bits) - 1)).astype(np.int64)
        assert (got_tri == np.argsort(codes >> np.uint64(fb), kind="stable")).all()
        # the radix tree masks the index bits: equal codes compare equal, told apart by position
        masked = packed[order] & ~np.uint64((1 << idx_bits) - 1)
        assert (np.diff(masked.astype(np.float64)) >= 0).all()
