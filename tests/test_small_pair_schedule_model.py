"""CPU model of K6 v5's two-warps-per-problem schedule (csrc/gpr_small_v5.cu): for every phase, the shared-memory tiles each
warp of the pair reads and writes between two consecutive pair barriers.  The check is the data-race condition itself: inside
one barrier interval no tile may be written by one warp and touched (read or written) by the other.  The intervals below are
a transcription of the kernel's loops (same ownership rule u = 2 * ul + half, same barrier positions), so a change of the
kernel's schedule has to be made here too -- and is then checked for every tile count NT = 1 .. 8."""
import itertools

import pytest


def cslot(NT, i, j):
    assert i >= j
    return j * NT - j * (j - 1) // 2 + (i - j)


class Interval:
    def __init__(self):
        self.r = [set(), set()]
        self.w = [set(), set()]

    def check(self, where):
        for h in (0, 1):
            o = 1 - h
            clash = self.w[h] & (self.r[o] | self.w[o])
            assert not clash, f"{where}: tiles {sorted(clash)} written by warp {h} and touched by warp {o} in one barrier interval"


def schedule(NT):
    """Yields (name, Interval) for one problem, in program order; a pair barrier separates consecutive intervals."""
    NTRI = NT * (NT + 1) // 2
    own_pairs = lambda h: [s for s0 in range(2 * h, NTRI, 4) for s in (s0, s0 + 1) if s < NTRI]
    # 1: assembly -- every warp writes its own tile pairs
    iv = Interval()
    for h in (0, 1):
        iv.w[h] |= set(own_pairs(h))
    assert iv.w[0] | iv.w[1] == set(range(NTRI)) and not (iv.w[0] & iv.w[1])
    yield "assembly", iv
    # discrepancy term: warp 0 read-modify-writes HF x HF entries anywhere
    iv = Interval()
    iv.w[0] |= set(range(NTRI))
    yield "delta term", iv
    # 2: Cholesky
    for kb in range(NT):
        cnt = NT - kb
        iv = Interval()  # update + write back own raw tiles + K-major reload + (warp 0) diagonal factorisation
        for h in (0, 1):
            for k in range(kb):
                iv.r[h].add(cslot(NT, kb, k))
                for u in range(h, cnt, 2):
                    iv.r[h].add(cslot(NT, kb + u, k))
            for u in range(h, cnt, 2):
                iv.r[h].add(cslot(NT, kb + u, kb))
                if u > 0:
                    iv.w[h].add(cslot(NT, kb + u, kb))
        iv.w[0].add(cslot(NT, kb, kb))  # inv(L_kk)
        yield f"chol update kb={kb}", iv
        iv = Interval()  # panel
        for h in (0, 1):
            iv.r[h].add(cslot(NT, kb, kb))
            for u in range(h, cnt, 2):
                if u > 0:
                    iv.w[h].add(cslot(NT, kb + u, kb))
        yield f"chol panel kb={kb}", iv
    # 3: W = L^-1
    for i in range(1, NT):
        iv = Interval()
        for h in (0, 1):
            for k in range(i):
                iv.r[h].add(cslot(NT, i, k))
                for j in range(1 - h, k + 1, 2):  # columns are split the other way round than the Cholesky's tiles
                    iv.r[h].add(cslot(NT, k, j))
        yield f"trtri accumulate i={i}", iv
        iv = Interval()
        for h in (0, 1):
            iv.r[h].add(cslot(NT, i, i))
            for j in range(1 - h, i, 2):
                iv.r[h].add(cslot(NT, i, j))
                iv.w[h].add(cslot(NT, i, j))
        yield f"trtri finish i={i}", iv
    # 4: mat-vecs only read tiles
    # 5: K^-1 in place
    for i in range(NT):
        iv = Interval()  # reads of iteration i (+ the G writes of iteration i - 1, which share the interval)
        for h in (0, 1):
            for k in range(i, NT):
                iv.r[h].add(cslot(NT, k, i))
                for j in range(1 - h, i + 1, 2):
                    iv.r[h].add(cslot(NT, k, j))
            if i > 0:
                for j in range(1 - h, i, 2):
                    iv.w[h].add(cslot(NT, i - 1, j))
        yield f"lauum reads i={i}", iv
    iv = Interval()
    for h in (0, 1):
        for j in range(1 - h, NT, 2):
            iv.w[h].add(cslot(NT, NT - 1, j))
    yield "lauum last writes", iv
    # 6: contraction -- warp 0 reads G at HF x HF pairs (anywhere); own-tile passes read (DS > 0) or overwrite (generic d)
    iv = Interval()
    iv.r[0] |= set(range(NTRI))
    for h in (0, 1):
        iv.r[h] |= set(own_pairs(h))
    yield "contraction (d = 5: G is read-only)", iv
    iv = Interval()  # generic d: an extra barrier separates the HF x HF reads from the in-place T^L = G o K^L
    for h in (0, 1):
        iv.r[h] |= set(own_pairs(h))
        iv.w[h] |= set(own_pairs(h))
    yield "contraction (generic d, after the extra barrier)", iv


@pytest.mark.parametrize("NT", range(1, 9))
def test_no_cross_warp_hazard_inside_a_barrier_interval(NT):
    names = []
    for name, iv in schedule(NT):
        iv.check(f"NT={NT} {name}")
        names.append(name)
    assert len(names) == 2 + 2 * NT + 2 * (NT - 1) + NT + 1 + 2


@pytest.mark.parametrize("NT", range(1, 9))
def test_every_tile_product_has_exactly_one_owner(NT):
    """Work split: each (row block, column) product of the three O(N^3) phases is done by exactly one warp of the pair."""
    chol = [(kb, u, h) for kb in range(NT) for h in (0, 1) for u in range(h, NT - kb, 2)]
    assert sorted((kb, u) for kb, u, _ in chol) == sorted((kb, u) for kb in range(NT) for u in range(NT - kb))
    trtri = [(i, j) for i in range(1, NT) for h in (0, 1) for j in range(1 - h, i, 2)]
    assert sorted(trtri) == [(i, j) for i in range(1, NT) for j in range(i)]
    lauum = [(i, j) for i in range(NT) for h in (0, 1) for j in range(1 - h, i + 1, 2)]
    assert sorted(lauum) == [(i, j) for i in range(NT) for j in range(i + 1)]
    # balance at NT = 7 (the HBS problem size), in DMMA instructions; warp 0 additionally factors the 7 diagonal tiles
    # (~130 FP64 instructions each = ~16 DMMA pipe-slots), which is why it gets the lighter column split in phases 3 and 5
    if NT == 7:
        cost = [0, 0]
        for kb, u, h in chol:
            cost[h] += 2 * kb + (2 if u > 0 else 0)
        for i in range(1, NT):
            for h in (0, 1):
                for j in range(1 - h, i, 2):
                    cost[h] += 2 * (i - j) + 2
        for i in range(NT):
            for h in (0, 1):
                for j in range(1 - h, i + 1, 2):
                    cost[h] += 2 * (NT - i)
        cost[0] += 7 * 16
        assert abs(cost[0] - cost[1]) < 0.25 * max(cost), cost
