"""Host restatement of the slot -> (owner lane, triangle) search of the pooled triangle phase (csrc/traverse.cuh, PT_COOP_LIST = 2):
the 3-bit table constant is read from the source, the popcount search and the binary search over the lanes' inclusive sums are
replayed in numpy and compared with the plain definition (the r-th set bit; the lane whose interval contains the slot)."""
import os
import re

import numpy as np

SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "rtx-pathtracer_b200", "csrc", "traverse.cuh")


def _lut():
    m = re.search(r"ti \+= uint32_t\(\((0x[0-9a-fA-F]+)ull >> \(\(t3 \* 3u \+ r\) \* 2u\)\) & 3ull\);", open(SRC).read())
    assert m, "the 3-bit table of the n-th-set-bit search moved"
    return int(m.group(1), 16)


def _popc(x):
    x = x - ((x >> 1) & 0x55555555)
    x = (x & 0x33333333) + ((x >> 2) & 0x33333333)
    x = (x + (x >> 4)) & 0x0F0F0F0F
    return ((x * 0x01010101) >> 24) & 0xFF


def _nth_set_bit(b, r, lut):
    b, r = b.astype(np.uint64), r.astype(np.uint64)
    ti = np.zeros(b.shape, np.uint64)
    for mask, step in ((0xFFF, 12), (0x3F, 6), (0x7, 3)):
        c = _popc((b >> ti) & np.uint64(mask))
        m = r >= c
        r = np.where(m, r - c, r)
        ti = np.where(m, ti + np.uint64(step), ti)
    t3 = (b >> ti) & np.uint64(7)
    return ti + ((np.uint64(lut) >> ((t3 * np.uint64(3) + r) * np.uint64(2))) & np.uint64(3))


def test_nth_set_bit_of_24_bit_masks():
    lut = _lut()
    rng = np.random.default_rng(3)
    masks = np.concatenate([np.arange(1, 1 << 12, dtype=np.uint64), np.arange(1, 1 << 12, dtype=np.uint64) << np.uint64(12),
                            rng.integers(1, 1 << 24, 400_000, dtype=np.uint64), np.array([0xFFFFFF, 0x800000, 0x800001, 0xE00000, 0x000007], np.uint64)])
    cnt = _popc(masks)
    for r0 in range(24):
        b = masks[cnt > r0]
        ti = _nth_set_bit(b, np.full(b.shape, r0), lut)
        assert (((b >> ti) & np.uint64(1)) == 1).all(), r0
        assert (_popc(b & ((np.uint64(1) << ti) - np.uint64(1))) == r0).all(), r0


def test_owner_search_over_inclusive_sums():
    rng = np.random.default_rng(5)
    for _ in range(300):
        cnt = rng.integers(0, 6, 32) * (rng.random(32) < rng.random())        # many empty lanes, a few with several triangles
        cnt = cnt.astype(np.int64)
        if rng.random() < 0.1:
            cnt[rng.integers(0, 32)] = 48                                     # one lane with two full leaf groups
        incl = np.cumsum(cnt)
        total = int(incl[-1])
        for i in range(total):
            lo, hi = 0, 31
            for _step in range(5):
                mid = (lo + hi) >> 1
                if incl[mid] > i:
                    hi = mid
                else:
                    lo = mid + 1
            owner = lo
            assert incl[owner] - cnt[owner] <= i < incl[owner], (cnt.tolist(), i, owner)
