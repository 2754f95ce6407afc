"""The plan of the region-sharded refit (k_plan, steps 2-3 of b200pt_guiding_update_all_ranks) on ONE GPU: for 1 / 2 / 4 / 8
ranks and ragged per-rank region counts, every rank's view of the plan is compared with a plain numpy restatement —
ownership by longest-processing-time-first, the (owner, region) layout of every rank's sorted buffer, the owner's
region-contiguous fit layout in rank order, and the copy segments (both exchange modes).  The multi-GPU execution itself
is checked bit for bit by tools/multi_gpu_guiding_check.py on 2-8 GPUs."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def reference_plan(counts, me, peer_mode):
    n, R = counts.shape
    total = counts.sum(0).astype(np.int64)
    order = sorted(range(R), key=lambda g: (-total[g], g))
    loads, owner, active = [0] * n, np.zeros(R, np.int64), []
    for g in order:
        if total[g] == 0:
            owner[g] = g % n
            continue
        d = min(range(n), key=lambda k: (loads[k], k))
        owner[g] = d
        loads[d] += int(total[g])
        if d == me:
            active.append(g)
    bins = sorted(range(R), key=lambda g: (owner[g], g))
    src_start = np.zeros((n, R), np.int64)
    for s in range(n):
        run = 0
        for g in bins:
            src_start[s, g] = run
            run += int(counts[s, g])
    mine = [g for g in bins if owner[g] == me]
    begin, run = np.zeros(R, np.int64), 0
    for g in mine:
        begin[g] = run
        run += int(total[g])
    slice_start = [sum(int(counts[s, g]) for g in range(R) if owner[g] < me) for s in range(n)]
    recv = [sum(int(counts[s, g]) for g in mine) for s in range(n)]
    stage, run = [0] * n, 0
    for s in range(n):
        stage[s] = run
        if s != me:
            run += recv[s]
    segs = []
    for g in mine:
        dst = int(begin[g])
        for s in range(n):
            off = int(src_start[s, g])
            if not peer_mode and s != me:
                off = stage[s] + off - slice_start[s]
            segs.append((s, off, dst, int(counts[s, g])))
            dst += int(counts[s, g])
    return dict(owner=owner, active=active, src_start=src_start, begin=begin, mine=mine, segs=np.array(segs, np.int64).reshape(-1, 4), loads=loads, total=total)


@pytest.mark.parametrize("nranks,splits,seed", [(1, 3, 0), (2, 5, 1), (4, 8, 2), (8, 8, 3), (8, 2, 4), (3, 6, 5), (2, 11, 6), (4, 10, 7)])
def test_plan_matches_numpy_restatement(nranks, splits, seed):
    P = helpers.pt()
    scene = P.Scene(helpers.scene_path("cornell-dielectric"))
    r = P.Renderer(32, 32, 0, splits)
    r.set_scene(scene)
    R = 1 << splits
    rng = np.random.default_rng(seed)
    counts = (rng.pareto(1.2, (nranks, R)) * 3000).astype(np.uint32)        # heavy-tailed: a few regions hold most samples
    counts[:, rng.integers(0, R, max(1, R // 8))] = 0                         # regions nobody has samples for
    if nranks > 1:
        counts[1:, rng.integers(0, R)] = 0                                    # a region only rank 0 sees
    covered = np.zeros(R, bool)
    for me in range(nranks):
        for peer_mode in (1, 0):
            got, ref = r.guiding_plan_debug(counts, me, peer_mode), reference_plan(counts, me, peer_mode)
            assert np.array_equal(got["owner"], ref["owner"]), me
            assert list(got["active"]) == ref["active"]                       # largest first, ties by region id
            assert np.array_equal(got["src_start"], ref["src_start"])
            assert got["num_owned"] == len(ref["mine"]) and got["owned_samples"] == ref["loads"][me] and got["total_samples"] == int(ref["total"].sum())
            assert got["local_valid"] == int(counts[me].sum())
            for g in ref["mine"]:
                assert got["region_begin"][g] == ref["begin"][g] and got["region_len"][g] == ref["total"][g]
            assert np.array_equal(got["region_len"][ref["owner"] != me], np.zeros((ref["owner"] != me).sum(), np.uint32))
            assert np.array_equal(got["segments"].astype(np.int64), ref["segs"]), (me, peer_mode)
        covered[ref["owner"] == me] = True
        # longest-processing-time-first keeps the ranks within one region of each other
        assert max(ref["loads"]) - min(ref["loads"]) <= int(ref["total"].max())
    assert covered.all()
