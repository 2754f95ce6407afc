#!/usr/bin/env python3
"""Generates tests/golden/guiding_golden.npz with the REFERENCE's own code (lightpmm + external/guiding headers compiled
from /root/reference into oracle/_ref/libguiding_ref.so; restated PathGuiding glue in oracle/guiding_ref.cpp).
Run in the build container, where /root/reference exists:   python tests/golden/make_guiding_golden.py

Contents
  fastexp_in / fastexp_out     lightpmm::exp (PMM_APPROX_EXP) known answers, bit patterns
  For the fit fixture (splits = 2 -> 4 regions of the cornell-dielectric scene box, seeds below, default parameters,
  three update rounds so fit, updateFit, split and (third round, > 8192 samples) merge run):
  aabbs, round{0,1,2}_K, _iters, _state (regions x 14 x 16), _vmms (raw VMM_Theta bytes), _sorted_offsets, _sorted (records)
The sample batches themselves are regenerated from tests/guiding_data.py with the same seeds (numpy's default_rng
stream is stable), so they are not stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import guiding_data
import helpers

SCENE_MIN, SCENE_MAX = helpers.scene_box("cornell-dielectric")   # (-1.686433, -0.0356, -1.686433) .. (1.686433, 3.386433, 1.686433)
SPLITS, PER_REGION, SEEDS = 2, 3000, (11, 12, 13)


def main():
    P, O = helpers.pt(), helpers.oracle()
    out = {}
    x = np.concatenate([np.linspace(-180.0, 2.0, 4001), -np.logspace(-7, 2, 2000), [0.0, -0.0, -87.0, -88.5, -126.0, -1e30],
                        # overflow / NaN / infinities: cvtps2dq returns the "integer indefinite" 0x80000000 (= -0.0f), _mm_max_ps passes a NaN
                        # in its second operand through — both reachable from the merge metric (VMFKernel::division with kappa = 50000)
                        np.linspace(60.0, 120.0, 241), [1e4, 1e30, np.inf, -np.inf, np.nan]]).astype(np.float32)
    out["fastexp_in"] = x
    out["fastexp_out"] = O.ref_fastexp(x).view(np.uint32)
    gp = P.default_guiding_params()
    g = O.GuidingRef(SPLITS, SCENE_MIN, SCENE_MAX, gp)
    aabbs = g.aabbs()
    out["aabbs"] = aabbs.view(np.float32).reshape(-1, 6)
    for rnd, seed in enumerate(SEEDS):
        batch = guiding_data.make_batch(aabbs, PER_REGION, seed)
        g.update(batch)
        states = [g.state(r) for r in range(g.region_count)]
        out["round%d_K" % rnd] = np.array([s["K"] for s in states], dtype=np.int32)
        out["round%d_iters" % rnd] = np.array([s["numEMIterations"] for s in states], dtype=np.int32)
        out["round%d_scalars" % rnd] = np.array([[s["sampleWeight"], s["numSamples"], s["totalNumSamples"]] for s in states], dtype=np.float64)
        out["round%d_state" % rnd] = np.stack([np.stack([s[f] for f in O.STATE_FIELDS]) for s in states])
        out["round%d_vmms" % rnd] = g.vmms().view(np.uint8).reshape(g.region_count, -1)
        srt, off = g.sorted_samples()
        out["round%d_sorted_offsets" % rnd] = off
        if rnd == 0:      # the sorted + pre-fitted records (SampleCollector::getSortedData + PathGuiding::preFit)
            out["round0_sorted"] = srt.view(np.uint8).reshape(len(srt), -1)
    path = os.path.join(HERE, "guiding_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; K:", out["round0_K"], out["round1_K"], "iters:", out["round0_iters"], out["round1_iters"], out["round2_iters"])


if __name__ == "__main__":
    main()
