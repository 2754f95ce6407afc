#!/usr/bin/env python3
"""Development aid (GPU box): reference build (oracle/_ref) vs host build of guiding_math.cuh (tests/harness) vs device
strict-order update on the same batches; prints per round / region the fields that differ most."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import guiding_data, helpers
from test_guiding_cpu import HostFit, FIELDS, SCENE_MIN, SCENE_MAX
P, O = helpers.pt(), helpers.oracle()
splits, per_region = int(sys.argv[1]), int(sys.argv[2])
scene = P.Scene(helpers.scene_path("cornell-dielectric"))
r = P.Renderer(64, 64, 0, splits); r.set_scene(scene)
gp = P.default_guiding_params()
g = O.GuidingRef(splits, SCENE_MIN, SCENE_MAX, gp)
aabbs = r.guiding_aabbs()
h = HostFit(aabbs, gp)
R = len(aabbs)
for rnd in range(3):
    counts = [per_region] * R
    counts[rnd % R] = 0
    counts[(rnd + 1) % R] = 9
    batch = guiding_data.make_batch(aabbs, counts, 1000 + rnd)
    g.update(batch); h.update(batch); r.guiding_update_host(batch, gp)
    for i in range(R):
        a, b, c = g.state(i), h.state(i), r.guiding_state(i)
        line = []
        for f in ("K", "numEMIterations", "sampleWeight"):
            if not (a[f] == b[f] == c[f]): line.append("%s ref %r host %r dev %r" % (f, a[f], b[f], c[f]))
        K = a["K"]
        for f in FIELDS:
            x, y, z = a[f][:K].astype(np.float64), b[f][:K].astype(np.float64), c[f][:K].astype(np.float64)
            with np.errstate(all="ignore"):
                eh = np.nanmax(np.abs(x - y) / np.maximum(np.abs(x), 1e-3)) if K else 0
                ed = np.nanmax(np.abs(x - z) / np.maximum(np.abs(x), 1e-3)) if K else 0
                ehd = np.nanmax(np.abs(y - z) / np.maximum(np.abs(y), 1e-3)) if K else 0
            if ed > 1e-4 or ehd > 1e-6: line.append("%s: ref-host %.1e ref-dev %.1e host-dev %.1e" % (f, eh, ed, ehd))
        if line: print("round %d region %d:" % (rnd, i), " | ".join(line))
print("done")
