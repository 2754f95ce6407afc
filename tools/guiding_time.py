#!/usr/bin/env python3
"""Run on the GPU box: device time of the guiding update on the BASELINE config-1 shape (256 regions x N samples),
fit (first update) then updateFit (second).  B200PT_LIB selects the library build; batches are cached in /tmp."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import guiding_data, helpers
P = helpers.pt()
per_region = int(sys.argv[1]) if len(sys.argv) > 1 else 57600
scene = P.Scene(helpers.scene_path("cornell-dielectric"))
r = P.Renderer(32, 32, 0, 8); r.set_scene(scene)
aabbs = r.guiding_aabbs()
batches = []
for seed in (1, 2):
    path = "/tmp/gbatch_%d_%d.npy" % (per_region, seed)
    if not os.path.exists(path):
        np.save(path, guiding_data.make_batch(aabbs, per_region, seed, invalid_fraction=0.0))
    batches.append(np.load(path))
gp = P.default_guiding_params()
out = []
for rep in range(3):
    r.guiding_reset(gp)
    for b in batches:
        r.stats_reset()
        r.guiding_update_host(b, gp)
        s = r.stats()
        out.append((s.ms_guiding_sort, s.ms_guiding_fit, s.guiding_em_sample_iterations))
best = [min(o[1] for o in out[i::2]) for i in (0, 1)]
print("%s: sort %.2f ms | fit %.2f ms (%.0f M sample-iters) | updateFit %.2f ms (%.0f M sample-iters)" %
      (os.path.basename(os.environ.get("B200PT_LIB", "default")), out[0][0], best[0], out[0][2] / 1e6, best[1], out[1][2] / 1e6))
