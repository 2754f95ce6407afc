#!/usr/bin/env python3
"""GPU box: same-seed device vs oracle renders; prints, per scene / mode, the share of pixels within 1e-4 relative and the
share that is bit-identical.  usage: parity_report.py [W H spp]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
P = helpers.pt()
NT = os.cpu_count() or 1
W, H, SPP = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (160, 90, 2)
CASES = [("cornell-dielectric", dict(enableNEE=1, enableMIS=1)), ("cornell-dielectric", dict(enableNEE=0)), ("veachMIS", dict(enableNEE=1, enableMIS=1)),
         ("veachMIS", dict(enableNEE=1, enableMIS=0)), ("veachMIS", dict(enableNEE=0)), ("miPhong", dict(enableNEE=1, enableMIS=1)),
         ("sponzaXML", dict(enableNEE=1, enableMIS=1, maxDepth=8)), ("test-scene", dict(enableNEE=1, enableMIS=1, maxDepth=8)),
         ("envMap", dict(enableNEE=1, enableMIS=1, maxDepth=6)), ("testSpheres", dict(enableNEE=1, enableMIS=1, maxDepth=6)), ("roughConductor", dict(enableNEE=1, enableMIS=1))]
for name, over in CASES:
    if not os.path.exists(helpers.scene_path(name)):
        continue
    scene, r, o = helpers.make_pair(name, W, H)
    pc = P.default_push_constants(randomUInt=P.tea(0, 0xC0FFEE), previousFrames=0, samplesPerPixel=SPP, **over)
    r.render_frame(pc)
    t = time.time(); o.render_region(pc, threads=NT); dt = time.time() - t
    g, c = r.read_image()[..., :3], o.image()[..., :3]
    rel = np.abs(g.astype(np.float64) - c) / np.maximum(np.abs(c), 1e-3)
    ok = (rel <= 1e-4).all(-1).mean()
    exact = (g.view(np.uint32) == c.view(np.uint32)).all(-1).mean()
    print("%-20s %-45s within 1e-4: %.5f  bit-identical: %.5f  worst rel %.2e  (oracle %.1f s)" % (name, over, ok, exact, rel.max(), dt), flush=True)
