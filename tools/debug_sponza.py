import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers
P = helpers.pt()
NT = os.cpu_count()
w, h = 128, 72
scene, r, o = helpers.make_pair("sponzaXML", w, h)
def run(**kw):
    pc = P.default_push_constants(randomUInt=P.tea(0, 0xC0FFEE), previousFrames=0, samplesPerPixel=1, **kw)
    r.render_frame(pc); o.render_region(pc, threads=NT)
    g = r.read_image()[..., :3].astype(np.float64); c = o.image()[..., :3].astype(np.float64)
    rel = np.abs(g - c) / np.maximum(np.abs(c), 1e-3)
    ok = (rel <= 1e-4).all(-1)
    print(kw, "frac %.4f" % ok.mean(), "mean g %.5f c %.5f" % (g.mean(), c.mean()), "median rel of bad %.3g" % (np.median(rel.max(-1)[~ok]) if (~ok).any() else 0), flush=True)
    return ok, g, c
for kw in (dict(enableNEE=0, maxDepth=0), dict(enableNEE=1, enableMIS=0, maxDepth=0), dict(enableNEE=1, enableMIS=1, maxDepth=0), dict(enableNEE=1, enableMIS=0, maxDepth=1),
           dict(enableNEE=1, enableMIS=0, maxDepth=2), dict(enableNEE=1, enableMIS=0, maxDepth=8), dict(enableNEE=0, maxDepth=8)):
    ok, g, c = run(**kw)
ok, g, c = run(enableNEE=1, enableMIS=0, maxDepth=1)
ys, xs = np.nonzero(~ok)
for y, x in list(zip(ys, xs))[:12]:
    print(x, y, g[y, x], c[y, x])
