import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers
P = helpers.pt()
name = sys.argv[1] if len(sys.argv) > 1 else "cornell-dielectric"
w, h = 160, 90
over = dict(enableNEE=1, enableMIS=1)
scene, r, o = helpers.make_pair(name, w, h)
pc = P.default_push_constants(randomUInt=P.tea(0, 0xC0FFEE), previousFrames=0, samplesPerPixel=2, **over)
r.render_frame(pc)
o.render_region(pc, threads=os.cpu_count())
g = r.read_image()[..., :3].astype(np.float64); c = o.image()[..., :3].astype(np.float64)
rel = np.abs(g - c) / np.maximum(np.abs(c), 1e-3)
bad = ~(rel <= 1e-4).all(-1)
s = r.stats(); oc = o.counters()
print("gpu extend", s.extend_rays, "shadow", s.shadow_rays, "oracle", oc)
print("bad pixels", bad.sum(), "of", bad.size)
ys, xs = np.nonzero(bad)
for y, x in list(zip(ys, xs))[:40]:
    print(x, y, g[y, x], c[y, x])
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/bad_%s.npy" % name, bad)
# 1 spp depth-limited variants to localise
for md in (0, 1, 2, 3):
    scene, r, o = helpers.make_pair(name, w, h)
    pc = P.default_push_constants(randomUInt=P.tea(0, 0xC0FFEE), previousFrames=0, samplesPerPixel=1, maxDepth=md, maxFollowDiscrete=0, **over)
    r.render_frame(pc); o.render_region(pc, threads=os.cpu_count())
    g = r.read_image()[..., :3].astype(np.float64); c = o.image()[..., :3].astype(np.float64)
    rel = np.abs(g - c) / np.maximum(np.abs(c), 1e-3)
    bad = ~(rel <= 1e-4).all(-1)
    s = r.stats(); oc = o.counters()
    print("maxDepth", md, "bad", bad.sum(), "gpu", s.extend_rays, s.shadow_rays, "oracle", oc["extend_rays"], oc["shadow_rays"])
