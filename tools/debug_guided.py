import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers
P = helpers.pt()
W, H = 96, 54
scene, r, o = helpers.make_pair("cornell-dielectric", W, H, guiding_splits=4)
o.set_guiding(r.guiding_aabbs(), r.guiding_get_vmms())
pc = P.default_push_constants(randomUInt=P.tea(0, 0xC0FFEE), previousFrames=0, samplesPerPixel=2, enableMIS=1, updateGuiding=1)
r.render_frame(pc); o.render_region(pc, threads=os.cpu_count())
g = r.guiding_get_samples().reshape(H * W, 16); c = o.samples(P.DIRECTIONAL_DATA_DTYPE).reshape(H * W, 16)
both = (g["flags"] != 0xFFFFFFFF) & (c["flags"] != 0xFFFFFFFF)
with np.errstate(invalid="ignore", divide="ignore"):
    d = np.abs(g["weight"] - c["weight"]) / np.maximum(np.abs(c["weight"]), 1e-3)
d = d[both]
print("weight rel err histogram:", np.histogram(d, bins=[0, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1, 1e9])[0])
gi, ci = r.read_image()[..., :3].astype(np.float64), o.image()[..., :3].astype(np.float64)
rel = np.abs(gi - ci) / np.maximum(np.abs(ci), 1e-3)
badpix = ~(rel <= 1e-4).all(-1).reshape(-1)
badw = (np.where(both, np.abs(g["weight"] - c["weight"]) / np.maximum(np.abs(c["weight"]), 1e-3) > 1e-4, False)).any(axis=1)
print("pixels with radiance mismatch", badpix.sum(), "pixels with weight mismatch", badw.sum(), "both", (badpix & badw).sum())
idx = np.nonzero(badw & ~badpix)[0][:6]
for i in idx:
    print("pixel", i, "flags", g["flags"][i][:6], c["flags"][i][:6]); print("  gpu w", g["weight"][i][:6]); print("  cpu w", c["weight"][i][:6]); print("  dist", g["distance"][i][:6], c["distance"][i][:6])
