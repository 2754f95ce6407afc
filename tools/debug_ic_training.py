"""Diagnostics of guiding training on irradiance-cache / ADRRS frames (tests/test_ic_gpu.py::
test_guiding_training_on_cache_frames_matches_oracle): per-frame agreement of the recorded samples and the image."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers
import test_ic_gpu as ti
P = helpers.pt()
W, H, INVALID = ti.W, ti.H, 0xFFFFFFFF
scene, r, o = helpers.make_pair(ti.SCENE, W, H, ic_size=ti.IC_SIZE, guiding_splits=3)
cache = ti._build_cache(P, o)
r.ic_put(*cache)
est = np.full((H, W, 4), 0.5, np.float32)
r.write_image(P.IMAGE_ESTIMATE, est); o.set_image(P.IMAGE_ESTIMATE, est)
o.set_guiding(r.guiding_aabbs(), r.guiding_get_vmms())
off = dict(irradianceCreateProb=0.0, irradianceUpdateProb=0.0)
frames = [dict(useIrradianceCache=1, useIrradianceCacheOnGlossy=1, **off), dict(useADRRS=1, adrrsSplit=0, adrrsS=5.0, **off),
          dict(useIrradianceCache=1, useIrradianceCacheOnGlossy=1, irradianceCreateProb=0.02, irradianceUpdateProb=0.005)]
for f, kw in enumerate(frames):
    pc = ti._pc(P, 80 + f, samplesPerPixel=2, updateGuiding=1, **kw)
    r.render_frame(pc); o.render_region(pc, threads=os.cpu_count())
    g = r.guiding_get_samples().reshape(H * W, 16); c = o.samples(P.DIRECTIONAL_DATA_DTYPE).reshape(H * W, 16)
    gv, cv = g["flags"] != INVALID, c["flags"] != INVALID
    both = gv & cv
    print("frame", f, kw)
    print("  valid gpu / cpu", gv.sum(), cv.sum(), "same slots", (gv == cv).all(axis=1).mean())
    for fld in ("pdf", "weight", "distance"):
        with np.errstate(invalid="ignore", divide="ignore"):
            d = np.abs(g[fld] - c[fld]) / np.maximum(np.abs(c[fld]), 1e-3)
        print("  %-8s rel err histogram" % fld, np.histogram(d[both], bins=[0, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1, 1e9])[0])
    print("  flags equal", (g["flags"] == c["flags"])[both].mean())
    gi, ci = r.read_image()[..., :3].astype(np.float64), o.image()[..., :3].astype(np.float64)
    rel = np.abs(gi - ci) / np.maximum(np.abs(ci), 1e-3)
    print("  image: pixels within 1e-4", (rel <= 1e-4).all(-1).mean(), "means", gi.mean(), ci.mean())
    bad = np.nonzero((gv != cv).any(axis=1))[0][:4]
    for i in bad:
        print("  pixel", i, "gpu flags", g["flags"][i][:5], "cpu flags", c["flags"][i][:5], "gpu w", g["weight"][i][:5], "cpu w", c["weight"][i][:5])
print("cache slots gpu / cpu", r.ic_get()[0].nextCacheSlot, o.ic_get(P)[0].nextCacheSlot)
