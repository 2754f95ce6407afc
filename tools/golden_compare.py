#!/usr/bin/env python3
"""GPU box: level-3 comparison against the reference tree's converged images (tests/golden/*_160x90.npy = the reference's
1280x720 EXRs box-filtered to 160x90): renders each scene at 1280x720, box-filters, prints relMSE and mean ratio.
usage: golden_compare.py [spp]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
P = helpers.pt()
SPP = int(sys.argv[1]) if len(sys.argv) > 1 else 256
PER = 16
for name in ("cornell-dielectric", "veachMIS", "miPhong", "envMap", "testSpheres", "irradianceCache", "sponzaXML"):
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + "_160x90.npy")).astype(np.float64)
    scene = P.Scene(helpers.scene_path(name))
    view, proj = scene.camera_matrices(1280 / 720)
    r = P.Renderer(1280, 720, 0, 0); r.set_scene(scene); r.set_camera(view, proj)
    t = time.time()
    for f in range(SPP // PER):
        r.render_frame(P.default_push_constants(randomUInt=P.tea(f, 0xC0FFEE), previousFrames=f, samplesPerPixel=PER, enableNEE=1, enableMIS=1))
    img = r.read_image()[..., :3].astype(np.float64).reshape(90, 8, 160, 8, 3).mean(axis=(1, 3))
    dt = time.time() - t
    rel = (img - gold) ** 2 / (gold ** 2 + 1e-2)
    per_px = rel.mean(-1)
    print("%-20s %4d spp %.1fs  relMSE %.3e  median px %.2e  p90 %.2e  p99 %.2e  mean ours/gold %.4f  per-channel ratio %s" %
          (name, SPP, dt, rel.mean(), np.median(per_px), np.quantile(per_px, 0.9), np.quantile(per_px, 0.99), img.mean() / gold.mean(),
           np.round(img.mean((0, 1)) / gold.mean((0, 1)), 3)), flush=True)
    # coarse map of where the error sits: 9 x 16 blocks
    blocks = per_px.reshape(9, 10, 16, 10).mean(axis=(1, 3))
    print("   worst blocks (row, col, relMSE):", [(int(i), int(j), float("%.2e" % blocks[i, j])) for i, j in zip(*np.unravel_index(np.argsort(-blocks.ravel())[:5], blocks.shape))])
    np.save("gpurun_out/l3_%s.npy" % name, img.astype(np.float32))
