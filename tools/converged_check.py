"""Level-3 check (north_star: converged 4096-spp images reach relMSE < 1e-3): the device path and the CPU oracle render
the SAME scene with DIFFERENT (independent) random streams; two unbiased estimators of the same integral must agree up
to Monte-Carlo noise.  Prints relMSE(device, oracle) next to the noise floor relMSE(device, device with other seeds).
usage: python tools/converged_check.py [scene] [W H] [spp]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers  # noqa: E402

P = helpers.pt()
name = sys.argv[1] if len(sys.argv) > 1 else "cornell-dielectric"
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (64, 36)
SPP = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
PER = 64
over = dict(enableNEE=1, enableMIS=1, samplesPerPixel=PER)
scene, r, o = helpers.make_pair(name, W, H)


def rel_mse(a, b, eps=1e-2):
    return float(((a - b) ** 2 / (b ** 2 + eps)).mean())


def device(seed):
    for f in range(SPP // PER):
        r.render_frame(P.default_push_constants(randomUInt=P.tea(f, seed), previousFrames=f, **over))
    return r.read_image()[..., :3].astype(np.float64)


t = time.time()
ga, gb = device(0xA11CE), device(0xB0B)
td = time.time() - t
t = time.time()
for f in range(SPP // PER):
    o.render_region(P.default_push_constants(randomUInt=P.tea(f, 0xC0FFEE), previousFrames=f, **over), threads=os.cpu_count() or 8)
c = o.image()[..., :3].astype(np.float64)
print(json.dumps(dict(scene=name, width=W, height=H, spp=SPP, relmse_device_vs_oracle=rel_mse(ga, c), relmse_device_vs_device=rel_mse(ga, gb),
                      mean_device=ga.mean(), mean_oracle=c.mean(), device_s=round(td, 2), oracle_s=round(time.time() - t, 2))))
