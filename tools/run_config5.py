"""BASELINE config 5 (stand-in scene: sponzaXML — fireplace_room.obj is missing from the reference checkout): 3840x2160,
GUIDING_SPLITS=8, 6 guiding optimisation frames (updateGuiding) through the frame driver, then guided frames
(useGuiding, guidingProb 0.5, parallax compensation).  Prints one JSON line.  Usage: python tools/run_config5.py [W H frames spp]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers  # noqa: E402

P = helpers.pt()
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
FRAMES = int(sys.argv[3]) if len(sys.argv) > 3 else 4
SPP = int(sys.argv[4]) if len(sys.argv) > 4 else 2
scene = P.Scene(helpers.scene_path("sponzaXML"))
view, proj = scene.camera_matrices(W / H)
r = P.Renderer(W, H, 0, 8)
r.set_scene(scene)
r.set_camera(view, proj)
app = P.App(r, accumulate=True, samplesPerPixel=SPP, enableNEE=1, enableMIS=1, updateGuiding=1, useParallaxCompensation=1)
out = dict(width=W, height=H, spp_per_frame=SPP, regions=r.guiding_region_count(), phases=[])


def phase(label, n):
    r.stats_reset()
    t = time.time()
    for k in range(n):
        app.draw_frame(P.tea(len(out["phases"]) * 100 + k, 0xC0FFEE))
    st = r.stats()
    out["phases"].append(dict(phase=label, frames=n, wall_ms=round((time.time() - t) * 1e3, 1), render_ms=round(st.ms_total, 1),
                              guiding_sort_ms=round(st.ms_guiding_sort, 1), guiding_fit_ms=round(st.ms_guiding_fit, 1), guiding_samples=st.guiding_samples,
                              Mrays_per_s=round((st.extend_rays + st.shadow_rays) / max(st.ms_total, 1e-6) / 1e3, 1)))


phase("training (updateGuiding, 6 refits)", app.state.numGuidingOptimizations + 1)
assert app.settings.updateGuiding == 0
app.settings.useGuiding = 1
app.settings.guidingProb = 0.5
app.input_changed()
phase("guided render", FRAMES)
img = r.read_image()[..., :3]
vm = r.guiding_get_vmms()
out.update(image_mean=float(img.mean()), finite=bool((img == img).all()), mean_components=float(vm["usedDistributions"].mean()))
print(json.dumps(out))
