"""BASELINE config 3: veachMIS 1280x720 (rough-conductor plates, alpha 0.02 / 0.06 / 0.1 / 0.2, four sphere lights) and
miPhong (the same set-up with Phong plates) rendered three ways — light sampling only (enableNEE=1, enableMIS=0), BSDF
sampling only (enableNEE=0 => addDirectLights) and MIS — with the ray throughput and the image means (the three
estimators are unbiased for the same integral, so the means must agree).  Prints one JSON line per scene.
Usage: python tools/run_config3.py [W H frames spp]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers  # noqa: E402

P = helpers.pt()
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1280, 720)
FRAMES = int(sys.argv[3]) if len(sys.argv) > 3 else 8
SPP = int(sys.argv[4]) if len(sys.argv) > 4 else 16
MODES = dict(nee=dict(enableNEE=1, enableMIS=0), bsdf=dict(enableNEE=0, enableMIS=0), mis=dict(enableNEE=1, enableMIS=1))
for name in ("veachMIS", "miPhong"):
    scene = P.Scene(helpers.scene_path(name))
    view, proj = scene.camera_matrices(W / H)
    r = P.Renderer(W, H, 0, 0)
    r.set_scene(scene)
    r.set_camera(view, proj)
    out = dict(scene=name, width=W, height=H, spp_per_frame=SPP, frames=FRAMES, modes={})
    for mode, over in MODES.items():
        pcs = [P.default_push_constants(randomUInt=P.tea(f, 0xC0FFEE), previousFrames=f, samplesPerPixel=SPP, **over) for f in range(FRAMES + 1)]
        r.render_frame(pcs[0])                      # warm-up (and previousFrames = 0 resets the accumulation)
        r.stats_reset()
        for pc in pcs[1:]:
            r.render_frame(pc)
        st = r.stats()
        img = r.read_image()[..., :3]
        out["modes"][mode] = dict(ms_per_frame=round(st.ms_total / FRAMES, 2), Mrays_per_s=round((st.extend_rays + st.shadow_rays) / max(st.ms_total, 1e-6) / 1e3, 1),
                                  spp_per_s=round(SPP * FRAMES / st.ms_total * 1e3, 1), extend_rays=int(st.extend_rays), shadow_rays=int(st.shadow_rays),
                                  image_mean=round(float(img.mean()), 5), finite=bool((img == img).all()))
    print(json.dumps(out))
