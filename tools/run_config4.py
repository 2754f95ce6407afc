"""BASELINE config 4: sponzaXML 1920x1080, IC_SIZE=10000, ADRRS with splitting — driven through the frame driver exactly
like RayTracingApp does (50 IC prepare frames at 1 spp, one estimate frame, then ADRRS frames); also an IC-only run and
a plain NEE+MIS run for comparison.  Prints one JSON line per run.  Usage: python tools/run_config4.py [W H frames spp]
Environment: RUN4_RUNS=plain,ic,adrrs (default all), RUN4_STAGE=0 turns the per-kernel events off (they cost ~10 %)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers  # noqa: E402

P = helpers.pt()
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
FRAMES = int(sys.argv[3]) if len(sys.argv) > 3 else 8
SPP = int(sys.argv[4]) if len(sys.argv) > 4 else 4
scene = P.Scene(helpers.scene_path("sponzaXML"))
view, proj = scene.camera_matrices(W / H)


def run(name, **settings):
    r = P.Renderer(W, H, 10000, 0)
    t0 = time.time()
    r.set_scene(scene)
    r.set_camera(view, proj)
    t_scene = time.time() - t0
    r.set_stage_timing(os.environ.get("RUN4_STAGE", "1") != "0")
    app = P.App(r, accumulate=True, samplesPerPixel=SPP, enableNEE=1, enableMIS=1, **settings)
    phases = []

    def phase(label, n):
        r.stats_reset()
        t = time.time()
        for _ in range(n):
            app.draw_frame(P.tea(len(phases) * 1000 + _, 0xC0FFEE))
        st = r.stats()
        hdr = r.ic_get()[0] if r.ic_size else None
        phases.append(dict(phase=label, frames=n, wall_ms=round((time.time() - t) * 1e3, 1), device_ms=round(st.ms_total, 1), trace_ms=round(st.ms_extend + st.ms_shadow, 1), shade_and_ic_ms=round(st.ms_shade, 1),
                           Mrays_per_s=round((st.extend_rays + st.shadow_rays) / max(st.ms_total, 1e-6) / 1e3, 1), extend=st.extend_rays,
                           shadow=st.shadow_rays, launches=st.kernel_launches, cache_entries=(hdr.nextCacheSlot if hdr else 0)))

    if settings.get("useIrradianceCache") or settings.get("useADRRS"):
        phase("prepare", app.state.irradianceCachePrepareFrames)
    if settings.get("useADRRS"):
        phase("estimate", 1)
    phase("render", FRAMES)
    img = r.read_image()[..., :3]
    print(json.dumps(dict(run=name, width=W, height=H, spp_per_frame=SPP, scene_setup_s=round(t_scene, 2), image_mean=float(img.mean()),
                          finite=bool((img == img).all()), phases=phases)))
    r.close()


RUNS = os.environ.get("RUN4_RUNS", "plain,ic,adrrs").split(",")
if "plain" in RUNS:
    run("plain NEE+MIS")
if "ic" in RUNS:
    run("IC", useIrradianceCache=1)
if "adrrs" in RUNS:
    run("ADRRS+split", useADRRS=1, adrrsSplit=1, adrrsS=5.0)
