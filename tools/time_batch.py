#!/usr/bin/env python
"""Run under gpurun: b200pt_render_frames vs frame-by-frame on the bench workload, with and without stage timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
P = bench._load("b200pt_binding", os.path.join(ROOT, "rtx-pathtracer_b200", "b200pt.py"))
scene = P.Scene(bench.SCENE)
view, proj = scene.camera_matrices(bench.WIDTH / bench.HEIGHT)
r = P.Renderer(bench.WIDTH, bench.HEIGHT, 0, 0)
r.set_scene(scene); r.set_camera(view, proj)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pcs = [bench.push_constants(P, P.tea(i, bench.SEED), i) for i in range(K)]
for timing in (False, True):
    r.set_stage_timing(timing)
    for rep in range(3):
        for mode in ("single", "batch"):
            r.stats_reset()
            t0 = time.perf_counter(); r.timer_start()
            if mode == "batch": r.render_frames(pcs)
            else:
                for pc in pcs: r.render_frame(pc)
            ms = r.timer_stop(); wall = (time.perf_counter() - t0) * 1e3
            st = r.stats()
            rays = int(st.extend_rays) + int(st.shadow_rays)
            print("timing=%d rep=%d %-6s dev_ms=%8.2f wall_ms=%8.2f Mrays/s=%7.1f iterations=%d" % (timing, rep, mode, ms, wall, rays / ms / 1e3, int(st.iterations)), flush=True)
