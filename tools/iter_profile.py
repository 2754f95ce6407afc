"""Queue sizes per wavefront iteration of one headline frame (B200PT_DUMP_ITERS): how much of a frame is tail."""
import os
import sys

os.environ["B200PT_DUMP_ITERS"] = "gpurun_out/iters.txt"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers  # noqa: E402
import numpy as np  # noqa: E402

P = helpers.pt()
W, H = 1280, 720
scene = P.Scene(helpers.scene_path("cornell-dielectric"))
view, proj = scene.camera_matrices(W / H)
r = P.Renderer(W, H, 0, 0)
r.set_scene(scene)
r.set_camera(view, proj)
for f in range(2):
    r.stats_reset()
    r.render_frame(P.default_push_constants(randomUInt=P.tea(f, 0xC0FFEE), previousFrames=f, samplesPerPixel=16, enableMIS=1))
st = r.stats()
r.close()
rows = np.loadtxt("gpurun_out/iters.txt", dtype=np.int64)
half = len(rows) // 2
rows = rows[half:]                      # second frame
n = rows[:, 1] + rows[:, 2] + rows[:, 3]
print("iterations", len(rows), "rays", int(n.sum()), "frame ms", st.ms_total)
N = W * H
for thr in (0.5, 0.25, 0.1, 0.03, 0.01, 0.001):
    m = rows[:, 1] < thr * N
    print("path queue < %5.1f %% of pixels: %4d iterations, %5.2f %% of the rays" % (thr * 100, int(m.sum()), 100.0 * n[m].sum() / n.sum()))
print("path-queue size every 20 iterations:", rows[::20, 1].tolist())
