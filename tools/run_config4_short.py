"""Short IC + ADRRS run for the ncu launch list (tools/profile_config4.sh): 6 prepare frames, estimate, 2 ADRRS frames."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers  # noqa: E402

P = helpers.pt()
W, H = 1920, 1080
scene = P.Scene(helpers.scene_path("sponzaXML"))
view, proj = scene.camera_matrices(W / H)
r = P.Renderer(W, H, 10000, 0)
r.set_scene(scene)
r.set_camera(view, proj)
app = P.App(r, accumulate=True, samplesPerPixel=4, enableNEE=1, enableMIS=1, useADRRS=1, adrrsSplit=1)
PREP = int(os.environ.get("RUN4_PREPARE", "6"))
app.state.irradianceCachePrepareFrames = PREP
for f in range(PREP + 1 + 2):
    app.draw_frame(P.tea(f, 0xC0FFEE))
print("cache entries", r.ic_get()[0].nextCacheSlot)
