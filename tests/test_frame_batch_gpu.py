"""b200pt_render_frames: a batch of frames (pixels walk from frame to frame without waiting for each other) must leave
exactly the images that the same frames leave when rendered one at a time (src/RayTracingApp.cpp:159-215 accumulation
loop), and must fall back to frame-after-frame rendering when the frames are not independent."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _pcs(P, n, spp, **kw):
    return [P.default_push_constants(randomUInt=P.tea(i, 0xC0FFEE), previousFrames=i, samplesPerPixel=spp, enableNEE=1, enableMIS=1,
                                     usePowerHeuristic=1, numNEE=1, maxDepth=12, maxFollowDiscrete=3, **kw) for i in range(n)]


@pytest.mark.parametrize("scene_name,average", [("cornell-dielectric", 0), ("cornell-dielectric", 1), ("veachMIS", 0), ("sponzaXML", 0)])
def test_batch_equals_frame_by_frame(scene_name, average):
    P = helpers.pt()
    w, h = 160, 90
    scene, r, _ = helpers.make_pair(scene_name, w, h)
    pcs = _pcs(P, 5, 3, enableAverageInsteadOfMix=average)
    for pc in pcs:
        r.render_frame(pc)
    one = {k: r.read_image(k).copy() for k in (P.IMAGE_OUTPUT, P.IMAGE_ACCUM)}
    st1 = r.stats()
    rays1 = (int(st1.extend_rays), int(st1.shadow_rays))
    r2 = P.Renderer(w, h, 0, 0)
    r2.set_scene(scene)
    r2.set_camera(*scene.camera_matrices(w / h))
    r2.render_frames(pcs)
    st2 = r2.stats()
    assert (int(st2.extend_rays), int(st2.shadow_rays)) == rays1
    assert int(st2.iterations) < int(st1.iterations)          # the frames overlapped
    for k in one:
        got = r2.read_image(k)
        assert np.array_equal(got.view(np.uint32), one[k].view(np.uint32)), "image %d differs" % k
    assert np.isfinite(one[P.IMAGE_OUTPUT][..., :3]).all() and one[P.IMAGE_OUTPUT][..., :3].mean() > 0.01


def test_batch_continues_an_accumulation():
    """previousFrames > 0 in the first frame of the batch: the batch extends what is already in the images."""
    P = helpers.pt()
    w, h = 128, 72
    scene, r, _ = helpers.make_pair("cornell-dielectric", w, h)
    pcs = _pcs(P, 6, 2)
    for pc in pcs:
        r.render_frame(pc)
    want = r.read_image(P.IMAGE_OUTPUT).copy()
    r2 = P.Renderer(w, h, 0, 0)
    r2.set_scene(scene)
    r2.set_camera(*scene.camera_matrices(w / h))
    r2.render_frames(pcs[:2])
    r2.render_frames(pcs[2:3])         # a batch of one = render_frame
    r2.render_frames(pcs[3:])
    assert np.array_equal(r2.read_image(P.IMAGE_OUTPUT).view(np.uint32), want.view(np.uint32))


def test_dependent_frames_run_one_after_the_other():
    """Frames that differ in more than seed / frame counter (here: maxDepth) are not walked; same images as the loop."""
    P = helpers.pt()
    w, h = 96, 54
    scene, r, _ = helpers.make_pair("cornell-dielectric", w, h)
    pcs = _pcs(P, 3, 2)
    pcs[1].maxDepth = 4
    for pc in pcs:
        r.render_frame(pc)
    want = r.read_image(P.IMAGE_OUTPUT).copy()
    it1 = int(r.stats().iterations)
    r2 = P.Renderer(w, h, 0, 0)
    r2.set_scene(scene)
    r2.set_camera(*scene.camera_matrices(w / h))
    r2.render_frames(pcs)
    assert int(r2.stats().iterations) == it1
    assert np.array_equal(r2.read_image(P.IMAGE_OUTPUT).view(np.uint32), want.view(np.uint32))
