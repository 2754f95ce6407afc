"""Native multi-GPU entry points of the C ABI (b200pt_comm_*, b200pt_reduce_image, b200pt_allgather_samples) with a
one-rank communicator: the single-GPU box of the test tier can check the plumbing; tools/multi_gpu_native_check.py runs the
same calls on 2+ GPUs against the torch.distributed path."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def test_one_rank_communicator_reduce_and_gather():
    P = helpers.pt()
    scene = P.Scene(helpers.scene_path("veachMIS"))
    w, h = 64, 36
    view, proj = scene.camera_matrices(w / h)
    r = P.Renderer(w, h, 0, 2)
    r.set_scene(scene)
    r.set_camera(view, proj)
    with pytest.raises(P.B200ptError):
        r.reduce_image()                      # no communicator yet
    uid = P.comm_unique_id()
    assert len(uid) == P.COMM_ID_BYTES and any(uid)
    r.comm_init(uid, 0, 1)
    with pytest.raises(P.B200ptError):
        r.comm_init(uid, 0, 1)                # already initialised
    r.render_frame(P.default_push_constants(randomUInt=P.tea(0, 9), previousFrames=0, samplesPerPixel=2, enableMIS=1, updateGuiding=1))
    before = r.read_image()
    r.reduce_image(P.IMAGE_OUTPUT, 3)         # (3 * image) / 3
    after = r.read_image()
    assert np.allclose(after, before, rtol=1e-6, atol=1e-7) and after[..., 3].min() == 1.0
    assert r.allgather_samples() == r.guiding_sample_capacity()
    # refit on the gathered records == refit on the context's own buffer
    r2 = P.Renderer(w, h, 0, 2)
    r2.set_scene(scene)
    r2.set_camera(view, proj)
    r2.guiding_put_samples(r.guiding_get_samples())
    r.guiding_update_all_ranks()
    r2.guiding_update()
    assert np.array_equal(r.guiding_get_vmms().view(np.uint8), r2.guiding_get_vmms().view(np.uint8))
    r.comm_destroy()
    with pytest.raises(P.B200ptError):
        r.allgather_samples()
