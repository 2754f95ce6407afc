"""Checkpoint / resume (b200pt_save_state / b200pt_load_state): a run that is saved, torn down and
resumed in a fresh context continues bit-identically — images, irradiance cache, guiding mixtures incl. adaptive splits."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
W, H = 64, 36


def _ctx(P, scene, ic_size, splits):
    view, proj = scene.camera_matrices(W / H)
    r = P.Renderer(W, H, ic_size, splits)
    r.set_scene(scene)
    r.set_camera(view, proj)
    return r


def _frames(P, r, app, first, count, gp):
    for f in range(first, first + count):
        app.renderer = r
        app.draw_frame(P.tea(f, 0xABCD), gp)


def test_resume_is_bit_identical(tmp_path):
    P = helpers.pt()
    scene = P.Scene(helpers.scene_path("irradianceCache"))
    gp = P.default_guiding_params(splitRegions=1, samplesForRegionSplit=300.0)
    settings = dict(samplesPerPixel=1, enableMIS=1, useIrradianceCache=1, irradianceCreateProb=0.02, irradianceUpdateProb=0.002)
    # uninterrupted: 3 prepare + 5 frames, the last 3 of them guiding-training frames with refits
    a = _ctx(P, scene, 256, 2)
    app_a = P.App(a, accumulate=True, **settings)
    app_a.state.irradianceCachePrepareFrames = 3
    _frames(P, a, app_a, 0, 5, gp)
    # interrupted after 5 frames
    b = _ctx(P, scene, 256, 2)
    app_b = P.App(b, accumulate=True, **settings)
    app_b.state.irradianceCachePrepareFrames = 3
    _frames(P, b, app_b, 0, 5, gp)
    ckpt = str(tmp_path / "run.b200pt")
    b.save_state(ckpt)
    saved_app = bytes(app_b.state)                         # the frame driver's state is a plain struct
    b.close()
    for app in (app_a, app_b):
        app.settings.useIrradianceCache = 0
        app.settings.updateGuiding = 1
        app.input_changed()
    _frames(P, a, app_a, 5, 3, gp)
    c = _ctx(P, scene, 256, 2)
    c.load_state(ckpt)
    app_c = P.App(c, accumulate=True)
    import ctypes as C
    C.memmove(C.addressof(app_c.state), saved_app, len(saved_app))
    app_c.settings.useIrradianceCache = 0
    app_c.settings.updateGuiding = 1
    app_c.input_changed()
    _frames(P, c, app_c, 5, 3, gp)
    assert np.array_equal(a.read_image(), c.read_image())
    ha, da, sa = a.ic_get(); hc, dc, sc = c.ic_get()
    assert (ha.nextCacheSlot, ha.nextUpdateSlot) == (hc.nextCacheSlot, hc.nextUpdateSlot) and ha.nextCacheSlot > 5
    assert np.array_equal(da.view(np.uint8), dc.view(np.uint8)) and np.array_equal(sa.view(np.uint8), sc.view(np.uint8))
    assert a.guiding_region_count() == c.guiding_region_count() and a.guiding_region_count() > 4
    assert np.array_equal(a.guiding_aabbs().view(np.uint8), c.guiding_aabbs().view(np.uint8))
    assert np.array_equal(a.guiding_get_vmms().view(np.uint8), c.guiding_get_vmms().view(np.uint8))


def test_checkpoint_errors(tmp_path):
    P = helpers.pt()
    scene = P.Scene(helpers.scene_path("veachMIS"))
    r = _ctx(P, scene, 0, 1)
    with pytest.raises(P.B200ptError):
        r.load_state(str(tmp_path / "missing"))
    bad = tmp_path / "bad"
    bad.write_bytes(b"not a checkpoint")
    with pytest.raises(P.B200ptError):
        r.load_state(str(bad))
    good = str(tmp_path / "good")
    r.save_state(good)
    other = _ctx(P, scene, 0, 2)                           # different guiding_splits
    with pytest.raises(P.B200ptError):
        other.load_state(good)


def test_corrupt_checkpoints_are_rejected_and_leave_the_context_untouched(tmp_path):
    """b200pt_load_state reads and validates the whole file before it uploads anything: a truncated file, a stale format
    version, or out-of-range slots / spawn links / component counts end in an error and the context keeps rendering the
    same frames as before.  b200pt_save_state writes <path>.tmp and renames."""
    import struct
    P = helpers.pt()
    scene = P.Scene(helpers.scene_path("cornell-dielectric"))
    r = _ctx(P, scene, 64, 2)
    for f in range(2):
        r.render_frame(P.default_push_constants(randomUInt=P.tea(f, 3), previousFrames=f, samplesPerPixel=1, enableMIS=1, updateGuiding=1))
        r.guiding_update()
    good = str(tmp_path / "state.ckpt")
    r.save_state(good)
    assert not os.path.exists(good + ".tmp")
    data = bytearray(open(good, "rb").read())
    before_img, before_vmm = r.read_image().copy(), r.guiding_get_vmms().copy()

    def rejected(name, mutate):
        d = bytearray(data)
        mutate(d)
        p = str(tmp_path / name)
        open(p, "wb").write(d)
        with pytest.raises(P.B200ptError):
            r.load_state(p)
        assert np.array_equal(r.read_image(), before_img) and np.array_equal(r.guiding_get_vmms().view(np.uint8), before_vmm.view(np.uint8)), name

    rejected("trunc_images.ckpt", lambda d: d.__delitem__(slice(len(d) // 3, None)))
    rejected("trunc_tail.ckpt", lambda d: d.__delitem__(slice(len(d) - 100, None)))
    rejected("version.ckpt", lambda d: struct.pack_into("<i", d, 8, 1))
    n = r.width * r.height
    ic_hdr = 8 + 20 + 3 * n * 16
    rejected("ic_slot.ckpt", lambda d: struct.pack_into("<I", d, ic_hdr, 1 << 30))                     # nextCacheSlot far beyond ic_size
    g = ic_hdr + 12 + 64 * (56 + 24)                                                                     # guiding section: 6 x int32 header
    regions = struct.unpack_from("<i", data, g + 4)[0]
    assert regions == 4
    spawn = g + 24 + C.sizeof(P.GuidingParams) + regions * 24
    rejected("spawn.ckpt", lambda d: struct.pack_into("<i", d, spawn, 1000))                            # spawn link out of range
    mix0 = spawn + 2 * regions * 4
    rejected("mixK.ckpt", lambda d: struct.pack_into("<i", d, mix0, 99))                                # K = 99 components
    r.load_state(good)                                                                                    # the intact file still loads
