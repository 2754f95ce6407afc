"""GPU parity of the irradiance cache and ADRRS against the oracle, through the C ABI.

Both sides use the race-free frame semantic described in oracle/tracer_oracle.cpp: lookups read the frame-start cache,
update / create slots are handed out in pixel order.  Cache entries are matched by their position (a hit point; the
traversal is bit-exact) and line up slot by slot; values are compared within 1e-4 relative (one entry sums 200 paths; both
sides use the elementary functions of include/b200pt_detmath.h, so the paths are the same and only summation order differs)."""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
NT = os.cpu_count() or 1
W, H = 96, 54
IC_SIZE = 512
SCENE = "irradianceCache"


def _pc(P, frame, **kw):
    base = dict(randomUInt=P.tea(frame, 77), previousFrames=0, samplesPerPixel=1, enableMIS=1)
    base.update(kw)
    return P.default_push_constants(**base)


def _prepare_pc(P, frame, **kw):
    # RayTracingApp::raytrace during the prepare frames (src/RayTracingApp.cpp:1130-1143)
    return _pc(P, frame, previousFrames=0xFFFFFFFF, useIrradianceCache=1, useIrradianceCacheOnGlossy=1, isIrradiancePrepareFrame=1,
               irradianceCreateProb=0.02, irradianceUpdateProb=0.005, **kw)


def _estimate_pc(P, frame):
    # RayTracingApp::setEstimateRTSettings (src/RayTracingApp.cpp:1206-1217)
    return _pc(P, frame, storeEstimate=1, samplesPerPixel=16, useIrradianceCache=1, useIrradianceCacheOnGlossy=1, enableNEE=1, maxDepth=1,
               maxFollowDiscrete=10, numNEE=5, useADRRS=0, irradianceCreateProb=0.0, irradianceUpdateProb=0.0)


def _compare_caches(P, dev, ref, what):
    _compare_caches_sized(P, dev, ref, IC_SIZE, what)


def _compare_caches_sized(P, dev, ref, ic_size, what):
    (hd, dd, sd), (hr, dr, sr) = dev, ref
    nd, nr = min(hd.nextCacheSlot, ic_size), min(hr.nextCacheSlot, ic_size)
    assert nr > 20, what
    assert abs(int(nd) - int(nr)) <= max(2, nr // 50), (what, nd, nr)
    # Slots are handed out in pixel order on both sides and the elementary functions are shared (b200pt_detmath.h), so the
    # caches line up entry by entry: same count, bit-identical centres and normals, and the values (sums over the 200 paths
    # of an entry) within the north-star 1e-4 — gradients within 1e-3 of the entry's scale (differences of nearly equal sums).
    assert nd == nr, (what, nd, nr)
    n = nr
    assert np.array_equal(sd["center"][:n], sr["center"][:n]) and np.array_equal(dd["normal"][:n], dr["normal"][:n]), what
    assert np.array_equal(dd["numUpdates"][:n], dr["numUpdates"][:n]), what
    scale = np.maximum(np.abs(dr["color"][:n]).max(axis=1), 1e-6)
    assert (np.abs(dd["color"][:n] - dr["color"][:n]).max(axis=1) <= 1e-4 * scale).all(), what
    assert (np.abs(dd["harmonicR"][:n] - dr["harmonicR"][:n]) <= 1e-4 * dr["harmonicR"][:n]).all(), what
    assert (np.abs(sd["radius"][:n] - sr["radius"][:n]) <= 1e-4 * sr["radius"][:n]).all(), what
    for g in ("rotGrad", "transGrad"):
        gs = np.maximum(np.maximum(np.abs(dr[g][:n]).max(axis=1), 1e-2 * scale), 1e-6)
        assert (np.abs(dd[g][:n] - dr[g][:n]).max(axis=1) <= 1e-3 * gs).all(), (what, g)


def _images_close(img, ref, what, frac_needed=1.0, mean_tol=1e-4):
    img, ref = img[..., :3].astype(np.float64), ref[..., :3].astype(np.float64)
    assert np.isfinite(img).all(), what
    rel = np.abs(img - ref) / np.maximum(np.abs(ref), 1e-3)
    frac = float((rel <= 1e-4).all(axis=-1).mean())       # north_star: per-path radiance within 1e-4 relative
    assert frac >= frac_needed, (what, frac)
    assert abs(img.mean() - ref.mean()) <= mean_tol * max(ref.mean(), 1e-6), (what, img.mean(), ref.mean())


def _build_cache(P, o, frames=4):
    """Reference cache for the lookup tests: a few prepare frames on the oracle."""
    for f in range(frames):
        o.render_region(_prepare_pc(P, f), threads=NT)
    return o.ic_get(P)


def test_cache_creation_and_update_match_oracle():
    P = helpers.pt()
    scene, r, o = helpers.make_pair(SCENE, W, H, ic_size=IC_SIZE)
    for f in range(3):
        pc = _prepare_pc(P, f)
        r.render_frame(pc)
        o.render_region(pc, threads=NT)
        dev, ref = r.ic_get(), o.ic_get(P)
        if f > 0:      # frame 0 only creates (empty cache: nothing to update)
            assert dev[0].nextUpdateSlot == ref[0].nextUpdateSlot, f
            assert (ref[1]["numUpdates"][:ref[0].nextCacheSlot] > 1).sum() > 0
        _compare_caches(P, dev, ref, "prepare frame %d" % f)
        r.ic_put(*ref)      # continue from identical caches
    st = r.stats()
    assert st.extend_rays > 0 and st.shadow_rays > 0


@pytest.mark.parametrize("kw", [dict(), dict(useIrradianceGradients=1), dict(irradianceCachePerformVisibilityCheck=1, useIrradianceCacheOnGlossy=1)])
def test_cache_lookup_render_matches_oracle(kw):
    P = helpers.pt()
    scene, r, o = helpers.make_pair(SCENE, W, H, ic_size=IC_SIZE)
    cache = _build_cache(P, o)
    r.ic_put(*cache)
    pc = _pc(P, 11, samplesPerPixel=2, useIrradianceCache=1, irradianceCreateProb=0.0, irradianceUpdateProb=0.0, **kw)
    r.render_frame(pc)
    o.render_region(pc, threads=NT)
    _images_close(r.read_image(), o.image(), "IC lookup %r" % (kw,))
    # nothing was created or updated
    hdr, data, spheres = r.ic_get()
    assert hdr.nextCacheSlot == cache[0].nextCacheSlot and np.array_equal(data.view(np.uint8), cache[1].view(np.uint8))


def test_estimate_frame_and_adrrs_match_oracle():
    P = helpers.pt()
    scene, r, o = helpers.make_pair(SCENE, W, H, ic_size=IC_SIZE)
    cache = _build_cache(P, o)
    r.ic_put(*cache)
    # estimate frame: 16 spp, depth 1, IC + 5 light samples, stored into the estimate image
    pc = _estimate_pc(P, 20)
    r.render_frame(pc)
    o.render_region(pc, threads=NT)
    est_ref = o.image(P.IMAGE_ESTIMATE)
    _images_close(r.read_image(P.IMAGE_ESTIMATE), est_ref, "estimate frame")
    r.write_image(P.IMAGE_ESTIMATE, est_ref)      # identical adjoint inputs for the ADRRS frames
    for f, kw in enumerate([dict(adrrsSplit=1), dict(adrrsSplit=0), dict(adrrsSplit=1, enableMIS=0, numNEE=2)]):
        pc = _pc(P, 30 + f, samplesPerPixel=2, useADRRS=1, adrrsS=5.0, irradianceCreateProb=0.0, irradianceUpdateProb=0.0, **kw)
        r.render_frame(pc)
        o.render_region(pc, threads=NT)
        _images_close(r.read_image(), o.image(), "ADRRS %r" % (kw,))


def test_adrrs_frames_keep_creating_cache_entries():
    P = helpers.pt()
    scene, r, o = helpers.make_pair(SCENE, W, H, ic_size=IC_SIZE)
    cache = _build_cache(P, o, frames=1)
    r.ic_put(*cache)
    est = np.full((H, W, 4), 0.5, np.float32)
    r.write_image(P.IMAGE_ESTIMATE, est)
    o.set_image(P.IMAGE_ESTIMATE, est)
    pc = _pc(P, 40, samplesPerPixel=1, useADRRS=1, adrrsSplit=1, irradianceCreateProb=0.05)
    r.render_frame(pc)
    o.render_region(pc, threads=NT)
    dev, ref = r.ic_get(), o.ic_get(P)
    assert ref[0].nextCacheSlot > cache[0].nextCacheSlot
    _compare_caches(P, dev, ref, "creation during an ADRRS frame")


def test_split_on_first_matches_oracle():
    P = helpers.pt()
    scene, r, o = helpers.make_pair(SCENE, W, H, ic_size=0)
    pc = _pc(P, 50, samplesPerPixel=2, splitOnFirst=1)
    r.render_frame(pc)
    o.render_region(pc, threads=NT)
    _images_close(r.read_image(), o.image(), "splitOnFirst")


def test_cache_fills_up_to_capacity_and_stops():
    P = helpers.pt()
    small = 40
    scene = P.Scene(helpers.scene_path(SCENE))
    view, proj = scene.camera_matrices(W / H)
    r = P.Renderer(W, H, small, 0)
    r.set_scene(scene)
    r.set_camera(view, proj)
    for f in range(2):
        r.render_frame(_prepare_pc(P, f, ))
    hdr, data, spheres = r.ic_get()
    assert hdr.nextCacheSlot == small + 1 and hdr.maxCaches == small      # the reference's off-by-one (quirk 11)
    assert (data["numUpdates"] >= 1).all() and (spheres["radius"] > 0).all()


def test_unsupported_modes_fail_loudly():
    P = helpers.pt()
    scene, r, o = helpers.make_pair(SCENE, 32, 18, ic_size=0)
    for kw in (dict(useIrradianceCache=1), dict(useADRRS=1), dict(showIrradianceCacheOnly=1), dict(visualizeMode=3)):
        with pytest.raises(P.B200ptError):
            r.render_frame(_pc(P, 0, **kw))


def test_guiding_training_returns_early_while_splitting():
    """rgen:1643-1650: with updateGuiding and a split mode the raygen only invalidates the pixel's samples."""
    P = helpers.pt()
    scene = P.Scene(helpers.scene_path(SCENE))
    view, proj = scene.camera_matrices(32 / 18)
    r = P.Renderer(32, 18, 0, 2)
    r.set_scene(scene)
    r.set_camera(view, proj)
    r.render_frame(_pc(P, 0))
    before = r.read_image()
    r.render_frame(_pc(P, 1, updateGuiding=1, splitOnFirst=1, previousFrames=1))
    assert np.array_equal(before, r.read_image())
    assert (r.guiding_get_samples()["flags"] == P.INVALID_REGION).all()


def test_lookup_order_and_regen_kernel_do_not_change_the_result(monkeypatch):
    """IC / ADRRS frames sort the path queue by grid cell for the lookups (k_icq_*), let k_shade walk it in that order and start
    the next paths of finished pixels in k_regen.  All three only decide WHICH thread serves which pixel: with the knobs off
    (queue order, next paths inside k_shade — the round-1 kernels) the cache and the lookup frame must be bit-identical, the
    frames of the split modes identical up to the order of a pixel's float atomics."""
    P = helpers.pt()
    scene = P.Scene(helpers.scene_path(SCENE))
    view, proj = scene.camera_matrices(W / H)

    def run(env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        r = P.Renderer(W, H, IC_SIZE, 0)          # the knobs are read when the context is created
        r.set_scene(scene)
        r.set_camera(view, proj)
        out = []
        for f in range(3):
            r.render_frame(_prepare_pc(P, f))
        out.append(r.ic_get()[1].view(np.uint8).copy())
        r.render_frame(_pc(P, 9, samplesPerPixel=2, useIrradianceCache=1, useIrradianceCacheOnGlossy=1, irradianceCreateProb=0.0, irradianceUpdateProb=0.0))
        out.append(r.read_image().copy())
        est = np.full((H, W, 4), 0.5, np.float32)
        r.write_image(P.IMAGE_ESTIMATE, est)
        r.render_frame(_pc(P, 10, samplesPerPixel=2, useADRRS=1, adrrsSplit=1, adrrsS=5.0, irradianceCreateProb=0.0, irradianceUpdateProb=0.0, enableMIS=0))
        out.append(r.read_image().copy())
        r.render_frame(_pc(P, 11, samplesPerPixel=2, splitOnFirst=1, enableMIS=0))
        out.append(r.read_image().copy())
        r.close()
        return out

    on = dict(B200PT_ICQ_SORT="1", B200PT_REGEN_SPLIT="1", B200PT_SHADE_SORTED="1", B200PT_IC_BUILD_GROUP="1", B200PT_IC_OVERLAP="1")
    new = run(on)
    # (the last two: cache entries built by one lane each instead of an eight-lane group sharing every ray; the cache update in
    # front of the frame's paths instead of beside them on a second stream)
    for env in (dict(on, B200PT_ICQ_SORT="0", B200PT_REGEN_SPLIT="0"), dict(on, B200PT_REGEN_SPLIT="0", B200PT_SHADE_SORTED="0"),
                dict(on, B200PT_ICQ_SORT="0"), dict(on, B200PT_IC_BUILD_GROUP="0"), dict(on, B200PT_IC_OVERLAP="0")):
        old = run(env)
        for i, (a, b) in enumerate(zip(new, old)):
            if i < 2:       # cache + lookup frame: one light sample per pixel and iteration, every sum has a fixed order
                assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), (env, i)
            else:           # split modes: a pixel can have two light samples in one trace pass, their float atomics commute only to the last bit
                assert np.allclose(a, b, rtol=2e-6, atol=1e-7), (env, i, float(np.abs(a - b).max()))


def test_guided_sampling_inside_ic_and_adrrs_frames_matches_oracle():
    """useGuiding together with the irradiance cache / ADRRS + splitting: the guided variants of the IC kernels
    (k_shade<GUIDE, IC> and k_regen<GUIDE>: region lookup + mixture sampling for the new direction of a path and of a drained
    split) against the oracle's megakernel with the same mixtures."""
    import test_guided_tracer_gpu as tg
    P = helpers.pt()
    scene, r, o = helpers.make_pair(SCENE, W, H, ic_size=IC_SIZE, guiding_splits=3)
    cache = _build_cache(P, o)
    r.ic_put(*cache)
    est = np.full((H, W, 4), 0.5, np.float32)
    r.write_image(P.IMAGE_ESTIMATE, est)
    o.set_image(P.IMAGE_ESTIMATE, est)
    vm = tg._synthetic_vmms(P, r.guiding_aabbs(), 9)
    r.guiding_put_vmms(vm)
    o.set_guiding(r.guiding_aabbs(), vm)
    frames = [dict(useADRRS=1, adrrsSplit=1, adrrsS=5.0), dict(useIrradianceCache=1, useIrradianceCacheOnGlossy=1), dict(splitOnFirst=1)]
    for f, kw in enumerate(frames):
        pc = _pc(P, 60 + f, samplesPerPixel=2, useGuiding=1, guidingProb=0.5, useParallaxCompensation=1, irradianceCreateProb=0.0, irradianceUpdateProb=0.0, **kw)
        r.render_frame(pc)
        o.render_region(pc, threads=NT)
        _images_close(r.read_image(), o.image(), "guided %r" % (kw,), frac_needed=0.985, mean_tol=5e-3)


def test_guiding_training_on_cache_frames_matches_oracle():
    """updateGuiding on frames that use the irradiance cache, or ADRRS without splitting (the reference allows both,
    rgen:1642-1650 only returns early for the split modes): the recorded DirectionalData of every pixel against the oracle.
    A path that ends on a cache record adds the record and its light samples to `result` without updateSamples()
    (rgen:1117-1124), so its earlier samples keep the light they had collected until then; Russian roulette ends a path
    the same way.  The third frame also creates and updates cache entries beside the training paths."""
    import test_guided_tracer_gpu as tg
    P = helpers.pt()
    scene, r, o = helpers.make_pair(SCENE, W, H, ic_size=IC_SIZE, guiding_splits=3)
    cache = _build_cache(P, o)
    r.ic_put(*cache)
    est = np.full((H, W, 4), 0.5, np.float32)
    r.write_image(P.IMAGE_ESTIMATE, est)
    o.set_image(P.IMAGE_ESTIMATE, est)
    o.set_guiding(r.guiding_aabbs(), r.guiding_get_vmms())
    off = dict(irradianceCreateProb=0.0, irradianceUpdateProb=0.0)
    frames = [(dict(useIrradianceCache=1, useIrradianceCacheOnGlossy=1, **off), 0.02),
              (dict(useADRRS=1, adrrsSplit=0, adrrsS=5.0, **off), 0.1),
              (dict(useIrradianceCache=1, useIrradianceCacheOnGlossy=1, irradianceCreateProb=0.02, irradianceUpdateProb=0.005), 0.02)]
    for f, (kw, min_valid) in enumerate(frames):
        pc = _pc(P, 80 + f, samplesPerPixel=2, updateGuiding=1, **kw)
        r.render_frame(pc)
        o.render_region(pc, threads=NT)
        tg.compare_recorded_samples(P, r, o, min_valid=min_valid)
        _images_close(r.read_image(), o.image(), "training %r" % (kw,), frac_needed=0.985, mean_tol=5e-3)
    _compare_caches(P, r.ic_get(), o.ic_get(P), "cache after a training frame that builds entries")
