"""The oracle's restatement of the shader megakernel (oracle/tracer_oracle.cpp) against the REFERENCE'S OWN SHADER SOURCE: raytrace.rgen
and the hit / miss / intersection shaders of /root/reference/shaders compiled as C++ into oracle/_ref/libshader_ref.so (oracle/shader_ref.cpp,
oracle/Makefile: the files go through sed where they lie, nothing is copied into this repository) and run on the CPU pixel by pixel.
What is not the reference's in that build is what the reference leaves to the driver: ray / primitive intersection and texture filtering
come from the oracle's traversal and sampler by callback, and the elementary functions are include/b200pt_detmath.h's (the same the
oracle and the kernels use).  Everything else — camera ray, bounce loop, light sampling, NEE + MIS, the seven BSDFs, irradiance-cache
lookup, ADRRS window / Russian roulette / splitting, guided sampling, sample recording, accumulation — is the reference's code, so a
frame must come out BIT-EQUAL.  SURVEY.md §8(c): this is what pins levels 1-2 of the tracer oracle.  CPU only; where the reference
tree is absent (GPU box) the committed frames of that build (tests/golden/shader_ref_frames.npz) are compared instead."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers

SHADER_REF = os.path.join(helpers.ROOT, "oracle", "_ref", "libshader_ref.so")
GOLDEN = os.path.join(helpers.ROOT, "tests", "golden", "shader_ref_frames.npz")
NT = os.cpu_count() or 1
W, H = 64, 36


class ShaderRef:
    """the compiled reference pipeline, fed from the same scene description, camera and oracle context"""

    def __init__(self, scene, view, proj, oracle, ic_size=0):
        P, O = helpers.pt(), helpers.oracle()
        self.S = C.CDLL(SHADER_REF)
        self.S.shader_ref_image.restype = C.POINTER(C.c_float)
        self.S.shader_ref_samples.restype = C.c_void_p
        self.S.shader_ref_init(W, H, ic_size)
        self.S.shader_ref_set_scene(C.byref(scene.desc))
        f = lambda a: np.ascontiguousarray(a, np.float32).ctypes.data_as(C.POINTER(C.c_float))
        self.S.shader_ref_set_camera(f(view), f(proj), f(P.mat4_inverse(view)), f(P.mat4_inverse(proj)))
        L = O.lib()
        self.S.shader_ref_set_callbacks(oracle._h, C.cast(L.oracle_trace_one, C.c_void_p), C.cast(L.oracle_texture, C.c_void_p))

    def render(self, pc):
        assert self.S.shader_ref_render(C.byref(pc), 0, 0, W, H) == 0

    def image(self, which=0):
        return np.ctypeslib.as_array(self.S.shader_ref_image(which), shape=(H, W, 4)).copy()

    def set_image(self, which, img):
        np.ctypeslib.as_array(self.S.shader_ref_image(which), shape=(H, W, 4))[:] = img

    def samples(self, dtype):
        n = W * H * 16
        buf = (C.c_char * (n * dtype.itemsize)).from_address(self.S.shader_ref_samples())
        return np.frombuffer(buf, dtype=dtype).copy()

    def set_guiding(self, aabbs, vmms):
        a, v = np.ascontiguousarray(aabbs), np.ascontiguousarray(vmms)
        self.S.shader_ref_set_guiding(a.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), len(a))

    def ic_put(self, hdr, data, spheres):
        d, s = np.ascontiguousarray(data), np.ascontiguousarray(spheres)
        self.S.shader_ref_ic_put(C.byref(hdr), d.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), len(d))


def _setup(scene_name, ic_size=0):
    P = helpers.pt()
    scene, _, o = helpers.make_pair(scene_name, W, H, ic_size=ic_size, gpu=False)
    view, proj = scene.camera_matrices(W / H)
    ref = ShaderRef(scene, view, proj, o, ic_size) if os.path.exists(SHADER_REF) else None
    return P, scene, o, ref


def _golden():
    return np.load(GOLDEN) if os.path.exists(GOLDEN) else {}


def _check(key, ours, ref_value):
    """bit-equality with the live build when it exists, and with its committed output"""
    ours = np.ascontiguousarray(ours)
    if ref_value is not None:
        assert np.array_equal(ours.view(np.uint8), np.ascontiguousarray(ref_value).view(np.uint8)), key
    g = _golden()
    assert key in g, "missing golden frame %s: run tests/golden/make_shader_ref_golden.py" % key
    assert np.array_equal(compact(ours), g[key].view(np.uint8).reshape(-1)), key


def compact(a):
    """what the golden file keeps: the bytes of an array, or their SHA-256 when it is large (the recorded sample buffers)"""
    import hashlib
    b = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    return b if b.nbytes <= 200000 else np.frombuffer(hashlib.sha256(b.tobytes()).digest(), np.uint8)


PLAIN = [("cornell-dielectric", dict(enableMIS=1)), ("cornell-dielectric", dict(enableNEE=0)), ("cornell-dielectric", dict(enableMIS=0, numNEE=3)),
         ("veachMIS", dict(enableMIS=1)), ("veachMIS", dict(enableMIS=1, usePowerHeuristic=0)), ("veachMIS", dict(enableNEE=0)), ("veachMIS", dict(enableMIS=0)),
         ("miPhong", dict(enableMIS=1)), ("miPhong", dict(enableNEE=0)),
         ("test-scene", dict(enableMIS=1)), ("test-scene", dict(enableNEE=0, maxDepth=3, maxFollowDiscrete=1)),
         ("envMap", dict(enableMIS=1)), ("envSynthetic", dict(enableMIS=1)), ("testSpheres", dict(enableMIS=1, useVisibleSphereSampling=1)),
         ("sponzaXML", dict(enableMIS=1, samplesPerPixel=1, maxDepth=4)), ("stackedCards", dict(enableMIS=1)), ("alphaLeaf", dict(enableMIS=1)),
         ("veachMIS", dict(enableMIS=1, enableAverageInsteadOfMix=0)), ("miPhong", dict(enableMIS=1, maxDepth=0, maxFollowDiscrete=0))]


def frame_key(scene_name, kw):
    return scene_name + ":" + ",".join("%s=%s" % kv for kv in sorted(kw.items()))


@pytest.mark.parametrize("scene_name,kw", PLAIN, ids=[frame_key(s, k) for s, k in PLAIN])
def test_path_tracing_frames_are_bit_equal_to_the_reference_shaders(scene_name, kw):
    """NEE / MIS / BSDF-only estimators, every material type, point / sphere / mesh / environment lights, textures, instanced JSON
    scenes: two accumulated frames (saveResult's running mean included)."""
    P, scene, o, ref = _setup(scene_name)
    base = dict(samplesPerPixel=2)
    base.update(kw)
    for f in range(2):
        pc = P.default_push_constants(randomUInt=P.tea(f, 0xC0FFEE), previousFrames=f, **base)
        o.render_region(pc, threads=NT)
        if ref:
            ref.render(pc)
    _check(frame_key(scene_name, kw), o.image(), ref.image() if ref else None)
    _check(frame_key(scene_name, kw) + ":accum", o.image(1), ref.image(1) if ref else None)


def _pc(P, frame, **kw):
    base = dict(randomUInt=P.tea(frame, 77), previousFrames=0, samplesPerPixel=2, enableMIS=1)
    base.update(kw)
    return P.default_push_constants(**base)


IC_CASES = (dict(useIrradianceCacheOnGlossy=1), dict(useIrradianceGradients=0), dict(irradianceCachePerformVisibilityCheck=1, irradianceA=0.5),
            dict(useIrradianceCacheOnGlossy=1, irradianceNumNEE=3, numNEE=2))
# RayTracingApp::setEstimateRTSettings (src/RayTracingApp.cpp:1206-1217), at 4 spp
ESTIMATE_FRAME = dict(storeEstimate=1, samplesPerPixel=4, useIrradianceCache=1, useIrradianceCacheOnGlossy=1, enableNEE=1, maxDepth=1, maxFollowDiscrete=10, numNEE=5, useADRRS=0)
ADRRS_CASES = (dict(useADRRS=1, adrrsSplit=1, adrrsS=5.0), dict(useADRRS=1, adrrsSplit=0, adrrsS=2.0), dict(splitOnFirst=1, enableMIS=0),
               dict(useADRRS=1, adrrsSplit=1, adrrsS=5.0, useIrradianceCache=1, useIrradianceCacheOnGlossy=1))
GUIDING_CASES = (dict(useGuiding=1, guidingProb=0.5, useParallaxCompensation=1), dict(useGuiding=1, guidingProb=1.0, useParallaxCompensation=0),
                 dict(updateGuiding=1), dict(updateGuiding=1, useGuiding=1, guidingProb=0.5, enableMIS=0))


def _cache_from_oracle(P, o, frames=4):
    for f in range(frames):       # RayTracingApp::raytrace during the prepare frames (src/RayTracingApp.cpp:1130-1143)
        o.render_region(_pc(P, f, previousFrames=0xFFFFFFFF, useIrradianceCache=1, useIrradianceCacheOnGlossy=1, isIrradiancePrepareFrame=1,
                            irradianceCreateProb=0.02, irradianceUpdateProb=0.005, samplesPerPixel=1), threads=NT)
    return o.ic_get(P)


def test_irradiance_cache_lookup_frames_are_bit_equal():
    """queryIrradianceCache through the reference's intersection + any-hit shaders (raytrace.irradiance.rint / .rahit: weights,
    gradients, visibility check) on a cache of a few hundred entries; entry creation and update are off (the reference races on
    them inside a frame; the oracle's frame semantic for those is its own and tested against the kernels)."""
    P, scene, o, ref = _setup("irradianceCache", ic_size=512)
    cache = _cache_from_oracle(P, o)
    assert cache[0].nextCacheSlot > 100
    if ref:
        ref.ic_put(*cache)
    off = dict(irradianceCreateProb=0.0, irradianceUpdateProb=0.0, useIrradianceCache=1)
    for i, kw in enumerate(IC_CASES):
        pc = _pc(P, 20 + i, **off, **kw)
        o.render_region(pc, threads=NT)
        if ref:
            ref.render(pc)
        _check("ic_lookup_%d" % i, o.image(), ref.image() if ref else None)


def test_adrrs_and_split_frames_are_bit_equal():
    """applyWeightWindow, Russian roulette, expected-value splitting with the split queue drained after the samples (rgen:1684-1708),
    splitOnFirst, and the estimate frame (storeEstimate, setEstimateRTSettings) that feeds the window."""
    P, scene, o, ref = _setup("irradianceCache", ic_size=512)
    cache = _cache_from_oracle(P, o)
    if ref:
        ref.ic_put(*cache)
    off = dict(irradianceCreateProb=0.0, irradianceUpdateProb=0.0)
    est_pc = _pc(P, 30, **ESTIMATE_FRAME, **off)
    o.render_region(est_pc, threads=NT)
    if ref:
        ref.render(est_pc)
    _check("estimate_frame", o.image(P.IMAGE_ESTIMATE), ref.image(2) if ref else None)
    for i, kw in enumerate(ADRRS_CASES):
        pc = _pc(P, 40 + i, **off, **kw)
        o.render_region(pc, threads=NT)
        if ref:
            ref.render(pc)
        _check("adrrs_%d" % i, o.image(), ref.image() if ref else None)


def _regions(P, scene_name, splits):
    O = helpers.oracle()
    if not O.ref_guiding_available():
        pytest.skip("oracle/_ref/libguiding_ref.so (the region tree) is not here")
    mn, mx = helpers.scene_box(scene_name)
    return O.GuidingRef(splits, mn, mx, P.default_guiding_params()).aabbs()


def test_guided_sampling_and_sample_recording_are_bit_equal():
    """getGuidingRegion through raytrace.guiding.rint / .rchit, sampleVMM / VMM pdf mixing in getNewDirection (useGuiding), and the
    recorded DirectionalData of training frames (updateGuiding: updateSamples / commitSamples, distances through specular chains,
    the `sampleOffset += currentSampleOffset` of rgen:1221) — every one of the W*H*16 records byte for byte."""
    import test_guided_tracer_gpu as tg
    P, scene, o, ref = _setup("cornell-dielectric")
    aabbs = _regions(P, "cornell-dielectric", 4)
    vmms = tg._synthetic_vmms(P, aabbs, 9)
    o.set_guiding(aabbs, vmms)
    if ref:
        ref.set_guiding(aabbs, vmms)
    for i, kw in enumerate(GUIDING_CASES):
        pc = _pc(P, 50 + i, numGuidingRegions=len(aabbs), **kw)
        o.render_region(pc, threads=NT)
        if ref:
            ref.render(pc)
        _check("guiding_%d" % i, o.image(), ref.image() if ref else None)
        if kw.get("updateGuiding"):
            ours = o.samples(P.DIRECTIONAL_DATA_DTYPE)
            assert (ours["flags"] != 0xFFFFFFFF).sum() > 0.3 * W * H
            _check("guiding_%d:samples" % i, ours.view(np.uint8), ref.samples(P.DIRECTIONAL_DATA_DTYPE).view(np.uint8) if ref else None)


def test_alpha_test_decisions_equal_the_reference_any_hit_shader():
    """raytrace.rahit (the stochastic alpha test of textured triangles) runs inside the traversal, i.e. on the oracle's side of the
    callback in the frames above — so it is compared on its own: 40 000 candidate hits on the alpha-textured leaf of scenes/alphaLeaf
    (random barycentrics, distances, origins, frame seeds; the texture has transparent, opaque and in-between texels), the oracle's
    accept / reject decision against the compiled shader's ignoreIntersectionEXT."""
    P, scene, o, ref = _setup("alphaLeaf")
    O = helpers.oracle().lib()
    O.oracle_alpha_rejects.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint32]
    rng = np.random.default_rng(21)
    n = 40000
    d = scene.desc
    prims = []                                    # (global primitive, instance, primitive inside the instance), instance order like set_scene
    for i in range(d.num_instances):
        for t in range(d.num_indices[d.instances[i].modelIndex] // 3):
            prims.append((len(prims), i, t))
    pick = rng.integers(0, len(prims), n)
    u = rng.random(n).astype(np.float32)
    v = (rng.random(n) * (1 - u)).astype(np.float32)
    ox = rng.normal(0, 3, n).astype(np.float32)
    t = (10 ** rng.uniform(-2, 2, n)).astype(np.float32)
    seeds = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    ours = np.array([O.oracle_alpha_rejects(o._h, prims[pick[k]][0], u[k], v[k], ox[k], t[k], int(seeds[k])) for k in range(n)], np.uint8)
    assert 0.05 < ours.mean() < 0.95              # both outcomes occur
    ref_value = None
    if ref:
        ref.S.shader_ref_alpha_rejects.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint32]
        ref_value = np.array([ref.S.shader_ref_alpha_rejects(prims[pick[k]][1], prims[pick[k]][2], u[k], v[k], ox[k], t[k], int(seeds[k])) for k in range(n)], np.uint8)
    _check("alpha_test", ours, ref_value)


def _ref_ic_get(P, ref, n=512):
    hdr, data, sp = P.CacheHeader(), np.zeros(n, dtype=P.CACHE_DATA_DTYPE), np.zeros(n, dtype=P.SPHERE_DTYPE)
    ref.S.shader_ref_ic_get(C.byref(hdr), data.ctypes.data_as(C.c_void_p), sp.ctypes.data_as(C.c_void_p), n)
    return hdr, data, sp


def _cache_bytes(c):
    hdr, data, sp = c
    n = min(int(hdr.nextCacheSlot), len(data))
    return np.concatenate([np.array([hdr.nextCacheSlot, hdr.maxCaches, hdr.nextUpdateSlot], np.uint32).view(np.uint8),
                           np.ascontiguousarray(data[:n]).view(np.uint8).reshape(-1), np.ascontiguousarray(sp[:n]).view(np.uint8).reshape(-1)])


def test_irradiance_cache_creation_and_update_are_bit_equal():
    """createIrradianceCache / calculateCacheData (20 x 10 stratified paths per entry, rotational and translational gradients, radius
    clamps) on whole prepare frames: slots are handed out in pixel order on both sides and new entries are not in the lookup structure
    before the next frame, so the caches — header, every entry, every sphere — and the images must be bit-equal.  updateIrradianceCache
    rewrites entries that lookups of the same frame can observe (the reference races there; the oracle reads the frame-start snapshot),
    so updates are compared on one-pixel frames, where the two semantics coincide: ten updates, the cache bit-equal after each."""
    P, scene, o, ref = _setup("irradianceCache", ic_size=512)
    prep = dict(previousFrames=0xFFFFFFFF, useIrradianceCache=1, useIrradianceCacheOnGlossy=1, isIrradiancePrepareFrame=1, samplesPerPixel=1)
    for f in range(3):      # (the third frame with tighter gradient / radius clamps)
        pc = _pc(P, f, irradianceCreateProb=0.02, irradianceUpdateProb=0.0, **prep, **(dict(irradianceGradientsMaxLength=0.05, irradianceCacheMinRadius=0.3) if f == 2 else {}))
        o.render_region(pc, threads=NT)
        if ref:
            ref.render(pc)
        _check("ic_create_%d:cache" % f, _cache_bytes(o.ic_get(P)), _cache_bytes(_ref_ic_get(P, ref)) if ref else None)
        _check("ic_create_%d" % f, o.image(), ref.image() if ref else None)
    assert o.ic_get(P)[0].nextCacheSlot > 60
    rng = np.random.default_rng(5)
    for k in range(10):
        x, y = int(rng.integers(0, W)), int(rng.integers(0, H))
        pc = _pc(P, 100 + k, irradianceCreateProb=0.0, irradianceUpdateProb=1.0, **prep)
        o.render_region(pc, x0=x, y0=y, x1=x + 1, y1=y + 1, threads=1)
        if ref:
            assert ref.S.shader_ref_render(C.byref(pc), x, y, x + 1, y + 1) == 0
        _check("ic_update_%d:cache" % k, _cache_bytes(o.ic_get(P)), _cache_bytes(_ref_ic_get(P, ref)) if ref else None)
    assert o.ic_get(P)[0].nextUpdateSlot == 10


def test_config2_rows_at_full_resolution_are_bit_equal():
    """BASELINE config 2 as the bench renders it (cornell-dielectric 1280x720, NEE + MIS, power heuristic, maxDepth 30, 16 spp): four rows
    spread over the frame — the launch size enters the per-pixel seed and the camera ray."""
    global W, H
    P = helpers.pt()
    old = (W, H)
    try:
        W, H = 1280, 720
        scene, _, o = helpers.make_pair("cornell-dielectric", W, H, gpu=False)
        view, proj = scene.camera_matrices(W / H)
        ref = ShaderRef(scene, view, proj, o) if os.path.exists(SHADER_REF) else None
        pc = P.default_push_constants(randomUInt=P.tea(0, 0xC0FFEE), previousFrames=0, samplesPerPixel=16, enableNEE=1, enableMIS=1, usePowerHeuristic=1, numNEE=1,
                                      maxDepth=30, maxFollowDiscrete=3)
        rows = (0, 241, 478, 719)
        for y in rows:
            o.render_region(pc, 0, y, W, y + 1, threads=NT)
            if ref:
                assert ref.S.shader_ref_render(C.byref(pc), 0, y, W, y + 1) == 0
        ours = np.stack([o.image()[y] for y in rows])
        assert np.abs(ours[..., :3]).sum() > 0
        _check("config2_rows", ours, np.stack([ref.image()[y] for y in rows]) if ref else None)
    finally:
        W, H = old
