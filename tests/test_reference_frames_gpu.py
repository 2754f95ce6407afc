"""The CUDA path against frames rendered by the REFERENCE'S OWN SHADER SOURCE (compiled as C++, oracle/shader_ref.cpp; the frames are
committed in tests/golden/shader_ref_frames.npz, see tests/test_shader_ref.py for how they are made and what in them is the
reference's).  Same scenes, cameras, push constants and seeds through the C ABI: every pixel within the north-star 1e-4 relative
(the device adds a pixel's light samples with float atomics, so the last bit may differ; everything else is the same arithmetic)."""
import numpy as np
import pytest

import helpers
import test_shader_ref as T

pytestmark = pytest.mark.gpu
W, H = T.W, T.H


def _golden_frame(key):
    g = T._golden()
    assert key in g, key
    return g[key].view(np.float32).reshape(H, W, 4)


def _close(img, ref, what):
    a, b = img[..., :3].astype(np.float64), ref[..., :3].astype(np.float64)
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
    assert np.isfinite(a).all(), what
    assert (rel <= 1e-4).all(), (what, float(rel.max()), float((rel > 1e-4).any(-1).mean()))


GPU_CASES = T.PLAIN[:17]      # the cases that have run on a B200 (profiles/r02c_gpu_calls.txt); later additions of T.PLAIN are oracle-vs-reference only so far


@pytest.mark.parametrize("scene_name,kw", GPU_CASES, ids=[T.frame_key(s, k) for s, k in GPU_CASES])
def test_device_frames_match_the_reference_shaders(scene_name, kw):
    P = helpers.pt()
    scene, r, _ = helpers.make_pair(scene_name, W, H)
    base = dict(samplesPerPixel=2)
    base.update(kw)
    for f in range(2):
        r.render_frame(P.default_push_constants(randomUInt=P.tea(f, 0xC0FFEE), previousFrames=f, **base))
    _close(r.read_image(), _golden_frame(T.frame_key(scene_name, kw)), scene_name)
    _close(r.read_image(P.IMAGE_ACCUM), _golden_frame(T.frame_key(scene_name, kw) + ":accum"), scene_name + " accumulation image")


def test_device_ic_and_adrrs_frames_match_the_reference_shaders():
    """irradiance-cache lookups, the estimate frame, ADRRS window / roulette / splitting and splitOnFirst (same flow as
    tests/test_shader_ref.py: the cache comes from prepare frames of the oracle, the estimate image from the estimate frame)."""
    P = helpers.pt()
    scene, r, o = helpers.make_pair("irradianceCache", W, H, ic_size=512)
    r.ic_put(*T._cache_from_oracle(P, o))
    off = dict(irradianceCreateProb=0.0, irradianceUpdateProb=0.0)
    for i, kw in enumerate(T.IC_CASES):
        r.render_frame(T._pc(P, 20 + i, useIrradianceCache=1, **off, **kw))
        _close(r.read_image(), _golden_frame("ic_lookup_%d" % i), "IC lookup %d" % i)
    r.render_frame(T._pc(P, 30, **T.ESTIMATE_FRAME, **off))
    _close(r.read_image(P.IMAGE_ESTIMATE), _golden_frame("estimate_frame"), "estimate frame")
    for i, kw in enumerate(T.ADRRS_CASES):
        r.render_frame(T._pc(P, 40 + i, **off, **kw))
        _close(r.read_image(), _golden_frame("adrrs_%d" % i), "ADRRS %d" % i)


def test_device_guided_and_training_frames_match_the_reference_shaders():
    import test_guided_tracer_gpu as tg
    P = helpers.pt()
    scene, r, _ = helpers.make_pair("cornell-dielectric", W, H, guiding_splits=4)
    aabbs = r.guiding_aabbs()
    r.guiding_put_vmms(tg._synthetic_vmms(P, aabbs, 9))
    for i, kw in enumerate(T.GUIDING_CASES):
        r.render_frame(T._pc(P, 50 + i, numGuidingRegions=len(aabbs), **kw))
        _close(r.read_image(), _golden_frame("guiding_%d" % i), "guiding %d" % i)
