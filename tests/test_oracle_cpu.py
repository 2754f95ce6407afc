"""CPU: the tracer oracle checked against what the reference pins (RNG constants) and against itself
(BVH-accelerated traversal == brute force, determinism, energy sanity)."""
import numpy as np
import pytest

import helpers


def test_tea_and_lcg_known_answers():
    P = helpers.pt()
    # TEA with 16 rounds and the constants of shaders/random.glsl:13-27 (Zafar et al.); python restatement vs C
    import ctypes as C
    L = helpers.oracle().lib()
    L.oracle_tea.restype = C.c_uint32
    L.oracle_tea.argtypes = [C.c_uint32, C.c_uint32]
    for a, b in [(0, 0), (1, 2), (921599, 0xC0FFEE), (0xFFFFFFFF, 0xFFFFFFFF), (12345, 67890)]:
        assert L.oracle_tea(a, b) == P.tea(a, b)      # two independent restatements (C and Python) agree
    assert len({P.tea(i, 7) for i in range(1000)}) == 1000
    # LCG: Numerical Recipes constants, low 24 bits / 2^24 (random.glsl:31-43)
    s = 1
    s = (1664525 * s + 1013904223) & 0xFFFFFFFF
    assert s == 1015568748
    assert (s & 0xFFFFFF) / float(1 << 24) == pytest.approx(8935788 / 16777216.0, abs=1e-9)


def test_oracle_accel_matches_brute_force():
    scene, _, acc = helpers.make_pair("cornell-dielectric", 48, 27, accel=True, gpu=False)
    _, _, brute = helpers.make_pair("cornell-dielectric", 48, 27, accel=False, gpu=False)
    prim = helpers.camera_rays(scene, 48, 27)
    sec = helpers.secondary_rays(acc, prim)[:1500]
    for rays in (prim, sec):
        a = acc.trace_rays(rays, threads=8)
        b = brute.trace_rays(rays, threads=8)
        assert np.array_equal(a["prim"], b["prim"])
        hit = a["prim"] != 0xFFFFFFFF
        assert np.array_equal(a["t"][hit], b["t"][hit]) and np.array_equal(a["u"][hit], b["u"][hit])
    sh = helpers.secondary_rays(acc, prim, shadow=True)[:1500]
    a = acc.trace_rays(sh, any_hit=True, threads=8)
    b = brute.trace_rays(sh, any_hit=True, threads=8)
    assert np.array_equal(a["prim"] != 0xFFFFFFFF, b["prim"] != 0xFFFFFFFF)
    assert 0.05 < np.mean(a["prim"] != 0xFFFFFFFF) < 0.95


def test_oracle_render_is_deterministic_and_thread_invariant():
    P = helpers.pt()
    _, _, o1 = helpers.make_pair("veachMIS", 32, 18, gpu=False)
    _, _, o2 = helpers.make_pair("veachMIS", 32, 18, gpu=False)
    pc = P.default_push_constants(randomUInt=1234, previousFrames=0, samplesPerPixel=2, enableMIS=1)
    o1.render_region(pc, threads=1)
    o2.render_region(pc, threads=8)
    a, b = o1.image(), o2.image()
    assert np.array_equal(a, b)
    assert np.isfinite(a).all() and a[..., :3].max() > 0
    assert o1.counters() == o2.counters() and o1.counters()["extend_rays"] > 32 * 18 * 2


def test_oracle_cornell_energy_sanity():
    """Closed diffuse box lit by a 15 W/m^2sr emitter: mean radiance is positive, finite and of order 0.1-1."""
    P = helpers.pt()
    _, _, o = helpers.make_pair("cornell-dielectric", 32, 18, gpu=False)
    pc = P.default_push_constants(randomUInt=7, previousFrames=0, samplesPerPixel=4, enableMIS=1)
    o.render_region(pc, threads=8)
    img = o.image()[..., :3]
    assert np.isfinite(img).all()
    assert 0.02 < img.mean() < 2.0
