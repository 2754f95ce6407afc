"""GPU parity tests (call through the C ABI):
   level 1 — hit primitive ids bit-exact against the brute-force oracle on identical rays (north_star);
   level 2 — per-pixel radiance of identical RNG streams within 1e-4 relative of the oracle's shader restatement."""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
MISS = 0xFFFFFFFF
NT = os.cpu_count() or 1


@pytest.mark.parametrize("scene_name", ["cornell-dielectric", "veachMIS", "miPhong", "sponzaXML", "test-scene", "envSynthetic", "envMap", "testSpheres", "stackedCards"])
def test_closest_hit_bit_exact(scene_name):
    w, h = 256, 144
    scene, r, o = helpers.make_pair(scene_name, w, h, accel=True)
    prim = helpers.camera_rays(scene, w, h)
    sec = helpers.secondary_rays(o, prim)
    for name, rays in (("primary", prim), ("secondary", sec)):
        g = r.trace_rays(rays)
        c = o.trace_rays(rays, threads=NT)
        assert np.array_equal(g["prim"], c["prim"]), "%s: %d primitive ids differ" % (name, np.sum(g["prim"] != c["prim"]))
        hit = c["prim"] != MISS
        assert hit.mean() > (0.05 if scene_name in ("testSpheres", "envMap", "envSynthetic") else 0.2)
        # t, u, v are computed with identical individually-rounded operations: bit-exact
        for f in ("t", "u", "v"):
            assert np.array_equal(g[f][hit].view(np.uint32), c[f][hit].view(np.uint32)), f


def test_closest_hit_vs_pure_brute_force_subset():
    """Same check against the O(N) brute force (no oracle BVH at all) on a subset."""
    w, h = 96, 54
    scene, r, o = helpers.make_pair("cornell-dielectric", w, h, accel=False)
    prim = helpers.camera_rays(scene, w, h, seed=5)
    rays = np.concatenate([prim[::3], helpers.secondary_rays(o, prim, seed=6)[:2500]])
    g = r.trace_rays(rays)
    c = o.trace_rays(rays, threads=NT)
    assert np.array_equal(g["prim"], c["prim"])
    hit = c["prim"] != MISS
    assert np.array_equal(g["t"][hit], c["t"][hit])


@pytest.mark.parametrize("scene_name", ["cornell-dielectric", "veachMIS", "stackedCards"])
def test_any_hit_matches_oracle(scene_name):
    w, h = 256, 144
    scene, r, o = helpers.make_pair(scene_name, w, h)
    prim = helpers.camera_rays(scene, w, h, seed=3)
    sh = helpers.secondary_rays(o, prim, seed=4, shadow=True)
    g = r.trace_rays(sh, any_hit=True)
    c = o.trace_rays(sh, any_hit=True, threads=NT)
    assert np.array_equal(g["prim"] != MISS, c["prim"] != MISS)
    assert 0.02 < np.mean(c["prim"] != MISS) < 0.98


def test_alpha_cutout_any_hit_shader():
    """raytrace.rahit:22-46: hits on a material whose diffuse texture has transparent texels are ignored stochastically
    (own RNG stream seeded from uv, ray origin, t and the frame seed) — for closest-hit and for shadow rays."""
    P = helpers.pt()
    w, h = 192, 108
    scene, r, o = helpers.make_pair("alphaLeaf", w, h)
    prim = helpers.camera_rays(scene, w, h, seed=9)
    down = prim.copy()                      # a second batch straight down through the card, starting above it
    rng = np.random.default_rng(3)
    down["origin"] = np.stack([rng.uniform(-1.6, 1.6, len(down)), np.full(len(down), 3.0), rng.uniform(-1.6, 1.6, len(down))], -1).astype(np.float32)
    down["dir"] = [0, -1, 0]
    for seed in (0, 0x12345678):           # the batch API uses the seed of the last rendered frame (0 before any)
        if seed:
            pc = P.default_push_constants(randomUInt=seed, previousFrames=0, samplesPerPixel=1)
            r.render_frame(pc)
            o.render_region(pc, x1=1, y1=1)
        for rays in (prim, down):
            for any_hit in (False, True):
                g, c = r.trace_rays(rays, any_hit=any_hit), o.trace_rays(rays, any_hit=any_hit, threads=NT)
                if any_hit:
                    assert np.array_equal(g["prim"] != MISS, c["prim"] != MISS)
                else:
                    assert np.array_equal(g["prim"], c["prim"])
    # the card (prims 2, 3) stops some of the vertical rays and lets most through to the floor (78 % of the texels are clear)
    g = r.trace_rays(down)
    card = np.isin(g["prim"], (2, 3)).mean()
    assert 0.05 < card < 0.5 and np.isin(g["prim"], (0, 1)).mean() > 0.4
    r2, o2, gi, ci = _render_both("alphaLeaf", 160, 90, samplesPerPixel=2, enableNEE=1, enableMIS=1, maxDepth=6)
    _assert_radiance_parity(gi, ci, PARITY)


def test_edge_cases_empty_and_degenerate_rays():
    P = helpers.pt()
    scene, r, o = helpers.make_pair("veachMIS", 16, 16)
    assert r.trace_rays(np.zeros(0, dtype=P.RAY_DTYPE)).shape == (0,)
    rays = np.zeros(6, dtype=P.RAY_DTYPE)
    rays["origin"] = [[0, 2, 15]] * 6
    rays["dir"] = [[0, 0, -1], [0, -1, 0], [1, 0, 0], [0, 0, 1], [0, -0.3, -1], [0, -0.3, -1]]
    rays["tmin"] = 1e-3
    rays["tmax"] = [1e6, 1e6, 1e6, 1e6, 1e6, 0.5]      # last one: interval too short to reach anything
    rays["dir"][4:] /= np.linalg.norm(rays["dir"][4:], axis=1, keepdims=True)
    g = r.trace_rays(rays)
    c = o.trace_rays(rays)
    assert np.array_equal(g["prim"], c["prim"])
    assert g["prim"][5] == MISS


def _render_both(scene_name, w, h, frames=1, **pc_over):
    P = helpers.pt()
    scene, r, o = helpers.make_pair(scene_name, w, h)
    imgs = []
    for f in range(frames):
        pc = P.default_push_constants(randomUInt=P.tea(f, 0xC0FFEE), previousFrames=f, **pc_over)
        r.render_frame(pc)
        o.render_region(pc, threads=NT)
    return r, o, r.read_image()[..., :3].astype(np.float64), o.image()[..., :3].astype(np.float64)


# Both sides evaluate the elementary functions with include/b200pt_detmath.h (bit-identical by construction), so a same-seed
# render follows the same path in every pixel: ALL pixels are held to the north-star tolerance (what remains, ~2e-7, is the
# order in which a pixel's light samples are added up).  Before that header the bar was 85 - 99.5 % of the pixels.
PARITY = 1.0


def _assert_radiance_parity(g, c, min_frac):
    assert np.isfinite(g).all()
    den = np.maximum(np.abs(c), 1e-3)           # 1e-4 relative, with an absolute floor of 1e-7 for black pixels
    rel = np.abs(g - c) / den
    ok = (rel <= 1e-4).all(axis=-1)
    frac = ok.mean()
    assert frac >= min_frac, "only %.4f of pixels within 1e-4 (worst rel %.3g)" % (frac, rel.max())
    # and the images agree in the mean far tighter than any visible difference
    assert abs(g.mean() - c.mean()) <= 2e-3 * c.mean() + 1e-6


@pytest.mark.parametrize("mode", ["nee_mis", "nee", "bsdf"])
def test_radiance_parity_cornell(mode):
    over = dict(nee_mis=dict(enableNEE=1, enableMIS=1), nee=dict(enableNEE=1, enableMIS=0), bsdf=dict(enableNEE=0))[mode]
    r, o, g, c = _render_both("cornell-dielectric", 160, 90, samplesPerPixel=2, **over)
    _assert_radiance_parity(g, c, PARITY)
    # identical paths => identical ray counts.  Shadow rays are only traced when dot(n, lightDir) > 0; for a vertex ON the
    # emitter that samples its own (coplanar) face this cosine is +-1 ulp around 0, so FMA contraction flips the sign for
    # ~0.2% of the NEE events — zero-radiance events, but they show up in the count.
    s, oc = r.stats(), o.counters()
    assert abs(int(s.extend_rays) - oc["extend_rays"]) <= 2e-4 * oc["extend_rays"]
    assert abs(int(s.shadow_rays) - oc["shadow_rays"]) <= 5e-3 * oc["shadow_rays"] + 2


@pytest.mark.parametrize("scene_name,mode", [("veachMIS", "nee_mis"), ("veachMIS", "nee"), ("veachMIS", "bsdf"), ("miPhong", "nee_mis")])
def test_radiance_parity_glossy(scene_name, mode):
    over = dict(nee_mis=dict(enableNEE=1, enableMIS=1), nee=dict(enableNEE=1, enableMIS=0), bsdf=dict(enableNEE=0))[mode]
    r, o, g, c = _render_both(scene_name, 160, 90, samplesPerPixel=2, **over)
    _assert_radiance_parity(g, c, PARITY)


def test_radiance_parity_sponza_textured():
    """66 445 triangles, 10 JPEG textures (bilinear, repeat, sRGB decode), one sphere light of radius 0.1 and radiance
    10 000 seventeen units above the floor.  Light samples of later vertices are ill-conditioned in this scene: the shadow
    ray is aimed at a point of a sphere 170 radii away, the sphere test (raytrace.sphere.rint:13-28) cancels two terms of
    size d^2 = 289 to get (r cos)^2 <= 0.01, so the last bit of the origin decides between "hits the light's own near side
    first" and "unoccluded" for ~1 % of the samples.  With CUDA's libm on one side and glibc's on the other 15 % of the
    pixels differed here; with the shared deterministic functions the origins are bit-identical and every pixel agrees."""
    r, o, g, c = _render_both("sponzaXML", 128, 72, samplesPerPixel=2, enableNEE=0, maxDepth=8)
    _assert_radiance_parity(g, c, PARITY)
    r, o, g, c = _render_both("sponzaXML", 128, 72, samplesPerPixel=2, enableNEE=1, enableMIS=1, maxDepth=0)
    _assert_radiance_parity(g, c, PARITY)
    assert c.mean() > 0
    r, o, g, c = _render_both("sponzaXML", 128, 72, samplesPerPixel=2, enableNEE=1, enableMIS=1, maxDepth=8)
    _assert_radiance_parity(g, c, PARITY)


@pytest.mark.parametrize("mode", ["nee_mis", "nee", "bsdf"])
def test_radiance_parity_json_scene_instances_point_and_mesh_lights(mode):
    """The reference's JSON test scene: 14 instances with rotation and non-uniform scale (normalTransform), two mesh
    lights (per-light face tables, quirk 5) and two point lights, MTL materials (diffuse, Phong, glass, mirror, light),
    PNG + JPEG textures."""
    over = dict(nee_mis=dict(enableNEE=1, enableMIS=1), nee=dict(enableNEE=1, enableMIS=0), bsdf=dict(enableNEE=0))[mode]
    r, o, g, c = _render_both("test-scene", 160, 90, samplesPerPixel=2, maxDepth=8, **over)
    _assert_radiance_parity(g, c, PARITY)
    assert c.mean() > 0


@pytest.mark.parametrize("mode", ["nee_mis", "nee", "bsdf"])
def test_radiance_parity_environment_map(mode):
    """Environment emitter: lat-long lookup on a miss (raytrace.rmiss), uniform-hemisphere light sampling with the stub
    pdf (quirk 7), MIS probe misses; plus a sphere light."""
    over = dict(nee_mis=dict(enableNEE=1, enableMIS=1), nee=dict(enableNEE=1, enableMIS=0), bsdf=dict(enableNEE=0))[mode]
    r, o, g, c = _render_both("envSynthetic", 160, 90, samplesPerPixel=2, maxDepth=6, **over)
    _assert_radiance_parity(g, c, PARITY)
    assert c.mean() > 0.05


@pytest.mark.parametrize("scene_name", ["envMap", "testSpheres"])
def test_radiance_parity_reference_envmap_scenes(scene_name):
    """The reference's own environment-map scenes (PIZ-compressed 512x256 lat-long map): a diffuse ball, and a lone
    conductor sphere — a scene without a single triangle."""
    r, o, g, c = _render_both(scene_name, 160, 90, samplesPerPixel=2, enableNEE=1, enableMIS=1, maxDepth=6)
    _assert_radiance_parity(g, c, PARITY)
    assert c.mean() > 0.05


def test_aov_layer_depth_and_split_statistics():
    """b200pt_read_aovs: maxReachedDepth / depthSum / depthsCounter / nextSplitSlot per pixel (rgen:1653-1655, :1677-1679,
    :1705-1707) — the data of the reference's depth and split debug views — against the oracle."""
    P = helpers.pt()
    scene, r, o = helpers.make_pair("cornell-dielectric", 128, 72)
    with pytest.raises(P.B200ptError):
        r.read_aovs()
    r.set_aovs(True)
    pc = P.default_push_constants(randomUInt=P.tea(2, 0xC0FFEE), previousFrames=0, samplesPerPixel=4, enableMIS=1, splitOnFirst=1)
    r.render_frame(pc)
    o.render_region(pc, threads=NT)
    g, c = r.read_aovs(), o.image(3)
    assert np.array_equal(g[..., 2], c[..., 2])                    # paths per pixel: 4 samples + the drained splits
    assert (g[..., 2] > 4).mean() > 0.5 and np.array_equal(g[..., 3], c[..., 3]) and g[..., 3].max() >= 4
    assert (g[..., 0] == c[..., 0]).mean() >= 0.995 and (g[..., 1] == c[..., 1]).mean() >= 0.995
    assert c[..., 0].max() > 5 and c[..., 1].mean() > c[..., 2].mean()      # depths are real
    r.set_aovs(False)
    r.render_frame(pc)
    with pytest.raises(P.B200ptError):
        r.read_aovs()


def test_accumulation_over_frames_matches_oracle():
    """previousFrames > 0: running mean mix(prev, x, 1/(n+1)) (rgen:1476-1483), and the sum/divide variant."""
    r, o, g, c = _render_both("veachMIS", 96, 54, frames=3, samplesPerPixel=1, enableMIS=1)
    _assert_radiance_parity(g, c, PARITY)
    r, o, g, c = _render_both("veachMIS", 96, 54, frames=3, samplesPerPixel=1, enableMIS=1, enableAverageInsteadOfMix=1)
    _assert_radiance_parity(g, c, PARITY)


def test_numNEE_balance_heuristic_and_depth_limits():
    r, o, g, c = _render_both("cornell-dielectric", 96, 54, samplesPerPixel=1, enableMIS=1, numNEE=3, usePowerHeuristic=0, maxDepth=4, maxFollowDiscrete=1)
    _assert_radiance_parity(g, c, PARITY)


def test_converges_to_reference_image():
    """Level 3 at test scale: a 256-spp 160x90 render against the box-filtered golden (Mitsuba) image of the reference.
    veachMIS has no missing asset; its rough-conductor plates use the reference's own (quirky) BSDF, so the tolerance is
    loose here and the strict relMSE < 1e-3 check is oracle-vs-kernel at equal spp (test_radiance_parity_*)."""
    P = helpers.pt()
    scene, r, o = helpers.make_pair("veachMIS", 160, 90)
    for f in range(16):
        r.render_frame(P.default_push_constants(randomUInt=P.tea(f, 1), previousFrames=f, samplesPerPixel=16, enableMIS=1))
    img = r.read_image()[..., :3].astype(np.float64)
    gold = np.load(os.path.join(helpers.ROOT, "tests", "golden", "veachMIS_160x90.npy")).astype(np.float64)
    assert abs(img.mean() - gold.mean()) < 0.05 * gold.mean()
    rel = ((img - gold) ** 2 / (gold ** 2 + 1e-2)).mean()
    assert rel < 0.05


@pytest.mark.parametrize("scene_name", ["cornell-dielectric", "veachMIS", "miPhong"])
def test_converged_4096spp_independent_streams(scene_name):
    """Level 3 proper (north_star: converged 4096-spp images reach relMSE < 1e-3).  Device and oracle render 4096 spp with
    INDEPENDENT random streams, so nothing but the expectation can make them agree: relMSE(device, oracle) has to sit
    at the Monte-Carlo noise floor, measured as relMSE(device, device with other seeds).  cornell-dielectric is the
    config-2 scene and meets the absolute 1e-3; the rough-conductor / Phong plates of veachMIS and miPhong are noisier
    at 4096 spp (floor ~1.3e-3 at 64x36), there the bound is relative to the floor (tools/converged_check.py)."""
    P = helpers.pt()
    w, h, spp, per = 64, 36, 4096, 64
    scene, r, o = helpers.make_pair(scene_name, w, h)
    over = dict(enableNEE=1, enableMIS=1, samplesPerPixel=per)

    def device(seed):
        for f in range(spp // per):
            r.render_frame(P.default_push_constants(randomUInt=P.tea(f, seed), previousFrames=f, **over))
        return r.read_image()[..., :3].astype(np.float64)

    def rel_mse(a, b):
        return float(((a - b) ** 2 / (b ** 2 + 1e-2)).mean())

    ga, gb = device(0xA11CE), device(0xB0B)
    for f in range(spp // per):
        o.render_region(P.default_push_constants(randomUInt=P.tea(f, 0xC0FFEE), previousFrames=f, **over), threads=NT)
    c = o.image()[..., :3].astype(np.float64)
    floor, err = rel_mse(ga, gb), rel_mse(ga, c)
    assert err < 1.6 * floor + 1e-4, (err, floor)
    assert abs(ga.mean() - c.mean()) < 0.03 * c.mean()
    if scene_name == "cornell-dielectric":
        assert err < 1e-3, err


def test_tie_in_t_resolves_to_lowest_primitive_id():
    """stackedCards: 150 coincident quads (a 300-way tie in t on every hit) — the closest-hit rule is the lexicographic
    minimum of (t, primitive id), whatever order the warp's pooled triangle tests finish in; every node visit hands a ray
    24 leaf triangles, so a warp's candidate list (768) needs several rounds of the cooperative triangle phase."""
    w, h = 192, 108
    scene, r, o = helpers.make_pair("stackedCards", w, h)
    prim = helpers.camera_rays(scene, w, h, seed=5)
    g = r.trace_rays(prim)
    stack = g["prim"] < 300
    assert stack.mean() > 0.15 and set(np.unique(g["prim"][stack])) <= {0, 1}
    c = o.trace_rays(prim, threads=NT)
    assert np.array_equal(g["prim"], c["prim"])
    for f in ("t", "u", "v"):
        assert np.array_equal(g[f][stack].view(np.uint32), c[f][stack].view(np.uint32)), f
