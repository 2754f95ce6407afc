"""GPU parity of guiding INSIDE the tracer against the oracle's restatement of the
megakernel: the recorded DirectionalData buffer (updateGuiding), guided direction sampling with given mixtures
(useGuiding), and the closed loop train -> refit on the device -> guided render."""
import os

import numpy as np
import pytest

import guiding_data
import helpers

pytestmark = pytest.mark.gpu
NT = os.cpu_count() or 1
INVALID = 0xFFFFFFFF
W, H = 96, 54
P_MAX_SPLITS = 11          # B200PT_MAX_GUIDING_SPLITS (include/b200pt.h)


def _pair(scene_name, splits):
    P = helpers.pt()
    scene, r, o = helpers.make_pair(scene_name, W, H, guiding_splits=splits)
    o.set_guiding(r.guiding_aabbs(), r.guiding_get_vmms())
    return P, scene, r, o


def compare_recorded_samples(P, r, o, min_valid=0.3):
    """every pixel's 16 DirectionalData slots of the last training frame: device (C ABI) against the oracle"""
    g = r.guiding_get_samples().reshape(H * W, 16)
    c = o.samples(P.DIRECTIONAL_DATA_DTYPE).reshape(H * W, 16)
    gv, cv = g["flags"] != INVALID, c["flags"] != INVALID
    assert cv.sum() > min_valid * H * W                        # the scene does produce samples
    same_slots = (gv == cv).all(axis=1)
    # pixels whose paths diverged at a stochastic branch (1-ulp libm differences) record different slots; the rest must agree
    assert same_slots.mean() >= 0.99, same_slots.mean()
    ok_pixels = np.ones(H * W, dtype=bool)
    frac = {}
    for f, tol in (("position", 1e-4), ("direction", 1e-4)):
        d = np.abs(g[f] - c[f]).max(axis=-1)
        okf = np.where(gv & cv, d <= tol, True).all(axis=1)
        frac[f] = okf.mean(); ok_pixels &= okf
    for f in ("pdf", "weight", "distance"):
        with np.errstate(invalid="ignore"):
            d = np.abs(g[f] - c[f]) / np.maximum(np.abs(c[f]), 1e-3)
        okf = np.where(gv & cv, (d <= 1e-4) | (g[f] == c[f]), True).all(axis=1)
        frac[f] = okf.mean(); ok_pixels &= okf
    okf = np.where(gv & cv, g["flags"] == c["flags"], True).all(axis=1)
    frac["flags"] = okf.mean(); ok_pixels &= okf
    # position / direction / pdf / distance / region agree in > 99.5 % of the pixels; the weight of a sample is the
    # radiance the REST of its path collected divided by its pdf — a per-path quantity without any pixel averaging —
    # so last-bit libm differences at later vertices show up at the 1e-5..1e-3 level in ~1 % of the pixels
    assert min(frac[f] for f in ("position", "direction", "pdf", "distance", "flags")) >= 0.995, frac
    assert (ok_pixels & same_slots).mean() >= 0.975, (frac, same_slots.mean())
    with np.errstate(invalid="ignore"):
        dw = np.abs(g["weight"] - c["weight"]) / np.maximum(np.abs(c["weight"]), 1e-3)
    both = gv & cv & same_slots[:, None]
    assert np.mean(((dw <= 1e-3) | (g["weight"] == c["weight"]))[both]) >= 0.99


@pytest.mark.parametrize("scene_name,spp", [("cornell-dielectric", 2), ("veachMIS", 2)])
def test_recorded_samples_match_oracle(scene_name, spp):
    """updateGuiding = 1 (training frame, unguided): every pixel's 16 DirectionalData slots against the oracle."""
    P, scene, r, o = _pair(scene_name, 4)
    pc = P.default_push_constants(randomUInt=P.tea(0, 0xC0FFEE), previousFrames=0, samplesPerPixel=spp, enableMIS=1, updateGuiding=1)
    r.render_frame(pc)
    o.render_region(pc, threads=NT)
    compare_recorded_samples(P, r, o)
    # radiance of the training frame is the ordinary unguided frame
    gi, ci = r.read_image()[..., :3].astype(np.float64), o.image()[..., :3].astype(np.float64)
    rel = np.abs(gi - ci) / np.maximum(np.abs(ci), 1e-3)
    assert (rel <= 1e-4).all(axis=-1).mean() >= 0.99


def test_recorded_regions_with_the_deepest_tree():
    """GUIDING_SPLITS = 11 (B200PT_MAX_GUIDING_SPLITS, 2048 regions): the region every recorded sample lands in against the
    oracle's point-in-box search, and one refit of those samples (sort by 2048 keys, plan, fit) leaves valid mixtures."""
    P, scene, r, o = _pair("cornell-dielectric", P_MAX_SPLITS)
    assert len(r.guiding_aabbs()) == 1 << P_MAX_SPLITS
    pc = P.default_push_constants(randomUInt=P.tea(4, 0xC0FFEE), previousFrames=0, samplesPerPixel=2, enableMIS=1, updateGuiding=1)
    r.render_frame(pc)
    o.render_region(pc, threads=NT)
    compare_recorded_samples(P, r, o)
    g = r.guiding_get_samples()
    valid = g["flags"] != INVALID
    assert valid.any() and g["flags"][valid].max() < (1 << P_MAX_SPLITS) and len(np.unique(g["flags"][valid])) > 200
    r.guiding_update(P.default_guiding_params())
    vm = r.guiding_get_vmms()
    K = vm["usedDistributions"]
    assert (K >= 1).all() and (K <= 16).all()
    for i in np.unique(g["flags"][valid])[:64]:
        assert abs(float(vm[i]["pi"][:K[i]].sum()) - 1.0) < 1e-4 and np.isfinite(vm[i]["thetas"]["k"][:K[i]]).all()
    with pytest.raises(P.B200ptError):
        P.Renderer(16, 16, 0, P_MAX_SPLITS + 1)


def _synthetic_vmms(P, aabbs, seed):
    """mixtures with 1..16 lobes, some kappa == 0 lobes and some with a parallax target, to exercise every branch"""
    rng = np.random.default_rng(seed)
    v = np.zeros(len(aabbs), dtype=P.VMM_THETA_DTYPE)
    for i in range(len(aabbs)):
        K = int(rng.integers(1, 17))
        pi = rng.dirichlet(np.ones(K)).astype(np.float32)
        mu = rng.normal(size=(K, 3)); mu /= np.linalg.norm(mu, axis=1, keepdims=True)
        k = (10 ** rng.uniform(-0.5, 3.0, K)).astype(np.float32)
        k[rng.random(K) < 0.15] = 0.0
        mean = 0.5 * (aabbs["min"][i] + aabbs["max"][i])
        dist = np.where(rng.random(K) < 0.5, rng.uniform(0.3, 3.0, K), -1.0).astype(np.float32)
        v[i]["usedDistributions"] = K
        v[i]["pi"][:K] = pi
        v[i]["meanPosition"] = mean
        v[i]["thetas"]["mu"][:K] = mu.astype(np.float32)
        v[i]["thetas"]["k"][:K] = k
        with np.errstate(divide="ignore", invalid="ignore"):
            v[i]["thetas"]["norm"][:K] = (k.astype(np.float64) / (2 * np.pi * (1 - np.exp(-2.0 * k)))).astype(np.float32)
        v[i]["thetas"]["eMin2K"][:K] = np.exp(-2.0 * k.astype(np.float64)).astype(np.float32)
        v[i]["thetas"]["distance"][:K] = dist
        v[i]["thetas"]["target"][:K] = (mean + dist[:, None] * mu).astype(np.float32)
    return v


@pytest.mark.parametrize("scene_name,over", [("cornell-dielectric", dict(useParallaxCompensation=1)), ("cornell-dielectric", dict(useParallaxCompensation=0, guidingProb=0.3)),
                                             ("veachMIS", dict(useParallaxCompensation=1, enableMIS=1))])
def test_guided_sampling_matches_oracle(scene_name, over):
    P, scene, r, o = _pair(scene_name, 5)
    vm = _synthetic_vmms(P, r.guiding_aabbs(), 9)
    r.guiding_put_vmms(vm)
    o.set_guiding(r.guiding_aabbs(), vm)
    pc = P.default_push_constants(randomUInt=P.tea(3, 0xC0FFEE), previousFrames=0, samplesPerPixel=2, useGuiding=1, **over)
    r.render_frame(pc)
    o.render_region(pc, threads=NT)
    gi, ci = r.read_image()[..., :3].astype(np.float64), o.image()[..., :3].astype(np.float64)
    assert np.isfinite(gi).all()
    rel = np.abs(gi - ci) / np.maximum(np.abs(ci), 1e-3)
    assert (rel <= 1e-4).all(axis=-1).mean() >= 0.985, (rel <= 1e-4).all(axis=-1).mean()
    assert abs(gi.mean() - ci.mean()) <= 5e-3 * ci.mean() + 1e-6
    s, oc = r.stats(), o.counters()
    assert abs(int(s.extend_rays) - oc["extend_rays"]) <= 2e-3 * oc["extend_rays"]


def test_guided_and_recording_together_match_oracle():
    """later training frames run with useGuiding and updateGuiding both on (RayTracingApp.cpp:126-141)"""
    P, scene, r, o = _pair("cornell-dielectric", 4)
    vm = _synthetic_vmms(P, r.guiding_aabbs(), 4)
    r.guiding_put_vmms(vm)
    o.set_guiding(r.guiding_aabbs(), vm)
    pc = P.default_push_constants(randomUInt=P.tea(1, 7), previousFrames=0, samplesPerPixel=2, enableMIS=1, useGuiding=1, updateGuiding=1)
    r.render_frame(pc)
    o.render_region(pc, threads=NT)
    g = r.guiding_get_samples().reshape(H * W, 16)
    c = o.samples(P.DIRECTIONAL_DATA_DTYPE).reshape(H * W, 16)
    gv, cv = g["flags"] != INVALID, c["flags"] != INVALID
    same = (gv == cv).all(axis=1)
    assert same.mean() >= 0.98
    both = gv & cv & same[:, None]
    with np.errstate(invalid="ignore"):
        d = np.abs(g["weight"] - c["weight"]) / np.maximum(np.abs(c["weight"]), 1e-3)
    assert np.where(both, (d <= 1e-4) | (g["weight"] == c["weight"]), True).all(axis=1).mean() >= 0.98


def test_region_lookup_agrees_with_exhaustive_scan():
    """getGuidingRegion through the recorded flags: every recorded position lies inside the AABB of its region, and it
    is the lowest-index region containing it (points on shared faces)."""
    P, scene, r, o = _pair("cornell-dielectric", 8)
    pc = P.default_push_constants(randomUInt=P.tea(0, 5), previousFrames=0, samplesPerPixel=1, updateGuiding=1)
    r.render_frame(pc)
    s = r.guiding_get_samples()
    s = s[s["flags"] != INVALID]
    assert len(s) > 1000
    aabbs = r.guiding_aabbs()
    lo, hi = aabbs["min"][s["flags"]], aabbs["max"][s["flags"]]
    assert np.all((s["position"] >= lo) & (s["position"] <= hi))
    sub = s[:: max(1, len(s) // 3000)]
    inside = np.all((sub["position"][:, None, :] >= aabbs["min"][None]) & (sub["position"][:, None, :] <= aabbs["max"][None]), axis=-1)
    assert np.array_equal(np.argmax(inside, axis=1), sub["flags"])


def test_closed_loop_training_then_guided_render_is_unbiased():
    """6 training frames with a device refit after each (numGuidingOptimizations = 6, RayTracingApp.h:205), then guided
    rendering: the guided image converges to the same mean as the unguided one."""
    P = helpers.pt()
    w, h = 128, 72
    scene, r, o = helpers.make_pair("cornell-dielectric", w, h, guiding_splits=6)
    for f in range(6):
        pc = P.default_push_constants(randomUInt=P.tea(f, 11), previousFrames=0, samplesPerPixel=4, enableMIS=1, updateGuiding=1, useGuiding=int(f > 0))
        r.render_frame(pc)
        r.guiding_update()
    st = r.stats()
    assert st.guiding_samples > 0 and st.guiding_regions_fit > 0
    vm = r.guiding_get_vmms()
    used = vm["usedDistributions"]
    assert used.min() >= 1 and used.max() <= 16
    imgs = {}
    for mode, guided in (("plain", 0), ("guided", 1)):
        for f in range(24):
            r.render_frame(P.default_push_constants(randomUInt=P.tea(100 + f, 3), previousFrames=f, samplesPerPixel=16, enableMIS=1, useGuiding=guided))
        imgs[mode] = r.read_image()[..., :3].astype(np.float64)
    assert np.isfinite(imgs["guided"]).all()
    assert abs(imgs["guided"].mean() - imgs["plain"].mean()) <= 0.02 * imgs["plain"].mean()
    # block-averaged images agree (unbiasedness region by region, not only in the global mean)
    blk = lambda a: a.reshape(h // 8, 8, w // 8, 8, 3).mean(axis=(1, 3))
    rel = np.abs(blk(imgs["guided"]) - blk(imgs["plain"])) / (blk(imgs["plain"]) + 1e-2)
    assert np.median(rel) < 0.03 and rel.mean() < 0.06


def test_region_lookup_after_adaptive_splits_matches_oracle():
    """After PathGuiding::splitRegion the regions are no longer the leaves of the halving tree: the tracer's lookup
    walks the tree to a base leaf and then through the regions cut off it; the oracle scans all boxes linearly."""
    P, scene, r, o = _pair("cornell-dielectric", 3)
    gp = P.default_guiding_params(splitRegions=1, samplesForRegionSplit=1500.0)
    for rnd in range(3):          # three refits on synthetic samples: 8 -> 16 -> ... regions
        aabbs = r.guiding_aabbs()
        r.guiding_update_host(guiding_data.make_batch(aabbs, [2000 if i % 2 == 0 else 900 for i in range(len(aabbs))], 50 + rnd), gp)
    aabbs = r.guiding_aabbs()
    assert len(aabbs) > 12
    r.guiding_put_vmms(_synthetic_vmms(P, aabbs, 11))
    o.set_guiding(aabbs, r.guiding_get_vmms())
    # training frame: the recorded region ids come from the lookup
    pc = P.default_push_constants(randomUInt=P.tea(3, 0xC0FFEE), previousFrames=0, samplesPerPixel=2, enableMIS=1, updateGuiding=1)
    r.render_frame(pc)
    o.render_region(pc, threads=NT)
    g = r.guiding_get_samples().reshape(H * W, 16)
    c = o.samples(P.DIRECTIONAL_DATA_DTYPE).reshape(H * W, 16)
    gv, cv = g["flags"] != INVALID, c["flags"] != INVALID
    same = (gv == cv).all(axis=1)
    assert same.mean() >= 0.99
    assert np.where(gv & cv, g["flags"] == c["flags"], True).all(axis=1)[same].mean() >= 0.999
    assert len(np.unique(c["flags"][cv])) > 8 and c["flags"][cv].max() >= 8          # spawned regions are in use
    # guided frame: sampling from the mixtures of the looked-up regions
    pc = P.default_push_constants(randomUInt=P.tea(4, 0xC0FFEE), previousFrames=0, samplesPerPixel=2, enableMIS=1, useGuiding=1, guidingProb=0.5)
    r.render_frame(pc)
    o.render_region(pc, threads=NT)
    gi, ci = r.read_image()[..., :3].astype(np.float64), o.image()[..., :3].astype(np.float64)
    rel = np.abs(gi - ci) / np.maximum(np.abs(ci), 1e-3)
    assert (rel <= 1e-4).all(axis=-1).mean() >= 0.985
