#!/usr/bin/env python3
"""Run on the GPU box: error statistics of the device guiding fit against the reference build (oracle/_ref) over many
regions / rounds, plus a first timing of the BASELINE config-1 shape.  Prints a table; not a test."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import guiding_data, helpers
from test_guiding_cpu import FIELDS, SCENE_MIN, SCENE_MAX

P, O = helpers.pt(), helpers.oracle()
scene = P.Scene(helpers.scene_path("cornell-dielectric"))


def vmm_pdf(st, dirs):
    K = st["K"]
    mu = np.stack([st["mux"][:K], st["muy"][:K], st["muz"][:K]], -1).astype(np.float64)
    k = st["kappa"][:K].astype(np.float64)
    norm = np.where(k > 0, k / (2 * np.pi * (1 - np.exp(-2 * k))), 1 / (4 * np.pi))
    c = dirs @ mu.T
    return (st["weight"][:K].astype(np.float64) * norm * np.exp(k * np.minimum(c - 1, 0))).sum(-1)


def parity(splits, per_region, rounds, kw):
    gp = P.default_guiding_params(**kw)
    r = P.Renderer(32, 32, 0, splits); r.set_scene(scene)
    g = O.GuidingRef(splits, SCENE_MIN, SCENE_MAX, gp)
    aabbs = r.guiding_aabbs(); R = len(aabbs)
    rng = np.random.default_rng(0)
    dirs = rng.normal(size=(4096, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    for rnd in range(rounds):
        batch = guiding_data.make_batch(aabbs, per_region, 500 + rnd)
        g.update(batch, threads=os.cpu_count()); r.guiding_update_host(batch, gp)
        errs = {f: [] for f in FIELDS}; pdf_err = []; struct = 0
        for i in range(R):
            a, b = g.state(i), r.guiding_state(i)
            if a["K"] != b["K"] or a["numEMIterations"] != b["numEMIterations"]:
                continue
            struct += 1
            K = a["K"]
            for f in FIELDS:
                x, y = a[f][:K].astype(np.float64), b[f][:K].astype(np.float64)
                with np.errstate(invalid="ignore"):
                    e = np.abs(x - y)
                    if f in ("kappa", "distance", "distSumW", "chi", "chiN", "covSumW"):
                        e = e / np.maximum(np.abs(x), 1e-3)
                errs[f].append(np.nanmax(np.where(np.isfinite(e), e, 0)))
            pa, pb = vmm_pdf(a, dirs), vmm_pdf(b, dirs)
            pdf_err.append(np.abs(pa - pb).max() / pa.max())
        print("splits %d N %d %s round %d: structurally equal %d/%d | mixture pdf rel-to-peak err: median %.1e p99 %.1e max %.1e" %
              (splits, per_region, kw, rnd, struct, R, np.median(pdf_err), np.percentile(pdf_err, 99), np.max(pdf_err)))
        print("    " + "  ".join("%s %.0e/%.0e" % (f, np.median(errs[f]), np.max(errs[f])) for f in FIELDS))


def timing(per_region):
    gp = P.default_guiding_params()
    r = P.Renderer(32, 32, 0, 8); r.set_scene(scene)
    aabbs = r.guiding_aabbs()
    t0 = time.time()
    b1 = guiding_data.make_batch(aabbs, per_region, 1, invalid_fraction=0.0)
    b2 = guiding_data.make_batch(aabbs, per_region, 2, invalid_fraction=0.0)
    print("generated 2 x %d samples in %.1f s" % (len(b1), time.time() - t0))
    for rep in range(2):
        r.guiding_reset(gp); r.stats_reset()
        for b in (b1, b2):
            t = time.time(); r.guiding_update_host(b, gp); dt = time.time() - t
            s = r.stats()
            print("GPU update: wall %.1f ms  sort %.2f ms fit %.2f ms  samples %d  em sample-iters %d" % (1e3 * dt, s.ms_guiding_sort, s.ms_guiding_fit, s.guiding_samples, s.guiding_em_sample_iterations))
            r.stats_reset()
    for threads in (os.cpu_count(), 1):
        g = O.GuidingRef(8, SCENE_MIN, SCENE_MAX, gp)
        for b in (b1, b2):
            t = time.time(); g.update(b, threads=threads); dt = time.time() - t
            print("CPU reference (%d threads): %.2f s  (%.2f Msamples/s)" % (threads, dt, len(b) / dt / 1e6))
        if per_region > 20000 and threads == os.cpu_count():
            pass


if __name__ == "__main__":
    parity(5, 1500, 3, {})
    parity(3, 6000, 3, {})
    parity(2, 1200, 2, {"useParallaxCompensation": 0})
    timing(int(sys.argv[1]) if len(sys.argv) > 1 else 57600)
