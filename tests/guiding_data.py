"""Synthetic radiance samples for the guiding-fit checks and the EM benchmark.

Per region: positions uniform in the region AABB; directions from a ground-truth mixture of 4 vMF lobes
(kappa in {4, 32, 256, 2048}, random mu, Dirichlet(1) weights) plus 10 % uniform sphere; pdf = 0.5 * cosine-hemisphere
(n = +y) + 0.5 * truth, clamped >= 1e-3; weight = lognormal(0, 0.75) * truth / pdf; distance uniform [0.5, 4] with
10 % zeros (= infinitely far); flags = region id.  Records are emitted in a shuffled order with INVALID slots mixed in,
like the W*H*16 buffer the tracer fills."""
import numpy as np

DD = np.dtype([("position", "<f4", 3), ("direction", "<f4", 3), ("weight", "<f4"), ("pdf", "<f4"), ("distance", "<f4"), ("flags", "<u4")])
INVALID = 0xFFFFFFFF
KAPPAS = (4.0, 32.0, 256.0, 2048.0)


def _sample_vmf(rng, mu, kappa, n):
    u = rng.random(n)
    w = 1.0 + np.log(u + (1.0 - u) * np.exp(-2.0 * kappa)) / kappa
    phi = 2.0 * np.pi * rng.random(n)
    s = np.sqrt(np.maximum(0.0, 1.0 - w * w))
    local = np.stack([s * np.cos(phi), s * np.sin(phi), w], -1)
    a = np.array([1.0, 0.0, 0.0]) if abs(mu[0]) < 0.9 else np.array([0.0, 1.0, 0.0])
    t = np.cross(mu, a); t /= np.linalg.norm(t)
    b = np.cross(mu, t)
    return local[:, :1] * t + local[:, 1:2] * b + local[:, 2:3] * mu


def _vmf_pdf(d, mu, kappa):
    return kappa / (2.0 * np.pi * (1.0 - np.exp(-2.0 * kappa))) * np.exp(kappa * (d @ mu - 1.0))


def region_samples(rng, aabb_min, aabb_max, region, n):
    mus = rng.normal(size=(4, 3)); mus /= np.linalg.norm(mus, axis=1, keepdims=True)
    pis = rng.dirichlet(np.ones(4))
    comp = rng.choice(5, size=n, p=np.append(0.9 * pis, 0.1))
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)      # the uniform 10 %
    for k in range(4):
        m = comp == k
        d[m] = _sample_vmf(rng, mus[k], KAPPAS[k], int(m.sum()))
    truth = 0.1 / (4.0 * np.pi) + 0.9 * sum(pis[k] * _vmf_pdf(d, mus[k], KAPPAS[k]) for k in range(4))
    pdf = np.maximum(0.5 * np.maximum(d[:, 1], 0.0) / np.pi + 0.5 * truth, 1e-3)
    out = np.zeros(n, dtype=DD)
    out["position"] = (aabb_min + rng.random((n, 3)) * (aabb_max - aabb_min)).astype(np.float32)
    out["direction"] = d.astype(np.float32)
    out["weight"] = (rng.lognormal(0.0, 0.75, n) * truth / pdf).astype(np.float32)
    out["pdf"] = pdf.astype(np.float32)
    dist = rng.uniform(0.5, 4.0, n)
    dist[rng.random(n) < 0.1] = 0.0
    out["distance"] = dist.astype(np.float32)
    out["flags"] = region
    return out


def make_batch(aabbs, per_region, seed, invalid_fraction=0.25, empty_regions=()):
    """aabbs: structured array with min/max; per_region: int or per-region list.  Returns a shuffled DD array."""
    rng = np.random.default_rng(seed)
    R = len(aabbs)
    counts = [per_region] * R if np.isscalar(per_region) else list(per_region)
    parts = []
    for r in range(R):
        if r in empty_regions or counts[r] == 0:
            rng.random(7)      # keep the stream position independent of which regions are empty
            continue
        parts.append(region_samples(rng, aabbs["min"][r].astype(np.float64), aabbs["max"][r].astype(np.float64), r, counts[r]))
    valid = np.concatenate(parts) if parts else np.zeros(0, dtype=DD)
    n_inv = int(len(valid) * invalid_fraction)
    inv = np.zeros(n_inv, dtype=DD)
    inv["flags"] = INVALID
    allrec = np.concatenate([valid, inv])
    return allrec[rng.permutation(len(allrec))]
