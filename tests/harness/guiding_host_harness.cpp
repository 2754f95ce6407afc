// TEST HARNESS — not part of the product.  Compiles the per-sample functions and the per-component logic of
// rtx-pathtracer_b200/csrc/guiding_math.cuh (the code the CUDA kernels of guiding_fit.cu run) for the host with a
// serial executor, so that the arithmetic can be checked against the reference's lightpmm build (oracle/_ref) in
// the CPU-only test run.  libb200pt.so does not contain or link this file; the product path is the CUDA path only.
#include "../../rtx-pathtracer_b200/csrc/guiding_math.cuh"
#include <vector>
#include <cstring>

using namespace b200pt;

namespace {
struct Sample { float dx, dy, dz, w, pdf, dist; };

struct SerialExec {
    const Sample *s; uint32_t N;
    EmAcc emAcc; StatAcc statAcc; DistAcc distAcc; GFrames fr; GFitState fs; float met[120];
    bool leader() const { return true; }
    int bcast(int v) { return v; }
    EmAcc &em() { return emAcc; }
    StatAcc &stat() { return statAcc; }
    DistAcc &dst() { return distAcc; }
    GFrames &frames() { return fr; }
    float *metric() { return met; }
    GFitState &fit() { return fs; }
    template <int KPAD> void emT(const GPacked &p, EmAcc &out) { memset(&out, 0, sizeof(out)); for (uint32_t i = 0; i < N; i++) gEmSample<KPAD>(p, s[i].dx, s[i].dy, s[i].dz, s[i].w, out); }
    template <int KPAD> void statT(const GPacked &p, const GFrames &f, StatAcc &out) { memset(&out, 0, sizeof(out)); for (uint32_t i = 0; i < N; i++) gStatSample<KPAD>(p, f, s[i].dx, s[i].dy, s[i].dz, s[i].w, s[i].pdf, out); }
    template <int KPAD> void distT(const GPacked &p, DistAcc &out) { memset(&out, 0, sizeof(out)); for (uint32_t i = 0; i < N; i++) gDistSample<KPAD>(p, s[i].dx, s[i].dy, s[i].dz, s[i].w, s[i].dist, out); }
    void emPass(const GMix &m, EmAcc &out) {
        GPacked p; gPack(m, p);
        switch (gKpad(m.K)) { case 4: emT<4>(p, out); break; case 8: emT<8>(p, out); break; case 12: emT<12>(p, out); break; default: emT<16>(p, out); }
    }
    void statPass(const GMix &m, const GFrames &f, StatAcc &out) {
        GPacked p; gPack(m, p);
        switch (gKpad(m.K)) { case 4: statT<4>(p, f, out); break; case 8: statT<8>(p, f, out); break; case 12: statT<12>(p, f, out); break; default: statT<16>(p, f, out); }
    }
    void distPass(const GMix &m, DistAcc &out) {
        GPacked p; gPack(m, p);
        switch (gKpad(m.K)) { case 4: distT<4>(p, out); break; case 8: distT<8>(p, out); break; case 12: distT<12>(p, out); break; default: distT<16>(p, out); }
    }
    void metricPass(const GMix &m, float *metric) {
        const int K = m.K; int p = 0;
        for (int a = 0; a < K - 1; a++) for (int b = a + 1; b < K; b++) metric[p++] = gMergeMetricPair(m, a, b);
    }
};

struct Harness {
    std::vector<b200pt_aabb> aabbs;
    std::vector<GMix> mixes;
    std::vector<b200pt_vmm_theta> thetas;
    b200pt_guiding_params gp;
    bool firstFit = true;
};
}  // namespace

extern "C" {
void *gharness_create(const b200pt_aabb *aabbs, int n, const b200pt_guiding_params *gp) {
    Harness *h = new Harness();
    h->aabbs.assign(aabbs, aabbs + n);
    h->gp = *gp;
    h->mixes.resize(n); h->thetas.resize(n);
    for (int r = 0; r < n; r++) { gInitialize(h->mixes[r], *gp); gPackTheta(h->mixes[r], gp->useParallaxCompensation != 0, h->thetas[r]); }
    return h;
}
void gharness_destroy(void *p) { delete static_cast<Harness *>(p); }
void gharness_update(void *p, const b200pt_directional_data *recs, int64_t n) {
    Harness *h = static_cast<Harness *>(p);
    const int R = int(h->aabbs.size());
    std::vector<std::vector<Sample>> per(R);
    for (int64_t i = 0; i < n; i++) {
        b200pt_directional_data d = recs[i];
        if (d.flags >= uint32_t(R)) continue;
        const b200pt_aabb &bb = h->aabbs[d.flags];
        float mean[3];
        for (int a = 0; a < 3; a++) mean[a] = bb.min[a] + 0.5f * (bb.max[a] - bb.min[a]);
        if (h->gp.useParallaxCompensation) gPrefitSample(d.position, d.direction, d.distance, mean);
        per[d.flags].push_back({d.direction[0], d.direction[1], d.direction[2], d.weight, d.pdf, d.distance});
    }
    for (int r = 0; r < R; r++) {
        if (per[r].empty()) continue;
        const b200pt_aabb &bb = h->aabbs[r];
        float mean[3];
        for (int a = 0; a < 3; a++) mean[a] = bb.min[a] + 0.5f * (bb.max[a] - bb.min[a]);
        SerialExec x; x.s = per[r].data(); x.N = uint32_t(per[r].size());
        uint64_t iters = 0;
        gUpdateRegion(x, h->mixes[r], h->gp, x.N, h->firstFit, mean, &iters);
        gPackTheta(h->mixes[r], h->gp.useParallaxCompensation != 0, h->thetas[r]);
    }
    h->firstFit = false;
}
void gharness_get_vmms(void *p, b200pt_vmm_theta *out) { Harness *h = static_cast<Harness *>(p); memcpy(out, h->thetas.data(), h->thetas.size() * sizeof(b200pt_vmm_theta)); }
void gharness_get_state(void *p, int region, float *sc, float *pc) {
    const GMix &m = static_cast<Harness *>(p)->mixes[region];
    sc[0] = float(m.K); sc[1] = m.sampleWeight; sc[2] = m.numSamples; sc[3] = float(m.totalNumSamples); sc[4] = float(m.numEMIterations);
    const float *src[14] = {m.w, m.kappa, m.r, m.mux, m.muy, m.muz, m.dist, m.distSumW, m.chi, m.chiN, m.covxx, m.covyy, m.covxy, m.covSumW};
    for (int f = 0; f < 14; f++) memcpy(pc + f * 16, src[f], 64);
}
void gharness_fastexp(const float *in, float *out, int n) { for (int i = 0; i < n; i++) out[i] = gFastExp(in[i]); }
}
