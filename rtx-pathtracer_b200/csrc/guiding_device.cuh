// Guiding inside the tracer: region lookup, vMF-mixture pdf / sampling and the per-pixel sample-recording state.
// Restates shaders/guiding.glsl:32-96 (vMF, VMM, sampleVMF, sampleVMM), shaders/raytrace.guiding.rint:12-19 +
// raytrace.rgen:904-921 (getGuidingRegion: point-in-AABB query on the guiding acceleration structure) and the sample
// bookkeeping of raytrace.rgen:966-990 (updateSamples / commitSamples).
#pragma once
#include "device_math.cuh"
#include "../../include/b200pt.h"

namespace b200pt {

// The reference finds the region with a TerminateOnFirstHit query against one AABB per region.  The regions come
// from recursive halving (PathGuiding::createRegions), so the lookup here walks that tree: level L holds 2^L boxes,
// node i has the children 2i and 2i+1.  Sibling boxes are computed independently (l.max -= h, r.min += h,
// src/Shapes.h:32-46) and may overlap or leave a gap of one ulp, so a plain descent could disagree with an exhaustive
// scan; the walk backtracks, which makes it return exactly the LOWEST-index leaf that contains the point (the rule the
// oracle's linear scan defines; the reference leaves the choice among overlapping boxes to the driver).
struct GuidingView {
    const b200pt_aabb *levels;        // all levels concatenated: level L starts at (2^L - 1)
    const float4 *levelSplits;        // per inner node (same numbering): split axis (int bits), left child's max / right child's min on that axis
    const b200pt_vmm_theta *vmms;     // binding 16
    int splits;
    // adaptive refinement (PathGuiding::splitRegion, src/PathGuiding.cpp:328-348): a split region keeps its index and the
    // left half of its box, the right half becomes a new region at the end of the array.  spawnFirst[r] / spawnNext[c]
    // chain the regions that were cut off region r (newest first); both nullptr until the first split.
    const b200pt_aabb *aabbs;         // binding 15: the current box of every region
    const int32_t *spawnFirst, *spawnNext;
};

__device__ __forceinline__ bool aabbContains(const b200pt_aabb &b, vec3 p) {   // raytrace.guiding.rint:15-17: min(aabb.min, p) == aabb.min && max(aabb.max, p) == aabb.max
    return fminf(b.min[0], p.x) == b.min[0] && fminf(b.min[1], p.y) == b.min[1] && fminf(b.min[2], p.z) == b.min[2] &&
           fmaxf(b.max[0], p.x) == b.max[0] && fmaxf(b.max[1], p.y) == b.max[1] && fmaxf(b.max[2], p.z) == b.max[2];
}

// region of a point inside base leaf `base` of the halving tree: the leaf itself, or — after adaptive splits — the
// lowest-index region among the leaf and everything that was cut off it whose current box contains the point
__device__ __forceinline__ uint32_t guidingLeafRegion(const GuidingView &g, uint32_t base, vec3 p) {
    if (!g.spawnFirst) return base;
    uint32_t best = B200PT_INVALID_REGION;
    int32_t stack[12];
    int sp = 0;
    stack[sp++] = int32_t(base);
    while (sp) {
        const int32_t r = stack[--sp];
        if (uint32_t(r) < best && aabbContains(g.aabbs[r], p)) best = uint32_t(r);
        for (int32_t c = g.spawnFirst[r]; c >= 0; c = g.spawnNext[c]) if (sp < 12) stack[sp++] = c;
    }
    return best;
}

// the backtracking walk: only needed when the greedy descent below runs into a dead end (the point lies in the one-ulp gap
// or overlap between two sibling boxes)
// (scalar arguments: an out-of-line call that passes a struct by reference costs its caller even when it is never executed)
__device__ __noinline__ uint32_t getGuidingRegionWalk(const b200pt_aabb *levels, const b200pt_aabb *aabbs, const int32_t *spawnFirst, const int32_t *spawnNext,
                                                      int splits, float px, float py, float pz) {
    GuidingView g; g.levels = levels; g.levelSplits = nullptr; g.vmms = nullptr; g.splits = splits; g.aabbs = aabbs; g.spawnFirst = spawnFirst; g.spawnNext = spawnNext;
    const vec3 p = V3(px, py, pz);
    int level = 0;
    uint32_t node = 0;
    uint32_t triedRight = 0;          // bit L: the right child of the node on the current path at level L was entered
    for (;;) {
        if (level == g.splits) {
            const uint32_t r = guidingLeafRegion(g, node, p);
            if (r != B200PT_INVALID_REGION) return r;
        } else {
            // descend: prefer the left child
            const b200pt_aabb *next = g.levels + ((1u << (level + 1)) - 1u);
            const uint32_t l = 2u * node, r = l + 1u;
            if (!((triedRight >> level) & 1u) && aabbContains(next[l], p)) { node = l; level++; continue; }
            if (!((triedRight >> level) & 1u) && aabbContains(next[r], p)) { triedRight |= 1u << level; node = r; level++; continue; }
        }
        // dead end (or both children exhausted): back up to the nearest ancestor whose right child is untried
        for (;;) {
            if (level == 0) return B200PT_INVALID_REGION;
            const bool cameFromLeft = (node & 1u) == 0u;
            triedRight &= ~(1u << level);
            level--; node >>= 1;
            if (cameFromLeft && !((triedRight >> level) & 1u)) {
                const b200pt_aabb *nx = g.levels + ((1u << (level + 1)) - 1u);
                triedRight |= 1u << level;
                if (aabbContains(nx[2u * node + 1u], p)) { node = 2u * node + 1u; level++; break; }
            }
        }
    }
}

// Greedy descent first: at every level take the left child if its box contains the point, else the right one — the same
// choices the walk above makes until its first dead end, but with one trip per level for every lane of the warp and selects
// instead of branches.  (As a per-lane walk with `continue` / `break` the lanes of a warp parted ways at the first level and
// never met again: 3 of 32 lanes per instruction, half of the guided shade kernel's instructions — profiles/r02b_shade_training_by_line.txt.)
// If some level offers no child, or the leaf has been split adaptively and rejects the point, the walk decides.
__device__ __forceinline__ uint32_t getGuidingRegion(const GuidingView &g, vec3 p) {
    if (!aabbContains(g.levels[0], p)) return B200PT_INVALID_REGION;
    uint32_t node = 0;
    bool ok = true;
    // The children of a box differ from it on the split axis only (splitAabb: l.max[axis] -= h, r.min[axis] += h), and p is inside
    // the parent by induction, so "the child contains p" (raytrace.guiding.rint:15-17 on all six faces) is one comparison:
    // fmax(l.max, p) == l.max <=> p <= l.max, fmin(r.min, p) == r.min <=> p >= r.min (p is finite here; +-0 compare equal both ways).
    // PT_REGION_SPLITREC = 1 does exactly that with one 16-byte record per inner node instead of the two child boxes (twelve loads per
    // level).  Same regions; measured on config 5 (gpurun_scratch/ab_rec.sh, two alternating repetitions): training render 4 169 ->
    // 4 117 ms, but guided frames 578 -> 610 ms per 4 — a net loss for anything but the seven training frames, so it stays off.
#ifndef PT_REGION_SPLITREC
#define PT_REGION_SPLITREC 0
#endif
    for (int level = 0; level < g.splits; level++) {
#if PT_REGION_SPLITREC
        const float4 sp = __ldg(&g.levelSplits[((1u << level) - 1u) + node]);
        const int axis = __float_as_int(sp.x);
        const float pa = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
        const bool inL = pa <= sp.y, inR = pa >= sp.z;
#else
        const b200pt_aabb *next = g.levels + ((1u << (level + 1)) - 1u);
        const bool inL = aabbContains(next[2u * node], p);
        const bool inR = aabbContains(next[2u * node + 1u], p);
#endif
        ok = ok && (inL || inR);
        node = 2u * node + (inL ? 0u : 1u);
    }
    if (ok) {
        if (!g.spawnFirst) return node;
        const uint32_t r = guidingLeafRegion(g, node, p);
        if (r != B200PT_INVALID_REGION) return r;
    }
    return getGuidingRegionWalk(g.levels, g.aabbs, g.spawnFirst, g.spawnNext, g.splits, p.x, p.y, p.z);
}

__device__ __forceinline__ float vmfPdf(vec3 wo, const b200pt_vmf_theta &th, vec3 worldPos, bool parallax) {   // guiding.glsl:32-44
    if (th.k == 0.0f) return 0.07957747155f;
    vec3 mu = V3(th.mu[0], th.mu[1], th.mu[2]);
    if (parallax && th.distance > 0.0f) mu = normalize(V3(th.target[0], th.target[1], th.target[2]) - worldPos);
    return th.norm * ptExpf(th.k * (dot(mu, wo) - 1.0f));
}
__device__ __forceinline__ float vmmPdf(vec3 wo, const b200pt_vmm_theta &vmm, vec3 worldPos, bool parallax) {   // guiding.glsl:53-60
    float res = 0.0f;
    for (int i = 0; i < vmm.usedDistributions; i++) res += vmm.pi[i] * vmfPdf(wo, vmm.thetas[i], worldPos, parallax);
    return res;
}
__device__ __forceinline__ vec3 sampleVmf(uint32_t &seed, const b200pt_vmf_theta &th, vec3 worldPos, bool parallax) {   // guiding.glsl:62-82
    if (th.k > 0.0f) {
        const float r1 = rnd(seed);
        const float r2 = rnd(seed);
        const float cosTheta = 1.0f + ptLogf(1.0f + th.eMin2K * r1 - r1) / th.k;
        const float sinTheta = 1.0f - cosTheta * cosTheta <= 0.0f ? 0.0f : sqrtf(1.0f - cosTheta * cosTheta);
        const float phi = 2.f * PT_PI * r2;
        float cosPhi, sinPhi;
        ptSinCosf(phi, &sinPhi, &cosPhi);
        vec3 mu = V3(th.mu[0], th.mu[1], th.mu[2]);
        if (parallax && th.distance > 0.0f) mu = normalize(V3(th.target[0], th.target[1], th.target[2]) - worldPos);
        return toWorld(V3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta), mu);
    }
    return randomOnUnitSphere(seed);
}
__device__ __forceinline__ vec3 sampleVmm(uint32_t &seed, const b200pt_vmm_theta &vmm, vec3 worldPos, bool parallax) {   // guiding.glsl:84-96
    const float rndDist = rnd(seed);
    int i = 0;
    const int maxDistribution = vmm.usedDistributions - 1;
    float piSum = vmm.pi[0];
    while (piSum < rndDist && i < maxDistribution) { i++; piSum += vmm.pi[i]; }
    return sampleVmf(seed, vmm.thetas[i], worldPos, parallax);
}

// per-pixel sample-recording state of the megakernel (rgen: sampleOffset, lightSums[], sampleThroughputs[] and the
// raytrace() locals currentSampleOffset / iUpdateDistance / distanceFactor), kept in global memory between the
// wavefront stages
struct GuidingRecord {
    b200pt_directional_data *samples;   // binding 18: W*H*16 records
    float4 *lightSums;                  // [pixel * 16 + i]
    float4 *sampleThr;                  // [pixel * 16 + i]
    int4 *state;                        // x sampleOffset, y currentSampleOffset, z iUpdateDistance, w pending commit (cso of the finished path, -1 none)
    float *distanceFactor;
    float4 *pathSum;                    // radiance of the pixel's current path (for `length(result) > 0`, rgen:1219-1222)
};

// updateSamples (rgen:973-978) on data in global memory.  ATOMIC: used by the shadow / probe resolvers, where several
// NEE results of one pixel may arrive concurrently.
template <bool ATOMIC>
__device__ __forceinline__ void updateSamples(const GuidingRecord &g, int pix, int sampleOffset, int currentSampleOffset, vec3 light) {
    for (int i = currentSampleOffset - 1; i >= sampleOffset; i--) {
        float *ls = reinterpret_cast<float *>(&g.lightSums[pix * B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL + i]);
        if (ATOMIC) { atomicAdd(ls + 0, light.x); atomicAdd(ls + 1, light.y); atomicAdd(ls + 2, light.z); }
        else { ls[0] += light.x; ls[1] += light.y; ls[2] += light.z; }
        light *= make_vec3(g.sampleThr[pix * B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL + i]);
    }
}

// commitSamples (rgen:980-990) followed by `if (length(result) > 0) sampleOffset += currentSampleOffset` (rgen:1219-1222, sic)
__device__ __forceinline__ void commitSamples(const GuidingRecord &g, int pix, int4 &st) {
    const int cso = st.w;
    for (int i = st.x; i < cso; i++) {
        b200pt_directional_data &dd = g.samples[pix * B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL + i];
        const float weight = length(make_vec3(g.lightSums[pix * B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL + i])) / dd.pdf;
        if (weight <= 0.0f) dd.flags = B200PT_INVALID_REGION;
        else dd.weight = weight;
    }
    if (length(make_vec3(g.pathSum[pix])) > 0.0f) st.x += cso;
    st.w = -1;
}

}  // namespace b200pt
