// Irradiance cache (IC) and ADRRS device code.
//   lookup      queryIrradianceCache        shaders/raytrace.rgen:748-773 + raytrace.irradiance.rint:11-21 + .rahit:18-51
//   build       calculateCacheData          rgen:1231-1316 (N=20 x M=10 stratified one-bounce paths, Ward–Heckbert gradients)
//               update / create             rgen:1334-1421, clampGradients :1318-1329
//   ADRRS       applyWeightWindow           rgen:830-869;  split bookkeeping :883-902
//
// The reference answers the lookup with a ray query against an acceleration structure of one AABB per cache sphere
// that the host refits once per frame (src/IrradianceCache.cpp:81-104).  B200 has no RT cores: the lookup structure
// here is a uniform grid over the scene box, rebuilt on the device at the start of every frame from a snapshot of the
// cache.  Every lookup of a frame reads that snapshot; entries created or updated during the frame are written to the
// live arrays in pixel order after the frame's last path (the reference races here, SURVEY quirk 11 — the oracle
// defines the same race-free semantic).  Cell lists hold cache indices in ascending order, so the weighted sum of a
// lookup is accumulated in index order like the oracle's linear scan (the any-hit order of the reference is the
// driver's).
#pragma once
#include "shading.cuh"
#include "guiding_device.cuh"

namespace b200pt {

#define IC_MAX_NEW 5          // MAX_NEW_IRRADIANCE_ENTRIES, rgen:65
#define IC_MAX_SPLITS 10      // MAX_SPLITS, rgen:82
#define IC_GRID_MAX 32        // cells per axis (at most)
#define IC_SPLIT_F4 5         // float4 per stored split record

struct ICView {
    const float4 *sphere;     // snapshot: center.xyz, radius          (binding 12)
    const float4 *normalR;    // snapshot: normal.xyz, harmonicR       (binding 13)
    const float4 *color;      // snapshot: color.xyz, numUpdates bits
    const float4 *rotGrad;    // snapshot: rotGrad.xyz
    const float4 *transGrad;  // snapshot: transGrad.xyz
    const uint32_t *cellStart;   // [cells + 1]
    const uint32_t *cellItems;   // cache indices, ascending inside a cell
    const float4 *cellSpheres;   // the same lists with the sphere (center, radius) in line: the containment test of a
                                 // lookup streams these without a dependent index -> sphere load
    float gmin[3], invCell[3];
    int dim[3];
};

// per-pixel state of the IC / ADRRS modes (rgen:60-92: newIrradianceCacheEntries[], splits[], estimate)
struct ICState {
    ICView view;
    const float4 *estimate;   // estimate image (binding 14), read by ADRRS
    float4 *queryResult;      // per path-queue entry of the current iteration: irradiance rgb, w = 1 when the lookup found something (k_ic_query)
    uint32_t *newCount;       // [pixel] nextNewIrradianceCacheSlot
    float4 *newEntries;       // [(pixel * IC_MAX_NEW + k) * 2]: origin, normal
    uint32_t *splitState;     // [pixel] nextSplitSlot | (index of the next split to drain) << 16; nullptr when no split mode is on
    float4 *splitData;        // [(pixel * IC_MAX_SPLITS + k) * IC_SPLIT_F4]
};

__device__ __forceinline__ int icCellCoord(const ICView &ic, int axis, float x) {
    int c = int(floorf((x - ic.gmin[axis]) * ic.invCell[axis]));
    return min(max(c, 0), ic.dim[axis] - 1);
}

__device__ __forceinline__ bool isICCapable(const b200pt_push_constants &pc, int matType) {   // rgen:806-813
    if (matType == B200PT_MAT_DIFFUSE || matType == B200PT_MAT_LIGHT) return true;
    return pc.useIrradianceCacheOnGlossy && !hasDiscreteDirection(matType);
}

__device__ __forceinline__ vec3 approxDiffuse(const DeviceScene &sc, const b200pt_material &mat, vec3 normal, vec3 wi, float u, float v) {   // rgen:815-828
    switch (mat.type) {
        case B200PT_MAT_DIFFUSE: case B200PT_MAT_LIGHT: case B200PT_MAT_PHONG: return matDiffuse(sc, mat, u, v);
        case B200PT_MAT_ROUGH_CONDUCTOR: { float c = dot(wi, normal); return V3(fresnelConductor(c, mat.eta, mat.k)) * 1.0f / PT_PI; }
    }
    return V3(-1.0f, -1.0f, -1.0f);
}

// rgen:748-773: weighted sum over the cache spheres that contain `origin`.
// Two passes per window of 32 list entries: (1) the containment test of every entry, four independent 128-bit loads in
// flight, result = one bit per entry; (2) the weight / colour arithmetic for the set bits only, in list order (ascending cache
// index, the oracle's summation order).  With one pass the ~100 instructions of (2) ran for every entry that ANY lane of the
// warp contained (10 of 32 lanes busy, profiles/r02b_ic_query_by_line.txt); now a lane's trip count is its own number of hits.
__device__ __forceinline__ bool queryIrradianceCache(const ICView &ic, const b200pt_push_constants &pc, vec3 origin, vec3 normal, vec3 &color) {
    const int cx = icCellCoord(ic, 0, origin.x), cy = icCellCoord(ic, 1, origin.y), cz = icCellCoord(ic, 2, origin.z);
    const int cell = (cz * ic.dim[1] + cy) * ic.dim[0] + cx;
    const uint32_t b = __ldg(&ic.cellStart[cell]), e = __ldg(&ic.cellStart[cell + 1]);
    vec3 cacheValueSum = V3(0.0f);
    float totalWeight = 0.0f;
    for (uint32_t w0 = b; w0 < e; w0 += 32u) {
        const uint32_t wn = min(32u, e - w0);
        uint32_t inside = 0u;
        for (uint32_t k0 = 0; k0 < wn; k0 += 4u) {
            float4 s4[4];
#pragma unroll
            for (int j = 0; j < 4; j++) s4[j] = __ldg(&ic.cellSpheres[w0 + min(k0 + j, wn - 1u)]);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float dist = length(origin - make_vec3(s4[j]));
                if (k0 + j < wn && dist <= s4[j].w) inside |= 1u << (k0 + j);                     // irradiance.rint:15-20
            }
        }
        while (inside) {
            const uint32_t k = w0 + uint32_t(__ffs(inside) - 1);
            inside &= inside - 1u;
            const float4 s = __ldg(&ic.cellSpheres[k]);
            const vec3 oc = origin - make_vec3(s);
            const float dist = length(oc);
            const uint32_t i = __ldg(&ic.cellItems[k]);
            const float4 nr = __ldg(&ic.normalR[i]);
            const vec3 cn = make_vec3(nr);
            float weight = 1.0f / (dist / nr.w + sqrtf(1.0f - dot(normal, cn)));         // irradiance.rahit:18-21
            if (isnan(weight) || isinf(weight)) weight = 1000000.0f;
            const bool vis = -0.001f <= dot(oc, (normal + cn) * 0.5f);                   // x / 2 == x * 0.5 for every float
            if (weight <= 1.0f / pc.irradianceA || (pc.irradianceCachePerformVisibilityCheck && !vis)) continue;
            const vec3 c = make_vec3(__ldg(&ic.color[i]));
            if (pc.useIrradianceGradients) {
                const float E = length(c);
                const vec3 col = E != 0.0f ? normalize(c) : V3(0.0f);
                const vec3 adjusted = col * (E + dot(cross(cn, normal), make_vec3(__ldg(&ic.rotGrad[i]))) + dot(oc, make_vec3(__ldg(&ic.transGrad[i]))));
                cacheValueSum += weight * adjusted;
            } else cacheValueSum += weight * c;
            totalWeight += weight;
        }
    }
    if (totalWeight > 0.0f) { color = cacheValueSum / totalWeight; return true; }
    return false;
}

// rgen:830-869
__device__ __forceinline__ float applyWeightWindow(const b200pt_push_constants &pc, uint32_t &seed, vec3 throughput, vec3 adjoint, vec3 estimate, int &n) {
    const float center = length(estimate / adjoint);
    const float lower = 2.0f * center / (1.0f + pc.adrrsS);
    const float upper = pc.adrrsS * lower;
    const float v = length(throughput);
    if (isnan(lower) || lower <= 0.0f) { n = 1; return 1.0f; }
    if (lower <= v && v <= upper) { n = 1; return 1.0f; }
    else if (v <= lower) { n = 1; return fmaxf(v / lower, 0.1f); }
    const float q = v / upper;
    n = int(q);
    if (rnd(seed) > (float(n + 1) - q)) n++;
    return q;
}

// rgen:883-902 — returns false when the pixel's split list is full
// (`store` = false: an eight-lane group of the cache build runs this redundantly, one lane writes — the __syncwarp orders the other lanes' read in front of it)
__device__ __forceinline__ bool splitPush(const ICState &ic, int pid, vec3 origin, vec3 normal, vec3 wi, vec3 throughput, float u, float v, int matIndex,
                                          int currentDepth, bool isFrontFace, bool store = true, unsigned syncMask = 0u) {
    if (!ic.splitState) return false;
    if (syncMask) __syncwarp(syncMask);
    const uint32_t ss = ic.splitState[pid];
    if (syncMask) __syncwarp(syncMask);
    const uint32_t next = ss & 0xffffu;
    if (next >= IC_MAX_SPLITS) return false;
    if (!store) return true;
    float4 *d = ic.splitData + (size_t(pid) * IC_MAX_SPLITS + next) * IC_SPLIT_F4;
    d[0] = make_f4(origin, u);
    d[1] = make_f4(normal, v);
    d[2] = make_f4(wi, __int_as_float(matIndex));
    d[3] = make_f4(throughput, __int_as_float(currentDepth));
    d[4] = make_float4(isFrontFace ? 1.0f : 0.0f, 0.0f, 0.0f, 0.0f);
    ic.splitState[pid] = ss + 1u;
    return true;
}

// getNewDirection, rgen:923-960 (GUIDE compiles the guided branch in)
template <bool GUIDE>
__device__ __forceinline__ float getNewDirection(const b200pt_push_constants &pc, const GuidingView &guide, uint32_t &seed, const b200pt_material &mat,
                                                 vec3 origin, vec3 normal, vec3 wi, bool isFrontFace, vec3 &newDirection) {
    if (GUIDE && pc.useGuiding && !hasDiscreteDirection(mat.type)) {
        const bool parallax = pc.useParallaxCompensation != 0;
        const uint32_t iRegion = getGuidingRegion(guide, origin);
        if (iRegion == B200PT_INVALID_REGION) return 0.0f;
        const b200pt_vmm_theta &vmm = guide.vmms[iRegion];
        float pdfMat;
        if (rnd(seed) < pc.guidingProb) {
            newDirection = sampleVmm(seed, vmm, origin, parallax);
            pdfMat = pdfBSDF(mat, normal, wi, newDirection);
        } else pdfMat = sampleBSDF(seed, mat, wi, normal, isFrontFace, newDirection);
        if (dot(newDirection, normal) < 0.0f || pdfMat <= 0.0f) return 0.0f;
        const float pdfGuiding = vmmPdf(newDirection, vmm, origin, parallax);
        if (isnan(pdfGuiding)) return sampleBSDF(seed, mat, wi, normal, isFrontFace, newDirection);
        return mixf(pdfMat, pdfGuiding, pc.guidingProb);
    }
    return sampleBSDF(seed, mat, wi, normal, isFrontFace, newDirection);
}

// ---------------------------------------------------------------------------------------------------------------
// In-line ("megakernel") path evaluation used by the cache build only: one thread walks the 200 one-bounce paths of
// one cache entry sequentially, because they share the pixel's LCG stream.  Rays are traced with
// the whole-ray wrapper; NEE shadow rays and MIS probes are resolved on the spot.
struct InlineTracer {
    const b200pt_push_constants &pc;
    const DeviceScene &sc;
    const ICState &ic;
    const GuidingView &guide;
    uint2 *stack;
    int stride;
    int pid;
    uint32_t seed;
    uint32_t extendRays, shadowRays, vertices;
    bool group;        // eight lanes run this tracer in lockstep on the same entry and share every ray (traceRayGroup); lane 0 of the group stores

    __device__ __forceinline__ InlineTracer(const b200pt_push_constants &pc_, const DeviceScene &sc_, const ICState &ic_, const GuidingView &g_,
                                            uint2 *stack_, int stride_, int pid_, uint32_t seed_, bool group_ = false)
        : pc(pc_), sc(sc_), ic(ic_), guide(g_), stack(stack_), stride(stride_), pid(pid_), seed(seed_), extendRays(0), shadowRays(0), vertices(0), group(group_) {}

    template <bool ANY>
    __device__ __forceinline__ void trace(vec3 o, vec3 d, float tmin, float tmax, HitRec &h) {
        if (group) traceRayGroup<ANY>(sc.trace, o, d, tmin, tmax, h, stack, stride);
        else traceRay<ANY>(sc.trace, o, d, tmin, tmax, h, stack, stride);
    }

    // nextEventEstimation, rgen:601-731
    __device__ vec3 nee(const b200pt_material &mat, vec3 origin, vec3 wi, vec3 normal, float tu, float tv, bool isFrontFace) {
        if (!neeSupported(mat.type)) return V3(0.0f);
        vec3 neeResult = V3(0.0f);
        vec3 lightDir, lightColor;
        float lightDistance;
        float pdfLights = sampleLights(sc, seed, pc.useVisibleSphereSampling != 0, origin, normal, lightDir, lightColor, lightDistance);
        bool isShadowed = true;
        const float cosThetaLight = dot(normal, lightDir);
        if (cosThetaLight > 0.0f && pdfLights > 0.0f) {
            HitRec sh;
            shadowRays++;
            trace<true>(origin, lightDir, PT_TMIN, lightDistance * (1 - 0.0001f), sh);
            if (sh.prim == PT_MISS) isShadowed = false;
        }
        if (!isShadowed) {
            if (pc.enableMIS) {
                const float pdfMatL = pdfBSDF(mat, normal, wi, lightDir);
                const float heuristic = pc.usePowerHeuristic ? powerHeuristic(pdfLights, pdfMatL) : balanceHeuristic(pdfLights, pdfMatL);
                if (isnan(heuristic)) return neeResult;
                neeResult += evalBsdf(sc, mat, tu, tv, normal, wi, lightDir, true) * lightColor * heuristic / pdfLights;
            } else neeResult = evalBsdf(sc, mat, tu, tv, normal, wi, lightDir, true) * lightColor / pdfLights;
        }
        if (pc.enableMIS) {
            vec3 bsdfDir = V3(0.0f);
            const float pdfMat = sampleBSDF(seed, mat, wi, normal, isFrontFace, bsdfDir);
            if (pdfMat > 0.0f) {
                HitRec h;
                extendRays++;
                trace<false>(origin, bsdfDir, PT_TMIN, PT_TMAX, h);
                if (h.prim == PT_MISS) {
                    lightColor = envColor(sc, bsdfDir);
                    pdfLights = 1.0f / (2.0f * PT_PI) / float(sc.numLights);
                    const float heuristic = pc.usePowerHeuristic ? powerHeuristic(pdfMat, pdfLights) : balanceHeuristic(pdfMat, pdfLights);
                    neeResult += evalBsdf(sc, mat, tu, tv, normal, wi, bsdfDir, true) * lightColor * heuristic / pdfMat;
                } else {
                    HitInfo info;
                    computeHitInfo(sc, h, origin, bsdfDir, info);
                    const b200pt_material *m = &sc.materials[info.matIndex];
                    if (m->type == B200PT_MAT_LIGHT) {
                        const int iLight = info.isSphere ? sc.spheres[info.instanceIndex].iLight : sc.instances[info.instanceIndex].iLight;
                        if (iLight >= 0) {
                            lightColor = V3(m->lightColor[0], m->lightColor[1], m->lightColor[2]);
                            pdfLights = pdfLight(sc.lights[iLight], bsdfDir, info.normal, info.t);
                            const float heuristic = pc.usePowerHeuristic ? powerHeuristic(pdfMat, pdfLights) : balanceHeuristic(pdfMat, pdfLights);
                            if (isnan(heuristic)) return neeResult;
                            neeResult += evalBsdf(sc, mat, tu, tv, normal, wi, bsdfDir, true) * lightColor * heuristic / pdfMat;
                        }
                    }
                }
            }
        }
        return neeResult;
    }
    __device__ __forceinline__ vec3 multipleNEE(const b200pt_material &mat, vec3 origin, vec3 wi, vec3 normal, float tu, float tv, bool isFrontFace, int numNEE) {   // rgen:871-877
        vec3 result = V3(0.0f);
        for (int i = 0; i < numNEE; i++) result += nee(mat, origin, wi, normal, tu, tv, isFrontFace);
        return result / float(numNEE);
    }

    // raytrace(), rgen:992-1226, as calculateCacheData calls it (rgen:1275-1277): throughput 1, addDirectLights = false,
    // addFirstHitLight = false, useNEE = true, useIC = true, createIC = false, useADRRS = false, saveSamples = false
    __device__ vec3 raytrace(vec3 origin, vec3 direction, int maxDepth, int maxFollowDiscrete, int numNEE, float &firstT) {
        bool follow = true;
        int followCount = 0;
        bool addNextDirectLights = false;
        int depth = 0;
        firstT = PT_TMAX;
        vec3 throughput = V3(1.0f);
        vec3 result = V3(0.0f);
        do {
            depth++;
            HitRec h;
            extendRays++;
            trace<false>(origin, direction, PT_TMIN, PT_TMAX, h);
            if (h.prim == PT_MISS) {
                if (addNextDirectLights) result += throughput * envColor(sc, direction);
                break;
            }
            vertices++;
            HitInfo info;
            computeHitInfo(sc, h, origin, direction, info);
            const b200pt_material mat = sc.materials[info.matIndex];
            origin = info.worldPos;
            const vec3 normal = info.normal;
            const vec3 wi = -direction;
            if (depth == 1) firstT = info.t;
            if (mat.type == B200PT_MAT_LIGHT && addNextDirectLights) result += throughput * V3(mat.lightColor[0], mat.lightColor[1], mat.lightColor[2]);
            if (hasDiscreteDirection(mat.type)) {
                addNextDirectLights = true;
                follow = true;
                if (depth >= maxDepth) followCount++;
            } else {
                addNextDirectLights = false;
                follow = false;
                if (isICCapable(pc, mat.type)) {
                    vec3 irradianceColor;
                    if (queryIrradianceCache(ic.view, pc, origin, normal, irradianceColor)) {
                        const vec3 diff = approxDiffuse(sc, mat, normal, wi, info.u, info.v);
                        result += throughput * diff * irradianceColor;
                        result += throughput * multipleNEE(mat, origin, wi, normal, info.u, info.v, info.isFrontFace, numNEE);
                        break;
                    }
                }
                if (pc.splitOnFirst && depth == 1) {
                    if (splitPush(ic, pid, origin, normal, wi, throughput * 0.5f, info.u, info.v, info.matIndex, depth, info.isFrontFace,
                                  !group || (threadIdx.x & 7u) == 0u, group ? 0xffu << (threadIdx.x & 24u) : 0u)) throughput *= 0.5f;
                }
                const vec3 neeLight = multipleNEE(mat, origin, wi, normal, info.u, info.v, info.isFrontFace, numNEE);
                result += throughput * neeLight;
            }
            vec3 newDirection = V3(0.0f);
            const float pdf = getNewDirection<true>(pc, guide, seed, mat, origin, normal, wi, info.isFrontFace, newDirection);
            if (pdf <= 0.0f) break;
            throughput *= evalBsdf(sc, mat, info.u, info.v, normal, wi, newDirection, info.isFrontFace) / pdf;
            direction = newDirection;
        } while (depth <= maxDepth || (follow && followCount <= maxFollowDiscrete));
        return result;
    }

    // calculateCacheData, rgen:1231-1316; returns the harmonic mean distance (-1: no surface seen)
    __device__ float calculateCacheData(vec3 origin, vec3 normal, vec3 &calculatedColor, vec3 &rotGrad, vec3 &transGrad) {
        const int N = 20, M = 10;
        const float M_HALF_PI = PT_PI / 2.0f;
        float invDistanceSum = 0.0f;
        int numDistances = 0;
        const int maxFollowDiscrete = 10;
        rotGrad = V3(0.0f); transGrad = V3(0.0f);
        vec3 color = V3(0.0f);
        float previousKLs[M], previousKRs[M];     // sic: declared inside the k loop in GLSL; the values persist in practice
        for (int j = 0; j < M; j++) { previousKLs[j] = 0.0f; previousKRs[j] = 0.0f; }
        for (int k = 0; k < N; k++) {
            const float phi = 2.0f * PT_PI * (float(k) + rnd(seed)) / float(N);
            const vec3 uk = toWorld(sphericalToCartesian(M_HALF_PI, phi), normal);
            const vec3 vk = toWorld(sphericalToCartesian(M_HALF_PI, phi + M_HALF_PI), normal);
            const vec3 previousVk = toWorld(sphericalToCartesian(M_HALF_PI, 2.0f * PT_PI * float(k) / float(N) + M_HALF_PI), normal);
            float previousJR = 0.0f, previousJL = 0.0f;
            for (int j = 0; j < M; j++) {
                const float theta = ptAsinf(sqrtf((float(j) + rnd(seed)) / float(M)));
                const vec3 direction = toWorld(sphericalToCartesian(theta, phi), normal);
                float r = PT_TMAX;
                const vec3 sampleColor = raytrace(origin, direction, 1, maxFollowDiscrete, pc.irradianceNumNEE, r);
                color += sampleColor;
                if (r < PT_TMAX) { invDistanceSum += 1.0f / r; numDistances++; }
                const float L = length(sampleColor);
                const float previousJTheta = ptAsinf(sqrtf(float(j) / float(M)));
                const float nextJTheta = ptAsinf(sqrtf(float(j + 1) / float(M)));
                float tanTheta = ptTanf(theta);
                if (isinf(tanTheta) || isnan(tanTheta)) tanTheta = 0.0f;
                rotGrad = rotGrad - tanTheta * L * vk;
                if (j > 0) {
                    const float cosPreviousTheta = ptCosf(previousJTheta);
                    transGrad += uk * 2.0f * PT_PI / float(N) * ptSinf(previousJTheta) * cosPreviousTheta * cosPreviousTheta / fminf(r, previousJR) * (L - previousJL);
                }
                if (k > 0) transGrad += previousVk * (ptSinf(nextJTheta) - ptSinf(previousJTheta)) / fminf(r, previousKRs[j]) * (L - previousKLs[j]);
                previousKLs[j] = L; previousKRs[j] = r; previousJL = L; previousJR = r;
            }
        }
        const float normFactor = PT_PI / float(M * N);
        calculatedColor = normFactor * color;
        rotGrad *= normFactor;
        if (invDistanceSum == 0.0f || numDistances == 0) return -1.0f;
        return 1.0f / (invDistanceSum / float(numDistances));
    }
};

__device__ __forceinline__ void clampGradients(float maxLength, vec3 &rotGrad, vec3 &transGrad) {   // rgen:1318-1329
    const float lr = length(rotGrad);
    if (lr > maxLength) rotGrad *= maxLength / lr;
    const float lt = length(transGrad);
    if (lt > maxLength) transGrad *= maxLength / lt;
}

}  // namespace b200pt
