// Shading math of the path tracer in CUDA: hit attributes, BSDF kit, light sampling, texture / env lookups.
// Restates (does not copy) the GLSL of the reference; each function names the shader lines it follows.
#pragma once
#include "device_math.cuh"
#include "traverse.cuh"
#include "../../include/b200pt.h"

namespace b200pt {

struct DeviceTexture { const float4 *texels; int width, height; };   // linear RGBA float (sRGB decoded at upload)

struct DeviceScene {
    TraceScene trace;
    const b200pt_vertex *vertices;       // all models concatenated
    const uint32_t *indices;             // all models concatenated (local vertex indices)
    const int32_t *modelVertexOffset;    // [num_models]
    const int32_t *modelIndexOffset;     // [num_models]
    const int4 *primVerts;               // per global triangle: global vertex ids v0,v1,v2 + instance index
    const b200pt_material *materials;
    const b200pt_instance *instances;
    const b200pt_light *lights;
    const int32_t *randomLightIndex;
    const b200pt_face_sample *randomTriIndex;
    const b200pt_sphere *spheres;
    const DeviceTexture *textures;
    int numLights, numFaceTables, numTextures, numInstances;
};

struct HitInfo {     // shaders/raycommon.glsl:1-13
    vec3 worldPos, normal;
    float u, v;      // textureUV
    int matIndex;
    float t;
    bool isFrontFace, isSphere;
    uint32_t instanceIndex;
};

// --- textures: sampler2D with linear filter + repeat addressing (src/SceneLoader.cpp:216-227) -----------------
__device__ __forceinline__ float4 sampleTexture(const DeviceScene &sc, int id, float u, float v) {
    const DeviceTexture tex = sc.textures[id];
    float x = u * float(tex.width) - 0.5f, y = v * float(tex.height) - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float ax = x - fx, ay = y - fy;
    int x0 = int(fx) % tex.width, y0 = int(fy) % tex.height;
    if (x0 < 0) x0 += tex.width;
    if (y0 < 0) y0 += tex.height;
    int x1 = x0 + 1 == tex.width ? 0 : x0 + 1, y1 = y0 + 1 == tex.height ? 0 : y0 + 1;
    float4 a = __ldg(&tex.texels[y0 * tex.width + x0]), b = __ldg(&tex.texels[y0 * tex.width + x1]);
    float4 c = __ldg(&tex.texels[y1 * tex.width + x0]), d = __ldg(&tex.texels[y1 * tex.width + x1]);
    float4 r;
    r.x = (a.x * (1 - ax) + b.x * ax) * (1 - ay) + (c.x * (1 - ax) + d.x * ax) * ay;
    r.y = (a.y * (1 - ax) + b.y * ax) * (1 - ay) + (c.y * (1 - ax) + d.y * ax) * ay;
    r.z = (a.z * (1 - ax) + b.z * ax) * (1 - ay) + (c.z * (1 - ax) + d.z * ax) * ay;
    r.w = (a.w * (1 - ax) + b.w * ax) * (1 - ay) + (c.w * (1 - ax) + d.w * ax) * ay;
    return r;
}

// raytrace.rmiss:17-29 — lat-long lookup into texture 0 (1x1 black when the scene has no env map)
__device__ __forceinline__ vec3 envColor(const DeviceScene &sc, vec3 dir, bool renormalize = true) {
    vec3 udir = renormalize ? normalize(dir) : dir;
    float at = ptAtan2f(udir.x, -udir.z);
    float u = at * 1.0f / (2.0f * PT_PI);      // `atan * M_INV_2PI` with the unparenthesised macro
    float v = PT_ACOSF(udir.y) / PT_PI;
    float4 c = sampleTexture(sc, 0, u, v);
    return V3(c.x, c.y, c.z);
}

// raytrace.rchit:16-58 and raytrace.sphere.rchit:15-34
__device__ __forceinline__ void computeHitInfo(const DeviceScene &sc, const HitRec &h, vec3 o, vec3 d, HitInfo &info) {
    info.t = h.t;
    if (h.prim >= sc.trace.numTris) {
        const uint32_t si = h.prim - sc.trace.numTris;
        const b200pt_sphere s = sc.spheres[si];
        vec3 pos = o + h.t * d;
        vec3 normal = normalize(pos - V3(s.center[0], s.center[1], s.center[2]));
        if (dot(d, normal) < 0.0f) { info.isFrontFace = true; info.normal = normal; }
        else { info.isFrontFace = false; info.normal = -normal; }
        info.worldPos = pos;
        info.u = 0.0f; info.v = 0.0f;
        info.matIndex = s.materialIndex;
        info.isSphere = true;
        info.instanceIndex = si;
        return;
    }
    const int4 pv = __ldg(&sc.primVerts[h.prim]);
    const b200pt_vertex *V = sc.vertices;
    const float4 p0 = __ldg(reinterpret_cast<const float4 *>(&V[pv.x]) + 0), n0 = __ldg(reinterpret_cast<const float4 *>(&V[pv.x]) + 1);
    const float4 t0 = __ldg(reinterpret_cast<const float4 *>(&V[pv.x]) + 2);
    const float4 p1 = __ldg(reinterpret_cast<const float4 *>(&V[pv.y]) + 0), n1 = __ldg(reinterpret_cast<const float4 *>(&V[pv.y]) + 1);
    const float4 t1 = __ldg(reinterpret_cast<const float4 *>(&V[pv.y]) + 2);
    const float4 p2 = __ldg(reinterpret_cast<const float4 *>(&V[pv.z]) + 0), n2 = __ldg(reinterpret_cast<const float4 *>(&V[pv.z]) + 1);
    const float4 t2 = __ldg(reinterpret_cast<const float4 *>(&V[pv.z]) + 2);
    const float bx = 1.0f - h.u - h.v, by = h.u, bz = h.v;
    vec3 normal = make_vec3(n0) * bx + make_vec3(n1) * by + make_vec3(n2) * bz;
    const uint32_t instIdx = uint32_t(pv.w) & PT_INSTANCE_MASK;
    vec3 worldPos = make_vec3(p0) * bx + make_vec3(p1) * by + make_vec3(p2) * bz;
    if (uint32_t(pv.w) & PT_INSTANCE_IDENTITY) {
        // both instance matrices are exactly the identity (every Mitsuba-XML scene): ((1*x + 0*y) + 0*z) + 0*w of finite
        // values is x with -0 folded onto +0, i.e. x + 0 — the same bits as the matrix product, 2 x 21 operations and two
        // 64-byte matrix loads cheaper (the products were 7.5 % of the shade kernel's instructions)
        normal = normalize(V3(normal.x + 0.0f, normal.y + 0.0f, normal.z + 0.0f));
        worldPos = V3(worldPos.x + 0.0f, worldPos.y + 0.0f, worldPos.z + 0.0f);
    } else {
        const b200pt_instance *inst = &sc.instances[instIdx];
        normal = normalize(mat4MulPoint(inst->normalTransform, normal, 0.0f));
        worldPos = mat4MulPoint(inst->transform, worldPos, 1.0f);
    }
    info.u = t0.x * bx + t1.x * by + t2.x * bz;
    info.v = t0.y * bx + t1.y * by + t2.y * bz;
    if (dot(d, normal) < 0.0f) { info.isFrontFace = true; info.normal = normal; }
    else { info.isFrontFace = false; info.normal = -normal; }
    info.worldPos = worldPos;
    info.matIndex = __float_as_int(t0.z);     // materialIndex of v0 (quirk 4)
    info.isSphere = false;
    info.instanceIndex = instIdx;
}

// material index of a hit without the rest of computeHitInfo (one index load + one vertex load instead of ten loads): for callers
// that only go on when the surface is of a certain material (the MIS probe: lights only)
__device__ __forceinline__ int hitMaterialIndex(const DeviceScene &sc, const HitRec &h) {
    if (h.prim >= sc.trace.numTris) return sc.spheres[h.prim - sc.trace.numTris].materialIndex;
    const int4 pv = __ldg(&sc.primVerts[h.prim]);
    return __float_as_int(__ldg(reinterpret_cast<const float4 *>(&sc.vertices[pv.x]) + 2).z);     // materialIndex of v0 (quirk 4)
}

// --- BSDF kit ------------------------------------------------------------------------------------------------
PT_NI_M3 float fresnelDielectric(float eta, float cosThetaI) {   // rgen:145-160
    float sinThetaSqr = eta * eta * (1 - cosThetaI * cosThetaI);
    if (sinThetaSqr > 1.0f) return 1.0f;
    float cosThetaT = sqrtf(1.0f - sinThetaSqr);
    float Rs = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    float Rp = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    return (Rs * Rs + Rp * Rp) / 2.0f;
}

PT_NI_M3 float fresnelConductor(float cosThetaI, float eta, float k) {   // rgen:163-175
    if (cosThetaI < 0.0f) cosThetaI = -cosThetaI;
    float Rs2 = ((eta * eta + k * k) * cosThetaI * cosThetaI - 2 * eta * cosThetaI + 1)
              / ((eta * eta + k * k) * cosThetaI * cosThetaI + 2 * eta * cosThetaI + 1);
    float Rp2 = ((eta * eta + k * k) - 2 * eta * cosThetaI + cosThetaI * cosThetaI)
              / ((eta * eta + k * k) + 2 * eta * cosThetaI + cosThetaI * cosThetaI);
    return (Rs2 + Rp2) / 2.0f;
}

__device__ __forceinline__ float beckmannD(const b200pt_material &mat, vec3 n, vec3 m) {   // rgen:179-202
    float cosTheta = dot(n, m);
    if (cosTheta <= 0) return 0.0f;
    float theta = PT_ACOSF(cosTheta);
    if (isnan(theta) || isinf(theta)) theta = 0;
    float tanTheta = PT_TANF(theta);
    if (isnan(tanTheta) || isinf(tanTheta)) tanTheta = 0;
    float alphaSqr = mat.roughness * mat.roughness;
    return PT_POWF(PT_E, -tanTheta * tanTheta / alphaSqr) / (PT_PI * alphaSqr * PT_POWF(cosTheta, 4.0f));
}

__device__ __forceinline__ float smithG1(const b200pt_material &mat, vec3 n, vec3 m, vec3 v) {   // rgen:204-222 (quirk 1 kept)
    float thetaV = dot(v, n);
    float c = dot(v, m) / thetaV;
    if (c <= 0) return 0;
    float a = 1.0f / (mat.roughness * PT_TANF(thetaV));
    if (a >= 1.6f) return 1.0f;
    float a2 = a * a;
    return (3.535f * a + 2.181f * a2) / (1 + 2.276f * a + 2.577f * a);
}

__device__ __forceinline__ float smithG(const b200pt_material &mat, vec3 i, vec3 o, vec3 n, vec3 m) {   // rgen:224-230
    float result = smithG1(mat, n, m, i) * smithG1(mat, n, m, o);
    if (result < 0) return 0;
    return result;
}

__device__ __forceinline__ vec3 matDiffuse(const DeviceScene &sc, const b200pt_material &mat, float u, float v) {   // rgen:232-239
    vec3 kd = V3(mat.diffuse[0], mat.diffuse[1], mat.diffuse[2]);
    if (mat.textureIdDiffuse != -1) {
        float4 t = sampleTexture(sc, mat.textureIdDiffuse, u, v);
        return kd * V3(t.x, t.y, t.z) / PT_PI;
    }
    return kd / PT_PI;
}

__device__ __forceinline__ vec3 phongBsdf(const DeviceScene &sc, const b200pt_material &mat, float u, float v, vec3 normal, vec3 wi, vec3 wo) {   // rgen:242-267
    vec3 res = V3(0.0f);
    float cosThetaWo = dot(wo, normal);
    if (cosThetaWo > 0) {
        res += matDiffuse(sc, mat, u, v);
        float dotReflDir = dot(reflect(-wo, normal), wi);
        if (dotReflDir > 0) {
            vec3 ks = V3(mat.specular[0], mat.specular[1], mat.specular[2]);
            vec3 s = (mat.specularHighlight + 2) / (2 * PT_PI) * ks * PT_POWF(dotReflDir, mat.specularHighlight);
            if (mat.textureIdSpecular != -1) {
                float4 t = sampleTexture(sc, mat.textureIdSpecular, u, v);
                s = s * V3(t.x, t.y, t.z);
            }
            res += s;
        }
    }
    return cosThetaWo * res;
}

__device__ __forceinline__ vec3 roughConductorBsdf(const b200pt_material &mat, vec3 normal, vec3 wi, vec3 wo) {   // rgen:269-280
    vec3 hr = normalize(wi + wo);
    float cosThetaIHr = dot(wi, hr);
    float cosThetaI = dot(wi, normal);
    float cosThetaO = dot(wo, normal);
    if (cosThetaI <= 0) return V3(0.0f);
    float f = cosThetaO * (fresnelConductor(cosThetaIHr, mat.eta, mat.k) * smithG(mat, wi, wo, normal, hr) * beckmannD(mat, normal, hr) / (4 * cosThetaI * cosThetaO));
    return V3(f);
}

PT_NI2 float pdfBSDF(const b200pt_material &mat, vec3 normal, vec3 wi, vec3 wo) {   // rgen:284-332
    switch (mat.type) {
        case B200PT_MAT_ROUGH_CONDUCTOR: {
            vec3 hr = normalize(wi + wo);
            float pm = beckmannD(mat, normal, hr) * fabsf(dot(hr, normal));
            if (pm <= 0 || dot(wo, hr) <= 0) return 0.0f;
            return pm / (4 * fabsf(dot(wo, hr)));
        }
        case B200PT_MAT_PHONG: {
            if (dot(normal, wo) < 0) return 0.0f;
            float lDiffuse = length(V3(mat.diffuse[0], mat.diffuse[1], mat.diffuse[2]));
            float lSpecular = length(V3(mat.specular[0], mat.specular[1], mat.specular[2]));
            float sumSpecDiff = lDiffuse + lSpecular;
            if (sumSpecDiff == 0) return 0.0f;
            vec3 reflected = reflect(-wi, normal);
            float highlight = mat.specularHighlight;
            float pdf = 0;
            if (dot(reflected, wo) > 0) {
                pdf = (highlight + 1) * PT_POWF(dot(reflected, wo), highlight) / (2 * PT_PI);
                pdf *= lSpecular / sumSpecDiff;
            }
            pdf += dot(wo, normal) / PT_PI * lDiffuse / sumSpecDiff;
            return pdf;
        }
        default:
            return dot(wo, normal) / PT_PI;
    }
}

// rgen:483-552. Returns the pdf (or the discrete probability); consumes RNG draws exactly like the shader.
PT_NI2 float sampleBSDF(uint32_t &seed, const b200pt_material &mat, vec3 wi, vec3 normal, bool frontFace, vec3 &newDirection) {
    switch (mat.type) {
        case B200PT_MAT_ROUGH_CONDUCTOR: {
            vec3 worldM = randomBeckmannNormal(seed, mat.roughness, normal);
            newDirection = reflect(-wi, worldM);
            return pdfBSDF(mat, normal, wi, newDirection);
        }
        case B200PT_MAT_PHONG: {
            float lDiffuse = length(V3(mat.diffuse[0], mat.diffuse[1], mat.diffuse[2]));
            float lSpecular = length(V3(mat.specular[0], mat.specular[1], mat.specular[2]));
            float sumSpecDiff = lDiffuse + lSpecular;
            if (sumSpecDiff == 0) return 0.0f;
            if (rnd(seed) * sumSpecDiff > lDiffuse) {
                vec3 reflected = reflect(-wi, normal);
                newDirection = randomInHemisphereCosinePower(seed, reflected, mat.specularHighlight);
                if (dot(normal, newDirection) < 0) return 0.0f;
                return pdfBSDF(mat, normal, wi, newDirection);
            }
            newDirection = randomInHemisphereCosine(seed, normal);
            return pdfBSDF(mat, normal, wi, newDirection);
        }
        case B200PT_MAT_SPECULAR:
        case B200PT_MAT_CONDUCTOR:
            newDirection = reflect(-wi, normal);
            return 1.0f;
        case B200PT_MAT_DIELECTRIC: {
            float eta = frontFace ? mat.refractionIndexInv : mat.refractionIndex;
            float cosThetaI = dot(wi, normal);
            float F = fresnelDielectric(eta, cosThetaI);
            if (rnd(seed) > F) { newDirection = refract(-wi, normal, eta); return 1.0f - F; }
            newDirection = reflect(-wi, normal);
            return F;
        }
        default:
            newDirection = randomInHemisphereCosine(seed, normal);
            return dot(newDirection, normal) / PT_PI;
    }
}

// rgen:557-599 — returns f * cos(theta_o)
PT_NI1 vec3 evalBsdf(const DeviceScene &sc, const b200pt_material &mat, float u, float v, vec3 normal, vec3 wi, vec3 wo, bool frontFace) {
    switch (mat.type) {
        case B200PT_MAT_DIFFUSE:
        case B200PT_MAT_LIGHT:
            return dot(wo, normal) * matDiffuse(sc, mat, u, v);
        case B200PT_MAT_PHONG:
            return phongBsdf(sc, mat, u, v, normal, wi, wo);
        case B200PT_MAT_ROUGH_CONDUCTOR: {
            vec3 r = roughConductorBsdf(mat, normal, wi, wo);
            if (isnan(r.x)) return V3(0.0f);
            return r;
        }
        case B200PT_MAT_DIELECTRIC: {
            float cosThetaI = dot(normal, wi);
            float eta = frontFace ? mat.refractionIndexInv : mat.refractionIndex;
            vec3 ks = V3(mat.specular[0], mat.specular[1], mat.specular[2]);
            if (dot(normal, wo) < 0) return ks * (1 - fresnelDielectric(eta, cosThetaI));
            return ks * fresnelDielectric(eta, cosThetaI);
        }
        case B200PT_MAT_CONDUCTOR:
            return V3(fresnelConductor(dot(wi, normal), mat.eta, mat.k));
        case B200PT_MAT_SPECULAR:
            return V3(mat.specular[0], mat.specular[1], mat.specular[2]);
        default:
            return V3(0.0f);
    }
}

__device__ __forceinline__ bool hasDiscreteDirection(int type) {   // rgen:733-742
    return type == B200PT_MAT_DIELECTRIC || type == B200PT_MAT_SPECULAR || type == B200PT_MAT_CONDUCTOR;
}
__device__ __forceinline__ bool neeSupported(int type) {           // rgen:602-611
    return type == B200PT_MAT_DIFFUSE || type == B200PT_MAT_PHONG || type == B200PT_MAT_LIGHT || type == B200PT_MAT_ROUGH_CONDUCTOR;
}
__device__ __forceinline__ bool isMatAlmostDiscrete(const b200pt_material &mat) {   // rgen:962-964
    return (mat.type == B200PT_MAT_ROUGH_CONDUCTOR && mat.roughness <= 0.3f) || (mat.type == B200PT_MAT_PHONG && mat.specularHighlight >= 250.0f);
}
__device__ __forceinline__ float powerHeuristic(float a, float b) { float s = a * a; return s / (s + b * b); }   // rgen:341-344
__device__ __forceinline__ float balanceHeuristic(float a, float b) { return a / (a + b); }                       // rgen:346-348
__device__ __forceinline__ float pdfLight(const b200pt_light &light, vec3 lightDir, vec3 lightNormal, float lightDistance) {   // rgen:334-339
    float cosThetaLight = dot(-lightDir, lightNormal);
    return light.sampleProb * lightDistance * lightDistance / cosThetaLight / light.area;
}

// rgen:358-473 — picks a light from the pre-drawn table and a point on it; returns the pdf
PT_NI3 float sampleLights(const DeviceScene &sc, uint32_t &seed, bool useVisibleSphereSampling, vec3 origin, vec3 normal,
                                              vec3 &lightDir, vec3 &lightColor, float &lightDistance) {
    int iRandomLight = rndInteger(seed, B200PT_SIZE_LIGHT_RANDOM - 1);
    int iLight = __ldg(&sc.randomLightIndex[iRandomLight]);
    if (iLight < 0 || iLight >= sc.numLights) {   // scene without lights: the reference reads a zeroed dummy record
        lightDir = V3(0.0f); lightColor = V3(0.0f); lightDistance = 0.0f;
        return 0.0f;
    }
    const b200pt_light light = sc.lights[iLight];
    if (light.type == B200PT_LIGHT_POINT) {
        vec3 toLight = V3(light.pos[0], light.pos[1], light.pos[2]) - origin;
        lightDistance = length(toLight);
        lightDir = toLight / lightDistance;
        lightColor = V3(light.color[0], light.color[1], light.color[2]) / (lightDistance * lightDistance);
        return light.sampleProb;
    } else if (light.type == B200PT_LIGHT_SPHERE) {
        const b200pt_sphere s = sc.spheres[light.instanceIndex];
        const b200pt_material *m = &sc.materials[s.materialIndex];
        lightColor = V3(m->lightColor[0], m->lightColor[1], m->lightColor[2]);
        vec3 sphereNormal, position;
        float area;
        vec3 center = V3(s.center[0], s.center[1], s.center[2]);
        if (useVisibleSphereSampling) {           // random.glsl:116-124
            area = light.area / 2.0f;
            sphereNormal = randomOnUnitSphere(seed);
            if (dot(normal, sphereNormal) > 0) sphereNormal *= -1.0f;
            position = center + sphereNormal * s.radius;
        } else {                                  // random.glsl:108-115
            area = light.area;
            sphereNormal = randomOnUnitSphere(seed);
            position = center + sphereNormal * s.radius;
        }
        vec3 toLight = position - origin;
        lightDistance = length(toLight);
        lightDir = normalize(toLight);
        float cosThetaLight = dot(-lightDir, sphereNormal);
        return light.sampleProb * lightDistance * lightDistance / (cosThetaLight * area);
    } else if (light.type == B200PT_LIGHT_ENV_MAP) {
        lightDir = randomInHemisphere(seed, normal);
        lightColor = envColor(sc, lightDir, false);   // same lat-long mapping without the normalize, rgen:411-416
        lightDistance = PT_TMAX;
        return light.sampleProb * 1.0f / (2.0f * PT_PI);
    }
    // area light
    const b200pt_instance *inst = &sc.instances[light.instanceIndex];
    int iModel = inst->modelIndex;
    int iRandomTri = rndInteger(seed, B200PT_SIZE_TRI_RANDOM - 1);
    int iTri = 0;
    if (iLight < sc.numFaceTables) iTri = sc.randomTriIndex[iLight * B200PT_SIZE_TRI_RANDOM + iRandomTri].index;   // quirk 5
    const uint32_t *idx = sc.indices + sc.modelIndexOffset[iModel] + 3 * iTri;
    const b200pt_vertex *vb = sc.vertices + sc.modelVertexOffset[iModel];
    const b200pt_vertex v0 = vb[idx[0]], v1 = vb[idx[1]], v2 = vb[idx[2]];
    float rx = rnd(seed), ry = rnd(seed);
    float sqrtx = sqrtf(rx);
    vec3 bary = V3(1.0f - sqrtx, sqrtx * (1.0f - ry), ry * sqrtx);
    vec3 P = V3(v0.pos[0], v0.pos[1], v0.pos[2]) * bary.x + V3(v1.pos[0], v1.pos[1], v1.pos[2]) * bary.y + V3(v2.pos[0], v2.pos[1], v2.pos[2]) * bary.z;
    vec3 N = V3(v0.normal[0], v0.normal[1], v0.normal[2]) * bary.x + V3(v1.normal[0], v1.normal[1], v1.normal[2]) * bary.y + V3(v2.normal[0], v2.normal[1], v2.normal[2]) * bary.z;
    if (inst->_pad[0]) {       // identity instance (flag set by b200pt_set_scene): see computeHitInfo
        P = V3(P.x + 0.0f, P.y + 0.0f, P.z + 0.0f);
        N = normalize(V3(N.x + 0.0f, N.y + 0.0f, N.z + 0.0f));
    } else {
        P = mat4MulPoint(inst->transform, P, 1.0f);
        N = normalize(mat4MulPoint(inst->normalTransform, N, 0.0f));
    }
    vec3 toLight = P - origin;
    lightDistance = length(toLight);
    lightDir = toLight / lightDistance;
    const b200pt_material *m = &sc.materials[v0.materialIndex];
    lightColor = V3(m->lightColor[0], m->lightColor[1], m->lightColor[2]);
    float cosThetaLight = dot(-lightDir, N);
    if (cosThetaLight < 0) cosThetaLight = -cosThetaLight;
    return light.sampleProb * lightDistance * lightDistance / cosThetaLight / light.area;
}

}  // namespace b200pt
