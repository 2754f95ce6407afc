// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing in the product (rtx-pathtracer_b200/, include/) may include, link or
// call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg do.
//
// CPU restatement of the reference's per-pixel ray-tracing megakernel: shaders/raytrace.rgen (+ .rchit, .rahit,
// .rmiss, .shadow.rmiss, .sphere.rint/.rchit, .irradiance.rint/.rahit, .guiding.rint/.rchit) and the includes
// random.glsl, transform.glsl, guiding.glsl, raycommon.glsl.  Every function cites the shader lines it follows
// (paths relative to /root/reference/shaders).  One pixel = one sequential program, exactly like a raygen invocation.
//
// PARITY PINNED (round 2c) against the reference's own shader source: the reference ships no golden vectors and its GLSL cannot
// run here as GLSL (no Vulkan, no glslc, no RT hardware — SURVEY.md §8(c)), but its shader files are C-like enough to be COMPILED
// AS C++ where they lie (oracle/Makefile, oracle/glsl_prelude*.h, oracle/shader_ref.cpp -> oracle/_ref/libshader_ref.so).  Fed with
// this file's hits, texels and elementary functions — the three things the reference leaves to the driver — that build renders
// frames that are BIT-EQUAL to this restatement's in every mode (tests/test_shader_ref.py: path tracing on seven scenes, irradiance
// cache lookups, ADRRS / splitting, guided sampling, every recorded DirectionalData record); function by function, with the C
// library's elementary functions instead, tests/test_oracle_cpu.py.  Also pinned: the light / face tables (the reference's
// WeightedSampler compiled into oracle/_ref/libhost_ref.so) and the converged images (scenes/*/**.exr) at the relMSE level.
//
// Ray traversal has no reference algorithm (driver / RT cores).  The oracle defines it as brute force over all
// primitives with individually rounded IEEE operations (compile with -ffp-contract=off):
//   triangles: Möller–Trumbore on world-space v0,e1,e2; accepted when tmin < t < tmax
//   spheres  : raytrace.sphere.rint:13-28, both roots, accepted when tmin <= t <= tmax
//   closest hit = lexicographic minimum of (t, primitive id); ids = triangles in instance order, then spheres.
// A median-split BVH (ORACLE_ACCEL) gives the same answers faster; tests check it against the brute force.
//
// Irradiance cache, frame semantic (ours — the reference races here): inside one frame every lookup
// sees the cache as it was when the frame started (in the reference new entries are not in the lookup acceleration
// structure before the host refits it for the next frame either; only the ~10 entries per frame that
// updateIrradianceCache rewrites can be observed half-way there).  Update slots (header.nextUpdateSlot++) and new cache
// slots (header.nextCacheSlot++) are handed out in pixel order, and the results are committed in that order after the
// last pixel: that is what the reference does when its invocations happen to run one after the other.
#include <cmath>
#include "../include/b200pt_detmath.h"   // sin / cos / pow / ...: the same deterministic kernels the device code uses (see the header)
namespace dm = b200pt_dm;
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <limits>
#include <thread>
#include <atomic>
#include <mutex>
#include "../include/b200pt.h"

namespace {

const float kPi = 3.14159265358979323846f;
const float kE = 2.71828182845904523536f;
const float tMin = 0.001f;          // rgen:52
const float tMax = 1000000.0f;      // rgen:53
const int MAX_NEW_IRRADIANCE_ENTRIES = 5;   // rgen:65
const int MAX_SPLITS = 10;                  // rgen:82
const int MAXD = B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL;

struct v3 {
    float x, y, z;
    v3() : x(0), y(0), z(0) {}
    v3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit v3(float s) : x(s), y(s), z(s) {}
    explicit v3(const float *p) : x(p[0]), y(p[1]), z(p[2]) {}
};
inline v3 operator+(v3 a, v3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline v3 operator-(v3 a, v3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline v3 operator-(v3 a) { return v3(-a.x, -a.y, -a.z); }
inline v3 operator*(v3 a, v3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline v3 operator*(v3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline v3 operator*(float s, v3 a) { return v3(a.x * s, a.y * s, a.z * s); }
inline v3 operator/(v3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
inline v3 operator/(v3 a, v3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline v3 &operator+=(v3 &a, v3 b) { a = a + b; return a; }
inline v3 &operator-=(v3 &a, v3 b) { a = a - b; return a; }
inline v3 &operator*=(v3 &a, v3 b) { a = a * b; return a; }
inline v3 &operator*=(v3 &a, float s) { a = a * s; return a; }
inline v3 &operator/=(v3 &a, float s) { a = a / s; return a; }
inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline v3 cross(v3 a, v3 b) { return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float length(v3 a) { return sqrtf(dot(a, a)); }
inline v3 normalize(v3 a) { return a / sqrtf(dot(a, a)); }
inline v3 reflect(v3 I, v3 N) { return I - 2.0f * dot(N, I) * N; }
inline v3 refract(v3 I, v3 N, float eta) {
    float d = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k < 0.0f) return v3(0.0f);
    return eta * I - (eta * d + sqrtf(k)) * N;
}
inline v3 mix(v3 x, v3 y, float a) { return x * (1.0f - a) + y * a; }
inline float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline bool isnanf_(float f) { return f != f; }
inline bool isinff_(float f) { return std::isinf(f); }

// column-major mat4 * (p, w); the evaluation order is the convention shared with the CUDA kernels
inline v3 mulPoint(const float *m, v3 p, float w) {
    return v3(((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12] * w,
              ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13] * w,
              ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14] * w);
}

// random.glsl:13-27
uint32_t tea(uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for (uint32_t n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
inline uint32_t floatToUintSat(float f) {   // GLSL uint(float) is undefined out of range; both sides saturate
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return uint32_t(f);
}

struct HitInfo {    // raycommon.glsl:1-13
    v3 worldPos, normal;
    float u = 0, v = 0;
    int matIndex = 0;
    float t = 0;
    bool isMiss = false, isFrontFace = false, isSphere = false;
    v3 missColor;
    uint32_t instanceIndex = 0;
};

struct Tex { int w = 0, h = 0; bool hasAlpha = false; std::vector<float> px; };   // linear RGBA float; hasAlpha: some texel is not opaque

struct AccelNode { float lo[3], hi[3]; int left, right; uint32_t first, count; };

struct SplitInfo {   // rgen:70-79
    v3 origin, normal, wi, throughput;
    float u, v;
    int matIndex, currentDepth;
    bool isFrontFace;
};
struct NewICInfo { v3 origin, normal; };

}  // namespace

struct oracle_ctx {
    int width = 0, height = 0;
    // scene (copies)
    std::vector<std::vector<b200pt_vertex>> vertices;
    std::vector<std::vector<uint32_t>> indices;
    std::vector<b200pt_material> mats;
    std::vector<b200pt_instance> instanceInfos;
    std::vector<b200pt_light> lights;
    std::vector<int32_t> randomLightIndex;
    std::vector<b200pt_face_sample> randomTriIndex;
    int numFaceTables = 0;
    std::vector<b200pt_sphere> spheres;
    std::vector<Tex> textures;
    // flattened world-space triangles: v0,e1,e2 (9 floats), and per global prim (instance, local prim)
    std::vector<float> tri;
    std::vector<uint32_t> primInstance, primLocal;
    uint32_t numTris = 0;
    bool anyTextured = false;
    // accel
    bool useAccel = false;
    std::vector<AccelNode> nodes;
    std::vector<uint32_t> order;
    // camera + push constants
    float view[16], proj[16], viewInverse[16], projInverse[16];
    b200pt_push_constants pushC;
    // images (RGBA)
    std::vector<float> image, accumulateImage, estimateImage;
    std::vector<float> aovImage;        // per pixel: maxReachedDepth, depthSum, depthsCounter, nextSplitSlot (rgen:1653-1655)
    // irradiance cache (bindings 10,12,13)
    b200pt_cache_header header{0, 0, 0};
    std::vector<b200pt_cache_data> cache;
    std::vector<b200pt_sphere> cacheSpheres;
    // frame-start snapshot (what every lookup of the frame reads) and the frame's deferred writes
    b200pt_cache_header snapHeader{0, 0, 0};
    std::vector<b200pt_cache_data> snapCache;
    std::vector<b200pt_sphere> snapSpheres;
    struct PendingEntry { bool valid; float harmonicR; float origin[3], normal[3], color[3], rotGrad[3], transGrad[3]; };
    std::vector<int32_t> updateSlotOfPixel;                 // per pixel of the region: cache index to update, -1 none
    std::vector<PendingEntry> pendingUpdate;                // per pixel of the region
    std::vector<std::vector<PendingEntry>> pendingCreate;   // per pixel of the region, in creation order
    int regionX0 = 0, regionY0 = 0, regionNX = 0;
    // guiding (bindings 15,16,18)
    std::vector<b200pt_aabb> guidingAabbs;
    std::vector<b200pt_vmm_theta> guidingVMM;
    std::vector<b200pt_directional_data> directionalData;
    // counters
    uint64_t extendRays = 0, shadowRays = 0, pathVertices = 0;
};

namespace {

struct Pixel {
    oracle_ctx &C;
    const b200pt_push_constants &pushC;
    uint32_t seed = 0;
    uint32_t px = 0, py = 0;
    HitInfo info;                     // rgen:136 (payload location 0, shared by every closest-hit trace)
    bool isShadowed = true;           // rgen:137
    NewICInfo newIrradianceCacheEntries[MAX_NEW_IRRADIANCE_ENTRIES];
    int nextNewIrradianceCacheSlot = 0;
    SplitInfo splits[MAX_SPLITS];
    int nextSplitSlot = 0;
    v3 estimate; float lengthEstimate = 0;
    v3 sampleThroughputs[MAXD], lightSums[MAXD];
    int sampleOffset = 0;
    uint64_t extendRays = 0, shadowRays = 0, pathVertices = 0;

    explicit Pixel(oracle_ctx &c) : C(c), pushC(c.pushC) {}

    // ---- random.glsl ------------------------------------------------------------------------------------------
    static float rndS(uint32_t &prev) {   // :31-43
        prev = 1664525u * prev + 1013904223u;
        return float(prev & 0x00FFFFFFu) / float(0x01000000);
    }
    float rnd() { return rndS(seed); }                                   // :45-47
    float getRandomNegPos() { return rnd() * 2 - 1; }                    // :50-52
    int getRandomInteger(int mx) { return int(rnd() * (mx + 1)); }       // :55-57
    v3 randomOnUnitSphere() {                                            // :60-68
        v3 res;
        do {
            float a = getRandomNegPos(), b = getRandomNegPos(), c = getRandomNegPos();
            res = v3(a, b, c);
        } while (length(res) > 1);
        return normalize(res);
    }
    v3 randomInHemisphere(v3 normal) {                                   // :71-83
        v3 res = randomOnUnitSphere();
        if (dot(normal, res) < 0) res = reflect(res, normal);
        return res;
    }
    // transform.glsl:7-42
    static void coordinateAxis(v3 z, v3 &x, v3 &y) {
        if (fabsf(z.x) > fabsf(z.y)) {
            float invLen = 1.0f / sqrtf(z.x * z.x + z.z * z.z);
            y = v3(z.z * invLen, 0.0f, -z.x * invLen);
        } else {
            float invLen = 1.0f / sqrtf(z.y * z.y + z.z * z.z);
            y = v3(0.0f, z.z * invLen, -z.y * invLen);
        }
        x = cross(y, z);
    }
    static v3 toWorld(v3 v, v3 n) { v3 x, y; coordinateAxis(n, x, y); return v.x * x + v.y * y + v.z * n; }
    static v3 sphericalToCartesian(float theta, float phi) { return v3(dm::sinF(theta) * dm::cosF(phi), dm::sinF(theta) * dm::sinF(phi), dm::cosF(theta)); }
    v3 randomInHemisphereCosine(v3 normal) {                             // random.glsl:86-94
        float u = rnd();
        float sqrt_u = sqrtf(u);
        float phi = 2 * kPi * rnd();
        return toWorld(v3(sqrt_u * dm::cosF(phi), sqrt_u * dm::sinF(phi), sqrtf(1 - u)), normal);
    }
    v3 randomInHemisphereCosinePower(v3 reflected, float p) {            // :97-106
        float u = rnd();
        float cosTheta = dm::powF(u, 1.0f / (p + 1));
        float phi = 2 * kPi * rnd();
        float sinTheta = sqrtf(1 - cosTheta * cosTheta);
        return toWorld(v3(sinTheta * dm::cosF(phi), sinTheta * dm::sinF(phi), cosTheta), reflected);
    }
    v3 randomOnSphere(const b200pt_sphere &s, v3 &normal) {              // :108-115
        normal = randomOnUnitSphere();
        return v3(s.center) + normal * s.radius;
    }
    v3 randomOnSphereVisible(const b200pt_sphere &s, v3 normal, v3 &sphereNormal) {   // :116-124
        sphereNormal = randomOnUnitSphere();
        if (dot(normal, sphereNormal) > 0) sphereNormal *= -1.0f;
        return v3(s.center) + sphereNormal * s.radius;
    }
    v3 randomBeckmannNormal(const b200pt_material &mat, v3 normal) {     // :126-136
        float thetaM = dm::atanF(sqrtf(-mat.roughness * mat.roughness * dm::logF(1 - rnd())));
        float phiM = 2 * kPi * rnd();
        float cosThetaNM = dm::cosF(thetaM);
        return toWorld(v3(dm::sinF(thetaM) * dm::cosF(phiM), dm::sinF(thetaM) * dm::sinF(phiM), cosThetaNM), normal);
    }

    // ---- texture(): linear filter, repeat addressing -------------------------------------------------------------
    void textureRGBA(int id, float u, float v, float out[4]) const {
        const Tex &t = C.textures[id];
        float x = u * float(t.w) - 0.5f, y = v * float(t.h) - 0.5f;
        float fx = floorf(x), fy = floorf(y);
        float ax = x - fx, ay = y - fy;
        int x0 = int(fx) % t.w, y0 = int(fy) % t.h;
        if (x0 < 0) x0 += t.w;
        if (y0 < 0) y0 += t.h;
        int x1 = x0 + 1 == t.w ? 0 : x0 + 1, y1 = y0 + 1 == t.h ? 0 : y0 + 1;
        const float *a = &t.px[4 * (size_t(y0) * t.w + x0)], *b = &t.px[4 * (size_t(y0) * t.w + x1)];
        const float *c = &t.px[4 * (size_t(y1) * t.w + x0)], *d = &t.px[4 * (size_t(y1) * t.w + x1)];
        for (int k = 0; k < 4; k++) out[k] = (a[k] * (1 - ax) + b[k] * ax) * (1 - ay) + (c[k] * (1 - ax) + d[k] * ax) * ay;
    }
    v3 textureRGB(int id, float u, float v) const { float c[4]; textureRGBA(id, u, v, c); return v3(c[0], c[1], c[2]); }

    // ---- traversal (brute force or oracle BVH) -------------------------------------------------------------------
    static bool triHit(const float *T, v3 o, v3 d, float &t, float &u, float &v) {
        const float *a = T, *b = T + 3, *c = T + 6;
        float px = d.y * c[2] - d.z * c[1], py = d.z * c[0] - d.x * c[2], pz = d.x * c[1] - d.y * c[0];
        float det = (b[0] * px + b[1] * py) + b[2] * pz;
        if (det == 0.0f) return false;
        float inv = 1.0f / det;
        float tx = o.x - a[0], ty = o.y - a[1], tz = o.z - a[2];
        u = ((tx * px + ty * py) + tz * pz) * inv;
        if (!(u >= 0.0f && u <= 1.0f)) return false;
        float qx = ty * b[2] - tz * b[1], qy = tz * b[0] - tx * b[2], qz = tx * b[1] - ty * b[0];
        v = ((d.x * qx + d.y * qy) + d.z * qz) * inv;
        if (!(v >= 0.0f && u + v <= 1.0f)) return false;
        t = ((c[0] * qx + c[1] * qy) + c[2] * qz) * inv;
        return true;
    }
    // raytrace.rahit:22-46 — stochastic alpha test, own RNG stream
    bool alphaRejects(uint32_t prim, float bu, float bv, v3 o, float t) const {
        uint32_t inst = C.primInstance[prim], lp = C.primLocal[prim];
        int iModel = C.instanceInfos[inst].modelIndex;
        const uint32_t *ind = &C.indices[iModel][3 * lp];
        const b200pt_vertex &v0 = C.vertices[iModel][ind[0]], &v1 = C.vertices[iModel][ind[1]], &v2 = C.vertices[iModel][ind[2]];
        const b200pt_material &mat = C.mats[v0.materialIndex];
        if (mat.textureIdDiffuse == -1) return false;
        // a fully opaque texture filters to alpha = 1 (give or take one rounding) and rnd() < 1: the test cannot reject
        // (except with probability 2^-24 per hit when the bilinear weights round to 1 - 2^-24); both sides skip it
        if (!C.textures[mat.textureIdDiffuse].hasAlpha) return false;
        float bx = 1.0f - bu - bv;
        float tu = v0.texCoord[0] * bx + v1.texCoord[0] * bu + v2.texCoord[0] * bv;
        float tv = v0.texCoord[1] * bx + v1.texCoord[1] * bu + v2.texCoord[1] * bv;
        uint32_t s = tea(floatToUintSat(tu * 100000000 + o.x * t), pushC.randomUInt);
        float c[4];
        textureRGBA(mat.textureIdDiffuse, tu, tv, c);
        return rndS(s) > c[3];
    }
    struct Cand { float t; uint32_t prim; float u, v; };
    inline void testTri(uint32_t prim, v3 o, v3 d, float tmin, float tmax, Cand &best) const {
        float t, u, v;
        if (!triHit(&C.tri[size_t(prim) * 9], o, d, t, u, v)) return;
        if (!(t > tmin && t < tmax)) return;
        if (!(t < best.t || (t == best.t && prim < best.prim))) return;
        if (C.anyTextured && alphaRejects(prim, u, v, o, t)) return;
        best.t = t; best.prim = prim; best.u = u; best.v = v;
    }
    // returns the closest accepted candidate (prim == MISS if none); anyHit: any accepted candidate will do
    Cand traverse(v3 o, v3 d, float tmin, float tmax, bool anyHit) const {
        Cand best{tmax, B200PT_MISS, 0, 0};
        for (uint32_t i = 0; i < C.spheres.size(); i++) {   // raytrace.sphere.rint:13-28
            const b200pt_sphere &s = C.spheres[i];
            float ox = o.x - s.center[0], oy = o.y - s.center[1], oz = o.z - s.center[2];
            float dotDOC = (d.x * ox + d.y * oy) + d.z * oz;
            float rootTerm = (dotDOC * dotDOC - ((ox * ox + oy * oy) + oz * oz)) + s.radius * s.radius;
            if (rootTerm < 0) continue;
            float root = sqrtf(rootTerm);
            float t1 = -dotDOC + root, t2 = -dotDOC - root;
            uint32_t id = C.numTris + i;
            for (float t : {t1, t2})
                if (t >= tmin && (t < best.t || (t == best.t && (best.prim == B200PT_MISS || id < best.prim)))) { best.t = t; best.prim = id; best.u = 0; best.v = 0; }
            if (anyHit && best.prim != B200PT_MISS) return best;
        }
        if (!C.useAccel) {
            for (uint32_t p = 0; p < C.numTris; p++) {
                testTri(p, o, d, tmin, tmax, best);
                if (anyHit && best.prim != B200PT_MISS) return best;
            }
            return best;
        }
        if (C.nodes.empty()) return best;
        float id3[3] = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z}, o3[3] = {o.x, o.y, o.z};
        int stack[128], sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const AccelNode &n = C.nodes[stack[--sp]];
            float t0 = 0.0f, t1 = best.t;
            bool miss = false;
            for (int a = 0; a < 3 && !miss; a++) {
                float ta = (n.lo[a] - o3[a]) * id3[a], tb = (n.hi[a] - o3[a]) * id3[a];
                if (ta != ta || tb != tb) continue;            // 0 * inf: origin on the slab plane, axis gives no bound
                if (ta > tb) std::swap(ta, tb);
                t0 = std::max(t0, ta); t1 = std::min(t1, tb);
                if (t0 > t1) miss = true;
            }
            if (miss) continue;
            if (n.left < 0) {
                for (uint32_t k = n.first; k < n.first + n.count; k++) {
                    testTri(C.order[k], o, d, tmin, tmax, best);
                    if (anyHit && best.prim != B200PT_MISS) return best;
                }
            } else { stack[sp++] = n.left; stack[sp++] = n.right; }
        }
        return best;
    }

    // texture 0 lat-long lookup — raytrace.rmiss:21-28
    v3 missColor(v3 dir) const {
        v3 udir = normalize(dir);
        float at = dm::atan2F(udir.x, -udir.z);
        float u = at * 1.0f / (2 * kPi);
        float v = dm::acosF(udir.y) / kPi;
        return textureRGB(0, u, v);
    }

    // traceRayEXT(topLevelAS, rayFlags, ..., PAYLOAD_HIT): raytrace.rchit / .sphere.rchit / .rmiss fill `info`
    void traceClosest(v3 origin, float tmin, v3 direction, float tmax) {
        extendRays++;
        Cand c = traverse(origin, direction, tmin, tmax, false);
        if (c.prim == B200PT_MISS) {                       // raytrace.rmiss:17-29 (isFrontFace etc. keep their old values)
            info.normal = -direction;
            info.t = tmax;
            info.isMiss = true;
            info.missColor = missColor(direction);
            info.matIndex = -1;
            return;
        }
        if (c.prim >= C.numTris) {                         // raytrace.sphere.rchit:15-34
            uint32_t si = c.prim - C.numTris;
            const b200pt_sphere &s = C.spheres[si];
            v3 pos = origin + c.t * direction;
            v3 normal = normalize(pos - v3(s.center));
            if (dot(direction, normal) < 0) { info.isFrontFace = true; info.normal = normal; }
            else { info.isFrontFace = false; info.normal = -normal; }
            info.worldPos = pos;
            info.u = 0; info.v = 0;
            info.matIndex = s.materialIndex;
            info.t = c.t;
            info.isMiss = false;
            info.isSphere = true;
            info.instanceIndex = si;
            return;
        }
        // raytrace.rchit:16-58
        uint32_t inst = C.primInstance[c.prim], lp = C.primLocal[c.prim];
        const b200pt_instance &ii = C.instanceInfos[inst];
        int iModel = ii.modelIndex;
        const uint32_t *ind = &C.indices[iModel][3 * lp];
        const b200pt_vertex &v0 = C.vertices[iModel][ind[0]], &v1 = C.vertices[iModel][ind[1]], &v2 = C.vertices[iModel][ind[2]];
        float bx = 1.0f - c.u - c.v, by = c.u, bz = c.v;
        v3 normal = v3(v0.normal) * bx + v3(v1.normal) * by + v3(v2.normal) * bz;
        normal = normalize(mulPoint(ii.normalTransform, normal, 0.0f));
        v3 worldPos = v3(v0.pos) * bx + v3(v1.pos) * by + v3(v2.pos) * bz;
        worldPos = mulPoint(ii.transform, worldPos, 1.0f);
        info.u = v0.texCoord[0] * bx + v1.texCoord[0] * by + v2.texCoord[0] * bz;
        info.v = v0.texCoord[1] * bx + v1.texCoord[1] * by + v2.texCoord[1] * bz;
        if (dot(direction, normal) < 0) { info.isFrontFace = true; info.normal = normal; }
        else { info.isFrontFace = false; info.normal = -normal; }
        info.worldPos = worldPos;
        info.matIndex = v0.materialIndex;
        info.t = c.t;
        info.isMiss = false;
        info.isSphere = false;
        info.instanceIndex = inst;
    }
    // shadow trace: flags TerminateOnFirstHit | SkipClosestHit, miss index 1 — raytrace.shadow.rmiss
    void traceShadow(v3 origin, float tmin, v3 direction, float tmax) {
        shadowRays++;
        Cand c = traverse(origin, direction, tmin, tmax, true);
        if (c.prim == B200PT_MISS) isShadowed = false;
    }

    // ---- BSDF kit ------------------------------------------------------------------------------------------------
    static float fresnel(float eta, float cosThetaI) {                   // rgen:145-160
        float sinThetaSqr = eta * eta * (1 - cosThetaI * cosThetaI);
        if (sinThetaSqr > 1.0f) return 1.0f;
        float cosThetaT = sqrtf(1.0f - sinThetaSqr);
        float Rs = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
        float Rp = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
        return (Rs * Rs + Rp * Rp) / 2.0f;
    }
    static float fresnelConductor(float cosThetaI, float eta, float k) {  // rgen:163-175
        if (cosThetaI < 0.0f) cosThetaI = -cosThetaI;
        float Rs2 = ((eta * eta + k * k) * cosThetaI * cosThetaI - 2 * eta * cosThetaI + 1)
                  / ((eta * eta + k * k) * cosThetaI * cosThetaI + 2 * eta * cosThetaI + 1);
        float Rp2 = ((eta * eta + k * k) - 2 * eta * cosThetaI + cosThetaI * cosThetaI)
                  / ((eta * eta + k * k) + 2 * eta * cosThetaI + cosThetaI * cosThetaI);
        return (Rs2 + Rp2) / 2.0f;
    }
    static float D(const b200pt_material &mat, v3 n, v3 m) {              // rgen:179-202
        float cosTheta = dot(n, m);
        if (cosTheta <= 0) return 0.0f;
        float theta = dm::acosF(cosTheta);
        if (isnanf_(theta) || isinff_(theta)) theta = 0;
        float tanTheta = dm::tanF(theta);
        if (isnanf_(tanTheta) || isinff_(tanTheta)) tanTheta = 0;
        float alphaSqr = mat.roughness * mat.roughness;
        return dm::powF(kE, -tanTheta * tanTheta / alphaSqr) / (kPi * alphaSqr * dm::powF(cosTheta, 4));
    }
    static float G1(const b200pt_material &mat, v3 n, v3 m, v3 v) {       // rgen:204-222
        float thetaV = dot(v, n);
        float c = dot(v, m) / thetaV;
        if (c <= 0) return 0;
        float a = 1.0f / (mat.roughness * dm::tanF(thetaV));
        if (a >= 1.6f) return 1.0f;
        float a2 = a * a;
        return (3.535f * a + 2.181f * a2) / (1 + 2.276f * a + 2.577f * a);
    }
    static float G(const b200pt_material &mat, v3 i, v3 o, v3 n, v3 m) {  // rgen:224-230
        float result = G1(mat, n, m, i) * G1(mat, n, m, o);
        return result < 0 ? 0 : result;
    }
    v3 diffuse(const b200pt_material &mat, float u, float v) const {     // rgen:232-239
        if (mat.textureIdDiffuse != -1) return v3(mat.diffuse) * textureRGB(mat.textureIdDiffuse, u, v) / kPi;
        return v3(mat.diffuse) / kPi;
    }
    v3 phong(const b200pt_material &mat, float u, float v, v3 normal, v3 wi, v3 wo) const {   // rgen:242-267
        v3 res(0.0f);
        float cosThetaWo = dot(wo, normal);
        if (cosThetaWo > 0) {
            res += diffuse(mat, u, v);
            float dotReflDir = dot(reflect(-wo, normal), wi);
            if (dotReflDir > 0) {
                v3 s = (mat.specularHighlight + 2) / (2 * kPi) * v3(mat.specular) * dm::powF(dotReflDir, mat.specularHighlight);
                if (mat.textureIdSpecular != -1) s = s * textureRGB(mat.textureIdSpecular, u, v);
                res += s;
            }
        }
        return cosThetaWo * res;
    }
    static v3 roughConductor(const b200pt_material &mat, v3 normal, v3 wi, v3 wo) {   // rgen:269-280
        v3 hr = normalize(wi + wo);
        float cosThetaIHr = dot(wi, hr), cosThetaI = dot(wi, normal), cosThetaO = dot(wo, normal);
        if (cosThetaI <= 0) return v3(0.0f);
        return v3(cosThetaO * (fresnelConductor(cosThetaIHr, mat.eta, mat.k) * G(mat, wi, wo, normal, hr) * D(mat, normal, hr) / (4 * cosThetaI * cosThetaO)));
    }
    static float pdfBSDF(const b200pt_material &mat, v3 normal, v3 wi, v3 wo) {       // rgen:284-332
        switch (mat.type) {
            case B200PT_MAT_ROUGH_CONDUCTOR: {
                v3 hr = normalize(wi + wo);
                float pm = D(mat, normal, hr) * fabsf(dot(hr, normal));
                if (pm <= 0 || dot(wo, hr) <= 0) return 0.0f;
                return pm / (4 * fabsf(dot(wo, hr)));
            }
            case B200PT_MAT_PHONG: {
                if (dot(normal, wo) < 0) return 0.0f;
                float lDiffuse = length(v3(mat.diffuse)), lSpecular = length(v3(mat.specular));
                float sumSpecDiff = lDiffuse + lSpecular;
                if (sumSpecDiff == 0) return 0.0f;
                v3 reflected = reflect(-wi, normal);
                float highlight = mat.specularHighlight;
                float pdf = 0;
                if (dot(reflected, wo) > 0) {
                    pdf = (highlight + 1) * dm::powF(dot(reflected, wo), highlight) / (2 * kPi);
                    pdf *= lSpecular / sumSpecDiff;
                }
                pdf += dot(wo, normal) / kPi * lDiffuse / sumSpecDiff;
                return pdf;
            }
            default: return dot(wo, normal) / kPi;
        }
    }
    static float pdfLight(const b200pt_light &light, v3 lightDir, v3 lightNormal, float lightDistance) {   // rgen:334-339
        float cosThetaLight = dot(-lightDir, lightNormal);
        return light.sampleProb * lightDistance * lightDistance / cosThetaLight / light.area;
    }
    static float powerHeuristic(float p1, float p2) { float s = p1 * p1; return s / (s + p2 * p2); }   // rgen:341-344
    static float balanceHeuristic(float p1, float p2) { return p1 / (p1 + p2); }                       // rgen:346-348

    float sampleLights(v3 origin, v3 normal, v3 &lightDir, v3 &lightColor, float &lightDistance) {     // rgen:358-473
        int iRandomLight = getRandomInteger(B200PT_SIZE_LIGHT_RANDOM - 1);
        int iLight = C.randomLightIndex[iRandomLight];
        if (iLight < 0 || iLight >= int(C.lights.size())) {   // light-less scene: out-of-bounds read in the reference
            lightDir = v3(0.0f); lightColor = v3(0.0f); lightDistance = 0;
            return 0.0f;
        }
        const b200pt_light &light = C.lights[iLight];
        if (light.type == B200PT_LIGHT_POINT) {
            v3 toLight = v3(light.pos) - origin;
            lightDistance = length(toLight);
            lightDir = toLight / lightDistance;
            lightColor = v3(light.color) / (lightDistance * lightDistance);
            return light.sampleProb;
        } else if (light.type == B200PT_LIGHT_SPHERE) {
            const b200pt_sphere &s = C.spheres[light.instanceIndex];
            lightColor = v3(C.mats[s.materialIndex].lightColor);
            v3 sphereNormal, position;
            float area;
            if (pushC.useVisibleSphereSampling) { area = light.area / 2.0f; position = randomOnSphereVisible(s, normal, sphereNormal); }
            else { area = light.area; position = randomOnSphere(s, sphereNormal); }
            v3 toLight = position - origin;
            lightDistance = length(toLight);
            lightDir = normalize(toLight);
            float cosThetaLight = dot(-lightDir, sphereNormal);
            return light.sampleProb * lightDistance * lightDistance / (cosThetaLight * area);
        } else if (light.type == B200PT_LIGHT_ENV_MAP) {
            lightDir = randomInHemisphere(normal);
            float at = dm::atan2F(lightDir.x, -lightDir.z);
            float u = at * 1.0f / (2 * kPi);
            float v = dm::acosF(lightDir.y) / kPi;
            lightColor = textureRGB(0, u, v);
            lightDistance = tMax;
            return light.sampleProb * 1.0f / (2 * kPi);
        }
        const b200pt_instance &ii = C.instanceInfos[light.instanceIndex];
        int iModel = ii.modelIndex;
        int iRandomTri = getRandomInteger(B200PT_SIZE_TRI_RANDOM - 1);
        int iTri = 0;   // randomTriIndex[iLight][..] with only mesh-light tables present (quirk 5); out of range reads 0
        if (iLight < C.numFaceTables) iTri = C.randomTriIndex[size_t(iLight) * B200PT_SIZE_TRI_RANDOM + iRandomTri].index;
        const uint32_t *ind = &C.indices[iModel][3 * iTri];
        const b200pt_vertex &v0 = C.vertices[iModel][ind[0]], &v1 = C.vertices[iModel][ind[1]], &v2 = C.vertices[iModel][ind[2]];
        float rx = rnd(), ry = rnd();
        float sqrtx = sqrtf(rx);
        v3 bary(1.0f - sqrtx, sqrtx * (1.0f - ry), ry * sqrtx);
        v3 P = v3(v0.pos) * bary.x + v3(v1.pos) * bary.y + v3(v2.pos) * bary.z;
        v3 N = v3(v0.normal) * bary.x + v3(v1.normal) * bary.y + v3(v2.normal) * bary.z;
        P = mulPoint(ii.transform, P, 1.0f);
        N = normalize(mulPoint(ii.normalTransform, N, 0.0f));
        v3 toLight = P - origin;
        lightDistance = length(toLight);
        lightDir = toLight / lightDistance;
        lightColor = v3(C.mats[v0.materialIndex].lightColor);
        float cosThetaLight = dot(-lightDir, N);
        if (cosThetaLight < 0) cosThetaLight = -cosThetaLight;
        return light.sampleProb * lightDistance * lightDistance / cosThetaLight / light.area;
    }

    float sampleBSDF(const b200pt_material &mat, v3 wi, v3 normal, bool frontFace, v3 &newDirection) {   // rgen:483-552
        switch (mat.type) {
            case B200PT_MAT_ROUGH_CONDUCTOR: {
                v3 worldM = randomBeckmannNormal(mat, normal);
                newDirection = reflect(-wi, worldM);
                return pdfBSDF(mat, normal, wi, newDirection);
            }
            case B200PT_MAT_PHONG: {
                float lDiffuse = length(v3(mat.diffuse)), lSpecular = length(v3(mat.specular));
                float sumSpecDiff = lDiffuse + lSpecular;
                if (sumSpecDiff == 0) return 0.0f;
                if (rnd() * sumSpecDiff > lDiffuse) {
                    v3 reflected = reflect(-wi, normal);
                    newDirection = randomInHemisphereCosinePower(reflected, mat.specularHighlight);
                    if (dot(normal, newDirection) < 0) return 0.0f;
                    return pdfBSDF(mat, normal, wi, newDirection);
                }
                newDirection = randomInHemisphereCosine(normal);
                return pdfBSDF(mat, normal, wi, newDirection);
            }
            case B200PT_MAT_SPECULAR:
            case B200PT_MAT_CONDUCTOR:
                newDirection = reflect(-wi, normal);
                return 1.0f;
            case B200PT_MAT_DIELECTRIC: {
                float eta = mat.refractionIndexInv;
                if (!frontFace) eta = mat.refractionIndex;
                float cosThetaI = dot(wi, normal);
                float F = fresnel(eta, cosThetaI);
                if (rnd() > F) { newDirection = refract(-wi, normal, eta); return 1.0f - F; }
                newDirection = reflect(-wi, normal);
                return F;
            }
            default:
                newDirection = randomInHemisphereCosine(normal);
                return dot(newDirection, normal) / kPi;
        }
    }

    v3 evalBsdf(const b200pt_material &mat, float u, float v, v3 normal, v3 wi, v3 wo, bool frontFace) const {   // rgen:557-599
        switch (mat.type) {
            case B200PT_MAT_DIFFUSE:
            case B200PT_MAT_LIGHT: return dot(wo, normal) * diffuse(mat, u, v);
            case B200PT_MAT_PHONG: return phong(mat, u, v, normal, wi, wo);
            case B200PT_MAT_ROUGH_CONDUCTOR: {
                v3 r = roughConductor(mat, normal, wi, wo);
                if (isnanf_(r.x)) return v3(0.0f);
                return r;
            }
            case B200PT_MAT_DIELECTRIC: {
                float cosThetaI = dot(normal, wi);
                float eta = mat.refractionIndexInv;
                if (!frontFace) eta = mat.refractionIndex;
                if (dot(normal, wo) < 0) return v3(mat.specular) * (1 - fresnel(eta, cosThetaI));
                return v3(mat.specular) * fresnel(eta, cosThetaI);
            }
            case B200PT_MAT_CONDUCTOR: return v3(fresnelConductor(dot(wi, normal), mat.eta, mat.k));
            case B200PT_MAT_SPECULAR: return v3(mat.specular);
            default: return v3(0.0f);
        }
    }

    v3 nextEventEstimation(const b200pt_material &mat, v3 origin, v3 wi, v3 normal, float tu, float tv) {   // rgen:601-731
        switch (mat.type) {
            case B200PT_MAT_DIFFUSE: case B200PT_MAT_PHONG: case B200PT_MAT_LIGHT: case B200PT_MAT_ROUGH_CONDUCTOR: break;
            default: return v3(0.0f);
        }
        v3 neeResult(0.0f);
        v3 lightDir, lightColor;
        float lightDistance;
        float pdfLights = sampleLights(origin, normal, lightDir, lightColor, lightDistance);
        isShadowed = true;
        float cosThetaLight = dot(normal, lightDir);
        if (cosThetaLight > 0 && pdfLights > 0) traceShadow(origin, tMin, lightDir, lightDistance * (1 - 0.0001f));
        if (!isShadowed) {
            if (pushC.enableMIS) {
                float pdfMat = pdfBSDF(mat, normal, wi, lightDir);
                float heuristic = pushC.usePowerHeuristic ? powerHeuristic(pdfLights, pdfMat) : balanceHeuristic(pdfLights, pdfMat);
                if (isnanf_(heuristic)) return neeResult;
                neeResult += evalBsdf(mat, tu, tv, normal, wi, lightDir, true) * lightColor * heuristic / pdfLights;
            } else {
                neeResult = evalBsdf(mat, tu, tv, normal, wi, lightDir, true) * lightColor / pdfLights;
            }
        }
        if (pushC.enableMIS) {
            v3 bsdfDir;
            float pdfMat = sampleBSDF(mat, wi, normal, info.isFrontFace, bsdfDir);
            if (pdfMat > 0) {
                traceClosest(origin, tMin, bsdfDir, tMax);     // overwrites the shared payload `info`
                if (info.isMiss) {
                    lightColor = info.missColor;
                    pdfLights = 1.0f / (2 * kPi) / float(C.lights.size());
                    float heuristic = pushC.usePowerHeuristic ? powerHeuristic(pdfMat, pdfLights) : balanceHeuristic(pdfMat, pdfLights);
                    neeResult += evalBsdf(mat, tu, tv, normal, wi, bsdfDir, true) * lightColor * heuristic / pdfMat;
                } else {
                    const b200pt_material &matSample = C.mats[info.matIndex];
                    if (matSample.type == B200PT_MAT_LIGHT) {
                        int iLight = info.isSphere ? C.spheres[info.instanceIndex].iLight : C.instanceInfos[info.instanceIndex].iLight;
                        if (iLight >= 0) {
                            const b200pt_light &light = C.lights[iLight];
                            lightColor = v3(matSample.lightColor);
                            pdfLights = pdfLight(light, bsdfDir, info.normal, info.t);
                            float heuristic = pushC.usePowerHeuristic ? powerHeuristic(pdfMat, pdfLights) : balanceHeuristic(pdfMat, pdfLights);
                            if (isnanf_(heuristic)) return neeResult;
                            neeResult += evalBsdf(mat, tu, tv, normal, wi, bsdfDir, true) * lightColor * heuristic / pdfMat;
                        }
                    }
                }
            }
        }
        return neeResult;
    }

    static bool hasDiscreteDirection(const b200pt_material &mat) {       // rgen:733-742
        return mat.type == B200PT_MAT_DIELECTRIC || mat.type == B200PT_MAT_SPECULAR || mat.type == B200PT_MAT_CONDUCTOR;
    }

    // ---- irradiance cache lookup: rgen:748-773 + raytrace.irradiance.rint:11-21 + .rahit:18-51 -------------------
    bool queryIrradianceCache(v3 origin, v3 normal, v3 &color) const {
        v3 cacheValueSum(0.0f);
        float totalWeight = 0;
        uint32_t n = std::min<uint32_t>(std::min<uint32_t>(uint32_t(C.snapSpheres.size()), C.snapHeader.maxCaches), C.snapHeader.nextCacheSlot);
        for (uint32_t i = 0; i < n; i++) {       // any-hit order is the driver's; ours: ascending cache index
            const b200pt_sphere &cs = C.snapSpheres[i];
            const b200pt_cache_data &cd = C.snapCache[i];
            v3 oc = origin - v3(cs.center);
            if (!(length(oc) <= cs.radius)) continue;          // .rint
            float weight = 1.0f / (length(origin - v3(cs.center)) / cd.harmonicR + sqrtf(1 - dot(normal, v3(cd.normal))));
            if (isnanf_(weight) || isinff_(weight)) weight = 1000000;
            bool vis = -0.001f <= dot(origin - v3(cs.center), (normal + v3(cd.normal)) / 2.0f);
            if (weight <= 1.0f / pushC.irradianceA || (pushC.irradianceCachePerformVisibilityCheck && !vis)) continue;
            if (pushC.useIrradianceGradients) {
                float E = length(v3(cd.color));
                v3 col = E != 0 ? normalize(v3(cd.color)) : v3(0.0f);
                v3 adjusted = col * (E + dot(cross(v3(cd.normal), normal), v3(cd.rotGrad)) + dot(origin - v3(cs.center), v3(cd.transGrad)));
                cacheValueSum += weight * adjusted;
            } else cacheValueSum += weight * v3(cd.color);
            totalWeight += weight;
        }
        if (totalWeight > 0) { color = cacheValueSum / totalWeight; return true; }
        return false;
    }
    bool isICCapable(const b200pt_material &mat) const {                 // rgen:806-813
        if (mat.type == B200PT_MAT_DIFFUSE || mat.type == B200PT_MAT_LIGHT) return true;
        if (pushC.useIrradianceCacheOnGlossy && !hasDiscreteDirection(mat)) return true;
        return false;
    }
    v3 approxDiffuse(const b200pt_material &mat, v3 normal, v3 wi, float u, float v) const {   // rgen:815-828
        switch (mat.type) {
            case B200PT_MAT_DIFFUSE: case B200PT_MAT_LIGHT: case B200PT_MAT_PHONG: return diffuse(mat, u, v);
            case B200PT_MAT_ROUGH_CONDUCTOR: { float c = dot(wi, normal); return v3(fresnelConductor(c, mat.eta, mat.k)) * 1.0f / kPi; }
        }
        return v3(-1, -1, -1);
    }
    float applyWeightWindow(v3 throughput, v3 adjoint, int &n) {         // rgen:830-869
        float center = length(estimate / adjoint);
        float lower = 2 * center / (1 + pushC.adrrsS);
        float upper = pushC.adrrsS * lower;
        float v = length(throughput);
        if (isnanf_(lower) || lower <= 0) { n = 1; return 1.0f; }
        if (lower <= v && v <= upper) { n = 1; return 1.0f; }
        else if (v <= lower) { n = 1; return std::max(v / lower, 0.1f); }
        float q = v / upper;
        n = int(q);
        if (rnd() > (n + 1 - q)) n++;
        return q;
    }
    v3 multipleNEE(const b200pt_material &mat, v3 origin, v3 wi, v3 normal, float u, float v, int numNEE) {   // rgen:871-877
        v3 result(0.0f);
        for (int i = 0; i < numNEE; i++) result += nextEventEstimation(mat, origin, wi, normal, u, v);
        return result / float(numNEE);
    }
    bool split(v3 origin, v3 normal, v3 wi, v3 throughput, float u, float v, int matIndex, int currentDepth, bool isFrontFace) {   // rgen:883-902
        if (nextSplitSlot >= MAX_SPLITS) return false;
        SplitInfo &s = splits[nextSplitSlot];
        s.origin = origin; s.normal = normal; s.wi = wi; s.throughput = throughput; s.u = u; s.v = v;
        s.matIndex = matIndex; s.currentDepth = currentDepth; s.isFrontFace = isFrontFace;
        nextSplitSlot++;
        return true;
    }

    // ---- guiding: rgen:904-960 + raytrace.guiding.rint:12-19 + guiding.glsl:32-96 --------------------------------
    uint32_t getGuidingRegion(v3 origin) const {
        for (size_t i = 0; i < C.guidingAabbs.size(); i++) {   // TerminateOnFirstHit: regions are disjoint up to shared faces → lowest index
            const b200pt_aabb &bb = C.guidingAabbs[i];
            bool in = true;
            const float o[3] = {origin.x, origin.y, origin.z};
            for (int a = 0; a < 3; a++) if (!(std::min(bb.min[a], o[a]) == bb.min[a] && std::max(bb.max[a], o[a]) == bb.max[a])) in = false;
            if (in) return uint32_t(i);
        }
        return 0xFFFFFFFFu;
    }
    static float vMF(v3 wo, const b200pt_vmf_theta &th, v3 worldPos, bool parallax) {   // guiding.glsl:32-44
        if (th.k == 0.0f) return 0.07957747155f;
        v3 mu(th.mu);
        if (parallax && th.distance > 0) mu = normalize(v3(th.target) - worldPos);
        return th.norm * dm::expF(th.k * (dot(mu, wo) - 1));
    }
    static float VMM(v3 wo, const b200pt_vmm_theta &vmm, v3 worldPos, bool parallax) {   // guiding.glsl:53-60
        float res = 0;
        for (int i = 0; i < vmm.usedDistributions; i++) res += vmm.pi[i] * vMF(wo, vmm.thetas[i], worldPos, parallax);
        return res;
    }
    v3 sampleVMF(const b200pt_vmf_theta &th, v3 worldPos, bool parallax) {               // guiding.glsl:62-82
        if (th.k > 0.0f) {
            const float r1 = rnd();
            const float r2 = rnd();
            const float cosTheta = 1.0f + dm::logF(1 + th.eMin2K * r1 - r1) / th.k;
            const float sinTheta = 1.0f - cosTheta * cosTheta <= 0.0f ? 0.0f : sqrtf(1.0f - cosTheta * cosTheta);
            const float phi = 2.f * kPi * r2;
            const float cosPhi = dm::cosF(phi), sinPhi = dm::sinF(phi);
            v3 mu(th.mu);
            if (parallax && th.distance > 0) mu = normalize(v3(th.target) - worldPos);
            return toWorld(v3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta), mu);
        }
        return randomOnUnitSphere();
    }
    v3 sampleVMM(const b200pt_vmm_theta &vmm, v3 worldPos, bool parallax) {              // guiding.glsl:84-96
        float rndDist = rnd();
        int iDistribution = 0, maxDistribution = vmm.usedDistributions - 1;
        float piSum = vmm.pi[0];
        while (piSum < rndDist && iDistribution < maxDistribution) { iDistribution++; piSum += vmm.pi[iDistribution]; }
        return sampleVMF(vmm.thetas[iDistribution], worldPos, parallax);
    }
    float getNewDirection(const b200pt_material &mat, v3 origin, v3 normal, v3 wi, v3 &newDirection) {   // rgen:923-960
        float pdf;
        if (pushC.useGuiding && !hasDiscreteDirection(mat)) {
            uint32_t iRegion = getGuidingRegion(origin);
            if (iRegion == 0xFFFFFFFFu) return 0.0f;
            const b200pt_vmm_theta &vmmTheta = C.guidingVMM[iRegion];
            float pdfMat;
            if (rnd() < pushC.guidingProb) {
                newDirection = sampleVMM(vmmTheta, origin, pushC.useParallaxCompensation != 0);
                pdfMat = pdfBSDF(mat, normal, wi, newDirection);
            } else pdfMat = sampleBSDF(mat, wi, normal, info.isFrontFace, newDirection);
            if (dot(newDirection, normal) < 0 || pdfMat <= 0) return 0.0f;
            float pdfGuiding = VMM(newDirection, vmmTheta, origin, pushC.useParallaxCompensation != 0);
            if (isnanf_(pdfGuiding)) return sampleBSDF(mat, wi, normal, info.isFrontFace, newDirection);
            pdf = mixf(pdfMat, pdfGuiding, pushC.guidingProb);
        } else pdf = sampleBSDF(mat, wi, normal, info.isFrontFace, newDirection);
        return pdf;
    }
    static bool isMatAlmostDiscrete(const b200pt_material &mat) {        // rgen:962-964
        return (mat.type == B200PT_MAT_ROUGH_CONDUCTOR && mat.roughness <= 0.3f) || (mat.type == B200PT_MAT_PHONG && mat.specularHighlight >= 250.0f);
    }
    uint32_t getBaseIndex() const { return (py * uint32_t(C.width) + px) * MAXD; }   // rgen:966-968
    void updateSamples(int currentSampleOffset, v3 light) {              // rgen:973-978
        for (int i = currentSampleOffset - 1; i >= sampleOffset; i--) { lightSums[i] += light; light *= sampleThroughputs[i]; }
    }
    void commitSamples(int currentSampleOffset) {                        // rgen:980-990
        for (int i = sampleOffset; i < currentSampleOffset; i++) {
            b200pt_directional_data &dd = C.directionalData[getBaseIndex() + i];
            float weight = length(lightSums[i]) / dd.pdf;
            if (weight <= 0) dd.flags = B200PT_INVALID_REGION;
            else dd.weight = weight;
        }
    }

    // ---- the bounce loop: rgen:992-1226 ----------------------------------------------------------------------------
    v3 raytrace(v3 origin, v3 direction, v3 throughput, int &currentDepth, int maxDepth, int maxFollowDiscrete, bool addDirectLights,
                bool addFirstHitLight, bool useNEE, int numNEE, bool useIC, bool createIC, bool useADRRS, bool saveSamples, float &firstT) {
        bool follow = true;
        int followCount = 0;
        bool addNextDirectLights = addFirstHitLight;
        int depth = currentDepth;
        v3 normal;
        firstT = tMax;
        int currentSampleOffset = sampleOffset;
        int iUpdateDistance = -1;
        float distanceFactor = 1.0f;
        v3 result(0.0f);
        do {
            depth++;
            traceClosest(origin, tMin, direction, tMax);
            if (info.isMiss) {
                if (addDirectLights || addNextDirectLights) {
                    result += throughput * info.missColor;
                    if (saveSamples) {
                        updateSamples(currentSampleOffset, info.missColor);
                        if (iUpdateDistance != -1) C.directionalData[getBaseIndex() + iUpdateDistance].distance = 0;
                    }
                }
                break;
            }
            pathVertices++;
            const b200pt_material mat = C.mats[info.matIndex];
            origin = info.worldPos;
            normal = info.normal;
            float tu = info.u, tv = info.v;
            if (depth == 1) firstT = info.t;
            if (mat.type == B200PT_MAT_LIGHT && (addDirectLights || addNextDirectLights)) {
                result += throughput * v3(mat.lightColor);
                if (saveSamples) updateSamples(currentSampleOffset, v3(mat.lightColor));
            }
            if (hasDiscreteDirection(mat)) {
                addNextDirectLights = true;
                follow = true;
                if (depth >= maxDepth) followCount++;
                if (saveSamples && iUpdateDistance != -1) C.directionalData[getBaseIndex() + iUpdateDistance].distance += info.t * distanceFactor;
            } else {
                addNextDirectLights = false;
                follow = false;
                if (useIC || (useADRRS && (depth > 1))) {
                    if (isICCapable(mat)) {
                        v3 irradianceColor;
                        if (queryIrradianceCache(origin, normal, irradianceColor)) {
                            v3 diff = approxDiffuse(mat, normal, -direction, tu, tv);
                            if (useADRRS) {
                                int n;
                                float q = applyWeightWindow(throughput, irradianceColor, n);
                                if (n == 1) {
                                    if (rnd() > q) break;
                                    throughput /= q;
                                } else if (pushC.adrrsSplit) {
                                    int possibleSplits = MAX_SPLITS - nextSplitSlot;
                                    if (n - 1 > possibleSplits) { n = possibleSplits + 1; q = float(n); }
                                    throughput /= q;
                                    for (int iSplit = 0; iSplit < n - 1; iSplit++) split(origin, normal, -direction, throughput, tu, tv, info.matIndex, depth, info.isFrontFace);
                                }
                            } else {
                                result += throughput * diff * irradianceColor;
                                result += throughput * multipleNEE(mat, origin, -direction, normal, tu, tv, numNEE);
                                break;
                            }
                        } else if (createIC && rnd() < pushC.irradianceCreateProb) {
                            if (nextNewIrradianceCacheSlot < MAX_NEW_IRRADIANCE_ENTRIES) {
                                newIrradianceCacheEntries[nextNewIrradianceCacheSlot].origin = origin;
                                newIrradianceCacheEntries[nextNewIrradianceCacheSlot].normal = normal;
                                nextNewIrradianceCacheSlot++;
                            }
                        }
                    }
                }
                if (pushC.splitOnFirst && depth == 1) {
                    if (split(origin, normal, -direction, throughput * 0.5f, tu, tv, info.matIndex, depth, info.isFrontFace)) throughput *= 0.5f;
                }
                if (saveSamples && iUpdateDistance != -1) {
                    C.directionalData[getBaseIndex() + iUpdateDistance].distance += info.t * distanceFactor;
                    if (!isMatAlmostDiscrete(mat)) iUpdateDistance = -1;
                }
                if (useNEE) {
                    v3 neeLight = multipleNEE(mat, origin, -direction, normal, tu, tv, numNEE);
                    result += throughput * neeLight;
                    if (saveSamples) updateSamples(currentSampleOffset, neeLight);
                }
            }
            v3 newDirection;
            float pdf = getNewDirection(mat, origin, normal, -direction, newDirection);
            if (pdf <= 0.0f) break;
            v3 throughputChange = evalBsdf(mat, tu, tv, normal, -direction, newDirection, info.isFrontFace) / pdf;
            throughput *= throughputChange;
            if (saveSamples && currentSampleOffset < MAXD) {
                if (hasDiscreteDirection(mat) || isMatAlmostDiscrete(mat)) {
                    C.directionalData[getBaseIndex() + currentSampleOffset].flags = B200PT_INVALID_REGION;
                } else {
                    b200pt_directional_data sd;
                    sd.position[0] = origin.x; sd.position[1] = origin.y; sd.position[2] = origin.z;
                    sd.direction[0] = newDirection.x; sd.direction[1] = newDirection.y; sd.direction[2] = newDirection.z;
                    sd.pdf = pdf; sd.weight = 0; sd.distance = 0;
                    sd.flags = getGuidingRegion(origin);
                    C.directionalData[getBaseIndex() + currentSampleOffset] = sd;
                    iUpdateDistance = currentSampleOffset;
                    distanceFactor = 1.0f;
                }
                lightSums[currentSampleOffset] = v3(0.0f);
                sampleThroughputs[currentSampleOffset] = throughputChange;
                currentSampleOffset++;
            } else if (iUpdateDistance != -1 && mat.type == B200PT_MAT_DIELECTRIC && dot(newDirection, normal) < 0) {
                float eta = mat.refractionIndex;
                if (!info.isFrontFace) eta = mat.refractionIndexInv;
                distanceFactor = fabsf(dot(normal, direction) / dot(normal, newDirection)) * eta;
            }
            direction = newDirection;
        } while (depth <= maxDepth || (follow && followCount <= maxFollowDiscrete));
        commitSamples(currentSampleOffset);
        if (length(result) > 0) sampleOffset += currentSampleOffset;     // sic: += (rgen:1221)
        currentDepth = depth;
        return result;
    }

    // ---- irradiance cache build: rgen:1231-1421 ----------------------------------------------------------------------
    float calculateCacheData(v3 origin, v3 normal, v3 &calculatedColor, v3 &rotGrad, v3 &transGrad) {
        const int N = 20, M = 10;
        const float M_HALF_PI = kPi / 2.0f;
        float invDistanceSum = 0;
        int numDistances = 0;
        int maxFollowDiscrete = 10;
        bool createIC = false;
        const bool addDirectLights = false;
        rotGrad = v3(0.0f); transGrad = v3(0.0f);
        v3 color(0.0f);
        float previousKLs[N] = {0}, previousKRs[N] = {0};   // sic: declared inside the k loop in GLSL (values persist in practice)
        for (int k = 0; k < N; k++) {
            float phi = 2 * kPi * (k + rnd()) / N;
            v3 uk = toWorld(sphericalToCartesian(M_HALF_PI, phi), normal);
            v3 vk = toWorld(sphericalToCartesian(M_HALF_PI, phi + M_HALF_PI), normal);
            v3 previousVk = toWorld(sphericalToCartesian(M_HALF_PI, 2 * kPi * k / N + M_HALF_PI), normal);
            float previousJR = 0, previousJL = 0;
            for (int j = 0; j < M; j++) {
                float theta = dm::asinF(sqrtf((j + rnd()) / M));
                v3 direction = toWorld(sphericalToCartesian(theta, phi), normal);
                float r = tMax;
                int currentDepth = 0;
                v3 sampleColor = raytrace(origin, direction, v3(1.0f), currentDepth, 1, maxFollowDiscrete, addDirectLights, false, true,
                                          pushC.irradianceNumNEE, true, createIC, false, false, r);
                color += sampleColor;
                if (r < tMax) { invDistanceSum += 1.0f / r; numDistances++; }
                float L = length(sampleColor);
                float previousJTheta = dm::asinF(sqrtf(j / float(M)));
                float nextJTheta = dm::asinF(sqrtf((j + 1) / float(M)));
                float tanTheta = dm::tanF(theta);
                if (isinff_(tanTheta) || isnanf_(tanTheta)) tanTheta = 0;
                rotGrad -= tanTheta * L * vk;
                if (j > 0) {
                    float cosPreviousTheta = dm::cosF(previousJTheta);
                    transGrad += uk * 2 * kPi / N * dm::sinF(previousJTheta) * cosPreviousTheta * cosPreviousTheta / std::min(r, previousJR) * (L - previousJL);
                }
                if (k > 0) transGrad += previousVk * (dm::sinF(nextJTheta) - dm::sinF(previousJTheta)) / std::min(r, previousKRs[j]) * (L - previousKLs[j]);
                previousKLs[j] = L; previousKRs[j] = r; previousJL = L; previousJR = r;
            }
        }
        float normFactor = kPi / (M * N);
        calculatedColor = normFactor * color;
        rotGrad *= normFactor;
        if (invDistanceSum == 0 || numDistances == 0) return -1;
        return 1.0f / (invDistanceSum / numDistances);
    }
    static void clampGradients(float maxLength, v3 &rotGrad, v3 &transGrad) {   // rgen:1318-1329
        float lr = length(rotGrad);
        if (lr > maxLength) rotGrad *= maxLength / lr;
        float lt = length(transGrad);
        if (lt > maxLength) transGrad *= maxLength / lt;
    }
    static void st3(float *d, v3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
    int regionIndex() const { return int(py - uint32_t(C.regionY0)) * C.regionNX + int(px - uint32_t(C.regionX0)); }
    void updateIrradianceCache() {                                       // rgen:1334-1381; slot handed out in pixel order, blend at commit
        const int cacheIndex = C.updateSlotOfPixel[regionIndex()];
        if (cacheIndex < 0) return;
        oracle_ctx::PendingEntry &pe = C.pendingUpdate[regionIndex()];
        v3 rotGrad, transGrad, calculatedColor;
        pe.harmonicR = calculateCacheData(v3(C.snapSpheres[cacheIndex].center), v3(C.snapCache[cacheIndex].normal), calculatedColor, rotGrad, transGrad);
        st3(pe.color, calculatedColor); st3(pe.rotGrad, rotGrad); st3(pe.transGrad, transGrad);
        pe.valid = true;
    }
    bool createIrradianceCache(v3 origin, v3 normal, v3 &calculatedColor) {   // rgen:1383-1421; slot handed out at commit
        if (C.snapHeader.nextCacheSlot > C.snapHeader.maxCaches) return false;
        oracle_ctx::PendingEntry pe;
        v3 rotGrad, transGrad;
        pe.harmonicR = calculateCacheData(origin, normal, calculatedColor, rotGrad, transGrad);
        pe.valid = !(pe.harmonicR < 0);
        st3(pe.origin, origin); st3(pe.normal, normal); st3(pe.color, calculatedColor); st3(pe.rotGrad, rotGrad); st3(pe.transGrad, transGrad);
        C.pendingCreate[regionIndex()].push_back(pe);
        return pe.valid;
    }

    void getCameraRay(float pcx, float pcy, v3 &origin, v3 &direction) {  // rgen:1487-1494
        float r1 = getRandomNegPos(), r2 = getRandomNegPos();
        float ux = (pcx + r1 / 2.0f) / float(C.width), uy = (pcy + r2 / 2.0f) / float(C.height);
        float dx = ux * 2.0f - 1.0f, dy = uy * 2.0f - 1.0f;
        origin = mulPoint(C.viewInverse, v3(0.0f), 1.0f);
        v3 target = mulPoint(C.projInverse, v3(dx, dy, 1.0f), 1.0f);
        direction = normalize(mulPoint(C.viewInverse, normalize(target), 0.0f));
    }

    void saveResult(v3 result) {                                         // rgen:1459-1485
        size_t p = (size_t(py) * C.width + px) * 4;
        float *img = &C.image[p], *acc = &C.accumulateImage[p];
        if (pushC.enableAverageInsteadOfMix) {
            if (pushC.previousFrames > 0) {
                v3 accumulated = v3(acc) + result;
                st3(img, accumulated / float(pushC.previousFrames + 1u)); img[3] = 1;
                st3(acc, accumulated); acc[3] = 1;
            } else { st3(acc, result); acc[3] = 1; st3(img, result); img[3] = 1; }
        } else {
            if (pushC.previousFrames > 0) result = mix(v3(img), result, 1.0f / float(pushC.previousFrames + 1u));
            st3(img, result); img[3] = 1;
        }
    }

    // rgen main(): 1623-1827 (visualisation modes other than VISU_RAYTRACE are out of scope)
    void run(uint32_t x, uint32_t y) {
        px = x; py = y;
        seed = tea(py * uint32_t(C.width) + px, pushC.randomUInt);
        v3 result(0.0f);
        const float pcx = float(px) + 0.5f, pcy = float(py) + 0.5f;
        if (pushC.useADRRS) {
            estimate = v3(&C.estimateImage[(size_t(py) * C.width + px) * 4]);
            lengthEstimate = length(estimate);
        }
        if (pushC.useIrradianceCache && rnd() < pushC.irradianceUpdateProb) updateIrradianceCache();
        if (pushC.updateGuiding) {
            for (int i = 0; i < MAXD; i++) C.directionalData[getBaseIndex() + i].flags = B200PT_INVALID_REGION;
            if ((pushC.useADRRS && pushC.adrrsSplit) || pushC.splitOnFirst) return;
        }
        const int numNEE = pushC.numNEE, maxDepth = pushC.maxDepth;
        int maxReachedDepth = 0, depthSum = 0, depthsCounter = 0;            // rgen:1653-1655
        float firstT;
        if (pushC.visualizeMode == 0 || pushC.storeEstimate) {
            const int samplesPerPixel = pushC.isIrradiancePrepareFrame ? 1 : pushC.samplesPerPixel;
            const int maxFollowDiscrete = pushC.maxFollowDiscrete;
            const bool useNEE = pushC.enableNEE != 0;
            const bool addDirectLights = !useNEE;
            const bool useIC = pushC.useIrradianceCache != 0;
            for (int iSample = 0; iSample < samplesPerPixel; ++iSample) {
                v3 origin, direction;
                getCameraRay(pcx, pcy, origin, direction);
                int currentDepth = 0;
                result += raytrace(origin, direction, v3(1.0f), currentDepth, maxDepth, maxFollowDiscrete, addDirectLights, true, useNEE, numNEE,
                                   useIC, true, pushC.useADRRS != 0, pushC.updateGuiding != 0, firstT);
                maxReachedDepth = std::max(maxReachedDepth, currentDepth); depthSum += currentDepth; depthsCounter++;      // rgen:1677-1679
            }
            for (int iSplit = 0; iSplit < nextSplitSlot; iSplit++) {
                SplitInfo sp = splits[iSplit];
                const b200pt_material mat = C.mats[sp.matIndex];
                if (useNEE) result += sp.throughput * multipleNEE(mat, sp.origin, sp.wi, sp.normal, sp.u, sp.v, numNEE);
                v3 direction;
                float pdf = getNewDirection(mat, sp.origin, sp.normal, sp.wi, direction);
                if (pdf <= 0.0f) break;
                v3 throughput = sp.throughput * evalBsdf(mat, sp.u, sp.v, sp.normal, sp.wi, direction, sp.isFrontFace) / pdf;
                result += raytrace(sp.origin, direction, throughput, sp.currentDepth, maxDepth, maxFollowDiscrete, addDirectLights, false, useNEE,
                                   numNEE, useIC, true, pushC.useADRRS != 0, pushC.updateGuiding != 0, firstT);
                maxReachedDepth = std::max(maxReachedDepth, sp.currentDepth); depthSum += sp.currentDepth; depthsCounter++;  // rgen:1705-1707
            }
            result /= float(samplesPerPixel);
        }
        { float *a = &C.aovImage[(size_t(py) * C.width + px) * 4]; a[0] = float(maxReachedDepth); a[1] = float(depthSum); a[2] = float(depthsCounter); a[3] = float(nextSplitSlot); }
        if (pushC.storeEstimate) { float *e = &C.estimateImage[(size_t(py) * C.width + px) * 4]; st3(e, result); e[3] = 1; }
        if (pushC.visualizeMode == 0) saveResult(result);
        for (int i = 0; i < nextNewIrradianceCacheSlot; i++) {
            v3 color;
            createIrradianceCache(newIrradianceCacheEntries[i].origin, newIrradianceCacheEntries[i].normal, color);
        }
    }
};

float srgbToLinear(uint8_t v) {
    float c = float(v) / 255.0f;
    return c <= 0.04045f ? c / 12.92f : dm::powF((c + 0.055f) / 1.055f, 2.4f);
}

int buildAccel(oracle_ctx &C, uint32_t first, uint32_t count, const std::vector<float> &lo, const std::vector<float> &hi, float pad, int depth) {
    int idx = int(C.nodes.size());
    C.nodes.emplace_back();
    AccelNode n;
    for (int a = 0; a < 3; a++) { n.lo[a] = std::numeric_limits<float>::infinity(); n.hi[a] = -n.lo[a]; }
    for (uint32_t i = first; i < first + count; i++)
        for (int a = 0; a < 3; a++) { n.lo[a] = std::min(n.lo[a], lo[3 * C.order[i] + a]); n.hi[a] = std::max(n.hi[a], hi[3 * C.order[i] + a]); }
    n.first = first; n.count = count; n.left = n.right = -1;
    if (count > 4 && depth < 60) {
        int axis = 0;
        float ext[3] = {n.hi[0] - n.lo[0], n.hi[1] - n.lo[1], n.hi[2] - n.lo[2]};
        if (ext[1] > ext[axis]) axis = 1;
        if (ext[2] > ext[axis]) axis = 2;
        uint32_t mid = first + count / 2;
        std::nth_element(C.order.begin() + first, C.order.begin() + mid, C.order.begin() + first + count,
                         [&](uint32_t a, uint32_t b) { return lo[3 * a + axis] + hi[3 * a + axis] < lo[3 * b + axis] + hi[3 * b + axis]; });
        int l = buildAccel(C, first, mid - first, lo, hi, pad, depth + 1);
        int r = buildAccel(C, mid, first + count - mid, lo, hi, pad, depth + 1);
        n.left = l; n.right = r;
    }
    for (int a = 0; a < 3; a++) { n.lo[a] -= pad; n.hi[a] += pad; }
    C.nodes[idx] = n;
    return idx;
}

}  // namespace

extern "C" {

oracle_ctx *oracle_create(int width, int height, int ic_size, int use_accel) {
    oracle_ctx *C = new oracle_ctx();
    C->width = width; C->height = height;
    size_t N = size_t(width) * height;
    C->image.assign(N * 4, 0.0f); C->accumulateImage.assign(N * 4, 0.0f); C->estimateImage.assign(N * 4, 0.0f); C->aovImage.assign(N * 4, 0.0f);
    C->header.maxCaches = uint32_t(ic_size);
    C->cache.assign(size_t(ic_size), b200pt_cache_data{});
    C->cacheSpheres.assign(size_t(ic_size), b200pt_sphere{});
    b200pt_directional_data z;
    memset(&z, 0, sizeof(z));
    C->directionalData.assign(N * MAXD, z);
    C->useAccel = use_accel != 0;
    memset(&C->pushC, 0, sizeof(C->pushC));
    return C;
}
void oracle_destroy(oracle_ctx *C) { delete C; }

int oracle_set_scene(oracle_ctx *C, const b200pt_scene_desc *s) {
    C->vertices.clear(); C->indices.clear();
    for (int m = 0; m < s->num_models; m++) {
        C->vertices.emplace_back(s->vertices[m], s->vertices[m] + s->num_vertices[m]);
        C->indices.emplace_back(s->indices[m], s->indices[m] + s->num_indices[m]);
    }
    C->mats.assign(s->materials, s->materials + s->num_materials);
    C->instanceInfos.assign(s->instances, s->instances + s->num_instances);
    C->lights.assign(s->lights, s->lights + s->num_lights);
    C->randomLightIndex.assign(s->random_light_index, s->random_light_index + B200PT_SIZE_LIGHT_RANDOM);
    C->numFaceTables = s->num_face_tables;
    C->randomTriIndex.assign(s->random_tri_index, s->random_tri_index + size_t(std::max(1, s->num_face_tables)) * B200PT_SIZE_TRI_RANDOM);
    C->spheres.assign(s->spheres, s->spheres + s->num_spheres);
    C->textures.clear();
    for (int t = 0; t < s->num_textures; t++) {
        Tex tx;
        tx.w = s->textures[t].width; tx.h = s->textures[t].height;
        size_t np = size_t(tx.w) * tx.h;
        tx.px.resize(np * 4);
        if (s->textures[t].format == B200PT_TEX_RGBA32F) memcpy(tx.px.data(), s->textures[t].pixels, np * 16);
        else {
            const uint8_t *p = static_cast<const uint8_t *>(s->textures[t].pixels);
            for (size_t i = 0; i < np; i++) {
                for (int k = 0; k < 3; k++) tx.px[4 * i + k] = srgbToLinear(p[4 * i + k]);
                tx.px[4 * i + 3] = float(p[4 * i + 3]) / 255.0f;
            }
        }
        C->textures.push_back(std::move(tx));
    }
    for (auto &t : C->textures) { t.hasAlpha = false; for (size_t p = 3; p < t.px.size(); p += 4) if (t.px[p] < 1.0f) { t.hasAlpha = true; break; } }
    C->anyTextured = false;
    for (auto &m : C->mats) if (m.textureIdDiffuse != -1 && m.textureIdDiffuse < int(C->textures.size()) && C->textures[m.textureIdDiffuse].hasAlpha) C->anyTextured = true;
    // world-space triangles: same formula and operation order as the product's host code (IEEE, no contraction)
    C->tri.clear(); C->primInstance.clear(); C->primLocal.clear();
    std::vector<float> lo, hi;
    for (int i = 0; i < s->num_instances; i++) {
        const b200pt_instance &inst = s->instances[i];
        int m = inst.modelIndex;
        int nt = s->num_indices[m] / 3;
        for (int t = 0; t < nt; t++) {
            float w[3][3];
            for (int k = 0; k < 3; k++) {
                const float *p = s->vertices[m][s->indices[m][3 * t + k]].pos;
                const float *M = inst.transform;
                for (int r = 0; r < 3; r++) w[k][r] = ((M[0 + r] * p[0] + M[4 + r] * p[1]) + M[8 + r] * p[2]) + M[12 + r];
            }
            for (int a = 0; a < 3; a++) C->tri.push_back(w[0][a]);
            for (int a = 0; a < 3; a++) C->tri.push_back(w[1][a] - w[0][a]);
            for (int a = 0; a < 3; a++) C->tri.push_back(w[2][a] - w[0][a]);
            for (int a = 0; a < 3; a++) { lo.push_back(std::min(w[0][a], std::min(w[1][a], w[2][a]))); hi.push_back(std::max(w[0][a], std::max(w[1][a], w[2][a]))); }
            C->primInstance.push_back(uint32_t(i)); C->primLocal.push_back(uint32_t(t));
        }
    }
    C->numTris = uint32_t(C->primInstance.size());
    C->nodes.clear(); C->order.clear();
    if (C->useAccel && C->numTris) {
        C->order.resize(C->numTris);
        for (uint32_t i = 0; i < C->numTris; i++) C->order[i] = i;
        float maxAbs = 1e-3f;
        for (float f : lo) maxAbs = std::max(maxAbs, fabsf(f));
        for (float f : hi) maxAbs = std::max(maxAbs, fabsf(f));
        buildAccel(*C, 0, C->numTris, lo, hi, maxAbs * 1e-4f, 0);
    }
    return 0;
}

void oracle_set_camera(oracle_ctx *C, const float view[16], const float proj[16], const float viewInv[16], const float projInv[16]) {
    memcpy(C->view, view, 64); memcpy(C->proj, proj, 64); memcpy(C->viewInverse, viewInv, 64); memcpy(C->projInverse, projInv, 64);
}

// render the pixels [x0,x1) x [y0,y1) of one frame (one raygen invocation each).  Pixels are independent: irradiance-
// cache lookups read the frame-start snapshot, cache writes are deferred and committed in pixel order afterwards.
int oracle_render_region(oracle_ctx *C, const b200pt_push_constants *pc, int x0, int y0, int x1, int y1, int num_threads) {
    C->pushC = *pc;
    int nx = x1 - x0, ny = y1 - y0;
    if (nx <= 0 || ny <= 0) return 0;
    uint64_t ext = 0, sh = 0, pv = 0;
    const int total = nx * ny;
    const int nthreads = num_threads < 1 ? 1 : num_threads;
    // ---- frame start: snapshot + update slots in pixel order (rgen:1334-1343) ----
    C->snapHeader = C->header; C->snapCache = C->cache; C->snapSpheres = C->cacheSpheres;
    C->regionX0 = x0; C->regionY0 = y0; C->regionNX = nx;
    C->updateSlotOfPixel.assign(size_t(total), -1);
    C->pendingUpdate.assign(size_t(total), oracle_ctx::PendingEntry{});
    C->pendingCreate.assign(size_t(total), {});
    if (pc->useIrradianceCache) {
        for (int i = 0; i < total; i++) {
            uint32_t x = uint32_t(x0 + i % nx), y = uint32_t(y0 + i / nx);
            uint32_t seed = tea(y * uint32_t(C->width) + x, pc->randomUInt);
            if (!(Pixel::rndS(seed) < pc->irradianceUpdateProb)) continue;
            uint32_t cacheIndex = C->header.nextUpdateSlot++;
            if (cacheIndex >= C->snapHeader.nextCacheSlot) { C->header.nextUpdateSlot = 0; continue; }
            if (cacheIndex >= C->snapCache.size() || C->snapCache[cacheIndex].numUpdates < 1) continue;
            C->updateSlotOfPixel[size_t(i)] = int32_t(cacheIndex);
        }
    }
    std::atomic<int> next(0);
    std::mutex mu;
    auto worker = [&]() {
        uint64_t e = 0, s2 = 0, v = 0;
        for (;;) {
            int begin = next.fetch_add(16);
            if (begin >= total) break;
            for (int i = begin; i < std::min(total, begin + 16); i++) {
                Pixel p(*C);
                p.run(uint32_t(x0 + i % nx), uint32_t(y0 + i / nx));
                e += p.extendRays; s2 += p.shadowRays; v += p.pathVertices;
            }
        }
        std::lock_guard<std::mutex> lock(mu);
        ext += e; sh += s2; pv += v;
    };
    if (nthreads == 1) worker();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; t++) pool.emplace_back(worker);
        for (auto &t : pool) t.join();
    }
    // ---- frame end: commit in pixel order ----
    for (int i = 0; i < total; i++) {                       // updateIrradianceCache, rgen:1351-1380 (blend against the snapshot)
        const oracle_ctx::PendingEntry &pe = C->pendingUpdate[size_t(i)];
        if (!pe.valid) continue;
        const int cacheIndex = C->updateSlotOfPixel[size_t(i)];
        const b200pt_cache_data &old = C->snapCache[size_t(cacheIndex)];
        b200pt_cache_data &cd = C->cache[size_t(cacheIndex)];
        uint32_t numUpdates = old.numUpdates;
        float a = std::min(numUpdates / float(numUpdates + 1), 0.95f);
        Pixel::st3(cd.color, mix(v3(pe.color), v3(old.color), a));
        v3 rotGrad = mix(v3(pe.rotGrad), v3(old.rotGrad), a);
        v3 transGrad = mix(v3(pe.transGrad), v3(old.transGrad), a);
        Pixel::clampGradients(pc->irradianceGradientsMaxLength, rotGrad, transGrad);
        Pixel::st3(cd.rotGrad, rotGrad); Pixel::st3(cd.transGrad, transGrad);
        float harmonicR = pe.harmonicR;
        if (harmonicR > 0) harmonicR = mixf(harmonicR, old.harmonicR, a);
        else harmonicR = old.harmonicR;
        harmonicR = std::max(harmonicR, pc->irradianceCacheMinRadius);
        cd.harmonicR = harmonicR;
        C->cacheSpheres[size_t(cacheIndex)].radius = pc->irradianceA * harmonicR;
        cd.numUpdates = numUpdates + 1;
    }
    for (int i = 0; i < total; i++) {                       // createIrradianceCache, rgen:1394-1420
        for (const oracle_ctx::PendingEntry &pe : C->pendingCreate[size_t(i)]) {
            if (!pe.valid) continue;
            if (C->header.nextCacheSlot > C->header.maxCaches) continue;
            uint32_t cacheIndex = C->header.nextCacheSlot++;
            if (cacheIndex > C->header.maxCaches) continue;
            if (cacheIndex >= C->cache.size()) continue;    // the reference writes one element past the buffer here (quirk 11)
            float harmonicR = std::max(pe.harmonicR, pc->irradianceCacheMinRadius);
            v3 rotGrad(pe.rotGrad), transGrad(pe.transGrad);
            Pixel::clampGradients(pc->irradianceGradientsMaxLength, rotGrad, transGrad);
            b200pt_cache_data &cd = C->cache[cacheIndex];
            memcpy(cd.normal, pe.normal, 12); memcpy(cd.color, pe.color, 12);
            cd.harmonicR = harmonicR;
            Pixel::st3(cd.rotGrad, rotGrad); Pixel::st3(cd.transGrad, transGrad);
            cd.numUpdates = 1;
            memcpy(C->cacheSpheres[cacheIndex].center, pe.origin, 12);
            C->cacheSpheres[cacheIndex].radius = pc->irradianceA * harmonicR;
        }
    }
    C->extendRays += ext; C->shadowRays += sh; C->pathVertices += pv;
    return 0;
}

int oracle_trace_rays(oracle_ctx *C, const b200pt_ray *rays, int64_t n, b200pt_hit *hits, int any_hit, int num_threads) {
    const int nthreads = num_threads < 1 ? 1 : num_threads;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        for (;;) {
            int64_t begin = next.fetch_add(256);
            if (begin >= n) break;
            for (int64_t i = begin; i < std::min<int64_t>(n, begin + 256); i++) {
                Pixel p(*C);
                Pixel::Cand c = p.traverse(v3(rays[i].origin), v3(rays[i].dir), rays[i].tmin, rays[i].tmax, any_hit != 0);
                hits[i].t = c.t; hits[i].prim = c.prim; hits[i].u = c.u; hits[i].v = c.v;
            }
        }
    };
    if (nthreads == 1) worker();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; t++) pool.emplace_back(worker);
        for (auto &t : pool) t.join();
    }
    return 0;
}

// One ray through the oracle's traversal, and one texture fetch, as plain C calls: oracle/shader_ref.cpp (the reference's shader
// megakernel compiled as C++) gets its hits and texels from here — intersection order, tie rule and the sampler's filter are
// defined by us (the reference leaves them to the driver), everything done WITH a hit or a texel is the reference's code.
// out: t, u, v;  ids: global primitive, instance (or sphere index), primitive inside the instance, is-sphere.  Returns 1 on a hit.
int oracle_trace_one(oracle_ctx *C, const float o[3], const float d[3], float tmin, float tmax, int any_hit, float out[3], uint32_t ids[4]) {
    Pixel p(*C);
    Pixel::Cand c = p.traverse(v3(o), v3(d), tmin, tmax, any_hit != 0);
    if (c.prim == B200PT_MISS) return 0;
    out[0] = c.t; out[1] = c.u; out[2] = c.v;
    ids[0] = c.prim;
    if (c.prim >= C->numTris) { ids[1] = c.prim - C->numTris; ids[2] = c.prim - C->numTris; ids[3] = 1u; }
    else { ids[1] = C->primInstance[c.prim]; ids[2] = C->primLocal[c.prim]; ids[3] = 0u; }
    return 1;
}
// the oracle's any-hit decision (raytrace.rahit) for one candidate hit on global primitive `prim`
int oracle_alpha_rejects(oracle_ctx *C, uint32_t prim, float u, float v, float origin_x, float t, uint32_t random_uint) {
    C->pushC.randomUInt = random_uint;
    Pixel p(*C);
    return p.alphaRejects(prim, u, v, v3(origin_x, 0.0f, 0.0f), t) ? 1 : 0;
}
void oracle_texture(oracle_ctx *C, int id, float u, float v, float out[4]) { Pixel p(*C); p.textureRGBA(id, u, v, out); }
uint32_t oracle_tea(uint32_t a, uint32_t b) { return tea(a, b); }
// Unit access to the restated shader functions of random.glsl / transform.glsl / guiding.glsl, for the comparison with the
// reference's own files compiled as C++ (oracle/glsl_ref.cpp: same function numbers and argument layout).
//   0 randomOnUnitSphere()            1 randomInHemisphere(n)            2 randomInHemisphereCosine(n)
//   3 randomInHemisphereCosinePower(reflected, p)                        4 randomOnSphere(center, radius) -> point, normal
//   5 randomOnSphereVisible(center, radius, n) -> point, normal          6 randomBeckmannNormal(n, roughness)
//   7 toWorld(v, n)                   9 sampleVMF(theta[10], worldPos, parallax)       10 vMF(theta[10], worldPos, parallax, wo)
//  11 sampleVMM(vmm[180], worldPos, parallax)                           12 VMM(vmm[180], worldPos, parallax, wo)
//  20 fresnel  21 fresnelConductor  22 evalBsdf  23 pdfBSDF  24 sampleBSDF -> pdf, direction  25 power / balance heuristic
//  26 applyWeightWindow(throughput, adjoint, estimate, adrrsS) -> q, n   27 approxDiffuse  28 pdfLight  29 material predicates
int oracle_unit_eval(int fn, uint32_t *seed_io, const float *in, float *out) {
    static oracle_ctx dummy;
    Pixel p(dummy);
    p.seed = *seed_io;
    auto put = [&](float *o, v3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; };
    switch (fn) {
        case 0: put(out, p.randomOnUnitSphere()); break;
        case 1: put(out, p.randomInHemisphere(v3(in))); break;
        case 2: put(out, p.randomInHemisphereCosine(v3(in))); break;
        case 3: put(out, p.randomInHemisphereCosinePower(v3(in), in[3])); break;
        case 4: { b200pt_sphere s; memset(&s, 0, sizeof(s)); memcpy(s.center, in, 12); s.radius = in[3]; v3 n; put(out, p.randomOnSphere(s, n)); put(out + 3, n); break; }
        case 5: { b200pt_sphere s; memset(&s, 0, sizeof(s)); memcpy(s.center, in, 12); s.radius = in[3]; v3 n; put(out, p.randomOnSphereVisible(s, v3(in + 4), n)); put(out + 3, n); break; }
        case 6: { b200pt_material m; memset(&m, 0, sizeof(m)); m.roughness = in[3]; put(out, p.randomBeckmannNormal(m, v3(in))); break; }
        case 7: put(out, Pixel::toWorld(v3(in), v3(in + 3))); break;
        case 9: { b200pt_vmf_theta t; memcpy(&t, in, sizeof(t)); put(out, p.sampleVMF(t, v3(in + 10), in[13] != 0.0f)); break; }
        case 10: { b200pt_vmf_theta t; memcpy(&t, in, sizeof(t)); out[0] = Pixel::vMF(v3(in + 14), t, v3(in + 10), in[13] != 0.0f); break; }
        case 11: { b200pt_vmm_theta t; memcpy(&t, in, sizeof(t)); put(out, p.sampleVMM(t, v3(in + 180), in[183] != 0.0f)); break; }
        case 12: { b200pt_vmm_theta t; memcpy(&t, in, sizeof(t)); out[0] = Pixel::VMM(v3(in + 184), t, v3(in + 180), in[183] != 0.0f); break; }
        // raytrace.rgen: material = b200pt_material (24 floats) at in[0], then normal, wi, wo, frontFace
        case 20: out[0] = Pixel::fresnel(in[0], in[1]); break;
        case 21: out[0] = Pixel::fresnelConductor(in[0], in[1], in[2]); break;
        case 22: { b200pt_material m; memcpy(&m, in, sizeof(m)); put(out, p.evalBsdf(m, 0.0f, 0.0f, v3(in + 24), v3(in + 27), v3(in + 30), in[33] != 0.0f)); break; }
        case 23: { b200pt_material m; memcpy(&m, in, sizeof(m)); out[0] = Pixel::pdfBSDF(m, v3(in + 24), v3(in + 27), v3(in + 30)); break; }
        case 24: { b200pt_material m; memcpy(&m, in, sizeof(m)); v3 d(0.0f); out[0] = p.sampleBSDF(m, v3(in + 27), v3(in + 24), in[33] != 0.0f, d); put(out + 1, d); break; }
        case 25: out[0] = Pixel::powerHeuristic(in[0], in[1]); out[1] = Pixel::balanceHeuristic(in[0], in[1]); break;
        case 26: { p.estimate = v3(in + 6); dummy.pushC.adrrsS = in[9]; int n = 0; out[0] = p.applyWeightWindow(v3(in), v3(in + 3), n); out[1] = float(n); break; }
        case 27: { b200pt_material m; memcpy(&m, in, sizeof(m)); put(out, p.approxDiffuse(m, v3(in + 24), v3(in + 27), 0.0f, 0.0f)); break; }
        case 28: { b200pt_light l; memset(&l, 0, sizeof(l)); l.sampleProb = in[0]; l.area = in[1]; out[0] = Pixel::pdfLight(l, v3(in + 2), v3(in + 5), in[8]); break; }
        case 29: { b200pt_material m; memcpy(&m, in, sizeof(m)); out[0] = float(p.hasDiscreteDirection(m)); out[1] = float(p.isMatAlmostDiscrete(m));
                   dummy.pushC.useIrradianceCacheOnGlossy = 0; out[2] = float(p.isICCapable(m)); dummy.pushC.useIrradianceCacheOnGlossy = 1; out[3] = float(p.isICCapable(m)); break; }
        default: return -1;
    }
    *seed_io = p.seed;
    return 0;
}
uint32_t oracle_lcg(uint32_t *prev) { Pixel::rndS(*prev); return *prev & 0x00FFFFFFu; }
float oracle_rnd(uint32_t *prev) { return Pixel::rndS(*prev); }

float *oracle_image(oracle_ctx *C, int which) {
    return which == B200PT_IMAGE_OUTPUT ? C->image.data() : which == B200PT_IMAGE_ACCUM ? C->accumulateImage.data() : which == 3 ? C->aovImage.data() : C->estimateImage.data();
}
void oracle_get_counters(oracle_ctx *C, uint64_t out[3]) { out[0] = C->extendRays; out[1] = C->shadowRays; out[2] = C->pathVertices; }
void oracle_reset_counters(oracle_ctx *C) { C->extendRays = C->shadowRays = C->pathVertices = 0; }
void oracle_set_guiding(oracle_ctx *C, const b200pt_aabb *aabbs, const b200pt_vmm_theta *vmms, int n) {
    C->guidingAabbs.assign(aabbs, aabbs + n);
    C->guidingVMM.assign(vmms, vmms + n);
}
b200pt_directional_data *oracle_samples(oracle_ctx *C) { return C->directionalData.data(); }
void oracle_ic_get(oracle_ctx *C, b200pt_cache_header *hdr, b200pt_cache_data *data, b200pt_sphere *spheres, int n) {
    if (hdr) *hdr = C->header;
    if (data) memcpy(data, C->cache.data(), size_t(n) * sizeof(*data));
    if (spheres) memcpy(spheres, C->cacheSpheres.data(), size_t(n) * sizeof(*spheres));
}
void oracle_ic_put(oracle_ctx *C, const b200pt_cache_header *hdr, const b200pt_cache_data *data, const b200pt_sphere *spheres, int n) {
    if (hdr) C->header = *hdr;
    if (data) memcpy(C->cache.data(), data, size_t(n) * sizeof(*data));
    if (spheres) memcpy(C->cacheSpheres.data(), spheres, size_t(n) * sizeof(*spheres));
}

}  // extern "C"
