// TEST INFRASTRUCTURE ONLY.  The reference's ray-tracing pipeline compiled as C++ and run on the CPU, one pixel after the other:
//   raytrace.rgen (the megakernel: camera ray, bounce loop, NEE + MIS, light sampling, irradiance-cache lookup, ADRRS window and
//   splits, guided sampling, sample recording, saveResult / saveEstimate), raytrace.rchit, raytrace.sphere.rchit, raytrace.rmiss,
//   raytrace.shadow.rmiss, raytrace.irradiance.rint / .rahit, raytrace.guiding.rint / .rchit
// — every line from /root/reference/shaders, piped through sed by oracle/Makefile into oracle/_ref/glsl/*.inc (build intermediates,
// deleted after the compile; see glsl_prelude.h for what the rewriting does).  Left out: the debug views (visualizeIC, the guiding
// visualisations and the visualizeMode switch cases other than VISU_RAYTRACE).
// What is NOT the reference's: traceRayEXT itself.  The reference hands rays to the driver's acceleration structure; here the
// closest / any hit comes from oracle/tracer_oracle.cpp's traversal through a callback (hit distance, barycentrics, instance and
// primitive ids — the inputs of the hit shaders), texels from its sampler, and the two procedural acceleration structures
// (irradiance cache spheres, guiding boxes) are walked in index order running the reference's intersection shaders.
// tests/test_shader_ref.py compares whole frames of this build with oracle/tracer_oracle.cpp's restatement.
#define GLSL_RT 1
#include "glsl_prelude_rt.h"
#include <string.h>
#include <stdio.h>
#include <vector>
#include "../include/b200pt.h"

namespace glsl {
struct texel4rt { union { struct { float x, y, z, w; }; struct { float r, g, b, a; }; xyz_t xyz; }; };
static sampler2D textureSamplers[256];
typedef void (*texture_fn)(void *ctx, int id, float u, float v, float out[4]);
typedef int (*trace_fn)(void *ctx, const float o[3], const float d[3], float tmin, float tmax, int any_hit, float out[3], uint32_t ids[4]);
static texture_fn g_texture; static trace_fn g_trace; static void *g_ctx;
static inline texel4rt texture(sampler2D s, vec2 uv) { texel4rt t; float o[4]; g_texture(g_ctx, s.id, uv.x, uv.y, o); t.x = o[0]; t.y = o[1]; t.z = o[2]; t.w = o[3]; return t; }
}
#include "glsl_macros.h"
namespace glsl {
#include "_ref/glsl/limits.inc"
#include "_ref/glsl/wavefront.inc"
#include "_ref/glsl/raycommon.inc"
// ---- descriptor sets 0 and 1 (raytrace.rgen:96-140), as plain objects
static pushConstant pushC;
static accelerationStructureEXT topLevelAS = {0}, irradianceAS = {1}, guidingAS = {2};
static image2D image, accumulateImage, estimateImage;
struct VerticesBuf { Vertex *v; }; struct IndicesBuf { uint *i; };
static VerticesBuf *vertices; static IndicesBuf *indices;
static Material *mats; static InstanceInfo *instanceInfos;
static int lightCount; static Light *lights;
static int *randomLigthIndex; typedef FaceSample FaceTable[SIZE_TRI_RANDOM]; static FaceTable *randomTriIndex;
static sphere *spheres, *cacheSpheres; static aabb *cacheAabbs, *guidingAabbs, *aabbs;
static cacheHeader header; static cacheData *cache;
struct CameraMatrices { mat4 view, proj, viewInverse, projInverse; }; static CameraMatrices cam;
// ---- ray payloads and the built-in variables of the stages
static hitInfo info; static shadowCheck shadowInfo; static cacheHits cacheInfo; static guidingInfo guidingInfos; static guidingVisuInfo guidingVisuInfos;
static uvec3 gl_LaunchIDEXT, gl_LaunchSizeEXT;
static int gl_InstanceID, gl_PrimitiveID;
static vec3 gl_WorldRayOriginEXT, gl_WorldRayDirectionEXT, attribs;
static float gl_HitTEXT, gl_RayTmaxEXT;
static bool g_reported; static float g_reportedT[4]; static int g_numReported;
static inline bool reportIntersectionEXT(float t, uint) { g_reported = true; if (g_numReported < 4) g_reportedT[g_numReported++] = t; return true; }
static void traceRayEXT(accelerationStructureEXT as, uint flags, uint cullMask, uint sbtOffset, uint sbtStride, uint missIndex, vec3 origin, float tmin, vec3 direction, float tmax, int payload);
static inline vec3 visualizeIC(vec3, vec3) { abort(); }      // debug view, not part of the comparison
#include "_ref/glsl/random_fwd.inc"
#include "_ref/glsl/transform_fwd.inc"
#include "_ref/glsl/transform.inc"
#include "_ref/glsl/guiding.inc"
static VMM_Theta *guidingVMM; static DirectionalData *directionalData;
#include "_ref/glsl/random.inc"
#include "_ref/glsl/rgen_all.inc"
#undef E          // raytrace.rgen's macro; raytrace.irradiance.rahit (its own translation unit in GLSL) has a local variable E
namespace rchit {
#include "_ref/glsl/stage_rchit.inc"
}
namespace sphere_rchit {
#include "_ref/glsl/stage_sphere_rchit.inc"
}
namespace rmiss {
#include "_ref/glsl/stage_rmiss.inc"
}
namespace shadow_rmiss {
#include "_ref/glsl/stage_shadow_rmiss.inc"
}
namespace irradiance_rint { static sphere *&spheres = glsl::cacheSpheres;
#include "_ref/glsl/stage_irradiance_rint.inc"
}
namespace irradiance_rahit {
#include "_ref/glsl/stage_irradiance_rahit.inc"
}
static bool g_ignored;
static inline void ignoreIntersectionEXT() { g_ignored = true; }
namespace rahit {
#include "_ref/glsl/stage_rahit.inc"
}
namespace guiding_rint {
#include "_ref/glsl/stage_guiding_rint.inc"
}
namespace guiding_rchit { static guidingInfo &info = glsl::guidingInfos;
#include "_ref/glsl/stage_guiding_rchit.inc"
}

static int g_numGuidingRegions = 0;
static uint g_icEntriesAtFrameStart = 0;
static unsigned long long g_raysTraced = 0;
static void traceRayEXT(accelerationStructureEXT as, uint flags, uint, uint, uint, uint missIndex, vec3 origin, float tmin, vec3 direction, float tmax, int) {
    gl_WorldRayOriginEXT = origin; gl_WorldRayDirectionEXT = direction; gl_RayTmaxEXT = tmax;
    if (as.id == 0) {                                   // topLevelAS: triangles + analytic spheres
        g_raysTraced++;
        const float o[3] = {origin.x, origin.y, origin.z}, d[3] = {direction.x, direction.y, direction.z};
        float out[3]; uint32_t ids[4];
        const bool any = (flags & gl_RayFlagsTerminateOnFirstHitEXT) != 0u;
        if (g_trace(g_ctx, o, d, tmin, tmax, any ? 1 : 0, out, ids)) {
            if (flags & gl_RayFlagsSkipClosestHitShaderEXT) return;
            gl_HitTEXT = out[0];
            if (ids[3]) { gl_PrimitiveID = int(ids[1]); sphere_rchit::main(); }
            else { gl_InstanceID = int(ids[1]); gl_PrimitiveID = int(ids[2]); attribs = vec3{out[1], out[2], 0.0f}; rchit::main(); }
        } else if (missIndex == 0u) rmiss::main();
        else shadow_rmiss::main();
    } else if (as.id == 1) {                            // irradianceAS: one box per cache entry, intersection + any-hit shader per candidate
        // the lookup structure is the one the host built before the frame (IrradianceCache::updateSpheres runs between frames,
        // src/IrradianceCache.cpp:81-104): entries created during this frame are not in it yet
        const uint n = g_icEntriesAtFrameStart;
        for (uint i = 0; i < n; i++) {
            gl_PrimitiveID = int(i); g_reported = false; g_numReported = 0;
            irradiance_rint::main();
            if (g_reported) irradiance_rahit::main();
        }
    } else {                                            // guidingAS: the first box (lowest index) that reports the point
        aabbs = guidingAabbs;
        for (int i = 0; i < g_numGuidingRegions; i++) {
            gl_PrimitiveID = i; g_reported = false; g_numReported = 0;
            guiding_rint::main();
            if (g_reported) { guiding_rchit::main(); return; }
        }
    }
}
}  // namespace glsl
using namespace glsl;

// ---- host side: scene upload in the natural C++ layout of the shader structs, frame loop -----------------------------------
static std::vector<std::vector<Vertex>> s_vertices; static std::vector<std::vector<uint>> s_indices;
static std::vector<VerticesBuf> s_vbuf; static std::vector<IndicesBuf> s_ibuf;
static std::vector<Material> s_mats; static std::vector<InstanceInfo> s_instances; static std::vector<Light> s_lights;
static std::vector<int> s_randomLight; static std::vector<FaceSample> s_faceTables; static std::vector<sphere> s_spheres;
static std::vector<float> s_image, s_accum, s_estimate;
static std::vector<sphere> s_cacheSpheres; static std::vector<cacheData> s_cache; static std::vector<aabb> s_cacheAabbs;
static std::vector<aabb> s_guidingAabbs; static std::vector<VMM_Theta> s_vmms; static std::vector<DirectionalData> s_samples;
static int s_width = 0, s_height = 0;
static vec3 v3(const float *p) { return vec3{p[0], p[1], p[2]}; }
static_assert(sizeof(VMM_Theta) == sizeof(b200pt_vmm_theta) && sizeof(DirectionalData) == sizeof(b200pt_directional_data), "scalar block layout");

extern "C" {
void shader_ref_set_callbacks(void *ctx, void *trace, void *tex) { g_ctx = ctx; g_trace = trace_fn(trace); g_texture = texture_fn(tex); }

int shader_ref_init(int width, int height, int ic_size) {
    s_width = width; s_height = height;
    const size_t n = size_t(width) * height * 4;
    s_image.assign(n, 0.0f); s_accum.assign(n, 0.0f); s_estimate.assign(n, 0.0f);
    image = image2D{s_image.data(), width}; accumulateImage = image2D{s_accum.data(), width}; estimateImage = image2D{s_estimate.data(), width};
    s_samples.clear(); directionalData = nullptr;          // binding 18 (W * H * 16 records) is allocated by the first training frame
    const size_t ic = size_t(ic_size > 0 ? ic_size : 1);
    s_cacheSpheres.assign(ic, sphere()); s_cache.assign(ic, cacheData()); s_cacheAabbs.assign(ic, aabb());
    cacheSpheres = s_cacheSpheres.data(); cache = s_cache.data(); cacheAabbs = s_cacheAabbs.data();
    header.nextCacheSlot = 0; header.maxCaches = uint(ic_size); header.nextUpdateSlot = 0;
    for (int i = 0; i < 256; i++) textureSamplers[i].id = i;
    gl_LaunchSizeEXT.x = uint(width); gl_LaunchSizeEXT.y = uint(height); gl_LaunchSizeEXT.z = 1u;
    return 0;
}

int shader_ref_set_scene(const b200pt_scene_desc *s) {
    s_vertices.assign(size_t(s->num_models), {}); s_indices.assign(size_t(s->num_models), {});
    s_vbuf.assign(size_t(s->num_models) + 1, VerticesBuf{nullptr}); s_ibuf.assign(size_t(s->num_models) + 1, IndicesBuf{nullptr});
    for (int m = 0; m < s->num_models; m++) {
        for (int k = 0; k < s->num_vertices[m]; k++) {
            const b200pt_vertex &b = s->vertices[m][k];
            Vertex v; v.pos = v3(b.pos); v.normal = v3(b.normal); v.texCoord = vec2{b.texCoord[0], b.texCoord[1]}; v.materialIndex = b.materialIndex;
            s_vertices[m].push_back(v);
        }
        s_indices[m].assign(s->indices[m], s->indices[m] + s->num_indices[m]);
        s_vbuf[m].v = s_vertices[m].data(); s_ibuf[m].i = s_indices[m].data();
    }
    vertices = s_vbuf.data(); indices = s_ibuf.data();
    s_mats.clear();
    for (int i = 0; i < s->num_materials; i++) {
        const b200pt_material &b = s->materials[i];
        Material m; m.lightColor = v3(b.lightColor); m.diffuse = v3(b.diffuse); m.specular = v3(b.specular); m.specularHighlight = b.specularHighlight;
        m.transparency = b.transparency; m.refractionIndex = b.refractionIndex; m.refractionIndexInv = b.refractionIndexInv; m.eta = b.eta; m.k = b.k;
        m.roughness = b.roughness; m.textureIdDiffuse = b.textureIdDiffuse; m.textureIdSpecular = b.textureIdSpecular; m.type = b.type;
        s_mats.push_back(m);
    }
    mats = s_mats.data();
    s_instances.clear();
    for (int i = 0; i < s->num_instances; i++) {
        InstanceInfo in; memcpy(in.transform.m, s->instances[i].transform, 64); memcpy(in.normalTransform.m, s->instances[i].normalTransform, 64);
        in.modelIndex = s->instances[i].modelIndex; in.iLight = s->instances[i].iLight;
        s_instances.push_back(in);
    }
    instanceInfos = s_instances.data();
    s_lights.clear();
    for (int i = 0; i < s->num_lights; i++) {
        const b200pt_light &b = s->lights[i];
        Light l; l.color = v3(b.color); l.pos = v3(b.pos); l.instanceIndex = b.instanceIndex; l.sampleProb = b.sampleProb; l.area = b.area; l.type = b.type;
        s_lights.push_back(l);
    }
    lights = s_lights.data(); lightCount = s->num_lights;
    s_randomLight.assign(s->random_light_index, s->random_light_index + SIZE_LIGHT_RANDOM);
    randomLigthIndex = s_randomLight.data();
    const size_t nf = size_t(s->num_face_tables > 0 ? s->num_face_tables : 1) * SIZE_TRI_RANDOM;
    s_faceTables.assign(nf, FaceSample());
    for (size_t i = 0; i < nf && s->random_tri_index; i++) { s_faceTables[i].index = s->random_tri_index[i].index; s_faceTables[i].sampleProb = s->random_tri_index[i].sampleProb; s_faceTables[i].faceArea = s->random_tri_index[i].faceArea; }
    randomTriIndex = reinterpret_cast<FaceTable *>(s_faceTables.data());
    s_spheres.clear();
    for (int i = 0; i < s->num_spheres; i++) { sphere sp; sp.center = v3(s->spheres[i].center); sp.radius = s->spheres[i].radius; sp.materialIndex = s->spheres[i].materialIndex; sp.iLight = s->spheres[i].iLight; s_spheres.push_back(sp); }
    if (s_spheres.empty()) s_spheres.push_back(sphere());
    spheres = s_spheres.data();
    return 0;
}

void shader_ref_set_camera(const float view[16], const float proj[16], const float viewInv[16], const float projInv[16]) {
    memcpy(cam.view.m, view, 64); memcpy(cam.proj.m, proj, 64); memcpy(cam.viewInverse.m, viewInv, 64); memcpy(cam.projInverse.m, projInv, 64);
}

void shader_ref_set_guiding(const b200pt_aabb *boxes, const b200pt_vmm_theta *vmms, int n) {
    s_guidingAabbs.assign(size_t(n), aabb()); s_vmms.assign(size_t(n), VMM_Theta());
    for (int i = 0; i < n; i++) { s_guidingAabbs[i].min = v3(boxes[i].min); s_guidingAabbs[i].max = v3(boxes[i].max); }
    memcpy(s_vmms.data(), vmms, size_t(n) * sizeof(VMM_Theta));
    guidingAabbs = s_guidingAabbs.data(); guidingVMM = s_vmms.data(); g_numGuidingRegions = n;
}

void shader_ref_ic_put(const b200pt_cache_header *h, const b200pt_cache_data *data, const b200pt_sphere *sp, int n) {
    header.nextCacheSlot = h->nextCacheSlot; header.maxCaches = h->maxCaches; header.nextUpdateSlot = h->nextUpdateSlot;
    for (int i = 0; i < n && size_t(i) < s_cache.size(); i++) {
        cacheData c; c.color = v3(data[i].color); c.normal = v3(data[i].normal); c.rotGrad = v3(data[i].rotGrad); c.transGrad = v3(data[i].transGrad);
        c.harmonicR = data[i].harmonicR; c.numUpdates = data[i].numUpdates;
        s_cache[i] = c;
        s_cacheSpheres[i].center = v3(sp[i].center); s_cacheSpheres[i].radius = sp[i].radius; s_cacheSpheres[i].materialIndex = sp[i].materialIndex; s_cacheSpheres[i].iLight = sp[i].iLight;
    }
}

// the any-hit shader of the triangle geometry (raytrace.rahit: the stochastic alpha test) on one candidate hit; returns 1 when it
// calls ignoreIntersectionEXT.  In a frame this decision is taken inside the traversal, i.e. on the oracle's side of the callback.
int shader_ref_alpha_rejects(int instance, int prim, float u, float v, float origin_x, float t, uint32_t random_uint) {
    gl_InstanceID = instance; gl_PrimitiveID = prim; attribs = vec3{u, v, 0.0f};
    gl_WorldRayOriginEXT = vec3{origin_x, 0.0f, 0.0f}; gl_HitTEXT = t; pushC.randomUInt = random_uint;
    g_ignored = false;
    rahit::main();
    return g_ignored ? 1 : 0;
}

void shader_ref_ic_get(b200pt_cache_header *h, b200pt_cache_data *data, b200pt_sphere *sp, int n) {
    h->nextCacheSlot = header.nextCacheSlot; h->maxCaches = header.maxCaches; h->nextUpdateSlot = header.nextUpdateSlot;
    for (int i = 0; i < n && size_t(i) < s_cache.size(); i++) {
        const cacheData &c = s_cache[i];
        memset(&data[i], 0, sizeof(data[i])); memset(&sp[i], 0, sizeof(sp[i]));
        for (int a = 0; a < 3; a++) { data[i].color[a] = (&c.color.x)[a]; data[i].normal[a] = (&c.normal.x)[a]; data[i].rotGrad[a] = (&c.rotGrad.x)[a]; data[i].transGrad[a] = (&c.transGrad.x)[a]; sp[i].center[a] = (&s_cacheSpheres[i].center.x)[a]; }
        data[i].harmonicR = c.harmonicR; data[i].numUpdates = c.numUpdates;
        sp[i].radius = s_cacheSpheres[i].radius; sp[i].materialIndex = s_cacheSpheres[i].materialIndex; sp[i].iLight = s_cacheSpheres[i].iLight;
    }
}

unsigned long long shader_ref_rays_traced(void) { return g_raysTraced; }     // extend + shadow rays handed to the callback so far

float *shader_ref_image(int which) { return which == 0 ? s_image.data() : which == 1 ? s_accum.data() : s_estimate.data(); }
void *shader_ref_samples(void) { return s_samples.data(); }

// one frame = traceRaysKHR(W, H, 1) over the pixels of [x0, x1) x [y0, y1), serially in row-major order
int shader_ref_render(const b200pt_push_constants *pc, int x0, int y0, int x1, int y1) {
    pushConstant &p = pushC;
#define PCF(f) p.f = pc->f
    PCF(randomUInt); PCF(previousFrames); PCF(maxDepth); PCF(maxFollowDiscrete); PCF(samplesPerPixel); PCF(enableRR); PCF(enableNEE); PCF(numNEE);
    PCF(enableAverageInsteadOfMix); PCF(enableMIS); PCF(usePowerHeuristic); PCF(storeEstimate); PCF(visualizeMode); PCF(showIrradianceCacheOnly);
    PCF(showIrradianceGradients); PCF(useIrradianceCache); PCF(highlightIrradianceCacheColor); PCF(irradianceA); PCF(irradianceUpdateProb);
    PCF(irradianceCreateProb); PCF(irradianceVisualizationScale); PCF(useIrradianceGradients); PCF(useIrradianceCacheOnGlossy);
    PCF(irradianceGradientsMaxLength); PCF(isIrradiancePrepareFrame); PCF(irradianceNumNEE); PCF(irradianceCacheMinRadius);
    PCF(irradianceCachePerformVisibilityCheck); PCF(useVisibleSphereSampling); PCF(useADRRS); PCF(adrrsS); PCF(adrrsSplit); PCF(splitOnFirst);
    PCF(useGuiding); PCF(guidingProb); PCF(guidingVisuScale); PCF(guidingVisuMax); PCF(guidingVisuIgnoreOcclusioon); PCF(updateGuiding);
    PCF(useParallaxCompensation); PCF(time); PCF(guidingVisuMove); PCF(guidingVisuPhiScale); PCF(guidingVisuThetaScale); PCF(numGuidingRegions);
    PCF(guidingPiPHighlightRegion); PCF(guidingPiPShowSpheres); PCF(guidingPiPSize);
#undef PCF
    if (!g_trace || !g_texture) return -1;
    if (pc->updateGuiding && s_samples.empty()) {
        s_samples.assign(size_t(s_width) * s_height * MAX_DIRECTIONAL_DATA_PER_PIXEL, DirectionalData());
        for (auto &d : s_samples) d.flags = INVALID;
        directionalData = s_samples.data();
    }
    g_icEntriesAtFrameStart = header.nextCacheSlot < header.maxCaches ? header.nextCacheSlot : header.maxCaches;
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            gl_LaunchIDEXT.x = uint(x); gl_LaunchIDEXT.y = uint(y); gl_LaunchIDEXT.z = 0u;
            nextNewIrradianceCacheSlot = 0; nextSplitSlot = 0; sampleOffset = 0;       // the initialisers of the shader's globals run per invocation
            glsl::main();
        }
    return 0;
}
}
