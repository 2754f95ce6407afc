/*
 * b200pt.h — C ABI of libb200pt.so, the B200-native path-tracing core.
 *
 * This is the drop-in boundary described in SURVEY.md §8(b).  The reference
 * (FelixFifi/rtx-pathtracer) has no FFI layer; its de-facto boundary is "what
 * the host binds to the Vulkan ray-tracing pipeline and what it reads back".
 * Every entry point below names the reference call site it replaces.
 *
 * Conventions
 *   - all functions return int: 0 = OK, negative = error (see B200PT_E_*);
 *     b200pt_last_error() returns a human-readable message for the calling thread
 *   - plain pointers and sizes only; the host keeps ownership of every input
 *     array, the library copies what it needs to the device
 *   - one context per GPU, not thread-safe per context
 *   - all structs are byte-identical to the GLSL / C++ layouts of the reference
 *     (little-endian float32 / int32), so buffers can be passed through unchanged
 */
#ifndef B200PT_H
#define B200PT_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200PT_OK            0
#define B200PT_E_INVALID    -1   /* bad argument */
#define B200PT_E_CUDA       -2   /* CUDA runtime error */
#define B200PT_E_IO         -3   /* file not found / parse error */
#define B200PT_E_STATE      -4   /* call order violated (e.g. render before set_scene) */
#define B200PT_E_NODEVICE   -5   /* no CUDA device: there is NO CPU fallback */

/* shaders/limits.glsl:1-6 */
#define B200PT_SIZE_LIGHT_RANDOM 10000
#define B200PT_SIZE_TRI_RANDOM   10000
#define B200PT_MAX_DISTRIBUTIONS 16
#define B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL 16
#define B200PT_INVALID_REGION 0xFFFFFFFFu      /* shaders/guiding.glsl:112 */

/* material types, shaders/raytrace.rgen:21-27 / src/SceneLoader.h:31-39 */
enum { B200PT_MAT_DIFFUSE = 0, B200PT_MAT_SPECULAR = 1, B200PT_MAT_DIELECTRIC = 2, B200PT_MAT_LIGHT = 3,
       B200PT_MAT_PHONG = 4, B200PT_MAT_CONDUCTOR = 5, B200PT_MAT_ROUGH_CONDUCTOR = 6 };
/* light types, shaders/raytrace.rgen:29-32 / src/SceneLoader.h:41-46 */
enum { B200PT_LIGHT_AREA = 0, B200PT_LIGHT_POINT = 1, B200PT_LIGHT_SPHERE = 2, B200PT_LIGHT_ENV_MAP = 3 };

/* ---- scene buffer layouts (SURVEY.md §8(a) layout table) ------------------------------- */

/* shaders/wavefront.glsl:1-7 (std140 array) / src/Model.h:24-33 — 48 B */
typedef struct b200pt_vertex {
    float pos[3];      float _pad0;
    float normal[3];   float _pad1;
    float texCoord[2];
    int32_t materialIndex;
    int32_t _pad2;
} b200pt_vertex;

/* shaders/wavefront.glsl:9-23 (std430) / src/SceneLoader.h:48-62 — 96 B */
typedef struct b200pt_material {
    float lightColor[3]; float _pad0;
    float diffuse[3];    float _pad1;
    float specular[3];
    float specularHighlight;
    float transparency;
    float refractionIndex;
    float refractionIndexInv;
    float eta;
    float k;
    float roughness;
    int32_t textureIdDiffuse;   /* -1 = none */
    int32_t textureIdSpecular;  /* -1 = none */
    int32_t type;
    int32_t _pad2[3];
} b200pt_material;

/* shaders/wavefront.glsl:25-30 (std140) / src/SceneLoader.h:64-69 — 144 B, matrices column-major */
typedef struct b200pt_instance {
    float transform[16];
    float normalTransform[16];
    int32_t modelIndex;
    int32_t iLight;             /* -1 = not a light */
    int32_t _pad[2];
} b200pt_instance;

/* shaders/wavefront.glsl:32-39 (scalar) / src/SceneLoader.h:71-78 — 40 B */
typedef struct b200pt_light {
    float color[3];
    float pos[3];
    uint32_t instanceIndex;     /* instance index (area) or sphere index (sphere light) */
    float sampleProb;
    float area;
    int32_t type;
} b200pt_light;

/* shaders/wavefront.glsl:41-45 / src/SceneLoader.h:80-84 — 12 B */
typedef struct b200pt_face_sample {
    int32_t index;
    float sampleProb;
    float faceArea;
} b200pt_face_sample;

/* shaders/raycommon.glsl:71-76 (scalar) / src/Shapes.h:65-75 — 24 B */
typedef struct b200pt_sphere {
    float center[3];
    float radius;
    int32_t materialIndex;
    int32_t iLight;             /* -1 = not a light */
} b200pt_sphere;

/* shaders/raycommon.glsl:78-81 / src/Shapes.h:19-22 — 24 B */
typedef struct b200pt_aabb {
    float min[3];
    float max[3];
} b200pt_aabb;

/* shaders/raycommon.glsl:83-87 / src/IrradianceCache.h — 12 B */
typedef struct b200pt_cache_header {
    uint32_t nextCacheSlot;
    uint32_t maxCaches;
    uint32_t nextUpdateSlot;
} b200pt_cache_header;

/* shaders/raycommon.glsl:89-96 (scalar) / src/IrradianceCache.h:17-30 — 56 B */
typedef struct b200pt_cache_data {
    float color[3];
    float normal[3];
    float rotGrad[3];
    float transGrad[3];
    float harmonicR;
    uint32_t numUpdates;
} b200pt_cache_data;

/* shaders/guiding.glsl:99-113 / external/lightpmm/include/pmm/DirectionalData.h:44-58 — 40 B */
typedef struct b200pt_directional_data {
    float position[3];
    float direction[3];
    float weight;
    float pdf;
    float distance;
    uint32_t flags;             /* region id or B200PT_INVALID_REGION */
} b200pt_directional_data;

/* shaders/guiding.glsl:14-21 / src/PathGuiding.h:36-56 — 40 B */
typedef struct b200pt_vmf_theta {
    float mu[3];
    float k;
    float norm;
    float eMin2K;
    float distance;
    float target[3];
} b200pt_vmf_theta;

/* shaders/guiding.glsl:46-51 / src/PathGuiding.h:58-75 — 720 B */
typedef struct b200pt_vmm_theta {
    b200pt_vmf_theta thetas[B200PT_MAX_DISTRIBUTIONS];
    float pi[B200PT_MAX_DISTRIBUTIONS];
    float meanPosition[3];
    int32_t usedDistributions;
} b200pt_vmm_theta;

/* push constants: shaders/raycommon.glsl:19-69 / src/RayTracingApp.h:116-165 — 192 B, same field order,
 * GLSL bools are 4-byte ints.  Defaults (b200pt_default_push_constants) are RayTracingApp.h's. */
typedef struct b200pt_push_constants {
    uint32_t randomUInt;
    uint32_t previousFrames;
    int32_t maxDepth;
    int32_t maxFollowDiscrete;
    int32_t samplesPerPixel;
    int32_t enableRR;                 /* dead in the reference (never read by a shader) */
    int32_t enableNEE;
    int32_t numNEE;
    int32_t enableAverageInsteadOfMix;
    int32_t enableMIS;
    int32_t usePowerHeuristic;
    int32_t storeEstimate;
    int32_t visualizeMode;
    int32_t showIrradianceCacheOnly;
    int32_t showIrradianceGradients;
    int32_t useIrradianceCache;
    int32_t highlightIrradianceCacheColor;
    float irradianceA;
    float irradianceUpdateProb;
    float irradianceCreateProb;
    float irradianceVisualizationScale;
    int32_t useIrradianceGradients;
    int32_t useIrradianceCacheOnGlossy;
    float irradianceGradientsMaxLength;
    int32_t isIrradiancePrepareFrame;
    int32_t irradianceNumNEE;
    float irradianceCacheMinRadius;
    int32_t irradianceCachePerformVisibilityCheck;
    int32_t useVisibleSphereSampling;
    int32_t useADRRS;
    float adrrsS;
    int32_t adrrsSplit;
    int32_t splitOnFirst;
    int32_t useGuiding;
    float guidingProb;
    float guidingVisuScale;
    float guidingVisuMax;
    int32_t guidingVisuIgnoreOcclusioon;
    int32_t updateGuiding;
    int32_t useParallaxCompensation;
    float time;
    int32_t guidingVisuMove;
    float guidingVisuPhiScale;
    float guidingVisuThetaScale;
    int32_t numGuidingRegions;
    int32_t guidingPiPHighlightRegion;
    int32_t guidingPiPShowSpheres;
    float guidingPiPSize;
} b200pt_push_constants;

/* textures: binding 7 of set 1 (src/SceneLoader.cpp:184-236, :370-396). Index 0 is the environment map. */
enum { B200PT_TEX_RGBA8_SRGB = 0, B200PT_TEX_RGBA32F = 1 };
typedef struct b200pt_texture {
    int32_t width, height;
    int32_t format;           /* B200PT_TEX_* */
    int32_t _pad;
    const void *pixels;       /* width*height*4 bytes (RGBA8) or floats (RGBA32F), row 0 first */
} b200pt_texture;

/* everything the reference binds per scene: descriptor set 1 bindings 1..8 (src/RayTracingApp.cpp:396-557,
 * src/SceneLoader.cpp:946-1099) */
typedef struct b200pt_scene_desc {
    int32_t num_models;
    const b200pt_vertex *const *vertices;   /* [num_models] */
    const int32_t *num_vertices;            /* [num_models] */
    const uint32_t *const *indices;         /* [num_models], 3 per triangle */
    const int32_t *num_indices;             /* [num_models] */
    int32_t num_materials;
    const b200pt_material *materials;
    int32_t num_instances;
    const b200pt_instance *instances;
    int32_t num_lights;
    const b200pt_light *lights;
    const int32_t *random_light_index;      /* [B200PT_SIZE_LIGHT_RANDOM] */
    int32_t num_face_tables;                /* one per mesh light, in light order (quirk 5) */
    const b200pt_face_sample *random_tri_index; /* [num_face_tables][B200PT_SIZE_TRI_RANDOM] */
    int32_t num_spheres;
    const b200pt_sphere *spheres;
    int32_t num_textures;
    const b200pt_texture *textures;         /* [0] = env map (1x1 black RGBA8 when the scene has none) */
    float scene_min[3], scene_max[3];       /* SceneLoader::calculateSceneSize, src/SceneLoader.cpp:350-368 */
} b200pt_scene_desc;

/* ray / hit records for the traversal-only parity hook (north_star level 1) */
typedef struct b200pt_ray {
    float origin[3]; float tmin;
    float dir[3];    float tmax;
} b200pt_ray;
typedef struct b200pt_hit {
    float t;            /* hit distance; undefined on miss */
    uint32_t prim;      /* global primitive id (triangles of instance 0.., then spheres); 0xFFFFFFFF = miss */
    float u, v;         /* barycentrics of vertices 1 and 2 (0,0 for spheres) */
} b200pt_hit;
#define B200PT_MISS 0xFFFFFFFFu

/* counters since the last b200pt_stats_reset */
typedef struct b200pt_stats {
    uint64_t extend_rays;      /* closest-hit path rays + MIS probe rays */
    uint64_t shadow_rays;      /* any-hit visibility rays */
    uint64_t path_vertices;    /* entries processed by the shade kernel */
    uint64_t samples;          /* camera paths started */
    uint64_t iterations;       /* wavefront iterations */
    uint64_t kernel_launches;  /* CUDA kernels launched by this library */
    uint64_t launches_extend;  /* launches of the closest-hit kernel */
    uint64_t launches_shadow;  /* launches of the any-hit kernel */
    uint64_t launches_shade;   /* launches of generate / shade / probe-resolve / accumulate */
    float ms_extend;           /* device time in the closest-hit kernel (CUDA events; only with stage timing on) */
    float ms_shadow;           /* device time in the any-hit kernel */
    float ms_shade;            /* device time in generate / shade / probe-resolve / accumulate */
    float ms_total;            /* device time of render_frame calls, first to last kernel (always measured) */
    /* guiding update (b200pt_guiding_update) */
    uint64_t guiding_samples;               /* valid DirectionalData records consumed */
    uint64_t guiding_em_sample_iterations;  /* sum over regions of N_r x EM iterations of fit / updateFit */
    uint64_t guiding_regions_fit;           /* non-empty regions processed */
    uint64_t launches_guiding;              /* sort + fit kernels launched */
    float ms_guiding_sort;                  /* device time: sort by region + preFit + SoA conversion */
    float ms_guiding_fit;                   /* device time: per-region EM / merge / split / statistics / pack */
    /* region-sharded refit across ranks (b200pt_guiding_update_all_ranks); single GPU: all_ranks == samples, rest 0 */
    uint64_t guiding_samples_all_ranks;     /* valid records of all ranks that took part in the updates */
    uint64_t guiding_bytes_received;        /* sample bytes this rank fetched from its peers' HBM (24 B per record) */
    float ms_guiding_exchange;              /* device time: barrier + record exchange into region-contiguous order */
    float ms_guiding_gather;                /* device time: all-gather of the fitted mixtures */
} b200pt_stats;

#define B200PT_MAX_GUIDING_SPLITS 11         /* 2048 initial regions; adaptive refinement (splitRegions) may double the initial count
                                              (per-region buffers: 1024 regions up to 9 splits, 4096 for 10 and 11) */
typedef struct b200pt_ctx b200pt_ctx;       /* opaque, one per GPU */
typedef struct b200pt_scene b200pt_scene;   /* opaque host-side scene (loader output) */

enum { B200PT_IMAGE_OUTPUT = 0, B200PT_IMAGE_ACCUM = 1, B200PT_IMAGE_ESTIMATE = 2 };

const char *b200pt_last_error(void);
int b200pt_device_count(void);

/* replaces RayTracingApp::RayTracingApp(width,height,icSize,guidingSplits,...) — src/RayTracingApp.cpp:11-53:
 * allocates output / accumulate / estimate images (:55-73), IC buffers (src/IrradianceCache.cpp:46-79),
 * the W*H*16 DirectionalData buffer (src/SampleCollector.cpp:29-51) and 2^splits guiding regions */
int b200pt_create(int device_ordinal, int width, int height, int ic_size, int guiding_splits, b200pt_ctx **out);
int b200pt_destroy(b200pt_ctx *ctx);        /* RayTracingApp::cleanup, src/RayTracingApp.cpp:1219-1241 */

/* replaces SceneLoader::createVulkanObjects (src/SceneLoader.cpp:67-87): buffer uploads + BLAS/TLAS build
 * (:1148-1193).  Builds a SAH BVH on the host, flattens it to compressed BVH8 nodes, uploads. Also rebuilds
 * the guiding region tree from the scene AABB (src/PathGuiding.cpp:11-24,81-104) and resets the IC. */
int b200pt_set_scene(b200pt_ctx *ctx, const b200pt_scene_desc *scene);

/* Host-only structural check of the acceleration structure b200pt_set_scene would build for `scene` (no context, no GPU, no
 * traversal): flattens the instances to world-space triangles like set_scene, builds the BVH8 and verifies what the
 * device traversal relies on — every primitive in exactly one leaf, every triangle inside the DEQUANTISED box of its slot
 * and of all its ancestors' slots, consistent child / triangle indexing, the recorded depth.  The reference has no
 * counterpart (its BLAS / TLAS come from the driver, src/SceneLoader.cpp:1101-1193); this is the CPU-side test hook of
 * our builder.  The first six fields are violation counts (all 0 for a valid tree), the rest statistics. */
typedef struct b200pt_bvh_report {
    uint32_t missing_prims, duplicate_prims, outside_box, bad_meta, depth_mismatch, unreachable_nodes;
    uint32_t num_nodes, num_tris, max_depth, inner_children, leaf_children;
} b200pt_bvh_report;
int b200pt_scene_bvh_check(const b200pt_scene_desc *scene, b200pt_bvh_report *report);

/* replaces RayTracingApp::updateUniformBuffer (src/RayTracingApp.cpp:559-585); column-major mat4, inverses are
 * computed inside like :578-579 */
int b200pt_set_camera(b200pt_ctx *ctx, const float view[16], const float proj[16]);

/* replaces pushConstants + traceRaysKHR(W,H,1) + waitIdle (src/RayTracingApp.cpp:1175-1201, :114).
 * pc->randomUInt is the frame seed (the reference draws it from glm::linearRand, quirk 10). */
int b200pt_render_frame(b200pt_ctx *ctx, const b200pt_push_constants *pc);

/* `count` consecutive frames in one call: the accumulation loop of the evaluation modes ("N samples -> time -> EXR",
 * src/RayTracingApp.cpp:93-97 and :159-215, where drawCallback runs frame after frame with a fresh randomUInt and
 * previousFrames + 1).  The images after the call are IDENTICAL to `count` calls of b200pt_render_frame.  When the
 * frames are plain path-tracing frames that differ only in randomUInt / previousFrames (no irradiance cache, ADRRS,
 * guiding, estimate, AOVs or alpha-tested textures) a pixel that has finished frame f starts frame f + 1 without
 * waiting for the other pixels, which removes the drain tail of every frame but the last; otherwise the frames run
 * one after the other. */
int b200pt_render_frames(b200pt_ctx *ctx, const b200pt_push_constants *pcs, int count);

/* replaces PostProcessing::saveOffscreenImage read-back (src/PostProcessing.cpp:310-338): W*H RGBA32F */
int b200pt_read_image(b200pt_ctx *ctx, int which, float *rgba_host);
int b200pt_write_image(b200pt_ctx *ctx, int which, const float *rgba_host);
/* same, but into / from DEVICE memory owned by the caller (e.g. a torch tensor used for an NCCL all-reduce) */
int b200pt_read_image_device(b200pt_ctx *ctx, int which, void *rgba_device);
int b200pt_write_image_device(b200pt_ctx *ctx, int which, const void *rgba_device);

/* level-1 parity hook: traversal only.  any_hit=0: closest hit (rgen:1011-1022 semantics);
 * any_hit=1: visibility (rgen:626-637): hit[i].prim != MISS means occluded. Host buffers. */
int b200pt_trace_rays(b200pt_ctx *ctx, const b200pt_ray *rays, int64_t n, b200pt_hit *hits, int any_hit);
/* same with DEVICE buffers, no copies: used for kernel-only timing */
int b200pt_trace_rays_device(b200pt_ctx *ctx, const void *rays_device, int64_t n, void *hits_device, int any_hit);

int b200pt_stats_get(b200pt_ctx *ctx, b200pt_stats *out);
/* per-kernel CUDA-event timing of the wavefront stages (replaces the reference's std::chrono prints):
 * 0 = off, 1 = every stage kernel (costs ~10 % of a frame: ~8 event records per wavefront iteration), 2 = only the
 * trace kernel (b200pt_stats.ms_extend / launches_extend stay valid, the other stage times read 0) */
int b200pt_set_stage_timing(b200pt_ctx *ctx, int enabled);
int b200pt_stats_reset(b200pt_ctx *ctx);
int b200pt_synchronize(b200pt_ctx *ctx);
/* CUDA-event stopwatch on the context's own stream (the stream every kernel of this library is launched on):
 * start records an event, stop records a second one, waits for it and returns the device time between them */
int b200pt_timer_start(b200pt_ctx *ctx);
int b200pt_timer_stop(b200pt_ctx *ctx, float *ms);

/* ---- guiding (PathGuiding / SampleCollector / lightpmm) ---------------------------------------- */

/* knobs of PathGuiding.h:121-129 and VMMFactoryProperties (VMMFactory.h:45-61, PathGuiding.cpp:33-40) */
typedef struct b200pt_guiding_params {
    int32_t useParallaxCompensation;   /* enableParallaxCompensationForOptimization, RayTracingApp.h:208 */
    int32_t splitAndMerge;             /* 1 */
    int32_t minSamplesForMerging;      /* 8192 */
    int32_t minSamplesForSplitting;    /* 4096 */
    int32_t minSamplesForPostSplitFitting; /* 4096 */
    float splitMinDivergence;          /* 0.5 */
    float mergeMaxDivergence;          /* 0.025 */
    int32_t numInitialComponents;      /* 8 */
    int32_t minItr;                    /* 1 */
    int32_t maxItr;                    /* 100 */
    float relLogLikelihoodThreshold;   /* 0.005 */
    float initKappa;                   /* 5 */
    float maxKappa;                    /* 50000 */
    float vPrior;                      /* 0.01 */
    float rPrior;                      /* 0 */
    float rPriorWeight;                /* 1 */
    int32_t splitRegions;              /* PathGuiding::splitRegions, src/PathGuiding.h:121 — 0: adaptive region refinement off */
    float samplesForRegionSplit;       /* 10000, src/PathGuiding.h:122: a region whose mixture saw more samples is halved */
} b200pt_guiding_params;
void b200pt_default_guiding_params(b200pt_guiding_params *p);

/* replaces PathGuiding::update(SampleCollector) (src/PathGuiding.cpp:276-312): sort by region, preFit,
 * fit/updateFit, merge/split, distance update, pack VMM_Theta — all on the device. */
int b200pt_guiding_update(b200pt_ctx *ctx, const b200pt_guiding_params *params);
int b200pt_guiding_region_count(b200pt_ctx *ctx, int *count);
int b200pt_guiding_get_aabbs(b200pt_ctx *ctx, b200pt_aabb *out, int n);
int b200pt_guiding_get_vmms(b200pt_ctx *ctx, b200pt_vmm_theta *out, int n);
int b200pt_guiding_put_vmms(b200pt_ctx *ctx, const b200pt_vmm_theta *in, int n);
/* parity hooks for the W*H*16 DirectionalData buffer (binding 18): raw order, host memory */
int b200pt_guiding_get_samples(b200pt_ctx *ctx, b200pt_directional_data *out, int64_t n);
int b200pt_guiding_put_samples(b200pt_ctx *ctx, const b200pt_directional_data *in, int64_t n);
int64_t b200pt_guiding_sample_capacity(b200pt_ctx *ctx);
/* the first n records of binding 18 into DEVICE memory owned by the caller (e.g. the send buffer of an NCCL all-gather
 * across the GPUs of a spp-sharded training run; the gathered records then go to b200pt_guiding_update_device) */
int b200pt_guiding_get_samples_device(b200pt_ctx *ctx, void *dst_device, int64_t n);
/* PathGuiding is rebuilt (regions kept, mixtures re-initialised with VMMFactory::initialize, firstFit = true) — what
 * RayTracingApp::sceneSwitcher does by constructing a new PathGuiding (src/RayTracingApp.cpp:327-367) */
int b200pt_guiding_reset(b200pt_ctx *ctx, const b200pt_guiding_params *params);
/* like b200pt_guiding_update, but on `n` caller-provided records in HOST memory instead of the context's own W*H*16
 * buffer (copied to a device staging buffer first): the reference-facing call for a host that keeps its own
 * SampleCollector, and the e2e leg of the EM benchmark */
int b200pt_guiding_update_host(b200pt_ctx *ctx, const b200pt_guiding_params *params, const b200pt_directional_data *samples, int64_t n);
/* same with `n` records already resident in DEVICE memory (no copy) */
int b200pt_guiding_update_device(b200pt_ctx *ctx, const b200pt_guiding_params *params, const void *samples_device, int64_t n);
/* Summation order of the per-region sample sums.  lightpmm adds the sufficient statistics sample after sample in float
 * (VMMFactory.h:497-536, incremental*.h); STRICT (default) keeps exactly that order on the device (parallel over the
 * 4K+2 running sums instead of over the samples), so a refit on the same records in the same order reproduces the
 * reference's mixtures to float rounding of the per-sample terms (1e-4, same EM iteration counts).  REORDERED sums
 * per-thread partitions block-parallel (thread-block cluster per region): results differ from the reference's like
 * the reference's differ from itself under its unstable sample sort (weights ~1e-4, kappa ~2e-3, rare iteration flips).
 * Env B200PT_GUIDING_ORDER=strict|reordered overrides. */
enum { B200PT_GUIDING_ORDER_STRICT = 0, B200PT_GUIDING_ORDER_REORDERED = 1 };
int b200pt_guiding_set_order(b200pt_ctx *ctx, int order);
/* parity hooks: the sorted + pre-fitted samples of the last update (what SampleCollector::getSortedData + preFit
 * produce; positions are the region's parallax mean) with region offsets [regions + 1]; either pointer may be NULL */
int64_t b200pt_guiding_sorted_count(b200pt_ctx *ctx);
int b200pt_guiding_get_sorted(b200pt_ctx *ctx, b200pt_directional_data *out, uint32_t *region_offsets);
/* full per-region fit state: scalars5 = K, sampleWeight, numSamples, totalNumSamples, numEMIterations; per_component =
 * 14 rows of 16: weight, kappa, r, mu x/y/z, distance, distance sumWeights, chi2 value, chi2 numSamples, cov xx/yy/xy,
 * cov sumWeights (lightpmm PMM + PMM_ExtraData, src/PathGuiding.h:77-85) */
int b200pt_guiding_get_state(b200pt_ctx *ctx, int region, float scalars5[5], float per_component[224]);
/* known-answer hook: lightpmm::exp (PMM_APPROX_EXP fastexp, pmm-vcl.h:157-184) as evaluated by the device code */
int b200pt_guiding_fastexp(b200pt_ctx *ctx, const float *in_host, float *out_host, int n);
/* self-test hook: the division inside fastexp (pmm-vcl.h:171, 27.7280233 / (4.84252568 - z)) is evaluated on the device
 * by a reciprocal + Newton + Markstein sequence instead of nvcc's guarded `/`; compares it bit for bit with IEEE division
 * for EVERY float divisor in [lo, hi] (fastexp reaches 2.84 < d < 4.85) */
int b200pt_guiding_selftest_division(b200pt_ctx *ctx, float lo, float hi, uint64_t *mismatches, uint64_t *tested);

/* ---- deterministic elementary functions (include/b200pt_detmath.h) --------------------------------
 * The kernels evaluate sin / cos / tan / asin / acos / atan / atan2 / pow / log / exp with double-precision kernels made of
 * IEEE add / mul / div / sqrt only (GLSL's results are the hardware's; CUDA's and glibc's libm differ from each other in the
 * last bit, which used to make 0.5-15 % of same-seed pixels differ between this library and its CPU checker).  Parity
 * hooks: `fn` over n inputs on the device / by the host compilation of the same header; b may be NULL for unary functions. */
enum { B200PT_DM_SIN = 0, B200PT_DM_COS, B200PT_DM_TAN, B200PT_DM_ASIN, B200PT_DM_ACOS, B200PT_DM_ATAN, B200PT_DM_ATAN2, B200PT_DM_POW,
       B200PT_DM_LOG, B200PT_DM_EXP,
       B200PT_DM_DIV3 /* component (i mod 3) of the kernels' vec3 / scalar: vec3(a) / b; on the host plain a / b */ };
int b200pt_detmath_eval(b200pt_ctx *ctx, int fn, const float *a, const float *b, float *out, int n);
int b200pt_detmath_eval_host(int fn, const float *a, const float *b, float *out, int n);

/* ---- AOVs ---------------------------------------------------------------
 * The quantities behind the reference's depth / split debug views (shaders/raytrace.rgen:1653-1655, :1677-1679,
 * :1705-1707, :1722-1741) as a per-pixel RGBA32F layer of the LAST rendered frame instead of a view mode:
 * x = maxReachedDepth, y = depthSum, z = depthsCounter (paths incl. drained splits), w = nextSplitSlot.
 * VISU_DEPTH_MAX = x / maxDepth, VISU_DEPTH_AVERAGE = y / z / maxDepth, VISU_SPLITS = w / 10; the estimate image is
 * B200PT_IMAGE_ESTIMATE.  Off by default (costs one read-modify-write per finished path). */
int b200pt_set_aovs(b200pt_ctx *ctx, int enabled);
int b200pt_read_aovs(b200pt_ctx *ctx, float *rgba);

/* ---- checkpoint / resume ------------------------------------------------
 * Everything a later frame depends on: the three images, the irradiance cache (header, data, spheres) and the guiding
 * state (region boxes incl. adaptive splits, mixtures with their running statistics, packed VMM_Thetas, firstFit).  The
 * frame driver's state is the plain struct b200pt_app, which the host stores itself.  A context that loads a
 * checkpoint must have been created with the same width, height, ic_size and guiding_splits and have its scene set;
 * frames rendered after b200pt_load_state are bit-identical to those of the run that saved it. */
int b200pt_save_state(b200pt_ctx *ctx, const char *path);
int b200pt_load_state(b200pt_ctx *ctx, const char *path);

/* ---- multi-GPU (ours: the reference is single-GPU) -----------------------------------------------
 * One context per GPU and per process; the image shards by sample index: every rank renders disjoint
 * frames of the full image, the running means are combined with one NCCL all-reduce, and guiding training frames
 * exchange their compacted DirectionalData so that every region is refitted once, by one rank, on all ranks' records.  NCCL is loaded
 * with dlopen("libnccl.so.2") on first use; B200PT_E_STATE when it is not installed. */
#define B200PT_COMM_ID_BYTES 128
int b200pt_comm_unique_id(char id[B200PT_COMM_ID_BYTES]);      /* ncclGetUniqueId: call on one rank, hand the bytes to the others */
int b200pt_comm_init(b200pt_ctx *ctx, const char id[B200PT_COMM_ID_BYTES], int rank, int nranks);
int b200pt_comm_destroy(b200pt_ctx *ctx);
/* image `which` of every rank becomes sum_r(frames_r * image_r) / sum_r(frames_r): the mean over all frames of all ranks */
int b200pt_reduce_image(b200pt_ctx *ctx, int which, int frames_local);
/* all-gather of the ranks' whole DirectionalData buffers (rank order) into the context; *total_out = records gathered.
 * Parity hook: the concatenation is what a single-GPU update would see (b200pt_guiding_update_device on it must give the
 * mixtures of b200pt_guiding_update_all_ranks bit for bit); the product path below never moves INVALID records. */
int b200pt_allgather_samples(b200pt_ctx *ctx, int64_t *total_out);
/* PathGuiding::update (src/PathGuiding.cpp:276-312) across the ranks of a spp-sharded training run, region-sharded:
 *   1. every rank sorts + compacts ITS records by region on the device (SampleCollector::getSortedData, :76-131);
 *   2. the per-region counts are all-gathered (regions x 4 B) and every rank derives the same plan: regions are dealt
 *      to ranks longest-first by total count, so each rank fits about 1/N of the samples;
 *   3. the owner of a region collects that region's valid records (24 B each) from the other ranks — one kernel that
 *      reads the peers' HBM over NVLink through CUDA-IPC mappings and writes them region-contiguous in rank order
 *      (fallback when IPC is unavailable or B200PT_EXCHANGE=nccl: grouped ncclSend / ncclRecv + the same kernel);
 *   4. every rank fits only its regions (updateRegion, :350-451);
 *   5. the fitted mixtures + packed VMM_Thetas are all-gathered (1.8 KB per region).
 * Result: bit-identical mixtures on every rank, equal to a single-GPU update on the concatenation of the ranks'
 * buffers in rank order.  Collective: every rank must call it, with the same params and the same buffer size. */
int b200pt_guiding_update_all_ranks(b200pt_ctx *ctx, const b200pt_guiding_params *params);
/* same on `n` caller-provided records in device memory per rank (n equal on all ranks; INVALID records allowed) */
int b200pt_guiding_update_all_ranks_device(b200pt_ctx *ctx, const b200pt_guiding_params *params, const void *samples_device, int64_t n);
/* parity hook: the plan of b200pt_guiding_update_all_ranks (steps 2-3) as rank `rank` of `nranks` would compute it from the
 * per-rank region counts counts[nranks][regions] — no communicator and no samples involved, so one GPU can check ownership
 * and layouts for any rank count.  Outputs (regions entries unless noted): owner, region_begin / region_len (this rank's fit
 * layout; len 0 for regions it does not own), src_start[nranks][regions] (where region g starts in rank s's sorted buffer),
 * active (this rank's non-empty regions, largest first; summary[0] entries), summary = {numActive, numOwned, numSegments,
 * localValid, ownedSamples, totalSamples}, segments[numOwned * nranks][4] = {src rank, src offset, dst offset, length}.
 * peer_mode 1: source offsets address the peers' buffers, 0: the staging buffer of the send/recv fallback. */
int b200pt_guiding_plan_debug(b200pt_ctx *ctx, const uint32_t *counts, int nranks, int rank, int peer_mode, uint8_t *owner, uint32_t *region_begin,
                              uint32_t *region_len, uint32_t *src_start, uint32_t *active, uint32_t summary[6], uint32_t *segments);
/* 0 = no communicator, 1 = records travel by ncclSend/ncclRecv, 2 = peers' buffers are read directly (CUDA IPC over NVLink);
 * decided collectively by the first b200pt_guiding_update_all_ranks */
int b200pt_comm_exchange_mode(b200pt_ctx *ctx);

/* ---- irradiance cache parity hooks (bindings 10,12,13) ---------------------------------------- */
int b200pt_ic_get(b200pt_ctx *ctx, b200pt_cache_header *hdr, b200pt_cache_data *data, b200pt_sphere *spheres, int n);
int b200pt_ic_put(b200pt_ctx *ctx, const b200pt_cache_header *hdr, const b200pt_cache_data *data,
                  const b200pt_sphere *spheres, int n);

/* ---- host side: the reference's scene surface (src/SceneLoader.*, src/MitsubaXML.h) ------------ */

void b200pt_default_push_constants(b200pt_push_constants *pc);   /* RayTracingApp.h:116-165 defaults */

/* SceneLoader(filepath): parses a Mitsuba-0.6 XML (or the reference's JSON) scene into the buffers above */
int b200pt_scene_load(const char *path, b200pt_scene **out);
int b200pt_scene_free(b200pt_scene *scene);
int b200pt_scene_get_desc(const b200pt_scene *scene, b200pt_scene_desc *out);
/* camera: origin, target, up, vfov as parsed (SceneLoader.h:147-150 defaults) */
int b200pt_scene_get_camera(const b200pt_scene *scene, float origin[3], float target[3], float up[3], float *vfov);
/* CameraController::lookAt + getViewMatrix + getProjMatrix (src/CameraController.cpp:76-107) */
void b200pt_camera_matrices(const float origin[3], const float target[3], const float up[3], float vfov_deg,
                            float aspect, float view[16], float proj[16]);
/* glm::inverse as used for viewInverse / projInverse (src/RayTracingApp.cpp:578-579); returns 0 when singular.
 * b200pt_set_camera applies exactly this function, so a checker can feed the same inverses to its own renderer. */
int b200pt_mat4_inverse(const float m[16], float out[16]);
/* CommonOps::writeEXR (src/CommonOps.cpp:12-38): 3 x FLOAT channels from RGBA32F */
int b200pt_write_exr(const char *path, const float *rgba, int width, int height);
/* CommonOps::readEXR (src/CommonOps.cpp:40-65): RGBA float (values pass through half like Imf::Rgba) */
int b200pt_read_exr(const char *path, float **rgba_out, int *width, int *height);
/* bitmap textures as SceneLoader::addTexture loads them (`stbi_load(path, &w, &h, &channels, 4)`,
 * src/SceneLoader.cpp:198-207): baseline JPEG or PNG -> RGBA8, row 0 first; release with b200pt_free */
int b200pt_read_image_file(const char *path, uint8_t **rgba_out, int *width, int *height);
void b200pt_free(void *p);

/* ---- frame driver: the per-frame host logic of RayTracingApp ------------------------------------
 * RayTracingApp::raytrace (src/RayTracingApp.cpp:1120-1170) decides each frame's push constants from the user's
 * settings: previousFrames (running mean), the irradiance-cache prepare frames (1 spp, IC on, ADRRS postponed),
 * the one estimate frame that ADRRS needs (setEstimateRTSettings, :1206-1217) and the restore afterwards;
 * RayTracingApp::drawCallback (:120-157) stops guiding training after numGuidingOptimizations updates and counts the
 * samples of the "collect N samples" evaluation (:159-186).  Pure host state, no device work: usable without a GPU. */
typedef struct b200pt_app {
    b200pt_push_constants settings;          /* rtPushConstants: what the user edits (the ImGui panel of the reference) */
    int32_t accumulateResults;               /* RayTracingApp.h:147 (the evaluation modes switch it on, :93,98) */
    int32_t hasInputChanged;                 /* set it after editing `settings`: restarts the running mean */
    int32_t irradianceCachePrepareFrames;    /* RayTracingApp.h:175, default 50 */
    int32_t currentPrepareFrames;
    int32_t numGuidingOptimizations;         /* RayTracingApp.h:180, default 6; -1 = keep training */
    int32_t currentGuidingOptimizations;     /* -1 = not training */
    int32_t loadBackupNextIteration;
    int32_t activateADRRSAfterPrepareFrames;
    int64_t evalCurrentSamples;              /* samples per pixel that count for the image (prepare frames do not) */
    b200pt_push_constants backupPushConstant;
} b200pt_app;
void b200pt_app_init(b200pt_app *app);                       /* reference defaults, previousFrames = -1 */
void b200pt_app_scene_switched(b200pt_app *app);             /* sceneSwitcher, :327-367: currentPrepareFrames = 0, input changed */
/* start of a frame: advances the state machine and returns the constants to render with (copy of app->settings) */
void b200pt_app_begin_frame(b200pt_app *app, uint32_t frame_seed, b200pt_push_constants *frame_pc);
/* end of a frame: guiding-optimisation bookkeeping and sample accounting; returns 1 when PathGuiding::update must
 * run on the frame's samples (drawCallback :143-154) */
int b200pt_app_end_frame(b200pt_app *app);
/* begin_frame + b200pt_render_frame + end_frame (+ b200pt_guiding_update with `guiding_params` when due) */
int b200pt_app_draw_frame(b200pt_app *app, b200pt_ctx *ctx, uint32_t frame_seed, const b200pt_guiding_params *guiding_params);

#ifdef __cplusplus
}
#endif
#endif /* B200PT_H */
