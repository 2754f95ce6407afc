// Wavefront path-tracing kernels: generate -> extend -> shade -> shadow-connect, with per-stage ray queues.
// The reference is a per-pixel megakernel (shaders/raytrace.rgen:1623-1827); here every pixel owns ONE path at a
// time (its samples are serialised exactly like the megakernel's spp loop, so the per-pixel RNG stream is consumed in
// the same order — SURVEY.md Appendix A) and the bounce loop is turned inside out into queue-driven kernels.
#pragma once
#include "shading.cuh"
#include "guiding_device.cuh"
#include "ic_device.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace b200pt {

struct FrameParams {
    b200pt_push_constants pc;
    float view[16], proj[16], viewInv[16], projInv[16];   // CameraMatrices UBO, rgen:128-134
    int width, height;
    int numPixels;
    int samplesPerPixel;     // pc.isIrradiancePrepareFrame ? 1 : pc.samplesPerPixel (rgen:1661)
};

// queue counters (device).  The host never needs them to size a launch: every stage kernel is persistent / grid-stride
// and reads its own element count from here, so a frame's iterations are issued back-to-back without host round trips.
enum { CNT_PATH0 = 0, CNT_PATH1 = 1, CNT_PROBE = 2, CNT_SHADOW = 3, CNT_INLINE_SHADOW = 4, CNT_WORK_TRACE = 5, CNT_SHADE_N = 6,
       CNT_TICKET = 7,      // k_probe_resolve_prep: blocks that have finished (the last one does the iteration bookkeeping)
       CNT_ICQ_N = 8,       // IC frames: queue entries that need a cache lookup (length of the cell-sorted order, k_icq_*)
       CNT_REGEN = 9,       // IC frames: pixels whose path ended in this shade pass (k_regen starts their next path)
       CNT_NUM = 10 };
// 64-bit statistics accumulated on the device by k_iter_prep
enum { DST_EXTEND = 0, DST_SHADOW = 1, DST_VERTICES = 2, DST_ITERATIONS = 3, DST_NUM = 4 };

struct TraceTuning { uint32_t chunk; int refillMin; };

// Frame batch (b200pt_render_frames).  Every pixel runs its samples one after the other, so a frame ends with a long
// tail of iterations in which only the pixels with long paths are alive (cornell-dielectric: 405 iterations per
// 16-spp frame, half of them with < 25 % of the paths; ~9 % of the frame time).  In a batch a pixel that has finished
// frame f starts frame f + 1 at once: it re-seeds with tea(pixel, frameSeed[f + 1]) exactly like k_generate and sums
// into the other half of the doubled pixelSum buffer (parity = f & 1), because the last light samples of frame f are
// still in the shadow / probe queues.  Frame f - 1 is folded into the image (accumulatePixel = k_accumulate's
// arithmetic) when the pixel finishes frame f, the last frame by k_accumulate after the queues have drained — per
// pixel the same operations in the same order as `count` calls of b200pt_render_frame, so the images are identical.
struct FrameBatch {
    const uint32_t *frameSeed;       // pushC.randomUInt of every frame
    const uint32_t *framePrev;       // pushC.previousFrames of every frame
    int numFrames;                   // 0 = no batch
    float4 *image, *accum;
};
#define ST_FRAME_SHIFT 16            // sampleIdx word: sample of the frame (low 16 bits) | frame of the batch

struct Wavefront {
    float4 *pathRayO[2];     // xyz origin, w = path id bits
    float4 *pathRayD[2];     // xyz direction
    float4 *pathHit;         // t, prim bits, u, v — indexed like the current path queue
    float4 *probeRayO;       // MIS probe rays (rgen:665-728): xyz origin
    float4 *probeRayD;       // xyz direction
    float4 *probeHit;
    float4 *probeA;          // bsdf value (f*cos) xyz, w = pdfMat
    float4 *probeB;          // path throughput xyz, w = pixel id bits
    float4 *shRayO;          // shadow rays: xyz origin, w = pixel id bits
    float4 *shRayD;          // xyz direction, w = tmax
    float4 *shC;             // contribution to add when unoccluded (path throughput applied)
    float4 *shG;             // guiding training only: the NEE value without the path throughput (xyz), w = currentSampleOffset bits
    uint32_t *seed;          // per-pixel LCG state
    float4 *thr;             // per-path throughput
    uint32_t *state;         // depth (0..15) | followCount (16..23) | flags (24..)
    uint32_t *sampleIdx;     // index of the sample the path is working on
    float4 *pixelSum;        // sum of the frame's sample radiances per pixel
    uint32_t *counters;      // CNT_*
    unsigned long long *dstats;   // DST_*
    GuidingView guide;       // region tree + mixtures (useGuiding / updateGuiding)
    GuidingRecord rec;       // sample-recording state (updateGuiding); rec.samples == nullptr when not recording
    ICState ic;              // irradiance cache / ADRRS state (useIrradianceCache / useADRRS / splitOnFirst frames)
    float4 *aov;             // optional per-pixel layer: max depth, depth sum, path count, split count (b200pt_set_aovs)
    FrameBatch batch;        // b200pt_render_frames: the frames a pixel walks through without waiting for the other pixels
    uint32_t *regenQ;        // IC / ADRRS frames: pixel ids handed from k_shade to k_regen (nullptr: next paths start inside k_shade)
    const uint32_t *shadeOrder;   // IC / ADRRS frames: the order in which k_shade<.,IC> walks the path queue (k_icq_*; nullptr: queue order)
};

#define ST_ADDNEXT (1u << 24)
#define ST_FOLLOW  (1u << 25)
#define ST_ODDFRAME (1u << 26)   // frame batches: the path belongs to an odd frame of the batch

// warp-aggregated queue append: the lanes that are converged at the call share ONE atomic and get consecutive slots
// (keeps the queues roughly path-ordered and the counter traffic 32x lower than per-thread atomics)
__device__ __forceinline__ uint32_t queuePush(uint32_t *counter) {
    // by hand: cooperative_groups' coalesced_group::shfl on a partial mask was 6 % of the shade kernel's instructions
    const unsigned mask = __activemask();
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (int(lane) == leader) base = atomicAdd(counter, uint32_t(__popc(mask)));
    base = __shfl_sync(mask, base, leader);
    return base + uint32_t(__popc(mask & ((1u << lane) - 1u)));
}

// rgen:1487-1494
__device__ __forceinline__ void cameraRay(const FrameParams &fp, uint32_t &seed, int px, int py, vec3 &origin, vec3 &direction) {
    float r1 = rndNegPos(seed), r2 = rndNegPos(seed);
    float ux = ((float(px) + 0.5f) + r1 / 2.0f) / float(fp.width);
    float uy = ((float(py) + 0.5f) + r2 / 2.0f) / float(fp.height);
    float dx = ux * 2.0f - 1.0f, dy = uy * 2.0f - 1.0f;
    origin = mat4MulPoint(fp.viewInv, V3(0.0f), 1.0f);
    vec3 target = mat4MulPoint(fp.projInv, V3(dx, dy, 1.0f), 1.0f);
    direction = normalize(mat4MulPoint(fp.viewInv, normalize(target), 0.0f));
}

// ---------------------------------------------------------------------------------------------------------------
// generate: rgen main() :1625 (seed) and the first getCameraRay of the spp loop (:1668-1670)
// slot: where the camera ray goes in path queue `q` (the pixel index for a full frame)
__device__ __forceinline__ void generatePixel(const FrameParams &fp, const Wavefront &wf, const int p, uint32_t seed, const int q, const uint32_t slot) {
    if (wf.aov) wf.aov[p] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (wf.ic.newCount) wf.ic.newCount[p] = 0u;
    if (wf.ic.splitState) wf.ic.splitState[p] = 0u;
    const int px = p % fp.width, py = p / fp.width;
    vec3 o, d;
    cameraRay(fp, seed, px, py, o, d);
    wf.pathRayO[q][slot] = make_f4(o, __int_as_float(p));
    wf.pathRayD[q][slot] = make_f4(d, 0.0f);
    wf.seed[p] = seed;
    wf.thr[p] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    wf.state[p] = ST_ADDNEXT;     // depth 0, addNextDirectLights = addFirstHitLight = true (rgen:995,1676)
    wf.sampleIdx[p] = 0;
    wf.pixelSum[p] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (wf.batch.numFrames) wf.pixelSum[p + fp.numPixels] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (fp.pc.updateGuiding) {           // resetSamplesToInvalid (rgen:1615-1621, :1642-1643) + fresh recording state
        for (int i = 0; i < B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL; i++) wf.rec.samples[p * B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL + i].flags = B200PT_INVALID_REGION;
        wf.rec.state[p] = make_int4(0, 0, -1, -1);
        wf.rec.distanceFactor[p] = 1.0f;
        wf.rec.pathSum[p] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}
// DEFER_UPDATED (irradiance-cache frames whose cache update runs on a second stream beside the wavefront loop, api.cu): the pixels
// the update draw selects are left out — their RNG stream continues where k_ic_update leaves it, so they start with
// k_generate_list once it has finished; the others are appended to the queue instead of sitting at their pixel index.
template <bool DEFER_UPDATED>
__global__ void __launch_bounds__(256) k_generate(FrameParams fp, Wavefront wf) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= fp.numPixels) return;
    uint32_t seed = tea(uint32_t(p), fp.pc.randomUInt);
    if (fp.pc.useIrradianceCache) {      // rgen:1639-1641: the update draw; k_ic_update has advanced the stream of the selected pixels
        if (rnd(seed) < fp.pc.irradianceUpdateProb) {
            if (DEFER_UPDATED) return;
            seed = wf.seed[p];
        }
    }
    generatePixel(fp, wf, p, seed, 0, DEFER_UPDATED ? queuePush(&wf.counters[CNT_PATH0]) : uint32_t(p));
}
__global__ void __launch_bounds__(256) k_generate_list(FrameParams fp, Wavefront wf, const uint32_t *__restrict__ list, const uint32_t *__restrict__ count, int q) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= *count) return;
    const int p = int(list[e]);
    generatePixel(fp, wf, p, wf.seed[p], q, queuePush(&wf.counters[CNT_PATH0 + q]));
}

// ---------------------------------------------------------------------------------------------------------------
// trace: ONE persistent launch per wavefront iteration traces all three ray queues — path rays and MIS probe rays
// (closest hit, rgen:1011-1022 / :671-682) and NEE shadow rays (any hit, rgen:622-638 + shadow.rmiss) whose
// contribution is splatted when unoccluded (rgen:657-663).  Ray index space: [path | probe | shadow].
template <bool REC>       // REC: guiding training frame — shadow hits also feed the recorded samples (kept out of the plain kernel: +7 registers)
struct WavefrontRayIO {
    const float4 *pathO, *pathD; float4 *pathHit;
    const float4 *probeO, *probeD; float4 *probeHit;
    const float4 *shO, *shD, *shC, *shG; float4 *pixelSum;
    GuidingRecord rec;
    uint32_t nPath, nProbe;
    __device__ __forceinline__ void load(uint32_t idx, vec3 &o, vec3 &d, float &tmin, float &tmax, bool &any) const {
        float4 ro, rd;
        tmin = PT_TMIN;
        if (idx < nPath) { ro = pathO[idx]; rd = pathD[idx]; tmax = PT_TMAX; any = false; }
        else if (idx < nPath + nProbe) { ro = probeO[idx - nPath]; rd = probeD[idx - nPath]; tmax = PT_TMAX; any = false; }
        else { ro = shO[idx - nPath - nProbe]; rd = shD[idx - nPath - nProbe]; tmax = rd.w; any = true; }
        o = make_vec3(ro); d = make_vec3(rd);
    }
    __device__ __forceinline__ void store(uint32_t idx, const HitRec &h, bool any) const {
        if (!any) {
            const float4 out = make_float4(h.t, __uint_as_float(h.prim), h.u, h.v);
            if (idx < nPath) pathHit[idx] = out; else probeHit[idx - nPath] = out;
        } else if (h.prim == PT_MISS) {
            const uint32_t k = idx - nPath - nProbe;
            const float4 c = shC[k];
            const int pix = __float_as_int(shO[k].w);
            float *dst = reinterpret_cast<float *>(&pixelSum[pix]);
            atomicAdd(dst + 0, c.x); atomicAdd(dst + 1, c.y); atomicAdd(dst + 2, c.z);
            if (REC) {               // the deferred half of `updateSamples(currentSampleOffset, neeLight)` (rgen:1164-1165)
                float *ps = reinterpret_cast<float *>(&rec.pathSum[pix]);
                atomicAdd(ps + 0, c.x); atomicAdd(ps + 1, c.y); atomicAdd(ps + 2, c.z);
                const float4 g = shG[k];
                updateSamples<true>(rec, pix, rec.state[pix].x, __float_as_int(g.w), make_vec3(g));
            }
        }
    }
};

template <bool REC, int ALPHA>
__global__ void __launch_bounds__(PT_TRACE_BLOCK, PT_TRACE_MIN_BLOCKS) k_trace(TraceScene sc, Wavefront wf, int cur, TraceTuning tune) {
    __shared__ uint2 stack[PT_STACK_SMEM * PT_TRACE_BLOCK];
    WavefrontRayIO<REC> io;
    io.nPath = wf.counters[CNT_PATH0 + cur]; io.nProbe = wf.counters[CNT_PROBE];
    const uint32_t total = io.nPath + io.nProbe + wf.counters[CNT_SHADOW];
    if (total == 0) return;
    io.pathO = wf.pathRayO[cur]; io.pathD = wf.pathRayD[cur]; io.pathHit = wf.pathHit;
    io.probeO = wf.probeRayO; io.probeD = wf.probeRayD; io.probeHit = wf.probeHit;
    io.shO = wf.shRayO; io.shD = wf.shRayD; io.shC = wf.shC; io.shG = wf.shG; io.pixelSum = wf.pixelSum;
    if (REC) io.rec = wf.rec;
    // small queues (the tail of a frame): shrink the chunk so the rays spread over all SMs
    const uint32_t warps = gridDim.x * (PT_TRACE_BLOCK / 32);
    uint32_t chunk = tune.chunk;
    while (chunk > 32 && total / chunk < warps) chunk >>= 1;
#if PT_COOP
    __shared__ CoopSmem coop;
    tracePersistentCoop<ALPHA>(sc, io, total, &wf.counters[CNT_WORK_TRACE], chunk, tune.refillMin, stack + threadIdx.x, coop);
#else
    tracePersistent<ALPHA>(sc, io, total, &wf.counters[CNT_WORK_TRACE], chunk, tune.refillMin, stack + threadIdx.x);
#endif
}

// generic ray batch for the traversal-only parity hook (b200pt_trace_rays): same persistent loop
struct BatchRayIO {
    const float4 *rays; float4 *hits; bool any;
    __device__ __forceinline__ void load(uint32_t idx, vec3 &o, vec3 &d, float &tmin, float &tmax, bool &a) const {
        const float4 ro = rays[2 * size_t(idx)], rd = rays[2 * size_t(idx) + 1];
        o = make_vec3(ro); d = make_vec3(rd); tmin = ro.w; tmax = rd.w; a = any;
    }
    __device__ __forceinline__ void store(uint32_t idx, const HitRec &h, bool) const { hits[idx] = make_float4(h.t, __uint_as_float(h.prim), h.u, h.v); }
};
template <int ALPHA>
__global__ void __launch_bounds__(PT_TRACE_BLOCK, PT_TRACE_MIN_BLOCKS) k_trace_batch(TraceScene sc, const float4 *__restrict__ rays, float4 *__restrict__ hits, uint32_t n,
                                                                int any, uint32_t *workCounter, TraceTuning tune) {
    __shared__ uint2 stack[PT_STACK_SMEM * PT_TRACE_BLOCK];
    BatchRayIO io; io.rays = rays; io.hits = hits; io.any = any != 0;
    const uint32_t warps = gridDim.x * (PT_TRACE_BLOCK / 32);
    uint32_t chunk = tune.chunk;
    while (chunk > 32 && n / chunk < warps) chunk >>= 1;
#if PT_COOP
    __shared__ CoopSmem coop;
    tracePersistentCoop<ALPHA>(sc, io, n, workCounter, chunk, tune.refillMin, stack + threadIdx.x, coop);
#else
    tracePersistent<ALPHA>(sc, io, n, workCounter, chunk, tune.refillMin, stack + threadIdx.x);
#endif
}

// between trace and shade of one iteration: fold the queue sizes into the 64-bit statistics, publish the shade count
// and reset the counters the shade kernel is about to fill
// hostSlot (optional): 8 words of mapped pinned host memory — the queue sizes this iteration started with go straight to
// the host (no copy-engine operation between the kernels of the loop), followed by a sequence number the host polls
__device__ __forceinline__ void iterPrep(const Wavefront &wf, int cur, volatile uint32_t *hostSlot, uint32_t seq) {
    uint32_t *c = wf.counters;
    const uint32_t nPath = c[CNT_PATH0 + cur], nProbe = c[CNT_PROBE], nShadow = c[CNT_SHADOW];
    if (hostSlot) {
        hostSlot[0] = nPath; hostSlot[1] = nProbe; hostSlot[2] = nShadow;
        __threadfence_system();
        hostSlot[7] = seq;
    }
    // (atomic: the cache update may be adding its rays to the same words from a second stream)
    atomicAdd(&wf.dstats[DST_EXTEND], (unsigned long long)nPath + nProbe);
    atomicAdd(&wf.dstats[DST_SHADOW], (unsigned long long)nShadow + c[CNT_INLINE_SHADOW]);
    atomicAdd(&wf.dstats[DST_VERTICES], (unsigned long long)nPath);
    if (nPath + nProbe + nShadow) atomicAdd(&wf.dstats[DST_ITERATIONS], 1ull);
    c[CNT_SHADE_N] = nPath;
    c[CNT_PATH0 + (1 - cur)] = 0; c[CNT_PROBE] = 0; c[CNT_SHADOW] = 0; c[CNT_INLINE_SHADOW] = 0; c[CNT_WORK_TRACE] = 0;
    c[CNT_REGEN] = 0;
}
__global__ void k_iter_prep(Wavefront wf, int cur, volatile uint32_t *hostSlot, uint32_t seq) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    iterPrep(wf, cur, hostSlot, seq);
}

// ---------------------------------------------------------------------------------------------------------------
// probe resolve: the BSDF-sampled direct-light check of MIS (rgen:684-727), run on the probe hits
__device__ __forceinline__ void probeResolveOne(const FrameParams &fp, const DeviceScene &sc, const Wavefront &wf, const uint32_t k) {
    const float4 hr = wf.probeHit[k];
    HitRec h; h.t = hr.x; h.prim = __float_as_uint(hr.y); h.u = hr.z; h.v = hr.w;
    // most probes end on an ordinary surface: look at the material before anything else of the record is read
    if (h.prim != PT_MISS && sc.materials[hitMaterialIndex(sc, h)].type != B200PT_MAT_LIGHT) return;
    const float4 A = wf.probeA[k], B = wf.probeB[k];
    const vec3 bsdf = make_vec3(A), T = make_vec3(B);
    const float pdfMat = A.w;
    const int pix = __float_as_int(B.w);
    const vec3 o = make_vec3(wf.probeRayO[k]), d = make_vec3(wf.probeRayD[k]);
    vec3 lightColor; float pdfLights;
    if (h.prim == PT_MISS) {
        lightColor = envColor(sc, d);
        pdfLights = 1.0f / (2.0f * PT_PI) / float(sc.numLights);
    } else {
        HitInfo info;
        computeHitInfo(sc, h, o, d, info);
        const b200pt_material *m = &sc.materials[info.matIndex];
        int iLight = info.isSphere ? sc.spheres[info.instanceIndex].iLight : sc.instances[info.instanceIndex].iLight;
        if (iLight < 0) return;
        lightColor = V3(m->lightColor[0], m->lightColor[1], m->lightColor[2]);
        pdfLights = pdfLight(sc.lights[iLight], d, info.normal, info.t);
    }
    const float heuristic = fp.pc.usePowerHeuristic ? powerHeuristic(pdfMat, pdfLights) : balanceHeuristic(pdfMat, pdfLights);
    if (h.prim != PT_MISS && isnan(heuristic)) return;
    vec3 nee = bsdf * lightColor * heuristic / pdfMat;
    if (fp.pc.numNEE != 1) nee = nee / float(fp.pc.numNEE);       // x / 1.0f == x: one out-of-line IEEE vec3 division less in the common case
    const vec3 c = T * nee;
    float *dst = reinterpret_cast<float *>(&wf.pixelSum[pix]);
    atomicAdd(dst + 0, c.x); atomicAdd(dst + 1, c.y); atomicAdd(dst + 2, c.z);
    if (wf.rec.samples) {
        float *ps = reinterpret_cast<float *>(&wf.rec.pathSum[pix]);
        atomicAdd(ps + 0, c.x); atomicAdd(ps + 1, c.y); atomicAdd(ps + 2, c.z);
        updateSamples<true>(wf.rec, pix, wf.rec.state[pix].x, __float_as_int(wf.probeRayO[k].w), nee);
    }
}
__global__ void __launch_bounds__(256) k_probe_resolve(FrameParams fp, DeviceScene sc, Wavefront wf) {
    const uint32_t n = wf.counters[CNT_PROBE];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) probeResolveOne(fp, sc, wf, k);
}
// the same with k_iter_prep folded in: the block that finishes last (ticket counter) does the iteration bookkeeping, one
// launch and one launch gap less per wavefront iteration.  Every block has read CNT_PROBE before it takes its ticket.
__global__ void __launch_bounds__(256) k_probe_resolve_prep(FrameParams fp, DeviceScene sc, Wavefront wf, int cur, volatile uint32_t *hostSlot, uint32_t seq) {
    const uint32_t n = wf.counters[CNT_PROBE];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) probeResolveOne(fp, sc, wf, k);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&wf.counters[CNT_TICKET], 1u) == gridDim.x - 1u) {
            wf.counters[CNT_TICKET] = 0u;
            iterPrep(wf, cur, hostSlot, seq);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// irradiance-cache lookups of one wavefront iteration (IC / ADRRS frames), run between trace and shade.  The lookup is a
// chain of dependent loads over a cell list of some tens of entries; inside the 128-register shade kernel (4 CTAs per
// SM) it was latency bound — issue slots 12 % busy, 60 % of the kernel's instructions (profiles/r01c_ncu_shade_ic.txt).
// As a kernel of its own it runs at full occupancy and the shade kernel just reads the result.
__device__ __forceinline__ void icQueryOne(const FrameParams &fp, const DeviceScene &sc, const Wavefront &wf, const int cur, const uint32_t qi) {
    const b200pt_push_constants &pc = fp.pc;
    const float4 hr = wf.pathHit[qi];
    HitRec h; h.t = hr.x; h.prim = __float_as_uint(hr.y); h.u = hr.z; h.v = hr.w;
    if (h.prim == PT_MISS) return;
    const float4 ro = wf.pathRayO[cur][qi], rd = wf.pathRayD[cur][qi];
    const int pid = __float_as_int(ro.w);
    const uint32_t depth = (wf.state[pid] & 0xffffu) + 1u;
    if (!(pc.useIrradianceCache || (pc.useADRRS && depth > 1))) return;
    HitInfo info;
    computeHitInfo(sc, h, make_vec3(ro), make_vec3(rd), info);
    const int type = sc.materials[info.matIndex].type;
    if (hasDiscreteDirection(type) || !isICCapable(pc, type)) return;
    vec3 irr = V3(0.0f);
    const bool found = queryIrradianceCache(wf.ic.view, pc, info.worldPos, info.normal, irr);
    wf.ic.queryResult[qi] = make_f4(irr, found ? 1.0f : 0.0f);
}
__global__ void __launch_bounds__(256) k_ic_query(FrameParams fp, DeviceScene sc, Wavefront wf, int cur) {
    const uint32_t n = wf.counters[CNT_SHADE_N];
    for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < n; qi += gridDim.x * blockDim.x) icQueryOne(fp, sc, wf, cur, qi);
}

// Cell-sorted lookups.  In queue order the lanes of a warp look into different grid cells: lists of 0-150 entries inside one
// warp, 10.5 of 32 lanes busy per instruction (profiles/r01c_ncu_ic_query.txt).  A counting sort of the queue entries that
// need a lookup by (approximate) grid cell puts lanes with the same list next to each other: same trip count, the list
// loads are broadcasts.  The order only says WHICH thread serves which queue entry; the result still goes to
// queryResult[queue index], every lookup walks its list in index order, so nothing changes per pixel.
struct ICQuerySort {
    uint32_t *key;        // per queue entry: cell, or ICQ_SKIP
    uint32_t *hist;       // per bin: entries (zero outside k_icq_count .. k_icq_scan)
    uint32_t *cursor;     // per bin: exclusive scan, advanced by k_icq_scatter
    uint32_t *order;      // queue indices, grouped by bin
    int numCells;         // bins: one per grid cell (entries that need a lookup), then ICQ_EXTRA_BINS classes of entries that do not
    int all;              // 1: the entries without a lookup are sorted behind the others (k_shade walks the whole order), 0: left out
};
#define ICQ_SKIP 0xffffffffu
#define ICQ_EXTRA_BINS 2      // numCells + 0: surface hit without a lookup, numCells + 1: miss
__global__ void __launch_bounds__(256) k_icq_count(FrameParams fp, Wavefront wf, ICQuerySort qs, int cur) {
    const uint32_t n = wf.counters[CNT_SHADE_N];
    const b200pt_push_constants &pc = fp.pc;
    const ICView &ic = wf.ic.view;
    const unsigned lane = threadIdx.x & 31u;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t qi = base + lane;
        uint32_t key = ICQ_SKIP;
        if (qi < n) {
            const float4 hr = wf.pathHit[qi];
            if (qs.all) key = uint32_t(qs.numCells) + (__float_as_uint(hr.y) != PT_MISS ? 0u : 1u);
            if (__float_as_uint(hr.y) != PT_MISS) {
                const float4 ro = wf.pathRayO[cur][qi], rd = wf.pathRayD[cur][qi];
                const uint32_t depth = (wf.state[__float_as_int(ro.w)] & 0xffffu) + 1u;
                if (pc.useIrradianceCache || (pc.useADRRS && depth > 1)) {
                    // the hit point up to rounding (the lookup itself takes the interpolated vertex position): good enough for a sort key
                    const int cx = icCellCoord(ic, 0, ro.x + hr.x * rd.x), cy = icCellCoord(ic, 1, ro.y + hr.x * rd.y), cz = icCellCoord(ic, 2, ro.z + hr.x * rd.z);
                    key = uint32_t((cz * ic.dim[1] + cy) * ic.dim[0] + cx);
                }
            }
            qs.key[qi] = key;
        }
        const unsigned m = __match_any_sync(0xffffffffu, key);
        if (key != ICQ_SKIP && int(lane) == __ffs(m) - 1) atomicAdd(&qs.hist[key], uint32_t(__popc(m)));
    }
}
// exclusive scan of the cell histogram by one block -> cursor, histogram back to zero, total -> CNT_ICQ_N
__global__ void __launch_bounds__(1024) k_icq_scan(ICQuerySort qs, uint32_t *counters) {
    __shared__ uint32_t partial[1024];
    const int n = qs.numCells + ICQ_EXTRA_BINS;
    const int per = (n + 1023) / 1024;
    const int b0 = threadIdx.x * per, b1 = min(n, b0 + per);
    uint32_t s = 0;
    for (int i = b0; i < b1; i++) s += qs.hist[i];
    partial[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        uint32_t v = threadIdx.x >= off ? partial[threadIdx.x - off] : 0u;
        __syncthreads();
        partial[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = threadIdx.x ? partial[threadIdx.x - 1] : 0u;
    for (int i = b0; i < b1; i++) {
        const uint32_t c = qs.hist[i]; qs.cursor[i] = run; qs.hist[i] = 0u;
        if (i == qs.numCells) counters[CNT_ICQ_N] = run;        // everything in front of the first extra bin needs a lookup
        run += c;
    }
}
__global__ void __launch_bounds__(256) k_icq_scatter(Wavefront wf, ICQuerySort qs) {
    const uint32_t n = wf.counters[CNT_SHADE_N];
    const unsigned lane = threadIdx.x & 31u;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t qi = base + lane;
        const uint32_t key = qi < n ? qs.key[qi] : ICQ_SKIP;
        const unsigned m = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(m) - 1;
        uint32_t pos = 0;
        if (key != ICQ_SKIP && int(lane) == leader) pos = atomicAdd(&qs.cursor[key], uint32_t(__popc(m)));
        pos = __shfl_sync(0xffffffffu, pos, leader);
        if (key != ICQ_SKIP) qs.order[pos + uint32_t(__popc(m & ((1u << lane) - 1u)))] = qi;
    }
}
__global__ void __launch_bounds__(256) k_ic_query_sorted(FrameParams fp, DeviceScene sc, Wavefront wf, ICQuerySort qs, int cur) {
    const uint32_t n = wf.counters[CNT_ICQ_N];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) icQueryOne(fp, sc, wf, cur, qs.order[i]);
}

// ---------------------------------------------------------------------------------------------------------------
// multipleNEE / nextEventEstimation (rgen:871-877, :601-731) in queue form: the shadow ray and the MIS probe ray of
// every light sample go to their queues with the contribution they carry; k_trace / k_probe_resolve finish them.
// T is the throughput the NEE result is multiplied with; cso = currentSampleOffset (guiding training only).
__device__ __forceinline__ void neeQueued(const FrameParams &fp, const DeviceScene &sc, const Wavefront &wf, uint32_t &seed, const b200pt_material &mat,
                                          const HitInfo &info, vec3 origin, vec3 normal, vec3 wi, vec3 T, int pid, bool saveSamples, int cso, uint2 *stack) {
    const b200pt_push_constants &pc = fp.pc;
    for (int iNee = 0; iNee < pc.numNEE; iNee++) {
        vec3 lightDir, lightColor; float lightDistance;
        const float pdfLights = sampleLights(sc, seed, pc.useVisibleSphereSampling != 0, origin, normal, lightDir, lightColor, lightDistance);
        const float cosThetaLight = dot(normal, lightDir);
        const bool traceShadow = cosThetaLight > 0 && pdfLights > 0;
        bool pushShadow = false, skipProbe = false;
        vec3 C = V3(0.0f);
        if (traceShadow) {
            const vec3 f = evalBsdf(sc, mat, info.u, info.v, normal, wi, lightDir, true);
            if (pc.enableMIS) {
                const float pdfMatL = pdfBSDF(mat, normal, wi, lightDir);
                const float heuristic = pc.usePowerHeuristic ? powerHeuristic(pdfLights, pdfMatL) : balanceHeuristic(pdfLights, pdfMatL);
                if (isnan(heuristic)) {
                    // rgen:653-655 returns before the BSDF probe, but only when the light is visible:
                    // resolve the visibility here so the RNG stream stays identical (rare)
                    HitRec sh;
                    atomicAdd(&wf.counters[CNT_INLINE_SHADOW], 1u);
                    traceRayInline<true>(sc.trace, origin, lightDir, PT_TMIN, lightDistance * (1 - 0.0001f), sh, stack, blockDim.x);
                    if (sh.prim == PT_MISS) skipProbe = true;
                } else {
                    C = f * lightColor * heuristic / pdfLights;
                    pushShadow = true;
                }
            } else {
                C = f * lightColor / pdfLights;
                pushShadow = true;
            }
        }
        if (pushShadow) {
            const uint32_t slotS = queuePush(&wf.counters[CNT_SHADOW]);
            wf.shRayO[slotS] = make_f4(origin, __int_as_float(pid));
            wf.shRayD[slotS] = make_f4(lightDir, lightDistance * (1 - 0.0001f));
            const vec3 Cn = pc.numNEE == 1 ? C : C / float(pc.numNEE);       // x / 1.0f == x
            wf.shC[slotS] = make_f4(T * Cn, 0.0f);
            if (saveSamples) wf.shG[slotS] = make_f4(Cn, __int_as_float(cso));
        }
        bool pushProbe = false;
        vec3 bsdfDir = V3(0.0f); float pdfMat = 0.0f;
        if (pc.enableMIS && !skipProbe) {         // rgen:665-728
            pdfMat = sampleBSDF(seed, mat, wi, normal, info.isFrontFace, bsdfDir);
            pushProbe = pdfMat > 0;
        }
        if (pushProbe) {
            const uint32_t slotP = queuePush(&wf.counters[CNT_PROBE]);
            const vec3 f = evalBsdf(sc, mat, info.u, info.v, normal, wi, bsdfDir, true);
            wf.probeRayO[slotP] = make_f4(origin, __int_as_float(cso));
            wf.probeRayD[slotP] = make_f4(bsdfDir, 0.0f);
            wf.probeA[slotP] = make_f4(f, pdfMat);
            wf.probeB[slotP] = make_f4(T, __int_as_float(pid));
        }
    }
}

// Out-of-line copies for the big shade variants (IC / guided).  Fully inlined, k_shade<0,1> is 40 640 SASS instructions
// (650 KB; the guided IC variant 55 296) — neeQueued alone is expanded three times, each with its own sampleLights,
// evalBsdf x2, pdfBSDF, sampleBSDF and the in-line visibility traversal — and ncu attributes 81 % of that kernel's warp
// stall cycles to "no instruction" (instruction-cache misses; issue slots 9.9 % busy, profiles/r01e_ncu_shade_ic.txt).
// One shared copy of the three heavy callees keeps the hot code inside the instruction cache.  The plain kernel (15 464
// instructions, 12 % no-instruction stalls) keeps everything in line: an ABI call costs it more than it saves.
__device__ __noinline__ void neeQueuedNI(const FrameParams &fp, const DeviceScene &sc, const Wavefront &wf, uint32_t &seed, const b200pt_material &mat,
                                         const HitInfo &info, vec3 origin, vec3 normal, vec3 wi, vec3 T, int pid, bool saveSamples, int cso, uint2 *stack) {
    neeQueued(fp, sc, wf, seed, mat, info, origin, normal, wi, T, pid, saveSamples, cso, stack);
}
template <bool GUIDE>
__device__ __noinline__ float getNewDirectionNI(const b200pt_push_constants &pc, const GuidingView &guide, uint32_t &seed, const b200pt_material &mat,
                                                vec3 origin, vec3 normal, vec3 wi, bool frontFace, vec3 &newDirection) {
    return getNewDirection<GUIDE>(pc, guide, seed, mat, origin, normal, wi, frontFace, newDirection);
}
__device__ __noinline__ vec3 evalBsdfNI(const DeviceScene &sc, const b200pt_material &mat, float u, float v, vec3 normal, vec3 wi, vec3 wo, bool frontFace) {
    return evalBsdf(sc, mat, u, v, normal, wi, wo, frontFace);
}
#ifndef PT_SHADE_NI
#define PT_SHADE_NI 1         // 0: everything in line in every variant (the state before round 1e)
#endif

// saveResult (rgen:1459-1485): fold one frame's per-pixel sum into the output / accumulation images
__device__ __forceinline__ void accumulatePixel(const b200pt_push_constants &pc, const int samplesPerPixel, const uint32_t prev, const float4 s,
                                                float4 *image, float4 *accum, float4 *estimate, const int p) {
    vec3 result = V3(s.x, s.y, s.z) / float(samplesPerPixel);
    if (pc.storeEstimate) estimate[p] = make_f4(result, 1.0f);
    if (pc.visualizeMode != 0) return;   // debug views are out of scope
    if (pc.enableAverageInsteadOfMix) {
        if (prev > 0) {
            vec3 a = make_vec3(accum[p]) + result;
            image[p] = make_f4(a / float(prev + 1u), 1.0f);
            accum[p] = make_f4(a, 1.0f);
        } else {
            accum[p] = make_f4(result, 1.0f);
            image[p] = make_f4(result, 1.0f);
        }
    } else {
        if (prev > 0) result = mix(make_vec3(image[p]), result, 1.0f / float(prev + 1u));
        image[p] = make_f4(result, 1.0f);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// shade: one bounce of rgen raytrace() (:1025-1215) for every path in the current queue
// GUIDE compiles in guided sampling (pc.useGuiding) and sample recording (pc.updateGuiding);
// IC compiles in the irradiance-cache lookup, ADRRS (weight window, Russian roulette, splitting) and the split drain
#ifndef PT_SHADE_MIN_BLOCKS
#define PT_SHADE_MIN_BLOCKS 5        // 96 registers, ~200 B of spills: 5 CTAs per SM beat 4 without spills by 5 % (tools/tune_variants.sh)
#endif
// BATCH compiles in the frame walk of b200pt_render_frames (plain frames only)
// DEFER (IC variants): a pixel whose path ended only goes to the regen queue; k_regen starts its next path (see there)
template <bool GUIDE, bool IC, bool BATCH = false, bool DEFER = false>
__global__ void __launch_bounds__(128, PT_SHADE_MIN_BLOCKS) k_shade(const __grid_constant__ FrameParams fp, const __grid_constant__ DeviceScene sc,
                                                                    const __grid_constant__ Wavefront wf, int cur) {
    __shared__ uint2 stack[PT_STACK_SMEM * 128];
#ifndef PT_SHADE_NI_GUIDE
#define PT_SHADE_NI_GUIDE 0       // the same for the guided-only variant: measured slower (config 5 training render 4 166 -> 4 895 ms, guided frames 578 -> 641)
#endif
    constexpr bool NI = PT_SHADE_NI && (IC || (PT_SHADE_NI_GUIDE && GUIDE));      // heavy callees out of line (see neeQueuedNI); guided-only: one expansion each, nothing to share
    #define NEE_QUEUED(...) do { if (NI) neeQueuedNI(__VA_ARGS__); else neeQueued(__VA_ARGS__); } while (0)
    #define GET_NEW_DIRECTION(...) (NI ? getNewDirectionNI<GUIDE>(__VA_ARGS__) : getNewDirection<GUIDE>(__VA_ARGS__))
    #define EVAL_BSDF(...) (NI ? evalBsdfNI(__VA_ARGS__) : evalBsdf(__VA_ARGS__))
    const uint32_t n = wf.counters[CNT_SHADE_N];
  for (uint32_t qiter = blockIdx.x * blockDim.x + threadIdx.x; qiter < n; qiter += gridDim.x * blockDim.x) {
    const uint32_t qi = (IC && wf.shadeOrder) ? wf.shadeOrder[qiter] : qiter;
    bool pushPath = false;
    vec3 outO = V3(0.0f), outD = V3(0.0f);
    int pid = 0;
    {
        const float4 ro = wf.pathRayO[cur][qi], rd = wf.pathRayD[cur][qi];
        const float4 hr = wf.pathHit[qi];
        pid = __float_as_int(ro.w);
        vec3 origin = make_vec3(ro), direction = make_vec3(rd);
        uint32_t seed = wf.seed[pid];
        vec3 T = make_vec3(wf.thr[pid]);
        uint32_t st = wf.state[pid];
        uint32_t depth = st & 0xffffu, followCount = (st >> 16) & 0xffu;
        bool addNext = (st & ST_ADDNEXT) != 0, follow = (st & ST_FOLLOW) != 0;
        bool oddFrame = BATCH && (st & ST_ODDFRAME) != 0;
        // where this path's radiance is summed: the pixel, or the pixel in the second half of pixelSum for odd frames of a batch
        const int sumIdx = oddFrame ? pid + fp.numPixels : pid;
        vec3 add = V3(0.0f);

        const b200pt_push_constants &pc = fp.pc;
        const bool useNEE = pc.enableNEE != 0;
        const bool addDirectLights = !useNEE;
        bool terminated = false;
        const bool saveSamples = GUIDE && pc.updateGuiding != 0;
        const int sbase = pid * B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL;
        int4 gst = make_int4(0, 0, -1, -1);      // sampleOffset, currentSampleOffset, iUpdateDistance, pending commit
        float distanceFactor = 1.0f;
        if (saveSamples) {
            gst = wf.rec.state[pid];
            distanceFactor = wf.rec.distanceFactor[pid];
            if (depth == 0) {        // first vertex of a new path: finish the previous path of this pixel (its last NEE
                                     // results have been resolved by now), then start the raytrace() locals afresh
                if (gst.w >= 0) commitSamples(wf.rec, pid, gst);
                gst.y = gst.x; gst.z = -1; distanceFactor = 1.0f;
                wf.rec.pathSum[pid] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        }
        depth++;

        HitRec h; h.t = hr.x; h.prim = __float_as_uint(hr.y); h.u = hr.z; h.v = hr.w;
        if (h.prim == PT_MISS) {                       // rgen:1025-1040
            if (addDirectLights || addNext) {
                const vec3 miss = envColor(sc, direction);
                add += T * miss;
                if (saveSamples) {                                    // rgen:1031-1037
                    updateSamples<false>(wf.rec, pid, gst.x, gst.y, miss);
                    if (gst.z != -1) wf.rec.samples[sbase + gst.z].distance = 0.0f;
                }
            }
            terminated = true;
        } else {
            HitInfo info;
            computeHitInfo(sc, h, origin, direction, info);
            const b200pt_material mat = sc.materials[info.matIndex];
            origin = info.worldPos;
            const vec3 normal = info.normal;
            const vec3 wi = -direction;

            if (mat.type == B200PT_MAT_LIGHT && (addDirectLights || addNext)) {     // rgen:1054-1060
                const vec3 Le = V3(mat.lightColor[0], mat.lightColor[1], mat.lightColor[2]);
                add += T * Le;
                if (saveSamples) updateSamples<false>(wf.rec, pid, gst.x, gst.y, Le);
            }

            if (hasDiscreteDirection(mat.type)) {      // rgen:1062-1075
                addNext = true;
                follow = true;
                if (int(depth) >= pc.maxDepth) followCount++;
                if (saveSamples && gst.z != -1) wf.rec.samples[sbase + gst.z].distance += info.t * distanceFactor;   // rgen:1073-1075
            } else {
                addNext = false;
                follow = false;
                if (IC && (pc.useIrradianceCache || (pc.useADRRS && depth > 1)) && isICCapable(pc, mat.type)) {      // rgen:1083-1133
                    const float4 icq = wf.ic.queryResult[qi];        // queryIrradianceCache, done by k_ic_query for this queue entry
                    const vec3 irradianceColor = make_vec3(icq);
                    if (icq.w != 0.0f) {
                        if (pc.useADRRS) {
                            int nSplit;
                            float q = applyWeightWindow(pc, seed, T, irradianceColor, make_vec3(wf.ic.estimate[pid]), nSplit);
                            if (nSplit == 1) {
                                if (rnd(seed) > q) terminated = true;                  // Russian roulette
                                else T = T / q;
                            } else if (pc.adrrsSplit) {
                                const int possibleSplits = IC_MAX_SPLITS - int(wf.ic.splitState[pid] & 0xffffu);
                                if (nSplit - 1 > possibleSplits) { nSplit = possibleSplits + 1; q = float(nSplit); }
                                T = T / q;
                                for (int iSplit = 0; iSplit < nSplit - 1; iSplit++)
                                    splitPush(wf.ic, pid, origin, normal, wi, T, info.u, info.v, info.matIndex, int(depth), info.isFrontFace);
                            }
                        } else {
                            add += T * approxDiffuse(sc, mat, normal, wi, info.u, info.v) * irradianceColor;
                            // rgen:1119-1123 adds this NEE to `result` only — no updateSamples() on the way out: on a training frame the light
                            // samples carry cso = sampleOffset, which makes the deferred update an empty loop while the path sum still counts them
                            if (neeSupported(mat.type)) NEE_QUEUED(fp, sc, wf, seed, mat, info, origin, normal, wi, T, pid, saveSamples, gst.x, stack + threadIdx.x);
                            terminated = true;
                        }
                    } else if (rnd(seed) < pc.irradianceCreateProb) {                  // createIC is true on every path of main()
                        const uint32_t k = wf.ic.newCount[pid];
                        if (k < IC_MAX_NEW) {
                            wf.ic.newEntries[(size_t(pid) * IC_MAX_NEW + k) * 2 + 0] = make_f4(origin, 0.0f);
                            wf.ic.newEntries[(size_t(pid) * IC_MAX_NEW + k) * 2 + 1] = make_f4(normal, 0.0f);
                            wf.ic.newCount[pid] = k + 1u;
                        }
                    }
                }
                if (!terminated) {
                    if (IC && pc.splitOnFirst && depth == 1) {                         // rgen:1135-1139
                        if (splitPush(wf.ic, pid, origin, normal, wi, T * 0.5f, info.u, info.v, info.matIndex, int(depth), info.isFrontFace)) T *= 0.5f;
                    }
                    if (saveSamples && gst.z != -1) {                     // rgen:1149-1156
                        wf.rec.samples[sbase + gst.z].distance += info.t * distanceFactor;
                        if (!isMatAlmostDiscrete(mat)) gst.z = -1;
                    }
                    if (useNEE && neeSupported(mat.type))                 // multipleNEE, rgen:871-877 / nextEventEstimation :601-731
                        NEE_QUEUED(fp, sc, wf, seed, mat, info, origin, normal, wi, T, BATCH ? sumIdx : pid, saveSamples, gst.y, stack + threadIdx.x);
                }
            }

            // getNewDirection (rgen:923-960) + throughput update (:1169-1177) + sample recording (:1179-1212)
            if (!terminated) {
                vec3 newDirection = V3(0.0f);
                const float pdf = GET_NEW_DIRECTION(pc, wf.guide, seed, mat, origin, normal, wi, info.isFrontFace, newDirection);
                if (pdf <= 0.0f) terminated = true;
                else {
                    const vec3 change = EVAL_BSDF(sc, mat, info.u, info.v, normal, wi, newDirection, info.isFrontFace) / pdf;
                    T *= change;
                    if (saveSamples && gst.y < B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL) {
                        b200pt_directional_data &sd = wf.rec.samples[sbase + gst.y];
                        if (hasDiscreteDirection(mat.type) || isMatAlmostDiscrete(mat)) sd.flags = B200PT_INVALID_REGION;
                        else {
                            sd.position[0] = origin.x; sd.position[1] = origin.y; sd.position[2] = origin.z;
                            sd.direction[0] = newDirection.x; sd.direction[1] = newDirection.y; sd.direction[2] = newDirection.z;
                            sd.weight = 0.0f; sd.pdf = pdf; sd.distance = 0.0f;
                            sd.flags = getGuidingRegion(wf.guide, origin);
                            gst.z = gst.y;
                            distanceFactor = 1.0f;
                        }
                        wf.rec.lightSums[sbase + gst.y] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        wf.rec.sampleThr[sbase + gst.y] = make_f4(change, 0.0f);
                        gst.y++;
                    } else if (gst.z != -1 && mat.type == B200PT_MAT_DIELECTRIC && dot(newDirection, normal) < 0.0f) {
                        float eta = mat.refractionIndex;
                        if (!info.isFrontFace) eta = mat.refractionIndexInv;
                        distanceFactor = fabsf(dot(normal, direction) / dot(normal, newDirection)) * eta;
                    }
                    direction = newDirection;
                }
            }
        }
        // do { ... } while (depth <= maxDepth || (follow && followCount <= maxFollowDiscrete))   rgen:1215
        if (!terminated && !(int(depth) <= pc.maxDepth || (follow && int(followCount) <= pc.maxFollowDiscrete))) terminated = true;

        if (add.x != 0.0f || add.y != 0.0f || add.z != 0.0f || isnan(add.x + add.y + add.z)) {
            float4 ps = wf.pixelSum[BATCH ? sumIdx : pid];
            ps.x += add.x; ps.y += add.y; ps.z += add.z;
            wf.pixelSum[BATCH ? sumIdx : pid] = ps;
            if (saveSamples) {
                float4 pp = wf.rec.pathSum[pid];
                pp.x += add.x; pp.y += add.y; pp.z += add.z;
                wf.rec.pathSum[pid] = pp;
            }
        }
        if (saveSamples) {
            if (terminated) gst.w = gst.y;       // commitSamples + the sampleOffset update wait for this path's last NEE results
            wf.rec.state[pid] = gst;
            wf.rec.distanceFactor[pid] = distanceFactor;
        }

        if (terminated) {
            if (wf.aov) {           // maxReachedDepth / depthSum / depthsCounter (rgen:1677-1679, :1705-1707), nextSplitSlot
                float4 a = wf.aov[pid];
                a.x = fmaxf(a.x, float(depth)); a.y += float(depth); a.z += 1.0f;
                if (IC && wf.ic.splitState) a.w = float(wf.ic.splitState[pid] & 0xffffu);
                wf.aov[pid] = a;
            }
            if (DEFER) {
                wf.regenQ[queuePush(&wf.counters[CNT_REGEN])] = uint32_t(pid);
            } else {
            // next sample of this pixel (rgen:1668-1681): same RNG stream, fresh path state
            uint32_t s = wf.sampleIdx[pid] + 1;
            if (BATCH && int(s & 0xffffu) >= fp.samplesPerPixel) {
                // frame f of the batch is finished for this pixel (its last light samples are still queued): fold frame
                // f - 1 into the image, hand its half of pixelSum to frame f + 1 and start that frame like k_generate
                const uint32_t f = s >> ST_FRAME_SHIFT;
                const int otherIdx = oddFrame ? pid : pid + fp.numPixels;
                if (f >= 1u) {
                    accumulatePixel(pc, fp.samplesPerPixel, wf.batch.framePrev[f - 1u], wf.pixelSum[otherIdx], wf.batch.image, wf.batch.accum, nullptr, pid);
                    wf.pixelSum[otherIdx] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
                if (int(f) + 1 < wf.batch.numFrames) {
                    s = (f + 1u) << ST_FRAME_SHIFT;
                    seed = tea(uint32_t(pid), wf.batch.frameSeed[f + 1u]);
                    oddFrame = !oddFrame;
                }
            }
            wf.sampleIdx[pid] = s;
            if (int(BATCH ? (s & 0xffffu) : s) < fp.samplesPerPixel) {
                cameraRay(fp, seed, pid % fp.width, pid / fp.width, outO, outD);
                T = V3(1.0f);
                depth = 0; followCount = 0; addNext = true; follow = false;
                pushPath = true;
            } else if (IC && wf.ic.splitState) {
                // all samples done: drain the pixel's split list, one split per finished path (rgen:1684-1708)
                const uint32_t ss = wf.ic.splitState[pid];
                const uint32_t nextSlot = ss & 0xffffu;
                uint32_t iSplit = ss >> 16;
                if (iSplit < nextSlot) {
                    const float4 *sp = wf.ic.splitData + (size_t(pid) * IC_MAX_SPLITS + iSplit) * IC_SPLIT_F4;
                    const float4 s0 = sp[0], s1 = sp[1], s2 = sp[2], s3 = sp[3], s4 = sp[4];
                    iSplit++;
                    HitInfo sinfo;
                    sinfo.worldPos = make_vec3(s0); sinfo.normal = make_vec3(s1); sinfo.u = s0.w; sinfo.v = s1.w;
                    sinfo.matIndex = __float_as_int(s2.w); sinfo.t = 0.0f; sinfo.isFrontFace = s4.x != 0.0f; sinfo.isSphere = false; sinfo.instanceIndex = 0u;
                    const vec3 swi = make_vec3(s2), sT = make_vec3(s3);
                    const b200pt_material smat = sc.materials[sinfo.matIndex];
                    if (useNEE && neeSupported(smat.type))      // the split happened before NEE: do it for the split's origin
                        NEE_QUEUED(fp, sc, wf, seed, smat, sinfo, sinfo.worldPos, sinfo.normal, swi, sT, pid, false, 0, stack + threadIdx.x);
                    vec3 newDirection = V3(0.0f);
                    const float pdf = GET_NEW_DIRECTION(pc, wf.guide, seed, smat, sinfo.worldPos, sinfo.normal, swi, sinfo.isFrontFace, newDirection);
                    if (pdf <= 0.0f) iSplit = nextSlot;      // `break`: the remaining splits are dropped
                    else {
                        T = sT * EVAL_BSDF(sc, smat, sinfo.u, sinfo.v, sinfo.normal, swi, newDirection, sinfo.isFrontFace) / pdf;
                        outO = sinfo.worldPos; outD = newDirection;
                        depth = uint32_t(__float_as_int(s3.w)); followCount = 0; addNext = false; follow = true;
                        pushPath = true;
                    }
                    wf.ic.splitState[pid] = (wf.ic.splitState[pid] & 0xffffu) | (iSplit << 16);
                }
            }
            }
        } else {
            outO = origin; outD = direction;
            pushPath = true;
        }
        wf.seed[pid] = seed;
        wf.thr[pid] = make_f4(T, 0.0f);
        wf.state[pid] = (depth & 0xffffu) | ((followCount & 0xffu) << 16) | (addNext ? ST_ADDNEXT : 0u) | (follow ? ST_FOLLOW : 0u) | (oddFrame ? ST_ODDFRAME : 0u);
    }
    if (pushPath) {
        const uint32_t slot = queuePush(&wf.counters[CNT_PATH0 + (1 - cur)]);
        wf.pathRayO[1 - cur][slot] = make_f4(outO, __int_as_float(pid));
        wf.pathRayD[1 - cur][slot] = make_f4(outD, 0.0f);
    }
  }
#undef NEE_QUEUED
#undef GET_NEW_DIRECTION
#undef EVAL_BSDF
}

// ---------------------------------------------------------------------------------------------------------------
// regen (IC / ADRRS frames): the tail of the megakernel's sample loop (rgen:1668-1708) for the pixels whose path ended in
// this iteration's shade pass — the next sample's camera ray, or, when all samples are done, the next record of the pixel's
// split list (its NEE, its new direction).  Inside k_shade<.,IC> this was a second copy of the heavy light-sampling /
// direction-sampling code that a warp ran AFTER its surviving lanes had been through the first one (13 000 SASS
// instructions, 61 % of the stall cycles "no instruction", profiles/r01e_shade_code_size.txt); as a kernel of its own each
// half fits the instruction cache and the lanes of a warp do the same thing.  Per pixel nothing changes: seed, throughput
// and state travel through the per-pixel arrays exactly as they do between two shade passes.
#ifndef PT_REGEN_MIN_BLOCKS
#define PT_REGEN_MIN_BLOCKS 5
#endif
template <bool GUIDE>
__global__ void __launch_bounds__(128, PT_REGEN_MIN_BLOCKS) k_regen(const __grid_constant__ FrameParams fp, const __grid_constant__ DeviceScene sc,
                                                                    const __grid_constant__ Wavefront wf, int cur) {
    __shared__ uint2 stack[PT_STACK_SMEM * 128];
    const uint32_t n = wf.counters[CNT_REGEN];
    const b200pt_push_constants &pc = fp.pc;
    const bool useNEE = pc.enableNEE != 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int pid = int(wf.regenQ[i]);
        uint32_t seed = wf.seed[pid];
        const uint32_t s = wf.sampleIdx[pid] + 1;
        wf.sampleIdx[pid] = s;
        bool pushPath = false;
        vec3 outO = V3(0.0f), outD = V3(0.0f), T = V3(1.0f);
        uint32_t depth = 0;
        bool addNext = true, follow = false;
        if (int(s) < fp.samplesPerPixel) {
            cameraRay(fp, seed, pid % fp.width, pid / fp.width, outO, outD);
            pushPath = true;
        } else if (wf.ic.splitState) {
            // all samples done: drain the pixel's split list, one split per finished path (rgen:1684-1708)
            const uint32_t ss = wf.ic.splitState[pid];
            const uint32_t nextSlot = ss & 0xffffu;
            uint32_t iSplit = ss >> 16;
            if (iSplit < nextSlot) {
                const float4 *sp = wf.ic.splitData + (size_t(pid) * IC_MAX_SPLITS + iSplit) * IC_SPLIT_F4;
                const float4 s0 = sp[0], s1 = sp[1], s2 = sp[2], s3 = sp[3], s4 = sp[4];
                iSplit++;
                HitInfo sinfo;
                sinfo.worldPos = make_vec3(s0); sinfo.normal = make_vec3(s1); sinfo.u = s0.w; sinfo.v = s1.w;
                sinfo.matIndex = __float_as_int(s2.w); sinfo.t = 0.0f; sinfo.isFrontFace = s4.x != 0.0f; sinfo.isSphere = false; sinfo.instanceIndex = 0u;
                const vec3 swi = make_vec3(s2), sT = make_vec3(s3);
                const b200pt_material smat = sc.materials[sinfo.matIndex];
                if (useNEE && neeSupported(smat.type))      // the split happened before NEE: do it for the split's origin
                    neeQueued(fp, sc, wf, seed, smat, sinfo, sinfo.worldPos, sinfo.normal, swi, sT, pid, false, 0, stack + threadIdx.x);
                vec3 newDirection = V3(0.0f);
                const float pdf = getNewDirection<GUIDE>(pc, wf.guide, seed, smat, sinfo.worldPos, sinfo.normal, swi, sinfo.isFrontFace, newDirection);
                if (pdf <= 0.0f) iSplit = nextSlot;      // `break`: the remaining splits are dropped
                else {
                    T = sT * evalBsdf(sc, smat, sinfo.u, sinfo.v, sinfo.normal, swi, newDirection, sinfo.isFrontFace) / pdf;
                    outO = sinfo.worldPos; outD = newDirection;
                    depth = uint32_t(__float_as_int(s3.w)); addNext = false; follow = true;
                    pushPath = true;
                }
                wf.ic.splitState[pid] = (ss & 0xffffu) | (iSplit << 16);
            }
        }
        wf.seed[pid] = seed;
        if (pushPath) {
            wf.thr[pid] = make_f4(T, 0.0f);
            wf.state[pid] = (depth & 0xffffu) | (addNext ? ST_ADDNEXT : 0u) | (follow ? ST_FOLLOW : 0u);
            const uint32_t slot = queuePush(&wf.counters[CNT_PATH0 + (1 - cur)]);
            wf.pathRayO[1 - cur][slot] = make_f4(outO, __int_as_float(pid));
            wf.pathRayD[1 - cur][slot] = make_f4(outD, 0.0f);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// accumulate: result /= spp, saveEstimate, saveResult (rgen:1710-1720, 1459-1485, 1602-1605)
__global__ void __launch_bounds__(256) k_accumulate(FrameParams fp, const float4 *__restrict__ pixelSum, float4 *image, float4 *accum, float4 *estimate,
                                                    GuidingRecord rec) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= fp.numPixels) return;
    if (fp.pc.updateGuiding && rec.samples) {     // the frame's last path of this pixel
        int4 gst = rec.state[p];
        if (gst.w >= 0) { commitSamples(rec, p, gst); rec.state[p] = gst; }
    }
    accumulatePixel(fp.pc, fp.samplesPerPixel, fp.pc.previousFrames, pixelSum[p], image, accum, estimate, p);
}

}  // namespace b200pt
