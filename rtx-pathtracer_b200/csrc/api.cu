// libb200pt.so — context management and the C ABI declared in include/b200pt.h.
// There is NO CPU fallback: every compute entry point needs a CUDA device and fails loudly without one.
#include "wavefront.cuh"
#include "ic_kernels.cuh"
#include "guiding_fit.cuh"
#include "../host/scene.h"
#include "../host/bvh.h"
#include <cstdio>
#include <unistd.h>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <stdexcept>
#include "comm.cuh"

using namespace b200pt;

static thread_local std::string g_lastError;
static int setError(int code, const std::string &msg) { g_lastError = msg; return code; }

#define CUDA_TRY(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            return setError(B200PT_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));            \
    } while (0)

namespace b200pt { NcclApi g_nccl; }

namespace {

template <typename T>
struct DevBuf {
    T *p = nullptr; size_t n = 0;
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        n = count;
        return cudaMalloc(reinterpret_cast<void **>(&p), std::max<size_t>(count, 1) * sizeof(T));
    }
    cudaError_t upload(const T *src, size_t count, cudaStream_t s) {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};
}  // namespace

struct b200pt_ctx {
    int device = 0, width = 0, height = 0, icSize = 0, guidingSplits = 0;
    int numPixels = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t evA = nullptr, evB = nullptr, evTimer0 = nullptr, evTimer1 = nullptr;
    bool hasScene = false, hasCamera = false;

    // scene
    DevBuf<float4> nodes, tris, sphereGeom, texels;
    DevBuf<b200pt_vertex> vertices;
    DevBuf<uint32_t> indices;
    DevBuf<int32_t> modelVertexOffset, modelIndexOffset, randomLightIndex;
    DevBuf<int4> primVerts;
    DevBuf<b200pt_material> materials;
    DevBuf<b200pt_instance> instances;
    DevBuf<b200pt_light> lights;
    DevBuf<b200pt_face_sample> randomTriIndex;
    DevBuf<b200pt_sphere> spheres;
    DevBuf<DeviceTexture> textures;
    DevBuf<AlphaTexture> alphaTextures;
    DevBuf<int> alphaMaterialTexture;
    DevBuf<AlphaScene> alphaScene;
    DeviceScene dscene{};
    float sceneMin[3], sceneMax[3];
    int bvhDepth = 0;

    // camera
    float view[16], proj[16], viewInv[16], projInv[16];

    // images
    DevBuf<float4> imgOutput, imgAccum, imgEstimate, imgAov;
    bool aovs = false;

    // wavefront
    Wavefront wf{};
    DevBuf<float4> pathRayO[2], pathRayD[2], pathHit, probeRayO, probeRayD, probeHit, probeA, probeB, shRayO, shRayD, shC, thr, pixelSum;
    DevBuf<uint32_t> seed, state, sampleIdx, counters;
    DevBuf<float4> shG, recLightSums, recSampleThr, recPathSum;
    DevBuf<int4> recState;
    DevBuf<float> recDistanceFactor;
    int queueNEE = 0;
    // pinned ring of queue-counter snapshots: the host looks at iteration i-LAG while iteration i is being issued
    enum { RING = 4, LAG = 2 };
    uint32_t *hostCounters = nullptr;   // pinned, RING x CNT_NUM
    // default: k_iter_prep stores the queue sizes into this mapped pinned ring and the host polls a sequence number —
    // no copy-engine operation inside the loop (B200PT_COUNTER_COPY=1 selects the memcpy + event variant)
    volatile uint32_t *hostRing = nullptr;
    uint32_t *hostRingDev = nullptr;
    uint32_t ringSeq = 0;
    bool counterCopy = false;
    bool fusePrep = true;         // k_iter_prep folded into k_probe_resolve (B200PT_FUSE_PREP=0: separate launch, for A/B)
    unsigned long long *hostDstats = nullptr;   // pinned, DST_NUM
    cudaEvent_t ringEvent[RING] = {nullptr, nullptr, nullptr, nullptr};
    DevBuf<unsigned long long> dstats;
    DevBuf<uint32_t> batchCounter;
    int icBuildBlocksPerSM = 2;
    bool icOverlap = true;        // B200PT_IC_OVERLAP=0: the cache update runs in front of the frame's paths on the same stream
    cudaStream_t stream2 = nullptr;
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    bool icBuildGroup = true;     // B200PT_IC_BUILD_GROUP=0: one lane per cache entry (the state before round 2b), for A/B
    // IC / ADRRS frames: lookups in grid-cell order (k_icq_*), next paths of finished pixels in a kernel of their own (k_regen)
    bool icqSort = true, regenSplit = true, shadeSorted = true;
    DevBuf<uint32_t> icqKey, icqHist, icqCursor, icqOrder, regenQ;
    int regenGrid = 0, regenGridGuided = 0, shadeGridICDefer = 0, shadeGridGuidedICDefer = 0;
    int numSMs = 0, traceGrid = 0, traceGridRec = 0, shadeGrid = 0, shadeGridGuided = 0, shadeGridIC = 0, shadeGridGuidedIC = 0, shadeGridBatch = 0, resolveGrid = 0, icQueryGrid = 0;
    TraceTuning tune{64u, 8};

    // guiding / IC state
    GuidingState guiding;
    DevBuf<b200pt_directional_data> samples, hostSamples;
    DevBuf<b200pt_cache_data> icData;
    DevBuf<b200pt_sphere> icSpheres;
    DevBuf<b200pt_cache_header> icHeader;
    // IC lookup snapshot + grid, per-pixel IC / split state, ordered-compaction scratch (allocated on first use)
    DevBuf<float4> icSnapSphere, icSnapNormalR, icSnapColor, icSnapRot, icSnapTrans, icPending, icNewEntries, icSplitData, icCellSpheres, icQueryResult;
    DevBuf<uint2> icRanges;
    DevBuf<uint32_t> icCellCount, icCellStart, icCellItems, icSnapHdr, icBlockCounts, icList, icNewCount, icSplitState, icValidFlags, icValidOffsets;
    DevBuf<int32_t> icUpdSlot;
    uint32_t *hostIcHdr = nullptr;      // pinned, ICH_NUM
    ICView icGrid{};
    int icNumCells = 0;

    // batch tracing scratch
    DevBuf<float4> batchRays, batchHits;

    // multi-GPU
    RankComm rc;
    DevBuf<float4> commImage;
    DevBuf<float> commCount;
    DevBuf<b200pt_directional_data> commSamples;

    b200pt_stats stats{};
    FILE *dumpIters = nullptr;          // B200PT_DUMP_ITERS=file: queue sizes after every wavefront iteration (tuning aid)

    // optional per-kernel timing: (kind, start event, stop event) triples resolved at the end of a frame
    int stageTiming = 0;             // 0 off, 1 every stage kernel, 2 the trace kernel only
    bool timed(int kind) const { return stageTiming == 1 || (stageTiming == 2 && kind == 0 /* KIND_EXTEND */); }
    DevBuf<uint32_t> batchSeeds, batchPrev;      // b200pt_render_frames: per-frame randomUInt / previousFrames
    std::vector<cudaEvent_t> eventPool;
    size_t eventsUsed = 0;
    struct Span { int kind; size_t a, b; };
    std::vector<Span> spans;
    cudaEvent_t nextEvent() {
        if (eventsUsed == eventPool.size()) { cudaEvent_t e; cudaEventCreate(&e); eventPool.push_back(e); }
        return eventPool[eventsUsed++];
    }
};

enum { KIND_EXTEND = 0, KIND_SHADOW = 1, KIND_SHADE = 2 };
struct StageTimer {   // records start/stop events around one launch when stage timing is on
    b200pt_ctx *c; int kind; size_t a = 0;
    StageTimer(b200pt_ctx *ctx, int k) : c(ctx), kind(k) {
        if (c->timed(kind)) { a = c->eventsUsed; cudaEventRecord(c->nextEvent(), c->stream); }
    }
    ~StageTimer() {
        if (c->timed(kind)) { size_t b = c->eventsUsed; cudaEventRecord(c->nextEvent(), c->stream); c->spans.push_back({kind, a, b}); }
        c->stats.kernel_launches++;
        if (kind == KIND_EXTEND) c->stats.launches_extend++; else if (kind == KIND_SHADOW) c->stats.launches_shadow++; else c->stats.launches_shade++;
    }
};
static void resolveStageTimes(b200pt_ctx *c) {
    for (const auto &sp : c->spans) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->eventPool[sp.a], c->eventPool[sp.b]) != cudaSuccess) continue;
        if (sp.kind == KIND_EXTEND) c->stats.ms_extend += ms; else if (sp.kind == KIND_SHADOW) c->stats.ms_shadow += ms; else c->stats.ms_shade += ms;
    }
    c->spans.clear();
    c->eventsUsed = 0;
}

static int ensureQueues(b200pt_ctx *c, int numNEE) {
    if (numNEE < 1) numNEE = 1;
    const size_t N = size_t(c->numPixels);
    for (int i = 0; i < 2; i++) { CUDA_TRY(c->pathRayO[i].alloc(N)); CUDA_TRY(c->pathRayD[i].alloc(N)); }
    CUDA_TRY(c->pathHit.alloc(N));
    CUDA_TRY(c->thr.alloc(N)); CUDA_TRY(c->pixelSum.alloc(N));
    CUDA_TRY(c->seed.alloc(N)); CUDA_TRY(c->state.alloc(N)); CUDA_TRY(c->sampleIdx.alloc(N));
    CUDA_TRY(c->counters.alloc(CNT_NUM)); CUDA_TRY(c->dstats.alloc(DST_NUM));
    if (numNEE > c->queueNEE) {
        const size_t M = N * size_t(numNEE);
        CUDA_TRY(c->probeRayO.alloc(M)); CUDA_TRY(c->probeRayD.alloc(M)); CUDA_TRY(c->probeHit.alloc(M));
        CUDA_TRY(c->probeA.alloc(M)); CUDA_TRY(c->probeB.alloc(M));
        CUDA_TRY(c->shRayO.alloc(M)); CUDA_TRY(c->shRayD.alloc(M)); CUDA_TRY(c->shC.alloc(M));
        c->queueNEE = numNEE;
    }
    Wavefront &w = c->wf;
    for (int i = 0; i < 2; i++) { w.pathRayO[i] = c->pathRayO[i].p; w.pathRayD[i] = c->pathRayD[i].p; }
    w.pathHit = c->pathHit.p;
    w.probeRayO = c->probeRayO.p; w.probeRayD = c->probeRayD.p; w.probeHit = c->probeHit.p; w.probeA = c->probeA.p; w.probeB = c->probeB.p;
    w.shRayO = c->shRayO.p; w.shRayD = c->shRayD.p; w.shC = c->shC.p;
    w.seed = c->seed.p; w.thr = c->thr.p; w.state = c->state.p; w.sampleIdx = c->sampleIdx.p; w.pixelSum = c->pixelSum.p;
    w.counters = c->counters.p; w.dstats = c->dstats.p;
    return B200PT_OK;
}

// IC / ADRRS buffers (allocated on the first frame that needs them) and the lookup-grid geometry
static int ensureIC(b200pt_ctx *c, bool needCache, bool splitMode) {
    const size_t N = size_t(c->numPixels);
    CUDA_TRY(c->icSnapHdr.alloc(ICH_NUM));
    CUDA_TRY(c->icBlockCounts.alloc(N / 256 + 2));
    CUDA_TRY(c->icList.alloc(N));
    if (c->regenSplit) CUDA_TRY(c->regenQ.alloc(N));
    if (needCache && c->icqSort) {
        const size_t cells = size_t(IC_GRID_MAX) * IC_GRID_MAX * IC_GRID_MAX + ICQ_EXTRA_BINS;
        CUDA_TRY(c->icqKey.alloc(N)); CUDA_TRY(c->icqOrder.alloc(N));
        if (!c->icqHist.p) {        // the histogram is zero between two sorts (k_icq_scan clears what k_icq_count added)
            CUDA_TRY(c->icqHist.alloc(cells));
            CUDA_TRY(cudaMemsetAsync(c->icqHist.p, 0, cells * sizeof(uint32_t), c->stream));
        }
        CUDA_TRY(c->icqCursor.alloc(cells));
    }
    if (needCache) {
        const size_t S = size_t(std::max(1, c->icSize));
        CUDA_TRY(c->icSnapSphere.alloc(S)); CUDA_TRY(c->icSnapNormalR.alloc(S)); CUDA_TRY(c->icSnapColor.alloc(S));
        CUDA_TRY(c->icSnapRot.alloc(S)); CUDA_TRY(c->icSnapTrans.alloc(S)); CUDA_TRY(c->icRanges.alloc(S));
        CUDA_TRY(c->icCellCount.alloc(size_t(IC_GRID_MAX) * IC_GRID_MAX * IC_GRID_MAX + 1));
        CUDA_TRY(c->icCellStart.alloc(size_t(IC_GRID_MAX) * IC_GRID_MAX * IC_GRID_MAX + 1));
        CUDA_TRY(c->icCellItems.alloc(S * 8)); CUDA_TRY(c->icCellSpheres.alloc(S * 8));
        CUDA_TRY(c->icNewCount.alloc(N));
        CUDA_TRY(c->icNewEntries.alloc(N * IC_MAX_NEW * 2));
        CUDA_TRY(c->icQueryResult.alloc(N));
        // uniform grid over the scene box (the same box the guiding regions start from), at most IC_GRID_MAX cells per axis
        float ext[3], mx = 0.0f;
        for (int a = 0; a < 3; a++) { ext[a] = c->sceneMax[a] - c->sceneMin[a]; mx = std::max(mx, ext[a]); }
        const float cell = mx > 0.0f ? mx / float(IC_GRID_MAX) : 1.0f;
        ICView &g = c->icGrid;
        c->icNumCells = 1;
        for (int a = 0; a < 3; a++) {
            g.gmin[a] = c->sceneMin[a];
            g.invCell[a] = 1.0f / cell;
            g.dim[a] = std::min(IC_GRID_MAX, std::max(1, int(ceilf(ext[a] / cell))));
            c->icNumCells *= g.dim[a];
        }
        g.sphere = c->icSnapSphere.p; g.normalR = c->icSnapNormalR.p; g.color = c->icSnapColor.p; g.rotGrad = c->icSnapRot.p; g.transGrad = c->icSnapTrans.p;
        g.cellStart = c->icCellStart.p; g.cellItems = c->icCellItems.p; g.cellSpheres = c->icCellSpheres.p;
    }
    if (splitMode) {
        CUDA_TRY(c->icSplitState.alloc(N));
        CUDA_TRY(c->icSplitData.alloc(N * IC_MAX_SPLITS * IC_SPLIT_F4));
    }
    return B200PT_OK;
}

static ICBuffers icBuffers(b200pt_ctx *c) {
    ICBuffers b{};
    b.header = c->icHeader.p; b.data = c->icData.p; b.spheres = c->icSpheres.p;
    b.snapSphere = c->icSnapSphere.p; b.snapNormalR = c->icSnapNormalR.p; b.snapColor = c->icSnapColor.p; b.snapRot = c->icSnapRot.p; b.snapTrans = c->icSnapTrans.p;
    b.ranges = c->icRanges.p; b.cellCount = c->icCellCount.p; b.cellStart = c->icCellStart.p; b.cellItems = c->icCellItems.p; b.cellSpheres = c->icCellSpheres.p;
    b.validFlags = c->icValidFlags.p; b.validOffsets = c->icValidOffsets.p;
    b.snapHdr = c->icSnapHdr.p; b.blockCounts = c->icBlockCounts.p; b.list = c->icList.p; b.updSlot = c->icUpdSlot.p; b.pending = c->icPending.p;
    b.icSize = c->icSize; b.numCells = c->icNumCells;
    return b;
}

static inline unsigned gridFor(uint64_t n, unsigned block) { return unsigned((n + block - 1) / block); }

// copies the IC snapshot header (entry count, grid size, list count) to the host; the few host decisions of an IC
// frame (buffer growth, launch shape of the cache-build kernels) are taken on it
static int readIcHeader(b200pt_ctx *c) {
    CUDA_TRY(cudaMemcpyAsync(c->hostIcHdr, c->icSnapHdr.p, ICH_NUM * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}

// ordered list of the pixels that satisfy `pred` -> c->icList, count in hostIcHdr[ICH_LIST_COUNT]
template <typename Pred>
static int compactPixels(b200pt_ctx *c, Pred pred) {
    const unsigned blocks = gridFor(uint64_t(c->numPixels), 256);
    k_flag_count<<<blocks, 256, 0, c->stream>>>(pred, c->numPixels, c->icBlockCounts.p);
    k_scan_single_block<<<1, 1024, 0, c->stream>>>(c->icBlockCounts.p, c->icBlockCounts.p, int(blocks), c->icSnapHdr.p + ICH_LIST_COUNT);
    c->stats.kernel_launches += 2;
    int rc = readIcHeader(c);
    if (rc != B200PT_OK) return rc;
    if (c->hostIcHdr[ICH_LIST_COUNT]) {
        k_flag_fill<<<blocks, 256, 0, c->stream>>>(pred, c->numPixels, c->icBlockCounts.p, c->icList.p);
        c->stats.kernel_launches++;
    }
    return B200PT_OK;
}

// launch shape of the cache-build kernels.  One entry = 200 sequential paths (the pixel's RNG stream is consumed in order), and the
// kernels hold 255 registers: 2 blocks of 4 warps per SM are resident.  Group mode (lanes = 8, ic_kernels.cuh): eight lanes share
// one entry and trace every ray together; an entry slot is 32, 16 or 8 threads wide — the sparsest packing that still runs in
// ONE wave, because a second wave costs a whole entry latency.  Lists too long even for four groups per warp (the first prepare
// frames create thousands of entries) fall back to one lane per entry, 8 / 16 / 32 entries per warp.
static void buildLaunchShape(const b200pt_ctx *c, uint32_t entries, int &grid, int &stride, int &lanes) {
    const uint32_t residentWarps = uint32_t(c->numSMs) * uint32_t(std::max(1, c->icBuildBlocksPerSM)) * 4u;
    uint32_t perWarp = 1;
    while (perWarp < 32u && (entries + perWarp - 1) / perWarp > residentWarps) perWarp *= 2;
    if (getenv("B200PT_IC_BUILD_LANES")) perWarp = uint32_t(std::max(1, std::min(32, atoi(getenv("B200PT_IC_BUILD_LANES")))));
    lanes = (c->icBuildGroup && perWarp <= 4u) ? 8 : 1;
    stride = int(32u / perWarp);
    grid = int(gridFor(uint64_t(entries) * uint64_t(stride), 128));
}

extern "C" {

const char *b200pt_last_error(void) { return g_lastError.c_str(); }

int b200pt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int b200pt_create(int device_ordinal, int width, int height, int ic_size, int guiding_splits, b200pt_ctx **out) {
    if (!out || width <= 0 || height <= 0 || ic_size < 0 || guiding_splits < 0)
        return setError(B200PT_E_INVALID, "b200pt_create: bad arguments");
    if (guiding_splits > B200PT_MAX_GUIDING_SPLITS)
        return setError(B200PT_E_INVALID, "b200pt_create: GUIDING_SPLITS > 11 is not supported (the device sort and plan handle at most 4096 regions, "
                                          "and adaptive region refinement needs head room above 2^splits)");
    int n = b200pt_device_count();
    if (n <= 0) return setError(B200PT_E_NODEVICE, "b200pt_create: no CUDA device visible (libb200pt has no CPU fallback)");
    if (device_ordinal < 0 || device_ordinal >= n) return setError(B200PT_E_INVALID, "b200pt_create: device ordinal out of range");
    CUDA_TRY(cudaSetDevice(device_ordinal));
    b200pt_ctx *c = new b200pt_ctx();
    c->device = device_ordinal; c->width = width; c->height = height; c->icSize = ic_size; c->guidingSplits = guiding_splits;
    c->numPixels = width * height;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&c->evA)); CUDA_TRY(cudaEventCreate(&c->evB));
    CUDA_TRY(cudaEventCreate(&c->evTimer0)); CUDA_TRY(cudaEventCreate(&c->evTimer1));
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void **>(&c->hostCounters), b200pt_ctx::RING * CNT_NUM * sizeof(uint32_t)));
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void **>(&c->hostDstats), DST_NUM * sizeof(unsigned long long)));
    for (auto &e : c->ringEvent) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    {
        void *hp = nullptr, *dp = nullptr;
        CUDA_TRY(cudaHostAlloc(&hp, b200pt_ctx::RING * 8 * sizeof(uint32_t), cudaHostAllocMapped));
        memset(hp, 0, b200pt_ctx::RING * 8 * sizeof(uint32_t));
        CUDA_TRY(cudaHostGetDevicePointer(&dp, hp, 0));
        c->hostRing = static_cast<volatile uint32_t *>(hp); c->hostRingDev = static_cast<uint32_t *>(dp);
        if (const char *e = getenv("B200PT_COUNTER_COPY")) c->counterCopy = atoi(e) != 0;
        if (const char *e = getenv("B200PT_FUSE_PREP")) c->fusePrep = atoi(e) != 0;
        if (const char *e = getenv("B200PT_ICQ_SORT")) c->icqSort = atoi(e) != 0;
        if (const char *e = getenv("B200PT_REGEN_SPLIT")) c->regenSplit = atoi(e) != 0;
        if (const char *e = getenv("B200PT_SHADE_SORTED")) c->shadeSorted = atoi(e) != 0;
        if (const char *e = getenv("B200PT_IC_BUILD_GROUP")) c->icBuildGroup = atoi(e) != 0;
        if (const char *e = getenv("B200PT_IC_OVERLAP")) c->icOverlap = atoi(e) != 0;
        int prLo = 0, prHi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prLo, &prHi));
        CUDA_TRY(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prHi));     // its few blocks should not queue behind a full wave
        CUDA_TRY(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
    }
    CUDA_TRY(c->batchCounter.alloc(1));
    // persistent launches: one full wave of resident CTAs (SM count x occupancy), sized once
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device_ordinal));
    c->numSMs = prop.multiProcessorCount;
    int occTrace = 0, occShade = 0, occResolve = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occTrace, k_trace<false, 0>, PT_TRACE_BLOCK, 0));
    int occTraceRec = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occTraceRec, k_trace<true, 0>, PT_TRACE_BLOCK, 0));
    c->traceGridRec = c->numSMs * std::max(1, occTraceRec);
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occShade, k_shade<false, false>, 128, 0));
    int occShadeGuided = 0, occShadeIC = 0, occShadeGuidedIC = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occShadeGuided, k_shade<true, false>, 128, 0));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occShadeIC, k_shade<false, true>, 128, 0));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occShadeGuidedIC, k_shade<true, true>, 128, 0));
    c->shadeGridGuided = c->numSMs * std::max(1, occShadeGuided);
    c->shadeGridIC = c->numSMs * std::max(1, occShadeIC);
    c->shadeGridGuidedIC = c->numSMs * std::max(1, occShadeGuidedIC);
    int occQuery = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occQuery, k_ic_query, 256, 0));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->icBuildBlocksPerSM, k_ic_create, 128, 0));
    {
        int o1 = 0, o2 = 0, o3 = 0, o4 = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, k_regen<false>, 128, 0));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, k_regen<true>, 128, 0));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o3, k_shade<false, true, false, true>, 128, 0));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o4, k_shade<true, true, false, true>, 128, 0));
        c->regenGrid = c->numSMs * std::max(1, o1); c->regenGridGuided = c->numSMs * std::max(1, o2);
        c->shadeGridICDefer = c->numSMs * std::max(1, o3); c->shadeGridGuidedICDefer = c->numSMs * std::max(1, o4);
    }
    c->icQueryGrid = c->numSMs * std::max(1, occQuery);
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void **>(&c->hostIcHdr), ICH_NUM * sizeof(uint32_t)));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occResolve, k_probe_resolve, 256, 0));
    int occShadeBatch = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occShadeBatch, k_shade<false, false, true>, 128, 0));
    c->shadeGridBatch = c->numSMs * std::max(1, occShadeBatch);
    c->traceGrid = c->numSMs * std::max(1, occTrace);
    c->shadeGrid = c->numSMs * std::max(1, occShade);
    c->resolveGrid = c->numSMs * std::max(1, occResolve);
    if (const char *e = getenv("B200PT_DUMP_ITERS")) c->dumpIters = fopen(e, "w");
    if (const char *e = getenv("B200PT_TRACE_CHUNK")) c->tune.chunk = uint32_t(std::max(32, atoi(e)));
    if (const char *e = getenv("B200PT_TRACE_REFILL")) c->tune.refillMin = std::max(1, std::min(32, atoi(e)));
    if (const char *e = getenv("B200PT_TRACE_GRID_PER_SM")) c->traceGrid = c->numSMs * std::max(1, atoi(e));
    const size_t N = size_t(c->numPixels);
    CUDA_TRY(c->imgOutput.alloc(N)); CUDA_TRY(c->imgAccum.alloc(N)); CUDA_TRY(c->imgEstimate.alloc(N));
    CUDA_TRY(cudaMemsetAsync(c->imgOutput.p, 0, N * sizeof(float4), c->stream));
    CUDA_TRY(cudaMemsetAsync(c->imgAccum.p, 0, N * sizeof(float4), c->stream));
    CUDA_TRY(cudaMemsetAsync(c->imgEstimate.p, 0, N * sizeof(float4), c->stream));
    // IC buffers: src/IrradianceCache.cpp:46-79 (header{0,maxCaches,0}, zeroed data, zero-radius spheres)
    CUDA_TRY(c->icData.alloc(size_t(ic_size))); CUDA_TRY(c->icSpheres.alloc(size_t(ic_size))); CUDA_TRY(c->icHeader.alloc(1));
    if (ic_size) {
        CUDA_TRY(cudaMemsetAsync(c->icData.p, 0, size_t(ic_size) * sizeof(b200pt_cache_data), c->stream));
        CUDA_TRY(cudaMemsetAsync(c->icSpheres.p, 0, size_t(ic_size) * sizeof(b200pt_sphere), c->stream));
    }
    b200pt_cache_header hdr{0u, uint32_t(ic_size), 0u};
    CUDA_TRY(cudaMemcpyAsync(c->icHeader.p, &hdr, sizeof(hdr), cudaMemcpyHostToDevice, c->stream));
    // DirectionalData buffer: src/SampleCollector.cpp:29-51 (zero-initialised records)
    CUDA_TRY(c->samples.alloc(N * B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL));
    CUDA_TRY(cudaMemsetAsync(c->samples.p, 0, N * B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL * sizeof(b200pt_directional_data), c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out = c;
    return B200PT_OK;
}

int b200pt_destroy(b200pt_ctx *c) {
    if (!c) return B200PT_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->nodes.release(); c->tris.release(); c->sphereGeom.release(); c->texels.release(); c->vertices.release(); c->indices.release();
    c->modelVertexOffset.release(); c->modelIndexOffset.release(); c->randomLightIndex.release(); c->primVerts.release();
    c->materials.release(); c->instances.release(); c->lights.release(); c->randomTriIndex.release(); c->spheres.release(); c->textures.release();
    c->alphaTextures.release(); c->alphaMaterialTexture.release(); c->alphaScene.release();
    c->imgOutput.release(); c->imgAccum.release(); c->imgEstimate.release(); c->imgAov.release();
    for (int i = 0; i < 2; i++) { c->pathRayO[i].release(); c->pathRayD[i].release(); }
    c->pathHit.release(); c->probeRayO.release(); c->probeRayD.release(); c->probeHit.release(); c->probeA.release(); c->probeB.release();
    c->shRayO.release(); c->shRayD.release(); c->shC.release(); c->thr.release(); c->pixelSum.release();
    c->shG.release(); c->recLightSums.release(); c->recSampleThr.release(); c->recPathSum.release(); c->recState.release(); c->recDistanceFactor.release();
    c->seed.release(); c->state.release(); c->sampleIdx.release(); c->counters.release(); c->dstats.release(); c->batchCounter.release();
    for (auto &e : c->ringEvent) if (e) cudaEventDestroy(e);
    if (c->hostDstats) cudaFreeHost(c->hostDstats);
    if (c->dumpIters) fclose(c->dumpIters);
    if (c->rc.comm && g_nccl.handle) g_nccl.CommDestroy(c->rc.comm);
    c->commImage.release(); c->commCount.release(); c->commSamples.release();
    c->samples.release(); c->hostSamples.release(); c->icData.release(); c->icSpheres.release(); c->icHeader.release();
    c->batchRays.release(); c->batchHits.release();
    c->icSnapSphere.release(); c->icSnapNormalR.release(); c->icSnapColor.release(); c->icSnapRot.release(); c->icSnapTrans.release();
    c->icPending.release(); c->icNewEntries.release(); c->icSplitData.release(); c->icRanges.release();
    c->icCellCount.release(); c->icCellStart.release(); c->icCellItems.release(); c->icSnapHdr.release(); c->icBlockCounts.release();
    c->icList.release(); c->icNewCount.release(); c->icSplitState.release(); c->icUpdSlot.release();
    c->icCellSpheres.release(); c->icValidFlags.release(); c->icValidOffsets.release(); c->icQueryResult.release();
    if (c->hostIcHdr) cudaFreeHost(c->hostIcHdr);
    c->guiding.release();
    for (cudaEvent_t e : c->eventPool) cudaEventDestroy(e);
    if (c->hostCounters) cudaFreeHost(c->hostCounters);
    if (c->hostRing) cudaFreeHost(const_cast<uint32_t *>(c->hostRing));
    if (c->evA) cudaEventDestroy(c->evA);
    if (c->evB) cudaEventDestroy(c->evB);
    if (c->evTimer0) cudaEventDestroy(c->evTimer0);
    if (c->evTimer1) cudaEventDestroy(c->evTimer1);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->evFork) cudaEventDestroy(c->evFork);
    if (c->evJoin) cudaEventDestroy(c->evJoin);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return B200PT_OK;
}

static float srgbToLinear(uint8_t v) {
    float c = float(v) / 255.0f;
    return c <= 0.04045f ? c / 12.92f : b200pt_dm::powF((c + 0.055f) / 1.055f, 2.4f);
}

// Shared by b200pt_set_scene and b200pt_scene_bvh_check: per-model buffers concatenated, instances flattened to world-space
// triangles (9 floats each); primVerts = global vertex ids + instance (| identity flag).  Returns what is wrong, or nullptr.
static const char *flattenScene(const b200pt_scene_desc *s, std::vector<b200pt_vertex> &verts, std::vector<uint32_t> &inds, std::vector<int32_t> &vOff,
                                std::vector<int32_t> &iOff, std::vector<float> &world, std::vector<int4> &primVerts) {
    vOff.assign(size_t(std::max(1, s->num_models)), 0); iOff.assign(size_t(std::max(1, s->num_models)), 0);
    for (int m = 0; m < s->num_models; m++) {
        vOff[m] = int32_t(verts.size()); iOff[m] = int32_t(inds.size());
        if (s->num_indices[m] % 3) return "index count not a multiple of 3";
        verts.insert(verts.end(), s->vertices[m], s->vertices[m] + s->num_vertices[m]);
        inds.insert(inds.end(), s->indices[m], s->indices[m] + s->num_indices[m]);
        for (int k = 0; k < s->num_indices[m]; k++)
            if (s->indices[m][k] >= uint32_t(s->num_vertices[m])) return "vertex index out of range";
    }
    for (int i = 0; i < s->num_instances; i++) {
        const b200pt_instance &inst = s->instances[i];
        int m = inst.modelIndex;
        if (m < 0 || m >= s->num_models) return "instance model index out of range";
        int nt = s->num_indices[m] / 3;
        bool identity = true;
        for (int k = 0; k < 16; k++) identity = identity && inst.transform[k] == (k % 5 == 0 ? 1.0f : 0.0f) && inst.normalTransform[k] == (k % 5 == 0 ? 1.0f : 0.0f);
        for (int t = 0; t < nt; t++) {
            int4 pv;
            int *pvp = &pv.x;
            for (int k = 0; k < 3; k++) {
                uint32_t li = s->indices[m][3 * t + k];
                float w[3];
                mat4TransformPoint(inst.transform, s->vertices[m][li].pos, w);
                // an overflowed coordinate or transform (1e40 in a file) would send the builder's binning out of its arrays
                if (!(std::isfinite(w[0]) && std::isfinite(w[1]) && std::isfinite(w[2]))) return "a vertex position is not finite (after the instance transform)";
                world.insert(world.end(), w, w + 3);
                pvp[k] = vOff[m] + int(li);
            }
            pv.w = int(uint32_t(i) | (identity ? PT_INSTANCE_IDENTITY : 0u));
            primVerts.push_back(pv);
        }
    }
    return nullptr;
}

int b200pt_scene_bvh_check(const b200pt_scene_desc *s, b200pt_bvh_report *report) {
    if (!s || !report) return setError(B200PT_E_INVALID, "b200pt_scene_bvh_check: null argument");
    try {
        std::vector<b200pt_vertex> verts; std::vector<uint32_t> inds;
        std::vector<int32_t> vOff, iOff;
        std::vector<float> world; std::vector<int4> primVerts;
        if (const char *bad = flattenScene(s, verts, inds, vOff, iOff, world, primVerts)) return setError(B200PT_E_INVALID, std::string("b200pt_scene_bvh_check: ") + bad);
        Bvh8 bvh;
        buildBvh8(world.data(), uint32_t(primVerts.size()), bvh);
        // test knob: damage the tree before checking it, to show that the check notices (tests/test_scene.py)
        if (const char *m = getenv("B200PT_BVH_CHECK_MUTATE")) {
            const int kind = atoi(m);
            Bvh8Node &n = bvh.nodes[bvh.nodes.size() / 2];
            if (kind == 1) { for (int sl = 0; sl < 8; sl++) if (n.meta[sl]) { n.qhi[0][sl] = n.qlo[0][sl]; n.qhi[1][sl] = n.qlo[1][sl]; } }      // boxes collapsed to a line
            else if (kind == 2 && bvh.tris.size() > 1) bvh.tris[1].prim = bvh.tris[0].prim;                    // one primitive twice, one lost
            else if (kind == 3) { for (int sl = 0; sl < 8; sl++) if (n.meta[sl] && !((n.imask >> sl) & 1)) { n.meta[sl] = uint8_t((5u << 5) | (n.meta[sl] & 31u)); break; } }   // bad unary count
            else if (kind == 4) bvh.maxDepth++;
        }
        Bvh8Report r;
        validateBvh8(bvh, world.data(), uint32_t(primVerts.size()), r);
        report->missing_prims = r.missingPrims; report->duplicate_prims = r.duplicatePrims; report->outside_box = r.outsideBox;
        report->bad_meta = r.badMeta; report->depth_mismatch = r.depthMismatch; report->unreachable_nodes = r.unreachableNodes;
        report->num_nodes = r.numNodes; report->num_tris = r.numTris; report->max_depth = r.maxDepth;
        report->inner_children = r.innerChildren; report->leaf_children = r.leafChildren;
    } catch (const std::exception &e) {
        return setError(B200PT_E_INVALID, std::string("b200pt_scene_bvh_check: ") + e.what());
    }
    return B200PT_OK;
}

int b200pt_set_scene(b200pt_ctx *c, const b200pt_scene_desc *s) {
    if (!c || !s) return setError(B200PT_E_INVALID, "b200pt_set_scene: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    try {
        // concatenate per-model buffers; flatten instances to world-space triangles (primitive ids follow instance order)
        std::vector<b200pt_vertex> verts; std::vector<uint32_t> inds;
        std::vector<int32_t> vOff, iOff;
        std::vector<float> world; std::vector<int4> primVerts;
        if (const char *bad = flattenScene(s, verts, inds, vOff, iOff, world, primVerts)) return setError(B200PT_E_INVALID, std::string("b200pt_set_scene: ") + bad);
        for (int i = 0; i < s->num_materials; i++)
            if ((s->materials[i].textureIdDiffuse >= s->num_textures) || (s->materials[i].textureIdSpecular >= s->num_textures))
                return setError(B200PT_E_INVALID, "b200pt_set_scene: texture id out of range");
        if (s->num_textures < 1) return setError(B200PT_E_INVALID, "b200pt_set_scene: texture 0 (env map slot) is required");
        uint32_t numTris = uint32_t(primVerts.size());
        Bvh8 bvh;
        buildBvh8(world.data(), numTris, bvh);
        if (bvh.maxDepth > PT_STACK_LOCAL) return setError(B200PT_E_INVALID, "b200pt_set_scene: BVH too deep for the traversal stack");
        c->bvhDepth = bvh.maxDepth;
        // alpha test (raytrace.rahit): a texture takes part when one of its texels is not opaque; triangles whose
        // material (of vertex 0, quirk 4) has such a diffuse texture carry a flag in the packed triangle
        std::vector<char> texHasAlpha(size_t(s->num_textures), 0);
        for (int t = 0; t < s->num_textures; t++) {
            const b200pt_texture &tx = s->textures[t];
            if (tx.width <= 0 || tx.height <= 0 || !tx.pixels) continue;
            const size_t np = size_t(tx.width) * tx.height;
            if (tx.format == B200PT_TEX_RGBA32F) { const float *px = static_cast<const float *>(tx.pixels); for (size_t p = 0; p < np; p++) if (px[4 * p + 3] < 1.0f) { texHasAlpha[t] = 1; break; } }
            else { const uint8_t *px = static_cast<const uint8_t *>(tx.pixels); for (size_t p = 0; p < np; p++) if (px[4 * p + 3] != 255) { texHasAlpha[t] = 1; break; } }
        }
        bool anyAlpha = false;
        std::vector<int> matTex(size_t(std::max(1, s->num_materials)), -1);
        for (int i = 0; i < s->num_materials; i++) {
            const int t = s->materials[i].textureIdDiffuse;
            if (t >= 0 && t < s->num_textures && texHasAlpha[t]) { matTex[i] = t; anyAlpha = true; }
        }
        if (anyAlpha) {
            for (PackedTri &pt : bvh.tris) {
                const int mi = verts[size_t(primVerts[pt.prim].x)].materialIndex;
                if (mi >= 0 && mi < s->num_materials && matTex[mi] != -1) { uint32_t one = 1u; memcpy(&pt.pad0, &one, 4); }
            }
        }

        cudaStream_t st = c->stream;
        CUDA_TRY(c->nodes.upload(reinterpret_cast<const float4 *>(bvh.nodes.data()), bvh.nodes.size() * 5, st));
        CUDA_TRY(c->tris.upload(reinterpret_cast<const float4 *>(bvh.tris.data()), bvh.tris.size() * 3, st));
        std::vector<float4> sg(size_t(s->num_spheres));
        for (int i = 0; i < s->num_spheres; i++) sg[i] = make_float4(s->spheres[i].center[0], s->spheres[i].center[1], s->spheres[i].center[2], s->spheres[i].radius);
        CUDA_TRY(c->sphereGeom.upload(sg.data(), sg.size(), st));
        CUDA_TRY(c->vertices.upload(verts.data(), verts.size(), st));
        CUDA_TRY(c->indices.upload(inds.data(), inds.size(), st));
        CUDA_TRY(c->modelVertexOffset.upload(vOff.data(), vOff.size(), st));
        CUDA_TRY(c->modelIndexOffset.upload(iOff.data(), iOff.size(), st));
        CUDA_TRY(c->primVerts.upload(primVerts.data(), primVerts.size(), st));
        CUDA_TRY(c->materials.upload(s->materials, size_t(s->num_materials), st));
        // device copy of the instance records: the std140 padding word carries "both matrices are exactly the identity"
        std::vector<b200pt_instance> devInstances(s->instances, s->instances + s->num_instances);
        for (b200pt_instance &di : devInstances) {
            bool identity = true;
            for (int k = 0; k < 16; k++) identity = identity && di.transform[k] == (k % 5 == 0 ? 1.0f : 0.0f) && di.normalTransform[k] == (k % 5 == 0 ? 1.0f : 0.0f);
            di._pad[0] = identity ? 1 : 0;
        }
        CUDA_TRY(c->instances.upload(devInstances.data(), devInstances.size(), st));
        CUDA_TRY(c->lights.upload(s->lights, size_t(s->num_lights), st));
        CUDA_TRY(c->randomLightIndex.upload(s->random_light_index, B200PT_SIZE_LIGHT_RANDOM, st));
        CUDA_TRY(c->randomTriIndex.upload(s->random_tri_index, size_t(std::max(1, s->num_face_tables)) * B200PT_SIZE_TRI_RANDOM, st));
        CUDA_TRY(c->spheres.upload(s->spheres, size_t(s->num_spheres), st));
        // textures → linear float4 texels (R8G8B8A8_SRGB decodes to linear before filtering)
        std::vector<float4> texels; std::vector<size_t> texOff;
        for (int t = 0; t < s->num_textures; t++) {
            const b200pt_texture &tx = s->textures[t];
            if (tx.width <= 0 || tx.height <= 0 || !tx.pixels) return setError(B200PT_E_INVALID, "b200pt_set_scene: bad texture");
            texOff.push_back(texels.size());
            size_t np = size_t(tx.width) * tx.height;
            if (tx.format == B200PT_TEX_RGBA32F) {
                const float *px = static_cast<const float *>(tx.pixels);
                for (size_t p = 0; p < np; p++) texels.push_back(make_float4(px[4 * p], px[4 * p + 1], px[4 * p + 2], px[4 * p + 3]));
            } else {
                const uint8_t *px = static_cast<const uint8_t *>(tx.pixels);
                for (size_t p = 0; p < np; p++)
                    texels.push_back(make_float4(srgbToLinear(px[4 * p]), srgbToLinear(px[4 * p + 1]), srgbToLinear(px[4 * p + 2]), float(px[4 * p + 3]) / 255.0f));
            }
        }
        CUDA_TRY(c->texels.upload(texels.data(), texels.size(), st));
        std::vector<DeviceTexture> dts(size_t(s->num_textures));
        for (int t = 0; t < s->num_textures; t++) { dts[t].texels = c->texels.p + texOff[t]; dts[t].width = s->textures[t].width; dts[t].height = s->textures[t].height; }
        CUDA_TRY(c->textures.upload(dts.data(), dts.size(), st));
        AlphaScene as{};
        if (anyAlpha) {
            std::vector<AlphaTexture> ats(size_t(s->num_textures));
            for (int t = 0; t < s->num_textures; t++) { ats[t].texels = dts[t].texels; ats[t].width = dts[t].width; ats[t].height = dts[t].height; }
            CUDA_TRY(c->alphaTextures.upload(ats.data(), ats.size(), st));
            CUDA_TRY(c->alphaMaterialTexture.upload(matTex.data(), matTex.size(), st));
            as.primVerts = c->primVerts.p; as.vertices = reinterpret_cast<const float4 *>(c->vertices.p);
            as.materialTexture = c->alphaMaterialTexture.p; as.textures = c->alphaTextures.p;
            CUDA_TRY(c->alphaScene.upload(&as, 1, st));
        }
        CUDA_TRY(cudaStreamSynchronize(st));

        DeviceScene &d = c->dscene;
        d.trace.nodes = c->nodes.p; d.trace.tris = c->tris.p; d.trace.spheres = c->sphereGeom.p;
        d.trace.numTris = numTris; d.trace.numSpheres = uint32_t(s->num_spheres);
        d.trace.alpha = anyAlpha ? c->alphaScene.p : nullptr; d.trace.alphaSeed = 0u;
        d.vertices = c->vertices.p; d.indices = c->indices.p; d.modelVertexOffset = c->modelVertexOffset.p; d.modelIndexOffset = c->modelIndexOffset.p;
        d.primVerts = c->primVerts.p; d.materials = c->materials.p; d.instances = c->instances.p; d.lights = c->lights.p;
        d.randomLightIndex = c->randomLightIndex.p; d.randomTriIndex = c->randomTriIndex.p; d.spheres = c->spheres.p; d.textures = c->textures.p;
        d.numLights = s->num_lights; d.numFaceTables = s->num_face_tables; d.numTextures = s->num_textures; d.numInstances = s->num_instances;
        memcpy(c->sceneMin, s->scene_min, 12); memcpy(c->sceneMax, s->scene_max, 12);
        c->hasScene = true;

        // guiding regions are rebuilt from the scene AABB (RayTracingApp::sceneSwitcher, src/RayTracingApp.cpp:327-367)
        int rc = c->guiding.init(c->guidingSplits, c->sceneMin, c->sceneMax, c->stream);
        if (rc != B200PT_OK) return setError(rc, "b200pt_set_scene: guiding init failed: " + c->guiding.error);
        // and the irradiance cache starts empty
        b200pt_cache_header hdr{0u, uint32_t(c->icSize), 0u};
        CUDA_TRY(cudaMemcpyAsync(c->icHeader.p, &hdr, sizeof(hdr), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    } catch (const std::exception &e) {
        return setError(B200PT_E_INVALID, std::string("b200pt_set_scene: ") + e.what());
    }
    return B200PT_OK;
}

int b200pt_set_camera(b200pt_ctx *c, const float view[16], const float proj[16]) {
    if (!c || !view || !proj) return setError(B200PT_E_INVALID, "b200pt_set_camera: null argument");
    memcpy(c->view, view, 64); memcpy(c->proj, proj, 64);
    if (!mat4Inverse(view, c->viewInv) || !mat4Inverse(proj, c->projInv)) return setError(B200PT_E_INVALID, "b200pt_set_camera: singular matrix");
    c->hasCamera = true;
    return B200PT_OK;
}

// one frame (batchCount <= 1) or a batch of plain frames that differ only in seed and previousFrames (pc = the first)
static int renderFrames(b200pt_ctx *c, const b200pt_push_constants *pc, const b200pt_push_constants *batchPcs, int batchCount);

int b200pt_render_frame(b200pt_ctx *c, const b200pt_push_constants *pc) {
    if (!c || !pc) return setError(B200PT_E_INVALID, "b200pt_render_frame: null argument");
    return renderFrames(c, pc, nullptr, 0);
}

int b200pt_render_frames(b200pt_ctx *c, const b200pt_push_constants *pcs, int count) {
    if (!c || !pcs || count < 1) return setError(B200PT_E_INVALID, "b200pt_render_frames: bad argument");
    // the frame walk needs frames that are independent of each other and identical up to the seed and the frame counter
    bool walk = count > 1 && count < 65536 && pcs[0].samplesPerPixel < 65536 && !c->aovs && c->hasScene && c->dscene.trace.alpha == nullptr &&
                !getenv("B200PT_NO_FRAME_WALK");
    const b200pt_push_constants &a = pcs[0];
    if (a.useIrradianceCache || a.useADRRS || a.splitOnFirst || a.useGuiding || a.updateGuiding || a.storeEstimate || a.isIrradiancePrepareFrame ||
        a.visualizeMode != 0 || a.showIrradianceCacheOnly) walk = false;
    for (int i = 1; walk && i < count; i++) {
        b200pt_push_constants b = pcs[i];
        b.randomUInt = a.randomUInt; b.previousFrames = a.previousFrames;
        if (memcmp(&a, &b, sizeof(b)) != 0) walk = false;
    }
    if (!walk) {
        for (int i = 0; i < count; i++) { const int rc = renderFrames(c, &pcs[i], nullptr, 0); if (rc != B200PT_OK) return rc; }
        return B200PT_OK;
    }
    return renderFrames(c, &pcs[0], pcs, count);
}

static int renderFrames(b200pt_ctx *c, const b200pt_push_constants *pc, const b200pt_push_constants *batchPcs, int batchCount) {
    if (!c->hasScene || !c->hasCamera) return setError(B200PT_E_STATE, "b200pt_render_frame: set_scene and set_camera must be called first");
    if (pc->showIrradianceCacheOnly || pc->visualizeMode != 0)
        return setError(B200PT_E_STATE, "b200pt_render_frame: the debug views (visualizeMode, showIrradianceCacheOnly) are not part of this library");
    const bool useCache = pc->useIrradianceCache || pc->useADRRS;
    const bool splitMode = (pc->useADRRS && pc->adrrsSplit) || pc->splitOnFirst;
    const bool icMode = useCache || splitMode;
    if (useCache && c->icSize < 1) return setError(B200PT_E_STATE, "b200pt_render_frame: irradiance cache / ADRRS need a context created with ic_size > 0");
    const bool guided = pc->useGuiding || pc->updateGuiding;
    if (guided && !c->guiding.ready) return setError(B200PT_E_STATE, "b200pt_render_frame: guiding needs a scene (region tree)");
    if (pc->numNEE < 1 || pc->samplesPerPixel < 1 || pc->maxDepth < 0 || pc->maxDepth > 60000)
        return setError(B200PT_E_INVALID, "b200pt_render_frame: numNEE, samplesPerPixel must be >= 1 and maxDepth in [0, 60000]");
    if (useCache && pc->irradianceNumNEE < 1) return setError(B200PT_E_INVALID, "b200pt_render_frame: irradianceNumNEE must be >= 1");
    CUDA_TRY(cudaSetDevice(c->device));
    // a pixel that finishes a path and starts a split in the same shade pass queues two rounds of light samples
    int rc = ensureQueues(c, (pc->enableNEE || useCache ? pc->numNEE : 1) * (splitMode ? 2 : 1));
    if (rc != B200PT_OK) return rc;
    if (icMode) { rc = ensureIC(c, useCache, splitMode); if (rc != B200PT_OK) return rc; }
    const bool batch = batchCount > 1;
    c->wf.batch = FrameBatch{};
    if (batch) {
        CUDA_TRY(c->pixelSum.alloc(size_t(c->numPixels) * 2));      // one half per frame parity
        c->wf.pixelSum = c->pixelSum.p;
        std::vector<uint32_t> seeds(size_t(batchCount), 0u), prevs(size_t(batchCount), 0u);
        for (int i = 0; i < batchCount; i++) { seeds[size_t(i)] = batchPcs[i].randomUInt; prevs[size_t(i)] = batchPcs[i].previousFrames; }
        CUDA_TRY(c->batchSeeds.upload(seeds.data(), seeds.size(), c->stream));
        CUDA_TRY(c->batchPrev.upload(prevs.data(), prevs.size(), c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));                 // the staging vectors go out of scope
        c->wf.batch.frameSeed = c->batchSeeds.p; c->wf.batch.framePrev = c->batchPrev.p; c->wf.batch.numFrames = batchCount;
        c->wf.batch.image = c->imgOutput.p; c->wf.batch.accum = c->imgAccum.p;
    }
    // samples are not collected while splitting: the raygen returns right after resetting them (rgen:1643-1650)
    const bool earlyReturn = pc->updateGuiding && splitMode;

    {   // guiding views: region tree + mixtures, and (training frames) the sample-recording state
        Wavefront &w = c->wf;
        w.guide.levels = c->guiding.levelAabbs; w.guide.levelSplits = c->guiding.levelSplits; w.guide.vmms = c->guiding.vmms; w.guide.splits = c->guiding.splits;
        w.guide.aabbs = c->guiding.aabbs;
        w.guide.spawnFirst = c->guiding.hasSpawns ? c->guiding.spawnFirst : nullptr;
        w.guide.spawnNext = c->guiding.hasSpawns ? c->guiding.spawnNext : nullptr;
        w.rec = GuidingRecord{};
        if (pc->updateGuiding) {
            const size_t N16 = size_t(c->numPixels) * B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL;
            CUDA_TRY(c->recLightSums.alloc(N16)); CUDA_TRY(c->recSampleThr.alloc(N16));
            CUDA_TRY(c->recState.alloc(size_t(c->numPixels))); CUDA_TRY(c->recDistanceFactor.alloc(size_t(c->numPixels)));
            CUDA_TRY(c->recPathSum.alloc(size_t(c->numPixels)));
            CUDA_TRY(c->shG.alloc(size_t(c->numPixels) * size_t(c->queueNEE)));
            w.rec.samples = c->samples.p; w.rec.lightSums = c->recLightSums.p; w.rec.sampleThr = c->recSampleThr.p;
            w.rec.state = c->recState.p; w.rec.distanceFactor = c->recDistanceFactor.p; w.rec.pathSum = c->recPathSum.p;
            w.shG = c->shG.p;
        }
        w.aov = nullptr;
        if (c->aovs) { CUDA_TRY(c->imgAov.alloc(size_t(c->numPixels))); w.aov = c->imgAov.p; }
        w.ic = ICState{};
        if (useCache) {
            w.ic.view = c->icGrid;
            w.ic.estimate = c->imgEstimate.p;
            w.ic.newCount = c->icNewCount.p; w.ic.newEntries = c->icNewEntries.p; w.ic.queryResult = c->icQueryResult.p;
        }
        if (splitMode) { w.ic.splitState = c->icSplitState.p; w.ic.splitData = c->icSplitData.p; }
        w.regenQ = (icMode && c->regenSplit) ? c->regenQ.p : nullptr;
        w.shadeOrder = (useCache && c->icqSort && c->shadeSorted) ? c->icqOrder.p : nullptr;
    }
    FrameParams fp;
    fp.pc = *pc;
    memcpy(fp.view, c->view, 64); memcpy(fp.proj, c->proj, 64); memcpy(fp.viewInv, c->viewInv, 64); memcpy(fp.projInv, c->projInv, 64);
    fp.width = c->width; fp.height = c->height; fp.numPixels = c->numPixels;
    fp.samplesPerPixel = pc->isIrradiancePrepareFrame ? 1 : pc->samplesPerPixel;

    cudaStream_t st = c->stream;
    const uint32_t N = uint32_t(c->numPixels);
    c->dscene.trace.alphaSeed = pc->randomUInt;
    CUDA_TRY(cudaEventRecord(c->evA, st));
    CUDA_TRY(cudaMemsetAsync(c->dstats.p, 0, DST_NUM * sizeof(unsigned long long), st));

    if (useCache && c->icqSort)      // the sort histogram is zero between two sorts — unless an earlier frame was abandoned between them
        CUDA_TRY(cudaMemsetAsync(c->icqHist.p, 0, (size_t(IC_GRID_MAX) * IC_GRID_MAX * IC_GRID_MAX + ICQ_EXTRA_BINS) * sizeof(uint32_t), st));
    if (useCache) {
        // frame-start snapshot of the cache + lookup grid (replaces IrradianceCache::updateSpheres, src/IrradianceCache.cpp:81-104)
        StageTimer t(c, KIND_SHADE);
        ICBuffers b = icBuffers(c);
        k_ic_snapshot<<<gridFor(uint64_t(std::max(1, c->icSize)), 256), 256, 0, st>>>(b, c->icGrid);
        k_ic_cells<false><<<gridFor(uint64_t(c->icNumCells), 256), 256, 0, st>>>(b, c->icGrid);
        k_scan_single_block<<<1, 1024, 0, st>>>(c->icCellCount.p, c->icCellStart.p, c->icNumCells, c->icSnapHdr.p + ICH_GRID_TOTAL);
        c->stats.kernel_launches += 2;
        rc = readIcHeader(c);
        if (rc != B200PT_OK) return rc;
        if (c->hostIcHdr[ICH_GRID_TOTAL] > c->icCellItems.n) {
            CUDA_TRY(c->icCellItems.alloc(size_t(c->hostIcHdr[ICH_GRID_TOTAL]) * 2));
            CUDA_TRY(c->icCellSpheres.alloc(size_t(c->hostIcHdr[ICH_GRID_TOTAL]) * 2));
            c->icGrid.cellItems = c->icCellItems.p; c->wf.ic.view.cellItems = c->icCellItems.p;
            c->icGrid.cellSpheres = c->icCellSpheres.p; c->wf.ic.view.cellSpheres = c->icCellSpheres.p;
            b = icBuffers(c);
        }
        if (c->hostIcHdr[ICH_GRID_TOTAL]) { k_ic_cells<true><<<gridFor(uint64_t(c->icNumCells), 256), 256, 0, st>>>(b, c->icGrid); c->stats.kernel_launches++; }
    }
    bool overlapUpdate = false;
    // an error return while the cache update is still running on the second stream must not leave it behind (the next call may
    // release the buffers it works on)
    struct JoinGuard { cudaStream_t s; bool armed; ~JoinGuard() { if (armed) cudaStreamSynchronize(s); } } joinGuard{c->stream2, false};
    if (pc->useIrradianceCache && pc->irradianceUpdateProb > 0.0f) {
        // updateIrradianceCache (rgen:1334-1381) for the pixels whose first random number selects them, before any path
        PredICUpdate pred{pc->randomUInt, pc->irradianceUpdateProb};
        rc = compactPixels(c, pred);
        if (rc != B200PT_OK) return rc;
        const uint32_t entries = c->hostIcHdr[ICH_LIST_COUNT];
        if (entries) {
            StageTimer t(c, KIND_SHADE);
            CUDA_TRY(c->icUpdSlot.alloc(entries)); CUDA_TRY(c->icPending.alloc(size_t(entries) * 3));
            ICBuffers b = icBuffers(c);
            int grid, stride, lanes;
            buildLaunchShape(c, entries, grid, stride, lanes);
            k_ic_update_assign<<<1, 32, 0, st>>>(b);
            // One entry is ~10 ms of dependent work on a handful of warps, whatever the list length, and nothing in the frame waits
            // for it except the selected pixels themselves (lookups read the frame-start snapshot; the blend is committed afterwards):
            // it runs on a second stream BESIDE the wavefront loop, the selected pixels start their paths when it is done.
            overlapUpdate = c->icOverlap && !earlyReturn && !batch;
            if (overlapUpdate) {
                CUDA_TRY(cudaEventRecord(c->evFork, st));
                CUDA_TRY(cudaStreamWaitEvent(c->stream2, c->evFork, 0));
                k_ic_update<<<grid, 128, 0, c->stream2>>>(fp, c->dscene, c->wf, b, stride, lanes);
                CUDA_TRY(cudaEventRecord(c->evJoin, c->stream2));
                joinGuard.armed = true;
                c->stats.kernel_launches += 1;
            } else {
                k_ic_update<<<grid, 128, 0, st>>>(fp, c->dscene, c->wf, b, stride, lanes);
                k_ic_update_commit<<<1, 32, 0, st>>>(fp, b);
                c->stats.kernel_launches += 2;
            }
        }
    }

    c->stats.samples += uint64_t(N) * uint64_t(batch ? batchCount : 1);
    uint32_t init[CNT_NUM] = {overlapUpdate ? 0u : N, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(c->counters.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    {
        StageTimer t(c, KIND_SHADE);
        if (overlapUpdate) k_generate<true><<<gridFor(N, 256), 256, 0, st>>>(fp, c->wf);
        else k_generate<false><<<gridFor(N, 256), 256, 0, st>>>(fp, c->wf);
    }
    // Wavefront loop.  Every kernel reads its queue sizes from device memory, so iterations are issued back-to-back;
    // the host only peeks at the counters of iteration i-LAG to learn when the queues have drained.
    int cur = 0;
    const uint64_t pathsPerPixel = uint64_t(fp.samplesPerPixel) + (splitMode ? IC_MAX_SPLITS : 0);
    const uint64_t maxIter = uint64_t(batch ? batchCount : 1) * pathsPerPixel * (uint64_t(pc->maxDepth) + uint64_t(std::max(0, pc->maxFollowDiscrete)) + 4) + 4 + b200pt_ctx::LAG;
    bool drained = earlyReturn;
    const uint32_t seqBase = c->ringSeq;          // sequence numbers never repeat across frames
    uint64_t iter = 0, passStart = 0;
  for (int pass = 0; pass < (overlapUpdate ? 2 : 1); pass++) {
    if (pass == 1) {
        // the cache update has finished: blend its results in and start the paths of the pixels it selected
        StageTimer t(c, KIND_SHADE);
        CUDA_TRY(cudaStreamWaitEvent(st, c->evJoin, 0));
        joinGuard.armed = false;          // from here on the frame's own stream is ordered behind it
        const ICBuffers b = icBuffers(c);
        k_ic_update_commit<<<1, 32, 0, st>>>(fp, b);
        k_generate_list<<<gridFor(uint64_t(c->hostIcHdr[ICH_LIST_COUNT]), 256), 256, 0, st>>>(fp, c->wf, c->icList.p, c->icSnapHdr.p + ICH_LIST_COUNT, cur);
        c->stats.kernel_launches++;
        drained = false;
        passStart = iter;
    }
    for (; !drained; iter++) {
        if (iter > maxIter + passStart) return setError(B200PT_E_STATE, "b200pt_render_frame: wavefront did not drain (internal error)");
        {
            StageTimer t(c, KIND_EXTEND);
            const bool alpha = c->dscene.trace.alpha != nullptr;     // some texture of the scene has a transparent texel
            if (pc->updateGuiding) {
                if (alpha) k_trace<true, 1><<<c->traceGridRec, PT_TRACE_BLOCK, 0, st>>>(c->dscene.trace, c->wf, cur, c->tune);
                else k_trace<true, 0><<<c->traceGridRec, PT_TRACE_BLOCK, 0, st>>>(c->dscene.trace, c->wf, cur, c->tune);
            } else {
                if (alpha) k_trace<false, 1><<<c->traceGrid, PT_TRACE_BLOCK, 0, st>>>(c->dscene.trace, c->wf, cur, c->tune);
                else k_trace<false, 0><<<c->traceGrid, PT_TRACE_BLOCK, 0, st>>>(c->dscene.trace, c->wf, cur, c->tune);
            }
        }
        const int slot = int(iter % b200pt_ctx::RING);
        const bool probes = (pc->enableNEE || useCache) && pc->enableMIS;
        volatile uint32_t *prepSlot = c->counterCopy ? nullptr : c->hostRingDev + slot * 8;
        const uint32_t prepSeq = c->counterCopy ? 0u : seqBase + uint32_t(iter) + 1u;
        if (probes && c->fusePrep) {
            StageTimer t(c, KIND_SHADE);
            k_probe_resolve_prep<<<c->resolveGrid, 256, 0, st>>>(fp, c->dscene, c->wf, cur, prepSlot, prepSeq);
        } else {
            if (probes) { StageTimer t(c, KIND_SHADE); k_probe_resolve<<<c->resolveGrid, 256, 0, st>>>(fp, c->dscene, c->wf); }
            k_iter_prep<<<1, 32, 0, st>>>(c->wf, cur, prepSlot, prepSeq);
            c->stats.kernel_launches++;
        }
        if (useCache && c->icqSort) {
            StageTimer t(c, KIND_SHADE);
            const ICQuerySort qs{c->icqKey.p, c->icqHist.p, c->icqCursor.p, c->icqOrder.p, c->icNumCells, c->shadeSorted ? 1 : 0};
            k_icq_count<<<c->icQueryGrid, 256, 0, st>>>(fp, c->wf, qs, cur);
            k_icq_scan<<<1, 1024, 0, st>>>(qs, c->counters.p);
            k_icq_scatter<<<c->icQueryGrid, 256, 0, st>>>(c->wf, qs);
            k_ic_query_sorted<<<c->icQueryGrid, 256, 0, st>>>(fp, c->dscene, c->wf, qs, cur);
            c->stats.kernel_launches += 3;
        } else if (useCache) { StageTimer t(c, KIND_SHADE); k_ic_query<<<c->icQueryGrid, 256, 0, st>>>(fp, c->dscene, c->wf, cur); }
        if (icMode && c->regenSplit) {
            StageTimer t(c, KIND_SHADE);
            if (guided) {
                k_shade<true, true, false, true><<<c->shadeGridGuidedICDefer, 128, 0, st>>>(fp, c->dscene, c->wf, cur);
                k_regen<true><<<c->regenGridGuided, 128, 0, st>>>(fp, c->dscene, c->wf, cur);
            } else {
                k_shade<false, true, false, true><<<c->shadeGridICDefer, 128, 0, st>>>(fp, c->dscene, c->wf, cur);
                k_regen<false><<<c->regenGrid, 128, 0, st>>>(fp, c->dscene, c->wf, cur);
            }
            c->stats.kernel_launches++;
        } else {
            StageTimer t(c, KIND_SHADE);
            if (guided && icMode) k_shade<true, true><<<c->shadeGridGuidedIC, 128, 0, st>>>(fp, c->dscene, c->wf, cur);
            else if (icMode) k_shade<false, true><<<c->shadeGridIC, 128, 0, st>>>(fp, c->dscene, c->wf, cur);
            else if (guided) k_shade<true, false><<<c->shadeGridGuided, 128, 0, st>>>(fp, c->dscene, c->wf, cur);
            else if (batch) k_shade<false, false, true><<<c->shadeGridBatch, 128, 0, st>>>(fp, c->dscene, c->wf, cur);
            else k_shade<false, false><<<c->shadeGrid, 128, 0, st>>>(fp, c->dscene, c->wf, cur);
        }
        cur = 1 - cur;
        uint32_t qPath = 1, qProbe = 0, qShadow = 0;
        bool have = false;
        if (c->counterCopy) {
            CUDA_TRY(cudaMemcpyAsync(c->hostCounters + slot * CNT_NUM, c->counters.p, CNT_NUM * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaEventRecord(c->ringEvent[slot], st));
            if (iter >= passStart + b200pt_ctx::LAG) {
                const int old = int((iter - b200pt_ctx::LAG) % b200pt_ctx::RING);
                CUDA_TRY(cudaEventSynchronize(c->ringEvent[old]));
                const uint32_t *hc = c->hostCounters + old * CNT_NUM;
                // after iteration j the live path queue is PATH[(j+1)&1]; all three empty => every later iteration is a no-op
                const int liveQ = int((iter - b200pt_ctx::LAG + 1) & 1);
                qPath = hc[CNT_PATH0 + liveQ]; qProbe = hc[CNT_PROBE]; qShadow = hc[CNT_SHADOW];
                have = true;
            }
        } else if (iter >= passStart + b200pt_ctx::LAG) {
            // queue sizes iteration j = iter - LAG STARTED with (= what the shade pass of j - 1 left), published by its k_iter_prep
            const uint64_t j = iter - b200pt_ctx::LAG;
            volatile uint32_t *hs = c->hostRing + (j % b200pt_ctx::RING) * 8;
            const uint32_t want = seqBase + uint32_t(j) + 1u;
            for (uint64_t spins = 0; hs[7] != want; spins++) {
                if ((spins & 0x3ff) == 0x3ff) {      // a faulted stream would never publish: do not spin forever
                    const cudaError_t q = cudaStreamQuery(st);
                    if (q != cudaErrorNotReady && hs[7] != want) {
                        if (q == cudaSuccess) return setError(B200PT_E_STATE, "b200pt_render_frame: iteration counters were not published (internal error)");
                        return setError(B200PT_E_CUDA, std::string("b200pt_render_frame: ") + cudaGetErrorString(q));
                    }
                }
            }
            __sync_synchronize();
            qPath = hs[0]; qProbe = hs[1]; qShadow = hs[2];
            have = true;
        }
        if (have) {
            if (qPath > N || qProbe > N * uint32_t(c->queueNEE) || qShadow > N * uint32_t(c->queueNEE))
                return setError(B200PT_E_STATE, "b200pt_render_frame: queue overflow (internal error)");
            drained = qPath == 0 && qProbe == 0 && qShadow == 0;
            if (c->dumpIters) fprintf(c->dumpIters, "%llu %u %u %u\n", (unsigned long long)(iter - b200pt_ctx::LAG), qPath, qProbe, qShadow);
        }
    }
  }
    c->ringSeq = seqBase + uint32_t(iter) + 8u;
    if (!earlyReturn) {
        StageTimer t(c, KIND_SHADE);
        if (batch) {      // the last frame of the batch (every earlier one was folded in by the pixels themselves)
            fp.pc.previousFrames = batchPcs[batchCount - 1].previousFrames;
            k_accumulate<<<gridFor(N, 256), 256, 0, st>>>(fp, c->wf.pixelSum + ((batchCount - 1) & 1) * size_t(N), c->imgOutput.p, c->imgAccum.p, c->imgEstimate.p, c->wf.rec);
        } else
        k_accumulate<<<gridFor(N, 256), 256, 0, st>>>(fp, c->wf.pixelSum, c->imgOutput.p, c->imgAccum.p, c->imgEstimate.p, c->wf.rec);
    }
    if (useCache && !earlyReturn) {
        // createIrradianceCache for the entries the frame's paths queued (rgen:1821-1827), slots in pixel order
        PredICCreate pred{c->icNewCount.p};
        rc = compactPixels(c, pred);
        if (rc != B200PT_OK) return rc;
        const uint32_t entries = c->hostIcHdr[ICH_LIST_COUNT];
        if (entries) {
            StageTimer t(c, KIND_SHADE);
            const size_t slots = size_t(entries) * IC_MAX_NEW;
            CUDA_TRY(c->icPending.alloc(slots * 3)); CUDA_TRY(c->icValidFlags.alloc(slots)); CUDA_TRY(c->icValidOffsets.alloc(slots + 1));
            ICBuffers b = icBuffers(c);
            int grid, stride, lanes;
            buildLaunchShape(c, entries, grid, stride, lanes);
            k_ic_create<<<grid, 128, 0, st>>>(fp, c->dscene, c->wf, b, stride, lanes);
            k_scan_single_block<<<1, 1024, 0, st>>>(c->icValidFlags.p, c->icValidOffsets.p, int(slots), nullptr);
            k_ic_create_commit<<<gridFor(slots, 256), 256, 0, st>>>(fp, c->wf, b);
            c->stats.kernel_launches += 2;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(c->hostDstats, c->dstats.p, DST_NUM * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(c->evB, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, c->evA, c->evB));
    c->stats.ms_total += ms;
    c->stats.extend_rays += c->hostDstats[DST_EXTEND];
    c->stats.shadow_rays += c->hostDstats[DST_SHADOW];
    c->stats.path_vertices += c->hostDstats[DST_VERTICES];
    c->stats.iterations += c->hostDstats[DST_ITERATIONS];
    resolveStageTimes(c);
    return B200PT_OK;
}

static float4 *imagePtr(b200pt_ctx *c, int which) {
    switch (which) {
        case B200PT_IMAGE_OUTPUT: return c->imgOutput.p;
        case B200PT_IMAGE_ACCUM: return c->imgAccum.p;
        case B200PT_IMAGE_ESTIMATE: return c->imgEstimate.p;
        default: return nullptr;
    }
}

int b200pt_read_image(b200pt_ctx *c, int which, float *rgba) {
    if (!c || !rgba || !imagePtr(c, which)) return setError(B200PT_E_INVALID, "b200pt_read_image: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(rgba, imagePtr(c, which), size_t(c->numPixels) * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}
int b200pt_write_image(b200pt_ctx *c, int which, const float *rgba) {
    if (!c || !rgba || !imagePtr(c, which)) return setError(B200PT_E_INVALID, "b200pt_write_image: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(imagePtr(c, which), rgba, size_t(c->numPixels) * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}
int b200pt_read_image_device(b200pt_ctx *c, int which, void *dst) {
    if (!c || !dst || !imagePtr(c, which)) return setError(B200PT_E_INVALID, "b200pt_read_image_device: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(dst, imagePtr(c, which), size_t(c->numPixels) * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}
int b200pt_write_image_device(b200pt_ctx *c, int which, const void *src) {
    if (!c || !src || !imagePtr(c, which)) return setError(B200PT_E_INVALID, "b200pt_write_image_device: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(imagePtr(c, which), src, size_t(c->numPixels) * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}

int b200pt_trace_rays_device(b200pt_ctx *c, const void *rays, int64_t n, void *hits, int any_hit) {
    if (!c || (n > 0 && (!rays || !hits)) || n < 0) return setError(B200PT_E_INVALID, "b200pt_trace_rays_device: bad argument");
    if (!c->hasScene) return setError(B200PT_E_STATE, "b200pt_trace_rays: set_scene must be called first");
    if (n == 0) return B200PT_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    if (n > int64_t(0x7fffffff)) return setError(B200PT_E_INVALID, "b200pt_trace_rays_device: at most 2^31-1 rays per call");
    CUDA_TRY(cudaMemsetAsync(c->batchCounter.p, 0, sizeof(uint32_t), c->stream));
    if (c->dscene.trace.alpha)
        k_trace_batch<1><<<c->traceGrid, PT_TRACE_BLOCK, 0, c->stream>>>(c->dscene.trace, static_cast<const float4 *>(rays), static_cast<float4 *>(hits), uint32_t(n),
                                                                           any_hit, c->batchCounter.p, c->tune);
    else
        k_trace_batch<0><<<c->traceGrid, PT_TRACE_BLOCK, 0, c->stream>>>(c->dscene.trace, static_cast<const float4 *>(rays), static_cast<float4 *>(hits), uint32_t(n),
                                                                            any_hit, c->batchCounter.p, c->tune);
    c->stats.kernel_launches++;
    if (any_hit) { c->stats.shadow_rays += uint64_t(n); c->stats.launches_shadow++; } else { c->stats.extend_rays += uint64_t(n); c->stats.launches_extend++; }
    CUDA_TRY(cudaGetLastError());
    return B200PT_OK;
}

int b200pt_trace_rays(b200pt_ctx *c, const b200pt_ray *rays, int64_t n, b200pt_hit *hits, int any_hit) {
    if (!c || (n > 0 && (!rays || !hits)) || n < 0) return setError(B200PT_E_INVALID, "b200pt_trace_rays: bad argument");
    if (!c->hasScene) return setError(B200PT_E_STATE, "b200pt_trace_rays: set_scene must be called first");
    if (n == 0) return B200PT_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(c->batchRays.alloc(size_t(n) * 2)); CUDA_TRY(c->batchHits.alloc(size_t(n)));
    CUDA_TRY(cudaMemcpyAsync(c->batchRays.p, rays, size_t(n) * sizeof(b200pt_ray), cudaMemcpyHostToDevice, c->stream));
    int rc = b200pt_trace_rays_device(c, c->batchRays.p, n, c->batchHits.p, any_hit);
    if (rc != B200PT_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(hits, c->batchHits.p, size_t(n) * sizeof(b200pt_hit), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}

int b200pt_stats_get(b200pt_ctx *c, b200pt_stats *out) {
    if (!c || !out) return setError(B200PT_E_INVALID, "b200pt_stats_get: null argument");
    *out = c->stats;
    return B200PT_OK;
}
int b200pt_stats_reset(b200pt_ctx *c) {
    if (!c) return setError(B200PT_E_INVALID, "b200pt_stats_reset: null argument");
    memset(&c->stats, 0, sizeof(c->stats));
    return B200PT_OK;
}
int b200pt_set_stage_timing(b200pt_ctx *c, int enabled) {
    if (!c) return setError(B200PT_E_INVALID, "b200pt_set_stage_timing: null argument");
    c->stageTiming = enabled < 0 || enabled > 2 ? 1 : enabled;
    return B200PT_OK;
}
int b200pt_timer_start(b200pt_ctx *c) {
    if (!c) return setError(B200PT_E_INVALID, "b200pt_timer_start: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventRecord(c->evTimer0, c->stream));
    return B200PT_OK;
}
int b200pt_timer_stop(b200pt_ctx *c, float *ms) {
    if (!c || !ms) return setError(B200PT_E_INVALID, "b200pt_timer_stop: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventRecord(c->evTimer1, c->stream));
    CUDA_TRY(cudaEventSynchronize(c->evTimer1));
    CUDA_TRY(cudaEventElapsedTime(ms, c->evTimer0, c->evTimer1));
    return B200PT_OK;
}
int b200pt_synchronize(b200pt_ctx *c) {
    if (!c) return setError(B200PT_E_INVALID, "b200pt_synchronize: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}

// ---- guiding ---------------------------------------------------------------------------------------------------
void b200pt_default_guiding_params(b200pt_guiding_params *p) {
    if (!p) return;
    p->useParallaxCompensation = 1; p->splitAndMerge = 1;
    p->minSamplesForMerging = 8192; p->minSamplesForSplitting = 4096; p->minSamplesForPostSplitFitting = 4096;
    p->splitMinDivergence = 0.5f; p->mergeMaxDivergence = 0.025f;
    p->numInitialComponents = 8; p->minItr = 1; p->maxItr = 100; p->relLogLikelihoodThreshold = 0.005f;
    p->initKappa = 5.0f; p->maxKappa = 50000.0f; p->vPrior = 0.01f; p->rPrior = 0.0f; p->rPriorWeight = 1.0f;
    p->splitRegions = 0; p->samplesForRegionSplit = 10000.0f;
}

int b200pt_guiding_update(b200pt_ctx *c, const b200pt_guiding_params *params) {
    if (!c || !params) return setError(B200PT_E_INVALID, "b200pt_guiding_update: null argument");
    if (!c->guiding.ready) return setError(B200PT_E_STATE, "b200pt_guiding_update: set_scene must be called first");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = c->guiding.update(c->samples.p, int64_t(c->numPixels) * B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL, *params, c->stream, &c->stats);
    if (rc != B200PT_OK) return setError(rc, "b200pt_guiding_update: " + c->guiding.error);
    return B200PT_OK;
}
int b200pt_guiding_reset(b200pt_ctx *c, const b200pt_guiding_params *params) {
    if (!c || !params) return setError(B200PT_E_INVALID, "b200pt_guiding_reset: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = c->guiding.reset(*params, c->stream);
    if (rc != B200PT_OK) return setError(rc, "b200pt_guiding_reset: " + c->guiding.error);
    return B200PT_OK;
}
int b200pt_guiding_update_device(b200pt_ctx *c, const b200pt_guiding_params *params, const void *samples_device, int64_t n) {
    if (!c || !params || n < 0 || (n > 0 && !samples_device)) return setError(B200PT_E_INVALID, "b200pt_guiding_update_device: bad argument");
    if (!c->guiding.ready) return setError(B200PT_E_STATE, "b200pt_guiding_update_device: set_scene must be called first");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = c->guiding.update(static_cast<b200pt_directional_data *>(const_cast<void *>(samples_device)), n, *params, c->stream, &c->stats);
    if (rc != B200PT_OK) return setError(rc, "b200pt_guiding_update_device: " + c->guiding.error);
    return B200PT_OK;
}
int b200pt_guiding_update_host(b200pt_ctx *c, const b200pt_guiding_params *params, const b200pt_directional_data *samples, int64_t n) {
    if (!c || !params || n < 0 || (n > 0 && !samples)) return setError(B200PT_E_INVALID, "b200pt_guiding_update_host: bad argument");
    if (!c->guiding.ready) return setError(B200PT_E_STATE, "b200pt_guiding_update_host: set_scene must be called first");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(c->hostSamples.alloc(size_t(n)));
    if (n) CUDA_TRY(cudaMemcpyAsync(c->hostSamples.p, samples, size_t(n) * sizeof(b200pt_directional_data), cudaMemcpyHostToDevice, c->stream));
    int rc = c->guiding.update(c->hostSamples.p, n, *params, c->stream, &c->stats);
    if (rc != B200PT_OK) return setError(rc, "b200pt_guiding_update_host: " + c->guiding.error);
    return B200PT_OK;
}
int b200pt_guiding_set_order(b200pt_ctx *c, int order) {
    if (!c || (order != B200PT_GUIDING_ORDER_STRICT && order != B200PT_GUIDING_ORDER_REORDERED)) return setError(B200PT_E_INVALID, "b200pt_guiding_set_order: bad argument");
    c->guiding.summationOrder = order;
    return B200PT_OK;
}
int64_t b200pt_guiding_sorted_count(b200pt_ctx *c) { return c ? int64_t(c->guiding.lastValidSamples) : 0; }
int b200pt_guiding_get_sorted(b200pt_ctx *c, b200pt_directional_data *out, uint32_t *region_offsets) {
    if (!c) return setError(B200PT_E_INVALID, "b200pt_guiding_get_sorted: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = c->guiding.getSorted(out, region_offsets, nullptr, c->stream);
    if (rc != B200PT_OK) return setError(rc, "b200pt_guiding_get_sorted: " + c->guiding.error);
    return B200PT_OK;
}
int b200pt_guiding_get_state(b200pt_ctx *c, int region, float scalars5[5], float per_component[224]) {
    if (!c || !scalars5 || !per_component) return setError(B200PT_E_INVALID, "b200pt_guiding_get_state: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = c->guiding.getState(region, scalars5, per_component, c->stream);
    if (rc != B200PT_OK) return setError(rc, "b200pt_guiding_get_state: " + c->guiding.error);
    return B200PT_OK;
}
int b200pt_guiding_fastexp(b200pt_ctx *c, const float *in, float *out, int n) {
    if (!c || n < 0 || (n > 0 && (!in || !out))) return setError(B200PT_E_INVALID, "b200pt_guiding_fastexp: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    std::string err;
    int rc = guidingFastExp(in, out, n, c->stream, err);
    if (rc != B200PT_OK) return setError(rc, "b200pt_guiding_fastexp: " + err);
    return B200PT_OK;
}
}  // extern "C"
__host__ __device__ static inline float detmathApply(int fn, float a, float b) {
    switch (fn) {
        case B200PT_DM_SIN: return b200pt_dm::sinF(a);
        case B200PT_DM_COS: return b200pt_dm::cosF(a);
        case B200PT_DM_TAN: return b200pt_dm::tanF(a);
        case B200PT_DM_ASIN: return b200pt_dm::asinF(a);
        case B200PT_DM_ACOS: return b200pt_dm::acosF(a);
        case B200PT_DM_ATAN: return b200pt_dm::atanF(a);
        case B200PT_DM_ATAN2: return b200pt_dm::atan2F(a, b);
        case B200PT_DM_POW: return b200pt_dm::powF(a, b);
        case B200PT_DM_LOG: return b200pt_dm::logF(a);
        default: return b200pt_dm::expF(a);
    }
}
__global__ void __launch_bounds__(256) k_detmath(int fn, const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (fn == B200PT_DM_DIV3) {            // the kernels' vec3 / scalar (device_math.cuh div3), one component per element
        const vec3 q = V3(a[i], a[i], a[i]) / (b ? b[i] : 1.0f);
        out[i] = i % 3 == 0 ? q.x : (i % 3 == 1 ? q.y : q.z);
    } else out[i] = detmathApply(fn, a[i], b ? b[i] : 0.0f);
}
extern "C" {
int b200pt_detmath_eval(b200pt_ctx *c, int fn, const float *a, const float *b, float *out, int n) {
    if (!c || fn < 0 || fn > B200PT_DM_DIV3 || n < 0 || (n > 0 && (!a || !out))) return setError(B200PT_E_INVALID, "b200pt_detmath_eval: bad argument");
    if (n == 0) return B200PT_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    DevBuf<float> da, db, dout;
    CUDA_TRY(da.upload(a, size_t(n), c->stream));
    if (b) CUDA_TRY(db.upload(b, size_t(n), c->stream));
    CUDA_TRY(dout.alloc(size_t(n)));
    k_detmath<<<gridFor(uint64_t(n), 256), 256, 0, c->stream>>>(fn, da.p, b ? db.p : nullptr, dout.p, n);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, dout.p, size_t(n) * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    da.release(); db.release(); dout.release();
    return B200PT_OK;
}
/* the same functions evaluated by the host compilation of the header (no device involved) */
int b200pt_detmath_eval_host(int fn, const float *a, const float *b, float *out, int n) {
    if (fn < 0 || fn > B200PT_DM_DIV3 || n < 0 || (n > 0 && (!a || !out))) return setError(B200PT_E_INVALID, "b200pt_detmath_eval_host: bad argument");
    for (int i = 0; i < n; i++) out[i] = fn == B200PT_DM_DIV3 ? a[i] / (b ? b[i] : 1.0f) : detmathApply(fn, a[i], b ? b[i] : 0.0f);
    return B200PT_OK;
}
int b200pt_guiding_selftest_division(b200pt_ctx *c, float lo, float hi, uint64_t *mismatches, uint64_t *tested) {
    if (!c || !mismatches || !tested) return setError(B200PT_E_INVALID, "b200pt_guiding_selftest_division: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    std::string err;
    unsigned long long m = 0, t = 0;
    int rc = guidingDivisionSelfTest(lo, hi, &m, &t, c->stream, err);
    if (rc != B200PT_OK) return setError(rc, "b200pt_guiding_selftest_division: " + err);
    *mismatches = m; *tested = t;
    return B200PT_OK;
}
int b200pt_guiding_region_count(b200pt_ctx *c, int *count) {
    if (!c || !count) return setError(B200PT_E_INVALID, "b200pt_guiding_region_count: null argument");
    *count = c->guiding.regionCount;
    return B200PT_OK;
}
int b200pt_guiding_get_aabbs(b200pt_ctx *c, b200pt_aabb *out, int n) {
    if (!c || !out || n < 0 || n > c->guiding.regionCount) return setError(B200PT_E_INVALID, "b200pt_guiding_get_aabbs: bad argument");
    memcpy(out, c->guiding.hostAabbs.data(), size_t(n) * sizeof(b200pt_aabb));
    return B200PT_OK;
}
int b200pt_guiding_get_vmms(b200pt_ctx *c, b200pt_vmm_theta *out, int n) {
    if (!c || !out || n < 0 || n > c->guiding.regionCount) return setError(B200PT_E_INVALID, "b200pt_guiding_get_vmms: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(out, c->guiding.vmms, size_t(n) * sizeof(b200pt_vmm_theta), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}
int b200pt_guiding_put_vmms(b200pt_ctx *c, const b200pt_vmm_theta *in, int n) {
    if (!c || !in || n < 0 || n > c->guiding.regionCount) return setError(B200PT_E_INVALID, "b200pt_guiding_put_vmms: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(c->guiding.vmms, in, size_t(n) * sizeof(b200pt_vmm_theta), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}
int64_t b200pt_guiding_sample_capacity(b200pt_ctx *c) { return c ? int64_t(c->numPixels) * B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL : 0; }
int b200pt_guiding_get_samples(b200pt_ctx *c, b200pt_directional_data *out, int64_t n) {
    if (!c || !out || n < 0 || n > b200pt_guiding_sample_capacity(c)) return setError(B200PT_E_INVALID, "b200pt_guiding_get_samples: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(out, c->samples.p, size_t(n) * sizeof(b200pt_directional_data), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}
int b200pt_guiding_get_samples_device(b200pt_ctx *c, void *dst_device, int64_t n) {
    if (!c || !dst_device || n < 0 || n > b200pt_guiding_sample_capacity(c)) return setError(B200PT_E_INVALID, "b200pt_guiding_get_samples_device: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(dst_device, c->samples.p, size_t(n) * sizeof(b200pt_directional_data), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}
int b200pt_guiding_put_samples(b200pt_ctx *c, const b200pt_directional_data *in, int64_t n) {
    if (!c || !in || n < 0 || n > b200pt_guiding_sample_capacity(c)) return setError(B200PT_E_INVALID, "b200pt_guiding_put_samples: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(c->samples.p, in, size_t(n) * sizeof(b200pt_directional_data), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}

int b200pt_set_aovs(b200pt_ctx *c, int enabled) {
    if (!c) return setError(B200PT_E_INVALID, "b200pt_set_aovs: null argument");
    c->aovs = enabled != 0;
    return B200PT_OK;
}
int b200pt_read_aovs(b200pt_ctx *c, float *rgba) {
    if (!c || !rgba) return setError(B200PT_E_INVALID, "b200pt_read_aovs: null argument");
    if (!c->aovs || !c->imgAov.p) return setError(B200PT_E_STATE, "b200pt_read_aovs: enable AOVs (b200pt_set_aovs) and render a frame first");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(rgba, c->imgAov.p, size_t(c->numPixels) * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}

// ---- checkpoint / resume -------------------------------------------------------------------------------------------
static const char kStateMagic[8] = {'B', '2', 'P', 'T', 'S', 'T', '0', '2'};
static const int32_t kStateVersion = 2;      // little-endian raw structs; bump on any layout change of GMix / b200pt_vmm_theta / the headers

int b200pt_save_state(b200pt_ctx *c, const char *path) {
    if (!c || !path) return setError(B200PT_E_INVALID, "b200pt_save_state: null argument");
    if (!c->hasScene) return setError(B200PT_E_STATE, "b200pt_save_state: set_scene must be called first");
    CUDA_TRY(cudaSetDevice(c->device));
    // written to <path>.tmp, flushed to disk and renamed: a crash in the middle leaves the previous checkpoint intact
    const std::string tmp = std::string(path) + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return setError(B200PT_E_IO, std::string("b200pt_save_state: cannot open ") + tmp);
    auto fail = [&](int code, const std::string &msg) { fclose(f); remove(tmp.c_str()); return setError(code, msg); };
    const int32_t hdr[5] = {kStateVersion, c->width, c->height, c->icSize, c->guidingSplits};
    const size_t N = size_t(c->numPixels), S = size_t(c->icSize);
    std::vector<float4> img(N);
    std::vector<b200pt_cache_data> icd(S);
    std::vector<b200pt_sphere> ics(S);
    b200pt_cache_header ich;
    bool ok = fwrite(kStateMagic, 8, 1, f) == 1 && fwrite(hdr, sizeof(hdr), 1, f) == 1;
    for (int which = 0; which < 3 && ok; which++) {
        if (cudaMemcpy(img.data(), imagePtr(c, which), N * sizeof(float4), cudaMemcpyDeviceToHost) != cudaSuccess) return fail(B200PT_E_CUDA, "b200pt_save_state: image read-back failed");
        ok = fwrite(img.data(), sizeof(float4), N, f) == N;
    }
    int rc = b200pt_ic_get(c, &ich, S ? icd.data() : nullptr, S ? ics.data() : nullptr, int(S));
    if (rc != B200PT_OK) { fclose(f); remove(tmp.c_str()); return rc; }
    ok = ok && fwrite(&ich, sizeof(ich), 1, f) == 1 && fwrite(icd.data(), sizeof(b200pt_cache_data), S, f) == S && fwrite(ics.data(), sizeof(b200pt_sphere), S, f) == S;
    if (ok) { rc = c->guiding.save(f, c->stream); if (rc != B200PT_OK) return fail(rc, "b200pt_save_state: " + c->guiding.error); }
    ok = ok && fflush(f) == 0 && fsync(fileno(f)) == 0;
    ok = (fclose(f) == 0) && ok;
    if (!ok) { remove(tmp.c_str()); return setError(B200PT_E_IO, std::string("b200pt_save_state: write failed: ") + tmp); }
    if (rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return setError(B200PT_E_IO, std::string("b200pt_save_state: cannot rename to ") + path); }
    return B200PT_OK;
}

int b200pt_load_state(b200pt_ctx *c, const char *path) {
    if (!c || !path) return setError(B200PT_E_INVALID, "b200pt_load_state: null argument");
    if (!c->hasScene) return setError(B200PT_E_STATE, "b200pt_load_state: set_scene must be called first");
    CUDA_TRY(cudaSetDevice(c->device));
    FILE *f = fopen(path, "rb");
    if (!f) return setError(B200PT_E_IO, std::string("b200pt_load_state: cannot open ") + path);
    auto fail = [&](int code, const std::string &msg) { fclose(f); return setError(code, msg); };
    char magic[8]; int32_t hdr[5];
    if (fread(magic, 8, 1, f) != 1 || memcmp(magic, kStateMagic, 8) != 0 || fread(hdr, sizeof(hdr), 1, f) != 1) return fail(B200PT_E_IO, "b200pt_load_state: not a b200pt checkpoint");
    if (hdr[0] != kStateVersion) return fail(B200PT_E_INVALID, "b200pt_load_state: checkpoint format version " + std::to_string(hdr[0]) + " is not supported (this build reads version " + std::to_string(kStateVersion) + ")");
    if (hdr[1] != c->width || hdr[2] != c->height || hdr[3] != c->icSize || hdr[4] != c->guidingSplits)
        return fail(B200PT_E_INVALID, "b200pt_load_state: the checkpoint was written by a context of a different size (width, height, ic_size, guiding_splits)");
    // The WHOLE file is read into host memory and validated before anything is uploaded: a truncated or corrupt file
    // leaves the context untouched, and no out-of-range slot, link or component count ever reaches the device.
    const size_t N = size_t(c->numPixels), S = size_t(c->icSize);
    std::vector<float4> img[3];
    for (int which = 0; which < 3; which++) {
        img[which].resize(N);
        if (fread(img[which].data(), sizeof(float4), N, f) != N) return fail(B200PT_E_IO, "b200pt_load_state: truncated checkpoint");
    }
    std::vector<b200pt_cache_data> icd(S);
    std::vector<b200pt_sphere> ics(S);
    b200pt_cache_header ich;
    if (fread(&ich, sizeof(ich), 1, f) != 1 || fread(icd.data(), sizeof(b200pt_cache_data), S, f) != S || fread(ics.data(), sizeof(b200pt_sphere), S, f) != S)
        return fail(B200PT_E_IO, "b200pt_load_state: truncated checkpoint");
    if (ich.maxCaches != uint32_t(c->icSize) || ich.nextCacheSlot > uint32_t(c->icSize) + 1u || ich.nextUpdateSlot > uint32_t(c->icSize) + 1u)      // (+1: quirk 11, the reference's off-by-one)
        return fail(B200PT_E_INVALID, "b200pt_load_state: irradiance-cache header out of range");
    GuidingCheckpoint gck;
    int rc = c->guiding.readCheckpoint(f, gck);
    if (rc != B200PT_OK) return fail(rc, "b200pt_load_state: " + c->guiding.error);
    fclose(f);
    for (int which = 0; which < 3; which++)
        if (cudaMemcpy(imagePtr(c, which), img[which].data(), N * sizeof(float4), cudaMemcpyHostToDevice) != cudaSuccess) return setError(B200PT_E_CUDA, "b200pt_load_state: image upload failed");
    rc = b200pt_ic_put(c, &ich, S ? icd.data() : nullptr, S ? ics.data() : nullptr, int(S));
    if (rc != B200PT_OK) return rc;
    rc = c->guiding.applyCheckpoint(gck, c->stream);
    if (rc != B200PT_OK) return setError(rc, "b200pt_load_state: " + c->guiding.error);
    return B200PT_OK;
}

// ---- multi-GPU -----------------------------------------------------------------------------------------------------
#define NCCL_TRY(expr)                                                                                     \
    do {                                                                                                   \
        ncclResult_t _r = (expr);                                                                          \
        if (_r != ncclSuccess) return setError(B200PT_E_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
    } while (0)

int b200pt_comm_unique_id(char id[B200PT_COMM_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) <= B200PT_COMM_ID_BYTES, "ncclUniqueId does not fit the ABI's id buffer");
    if (!id) return setError(B200PT_E_INVALID, "b200pt_comm_unique_id: null argument");
    if (!g_nccl.load()) return setError(B200PT_E_STATE, "b200pt_comm_unique_id: libnccl.so.2 not found");
    ncclUniqueId uid;
    NCCL_TRY(g_nccl.GetUniqueId(&uid));
    memset(id, 0, B200PT_COMM_ID_BYTES);
    memcpy(id, &uid, sizeof(uid));
    return B200PT_OK;
}
int b200pt_comm_init(b200pt_ctx *c, const char id[B200PT_COMM_ID_BYTES], int rank, int nranks) {
    if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks) return setError(B200PT_E_INVALID, "b200pt_comm_init: bad argument");
    if (!g_nccl.load()) return setError(B200PT_E_STATE, "b200pt_comm_init: libnccl.so.2 not found");
    if (c->rc.comm) return setError(B200PT_E_STATE, "b200pt_comm_init: communicator already initialised");
    CUDA_TRY(cudaSetDevice(c->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    NCCL_TRY(g_nccl.CommInitRank(&c->rc.comm, nranks, uid, rank));
    c->rc.rank = rank; c->rc.nranks = nranks;
    if (nranks > 1 && c->guiding.ready) {
        // one-time work of the region-sharded refit, done here instead of inside the first update: the sorted-sample buffers
        // for this context's W*H*16 records, their CUDA-IPC mappings on every peer (collective), NCCL's lazy connection set-up
        int rc2 = c->guiding.ensureCapacity(int64_t(c->numPixels) * B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL);
        if (rc2 == B200PT_OK) rc2 = c->guiding.setupPeers(c->rc, c->stream);
        if (rc2 != B200PT_OK) return setError(rc2, "b200pt_comm_init: " + c->guiding.error);
    }
    return B200PT_OK;
}
int b200pt_comm_destroy(b200pt_ctx *c) {
    if (!c) return setError(B200PT_E_INVALID, "b200pt_comm_destroy: null argument");
    if (c->rc.comm) {
        CUDA_TRY(cudaSetDevice(c->device)); CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->guiding.closePeers(); c->guiding.peerTried = false; c->rc.peerMode = false;
        NCCL_TRY(g_nccl.CommDestroy(c->rc.comm)); c->rc.comm = nullptr;
    }
    c->rc.rank = 0; c->rc.nranks = 1;
    return B200PT_OK;
}

}  // extern "C"
__global__ void __launch_bounds__(256) k_comm_scale(const float4 *__restrict__ img, float4 *__restrict__ out, int n, float w, float *count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count = w;
    if (i < n) { const float4 v = img[i]; out[i] = make_float4(v.x * w, v.y * w, v.z * w, v.w * w); }
}
__global__ void __launch_bounds__(256) k_comm_normalise(const float4 *__restrict__ acc, float4 *__restrict__ img, int n, const float *count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float t = *count;
    if (i < n) { const float4 v = acc[i]; img[i] = make_float4(v.x / t, v.y / t, v.z / t, v.w / t); }
}
extern "C" {

int b200pt_reduce_image(b200pt_ctx *c, int which, int frames_local) {
    if (!c || !imagePtr(c, which) || frames_local < 0) return setError(B200PT_E_INVALID, "b200pt_reduce_image: bad argument");
    if (!c->rc.comm) return setError(B200PT_E_STATE, "b200pt_reduce_image: b200pt_comm_init must be called first");
    CUDA_TRY(cudaSetDevice(c->device));
    const int n = c->numPixels;
    CUDA_TRY(c->commImage.alloc(size_t(n))); CUDA_TRY(c->commCount.alloc(1));
    k_comm_scale<<<gridFor(uint64_t(n), 256), 256, 0, c->stream>>>(imagePtr(c, which), c->commImage.p, n, float(frames_local), c->commCount.p);
    NCCL_TRY(g_nccl.GroupStart());
    NCCL_TRY(g_nccl.AllReduce(c->commImage.p, c->commImage.p, size_t(n) * 4, ncclFloat, ncclSum, c->rc.comm, c->stream));
    NCCL_TRY(g_nccl.AllReduce(c->commCount.p, c->commCount.p, 1, ncclFloat, ncclSum, c->rc.comm, c->stream));
    NCCL_TRY(g_nccl.GroupEnd());
    k_comm_normalise<<<gridFor(uint64_t(n), 256), 256, 0, c->stream>>>(c->commImage.p, imagePtr(c, which), n, c->commCount.p);
    c->stats.kernel_launches += 2;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaGetLastError());
    return B200PT_OK;
}
int b200pt_allgather_samples(b200pt_ctx *c, int64_t *total_out) {
    if (!c) return setError(B200PT_E_INVALID, "b200pt_allgather_samples: null argument");
    if (!c->rc.comm) return setError(B200PT_E_STATE, "b200pt_allgather_samples: b200pt_comm_init must be called first");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t per = size_t(c->numPixels) * B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL;      // every rank renders the same resolution
    CUDA_TRY(c->commSamples.alloc(per * size_t(c->rc.nranks)));
    NCCL_TRY(g_nccl.AllGather(c->samples.p, c->commSamples.p, per * sizeof(b200pt_directional_data), ncclChar, c->rc.comm, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (total_out) *total_out = int64_t(per * size_t(c->rc.nranks));
    return B200PT_OK;
}
int b200pt_guiding_update_all_ranks(b200pt_ctx *c, const b200pt_guiding_params *params) {
    if (!c || !params) return setError(B200PT_E_INVALID, "b200pt_guiding_update_all_ranks: null argument");
    if (!c->guiding.ready) return setError(B200PT_E_STATE, "b200pt_guiding_update_all_ranks: set_scene must be called first");
    if (!c->rc.comm) return setError(B200PT_E_STATE, "b200pt_guiding_update_all_ranks: b200pt_comm_init must be called first");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = c->guiding.update(c->samples.p, int64_t(c->numPixels) * B200PT_MAX_DIRECTIONAL_DATA_PER_PIXEL, *params, c->stream, &c->stats, &c->rc);
    if (rc != B200PT_OK) return setError(rc, "b200pt_guiding_update_all_ranks: " + c->guiding.error);
    return B200PT_OK;
}
int b200pt_guiding_update_all_ranks_device(b200pt_ctx *c, const b200pt_guiding_params *params, const void *samples_device, int64_t n) {
    if (!c || !params || n < 0 || (n > 0 && !samples_device)) return setError(B200PT_E_INVALID, "b200pt_guiding_update_all_ranks_device: bad argument");
    if (!c->guiding.ready) return setError(B200PT_E_STATE, "b200pt_guiding_update_all_ranks_device: set_scene must be called first");
    if (!c->rc.comm) return setError(B200PT_E_STATE, "b200pt_guiding_update_all_ranks_device: b200pt_comm_init must be called first");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = c->guiding.update(static_cast<b200pt_directional_data *>(const_cast<void *>(samples_device)), n, *params, c->stream, &c->stats, &c->rc);
    if (rc != B200PT_OK) return setError(rc, "b200pt_guiding_update_all_ranks_device: " + c->guiding.error);
    return B200PT_OK;
}
int b200pt_guiding_plan_debug(b200pt_ctx *c, const uint32_t *counts, int nranks, int rank, int peer_mode, uint8_t *owner, uint32_t *region_begin, uint32_t *region_len,
                              uint32_t *src_start, uint32_t *active, uint32_t summary[6], uint32_t *segments) {
    if (!c || !counts || !owner || !region_begin || !region_len || !src_start || !active || !summary || !segments) return setError(B200PT_E_INVALID, "b200pt_guiding_plan_debug: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = c->guiding.planDebug(counts, nranks, rank, peer_mode, owner, region_begin, region_len, src_start, active, summary, segments, c->stream);
    if (rc != B200PT_OK) return setError(rc, "b200pt_guiding_plan_debug: " + c->guiding.error);
    return B200PT_OK;
}
int b200pt_comm_exchange_mode(b200pt_ctx *c) { return c && c->rc.comm ? (c->rc.peerMode ? 2 : 1) : 0; }

// ---- irradiance cache parity hooks -------------------------------------------------------------------------------
int b200pt_ic_get(b200pt_ctx *c, b200pt_cache_header *hdr, b200pt_cache_data *data, b200pt_sphere *spheres, int n) {
    if (!c || n < 0 || n > c->icSize) return setError(B200PT_E_INVALID, "b200pt_ic_get: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    if (hdr) CUDA_TRY(cudaMemcpyAsync(hdr, c->icHeader.p, sizeof(*hdr), cudaMemcpyDeviceToHost, c->stream));
    if (data && n) CUDA_TRY(cudaMemcpyAsync(data, c->icData.p, size_t(n) * sizeof(*data), cudaMemcpyDeviceToHost, c->stream));
    if (spheres && n) CUDA_TRY(cudaMemcpyAsync(spheres, c->icSpheres.p, size_t(n) * sizeof(*spheres), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}
int b200pt_ic_put(b200pt_ctx *c, const b200pt_cache_header *hdr, const b200pt_cache_data *data, const b200pt_sphere *spheres, int n) {
    if (!c || n < 0 || n > c->icSize) return setError(B200PT_E_INVALID, "b200pt_ic_put: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    if (hdr) CUDA_TRY(cudaMemcpyAsync(c->icHeader.p, hdr, sizeof(*hdr), cudaMemcpyHostToDevice, c->stream));
    if (data && n) CUDA_TRY(cudaMemcpyAsync(c->icData.p, data, size_t(n) * sizeof(*data), cudaMemcpyHostToDevice, c->stream));
    if (spheres && n) CUDA_TRY(cudaMemcpyAsync(c->icSpheres.p, spheres, size_t(n) * sizeof(*spheres), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}

// ---- host side ---------------------------------------------------------------------------------------------------
void b200pt_default_push_constants(b200pt_push_constants *pc) {   // src/RayTracingApp.h:116-165
    if (!pc) return;
    memset(pc, 0, sizeof(*pc));
    pc->previousFrames = 0xFFFFFFFFu;
    pc->maxDepth = 30; pc->maxFollowDiscrete = 3; pc->samplesPerPixel = 1; pc->enableNEE = 1; pc->numNEE = 1;
    pc->usePowerHeuristic = 1;
    pc->irradianceA = 0.2f; pc->irradianceUpdateProb = 0.00001f; pc->irradianceCreateProb = 0.0005f; pc->irradianceVisualizationScale = 1.0f;
    pc->irradianceGradientsMaxLength = 5; pc->irradianceNumNEE = 1; pc->irradianceCacheMinRadius = 0.1f;
    pc->adrrsS = 5; pc->adrrsSplit = 1; pc->guidingProb = 0.5f; pc->guidingVisuScale = 0.5f; pc->guidingVisuMax = 1.0f;
    pc->useParallaxCompensation = 1; pc->guidingVisuPhiScale = 0.004f; pc->guidingVisuThetaScale = 0.01f;
    pc->guidingPiPShowSpheres = 1; pc->guidingPiPSize = 0.3f;
}

struct b200pt_scene { Scene scene; };

int b200pt_scene_load(const char *path, b200pt_scene **out) {
    if (!path || !out) return setError(B200PT_E_INVALID, "b200pt_scene_load: null argument");
    b200pt_scene *s = new b200pt_scene();
    try { s->scene.loadFile(path); }
    catch (const std::exception &e) { delete s; return setError(B200PT_E_IO, std::string("b200pt_scene_load: ") + e.what()); }
    *out = s;
    return B200PT_OK;
}
int b200pt_scene_free(b200pt_scene *s) { delete s; return B200PT_OK; }
int b200pt_scene_get_desc(const b200pt_scene *s, b200pt_scene_desc *out) {
    if (!s || !out) return setError(B200PT_E_INVALID, "b200pt_scene_get_desc: null argument");
    s->scene.fillDesc(out);
    return B200PT_OK;
}
int b200pt_scene_get_camera(const b200pt_scene *s, float origin[3], float target[3], float up[3], float *vfov) {
    if (!s) return setError(B200PT_E_INVALID, "b200pt_scene_get_camera: null argument");
    if (origin) memcpy(origin, s->scene.origin, 12);
    if (target) memcpy(target, s->scene.target, 12);
    if (up) memcpy(up, s->scene.upDir, 12);
    if (vfov) *vfov = s->scene.vfov;
    return B200PT_OK;
}

// CameraController::lookAt → quatLookAt(viewDirection, up) (conjugated) → getViewMatrix = translate(mat4_cast(q), -pos);
// getProjMatrix = glm::perspective(radians(vfov), aspect, 0.1, 1000) with [1][1] *= -1   (src/CameraController.cpp:76-107).
// The rotation is built directly as the matrix quatLookAtRH forms (no quaternion round trip).
void b200pt_camera_matrices(const float origin[3], const float target[3], const float up[3], float vfov_deg, float aspect, float view[16], float proj[16]) {
    float dir[3] = {target[0] - origin[0], target[1] - origin[1], target[2] - origin[2]};
    float l = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    for (float &v : dir) v /= l;
    float c2[3] = {-dir[0], -dir[1], -dir[2]};                               // Result[2] = -direction
    float right[3] = {up[1] * c2[2] - c2[1] * up[2], up[2] * c2[0] - c2[2] * up[0], up[0] * c2[1] - c2[0] * up[1]};   // cross(up, Result[2])
    float rl = 1.0f / sqrtf(fmaxf(0.00001f, right[0] * right[0] + right[1] * right[1] + right[2] * right[2]));
    float c0[3] = {right[0] * rl, right[1] * rl, right[2] * rl};
    float c1[3] = {c2[1] * c0[2] - c0[1] * c2[2], c2[2] * c0[0] - c0[2] * c2[0], c2[0] * c0[1] - c0[0] * c2[1]};       // cross(Result[2], Result[0])
    // view rotation = transpose of [c0 c1 c2]; column-major storage
    float R[16];
    mat4Identity(R);
    for (int a = 0; a < 3; a++) { R[a * 4 + 0] = c0[a]; R[a * 4 + 1] = c1[a]; R[a * 4 + 2] = c2[a]; }
    for (int r = 0; r < 4; r++) R[12 + r] = R[0 + r] * -origin[0] + R[4 + r] * -origin[1] + R[8 + r] * -origin[2] + R[12 + r];
    memcpy(view, R, 64);
    float fovy = vfov_deg * 0.01745329251994329576923690768489f;
    float tanHalf = tanf(fovy / 2.0f);
    const float zn = 0.1f, zf = 1000.0f;
    float P[16];
    memset(P, 0, sizeof(P));
    P[0] = 1.0f / (aspect * tanHalf);
    P[5] = -(1.0f / tanHalf);
    P[10] = zf / (zn - zf);              // GLM_FORCE_DEPTH_ZERO_TO_ONE (src/Model.h:13); irrelevant for ray directions
    P[11] = -1.0f;
    P[14] = -(zf * zn) / (zf - zn);
    memcpy(proj, P, 64);
}

int b200pt_mat4_inverse(const float m[16], float out[16]) { return (m && out && mat4Inverse(m, out)) ? 1 : 0; }

int b200pt_write_exr(const char *path, const float *rgba, int width, int height) {
    if (!path || !rgba || width <= 0 || height <= 0) return setError(B200PT_E_INVALID, "b200pt_write_exr: bad argument");
    try { writeExrRGB(path, rgba, width, height); }
    catch (const std::exception &e) { return setError(B200PT_E_IO, std::string("b200pt_write_exr: ") + e.what()); }
    return B200PT_OK;
}
int b200pt_read_exr(const char *path, float **rgba_out, int *width, int *height) {
    if (!path || !rgba_out || !width || !height) return setError(B200PT_E_INVALID, "b200pt_read_exr: bad argument");
    try {
        std::vector<float> px;
        readExrRGBA(path, px, *width, *height, true);
        *rgba_out = static_cast<float *>(malloc(px.size() * sizeof(float)));
        memcpy(*rgba_out, px.data(), px.size() * sizeof(float));
    } catch (const std::exception &e) { return setError(B200PT_E_IO, std::string("b200pt_read_exr: ") + e.what()); }
    return B200PT_OK;
}
int b200pt_read_image_file(const char *path, uint8_t **rgba_out, int *width, int *height) {
    if (!path || !rgba_out || !width || !height) return setError(B200PT_E_INVALID, "b200pt_read_image_file: bad argument");
    try {
        std::vector<uint8_t> px;
        decodeImageFile(path, *width, *height, px);
        *rgba_out = static_cast<uint8_t *>(malloc(px.size()));
        memcpy(*rgba_out, px.data(), px.size());
    } catch (const std::exception &e) { return setError(B200PT_E_IO, std::string("b200pt_read_image_file: ") + e.what()); }
    return B200PT_OK;
}
void b200pt_free(void *p) { free(p); }

}  // extern "C"
