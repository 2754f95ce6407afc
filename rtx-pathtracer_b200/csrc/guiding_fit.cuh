// Guiding state on the device: region tree + per-region vMF mixtures (PathGuiding, src/PathGuiding.{h,cpp}) and the
// scratch buffers of the on-device sample sort (SampleCollector::getSortedData, src/SampleCollector.cpp:76-131).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/b200pt.h"
#include "comm.cuh"

namespace b200pt {

struct GMix;
struct GSegment;
struct GPlanSummary;
struct GPass;

struct GuidingCheckpoint {             // host copy of a checkpoint's guiding section, validated before it is applied
    int regionCount = 0;
    bool firstFit = true, hasSpawns = false;
    b200pt_guiding_params params{};
    std::vector<b200pt_aabb> aabbs;
    std::vector<int32_t> spawnFirst, spawnNext;
    std::vector<char> mixes;
    std::vector<b200pt_vmm_theta> vmms;
};

struct GuidingState {
    bool ready = false;
    bool firstFit = true;                  // PathGuiding::firstFit (global: cleared by the first update)
    int regionCount = 0;                   // grows when regions are split adaptively (PathGuiding::splitRegion)
    int maxRegions = 0;                    // capacity of every per-region buffer
    std::vector<b200pt_aabb> hostAabbs;
    std::vector<int32_t> hostSpawnFirst, hostSpawnNext;
    int32_t *spawnFirst = nullptr, *spawnNext = nullptr;   // device; consulted by the tracer once hasSpawns
    bool hasSpawns = false;
    int2 *splitPairs = nullptr;            // device scratch: (source region, new region) of one update's splits
    int splits = 0;
    b200pt_aabb *aabbs = nullptr;          // device, binding 15
    b200pt_aabb *levelAabbs = nullptr;     // device: all 2^(splits+1)-1 boxes of the halving tree, level by level
    float4 *levelSplits = nullptr;         // device: per inner node of that tree {split axis (int bits), left child's max, right child's min on that axis, -}
    b200pt_vmm_theta *vmms = nullptr;      // device, binding 16
    GMix *mixes = nullptr;                 // device: lightpmm PMM + PMM_ExtraData per region
    // sort scratch
    uint32_t *regionTotal = nullptr, *regionOffset = nullptr, *activeRegions = nullptr, *tileCounts = nullptr, *srcIndex = nullptr;
    float4 *dirw = nullptr;                // sorted samples: direction (after preFit) + weight
    float2 *pdfDist = nullptr;             // sorted samples: pdf, distance (after preFit)
    int64_t capacity = 0;
    uint32_t lastValidSamples = 0;
    unsigned long long *devScalars = nullptr, *hostScalars = nullptr;   // [0] EM sample-iterations
    // plan of an update (k_plan): ownership, layouts, copy segments; see guiding_fit.cu
    uint32_t *allCounts = nullptr, *srcStart = nullptr;                 // [ranks][maxRegions]
    uint32_t *regionBegin = nullptr, *regionLen = nullptr, *totalAll = nullptr;
    uint8_t *owner = nullptr;
    uint32_t *regionSlot = nullptr;                                     // position of a region among this rank's regions
    // work-sharing fit kernel: pass descriptors, per-chunk partial sums, {next region, regions finished}
    GPass *passes = nullptr;
    float *partials = nullptr;
    int64_t partialRows = 0;
    uint32_t *fitControl = nullptr;
    int sharedGrid = 0;
    GSegment *segments = nullptr;                                       // [maxRegions * ranks]
    GPlanSummary *planDev = nullptr, *planHost = nullptr;               // device / pinned host copy
    int planRanks = 0;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // start, sorted, planned, exchanged, fitted, gathered
    // region-sharded refit across ranks
    float4 *fitDirw = nullptr, *stageDirw = nullptr;                    // owned regions, contiguous / NCCL staging
    float2 *fitPdfDist = nullptr, *stagePdfDist = nullptr;
    int64_t fitCapacity = 0, stageCapacity = 0;
    GMix *gatherMix = nullptr;
    b200pt_vmm_theta *gatherVmm = nullptr;
    float *barrierWord = nullptr;
    const float4 *peerDirw[B200PT_MAX_RANKS] = {};
    const float2 *peerPdfDist[B200PT_MAX_RANKS] = {};
    int64_t peerCapacity = -1;                                          // capacity the peer mappings were made for
    bool peerTried = false;
    int peerSelf = -1;                                                  // own entries alias dirw / pdfDist: never IPC-closed
    b200pt_guiding_params lastParams{};
    int summationOrder = 1;                // 0: the reference's sequential float sums (strict), 1: block-parallel sums (reordered, the fast default)
    std::string error;

    int init(int splits, const float sceneMin[3], const float sceneMax[3], cudaStream_t stream);
    int reset(const b200pt_guiding_params &params, cudaStream_t stream);
    int ensureCapacity(int64_t numSamples);
    // rc == nullptr or one rank: PathGuiding::update on this GPU's records.  Several ranks: every rank passes its own
    // records; regions are fitted by their owner on the records of ALL ranks and the results are shared (identical
    // mixtures everywhere, equal to a single-GPU update on the concatenation of the ranks' buffers in rank order).
    int update(b200pt_directional_data *samples, int64_t numSamples, const b200pt_guiding_params &params, cudaStream_t stream, b200pt_stats *stats,
               RankComm *rc = nullptr);
    int ensurePlan(int ranks);
    int planDebug(const uint32_t *counts, int N, int me, int peerMode, uint8_t *ownerOut, uint32_t *regionBeginOut, uint32_t *regionLenOut, uint32_t *srcStartOut,
                  uint32_t *activeOut, uint32_t *summaryOut, uint32_t *segmentsOut, cudaStream_t stream);
    int setupPeers(RankComm &rc, cudaStream_t stream);
    void closePeers();
    int splitRegions(const b200pt_guiding_params &params, cudaStream_t stream);
    int save(FILE *f, cudaStream_t stream);          // checkpoint: regions, spawn chains, mixtures, packed VMMs
    int readCheckpoint(FILE *f, GuidingCheckpoint &ck);          // read + validate, host only
    int applyCheckpoint(const GuidingCheckpoint &ck, cudaStream_t stream);
    int getState(int region, float scalars5[5], float perComponent[14 * 16], cudaStream_t stream);
    int getSorted(b200pt_directional_data *out, uint32_t *offsets, const b200pt_directional_data *rawDevice, cudaStream_t stream);
    void release();
};

int guidingDivisionSelfTest(float lo, float hi, unsigned long long *mismatchesOut, unsigned long long *testedOut, cudaStream_t stream, std::string &error);
int guidingFastExp(const float *hostIn, float *hostOut, int n, cudaStream_t stream, std::string &error);

}  // namespace b200pt
