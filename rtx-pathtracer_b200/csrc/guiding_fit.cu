// Guiding update on the device: PathGuiding::update(SampleCollector) (src/PathGuiding.cpp:276-312) without the
// 0.6-5.3 GB read-back and the CPU sort/fit of the reference.
//
//   1. sort      SampleCollector::getSortedData (src/SampleCollector.cpp:76-131): stable counting sort of the W*H*16
//                DirectionalData records by region id — k_sort_count (per-warp histograms with __match_any_sync),
//                k_sort_scan_tiles / k_sort_scan_regions (offsets), k_sort_scatter.  The scatter also applies the
//                sample part of PathGuiding::preFit (:386-402, re-anchoring at the region centre) and writes the
//                samples as SoA (float4 dir+weight | float2 pdf+distance): the EM passes then stream 16 B per sample.
//   2. fit       k_guiding_update: ONE thread block per non-empty region runs the whole updateRegion sequence
//                (fit | updateFit, mergeAll, chi^2 / covariance statistics, splitAll with masked fits, distance update)
//                on a shared-memory copy of the mixture; every pass over the samples is a block-wide loop with a
//                warp-shuffle + shared-memory reduction, the per-component logic runs on one thread (guiding_math.cuh).
//   3. pack      the region's VMM_Theta (binding 16) is rewritten in place — no host round trip.
// Region creation follows PathGuiding::createRegions (src/PathGuiding.cpp:81-104) with Aabb::addEpsilon / splitAabb
// (src/Shapes.h:32-53).
#include "guiding_fit.cuh"
#include "guiding_math.cuh"
#include "comm.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

namespace b200pt {

#ifndef G_BLOCK
#define G_BLOCK 512                 // threads of the per-region block
#endif
#ifndef G_BLOCKS_PER_SM
#define G_BLOCKS_PER_SM 1
#endif
#define G_WARPS (G_BLOCK / 32)
// The lobe parameters live in shared memory (GPacked) and are re-read for every sample with 128-/64-bit broadcast
// loads.  Without this compiler barrier nvcc hoists all 96 loop-invariant loads into registers and spills the
// accumulators to local memory instead.
#define G_NO_HOIST() asm volatile("" ::: "memory")
#define SORT_TILE 2048              // records per warp in the counting sort
#define SORT_WARPS 8                // warps per block in the counting sort (at most; GuidingState::update uses fewer when R is large)
#define G_PLAN_MAX_REGIONS 4096     // what k_plan's shared arrays hold = the largest maxRegions

static inline float __int_as_float_host(int i) { float f; memcpy(&f, &i, sizeof(f)); return f; }
static void splitAabb(const b200pt_aabb &b, b200pt_aabb &l, b200pt_aabb &r) {   // src/Shapes.h:32-46, axis = largest
    float size[3] = {b.max[0] - b.min[0], b.max[1] - b.min[1], b.max[2] - b.min[2]};
    int axis = size[0] > size[1] ? (size[0] > size[2] ? 0 : 2) : (size[1] > size[2] ? 1 : 2);
    l = b; r = b;
    l.max[axis] -= 0.5f * size[axis];
    r.min[axis] += 0.5f * size[axis];
}

// ---- sort ------------------------------------------------------------------------------------------------------------
// warp-level stable ranking: lanes holding the same key get consecutive ranks in lane order; `hist` is the warp's
// running histogram in shared memory (R + 1 bins, bin R = INVALID)
__device__ __forceinline__ uint32_t warpRank(uint32_t *hist, uint32_t key, bool valid, unsigned lane) {
    const unsigned peers = __match_any_sync(0xffffffffu, valid ? key : 0xffffffffu);
    uint32_t base = 0;
    const int leader = __ffs(peers) - 1;
    if (valid && int(lane) == leader) { base = hist[key]; hist[key] = base + __popc(peers); }
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(peers & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(SORT_WARPS * 32) k_sort_count(const b200pt_directional_data *__restrict__ recs, uint64_t n, uint32_t R,
                                                                uint32_t *__restrict__ tileCounts /* [tiles][R] */, uint32_t numTiles) {
    extern __shared__ uint32_t hist[];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x * (blockDim.x >> 5) + warp;      // 8 warps per block, fewer when R histograms of 8 warps exceed 48 KB
    uint32_t *h = hist + warp * R;
    for (uint32_t i = lane; i < R; i += 32) h[i] = 0;
    __syncwarp();
    if (tile < numTiles) {
        const uint64_t begin = uint64_t(tile) * SORT_TILE;
        for (uint32_t k = 0; k < SORT_TILE; k += 32) {
            const uint64_t i = begin + k + lane;
            const uint32_t key = i < n ? recs[i].flags : 0xffffffffu;
            warpRank(h, key, key < R, lane);
        }
        __syncwarp();
        for (uint32_t i = lane; i < R; i += 32) tileCounts[uint64_t(tile) * R + i] = h[i];
    }
}

// per region: exclusive scan of its counts over the tiles (in place) and the region total
__global__ void __launch_bounds__(256) k_sort_scan_tiles(uint32_t *tileCounts, uint32_t numTiles, uint32_t R, uint32_t *regionTotal) {
    __shared__ uint32_t warpSum[8];
    __shared__ uint32_t running;
    const uint32_t r = blockIdx.x;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (uint32_t base = 0; base < numTiles; base += 256) {
        const uint32_t t = base + threadIdx.x;
        const uint32_t v = t < numTiles ? tileCounts[uint64_t(t) * R + r] : 0u;
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (int(lane) >= o) incl += u; }
        if (lane == 31) warpSum[warp] = incl;
        __syncthreads();
        uint32_t before = running;
        for (unsigned w = 0; w < warp; w++) before += warpSum[w];
        if (t < numTiles) tileCounts[uint64_t(t) * R + r] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 255) running = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) regionTotal[r] = running;
}

// ---- plan: who fits which region, and where every record goes -------------------------------------------------------
// One block.  Input: the per-region counts of every rank (allCounts[s][g]; a single-GPU update is the case N = 1).
//   * owner[g]: longest-processing-time-first over the regions sorted by total count (ties: lower region id), so every
//     rank fits about total/N samples whatever the distribution over regions; deterministic, and identical on every
//     rank because all of them see the same counts.
//   * every rank sorts its own records by (owner, region): the records for rank d are ONE contiguous slice of its
//     buffer (srcStart[s][g] = where region g starts in rank s's sorted buffer).
//   * the owner lays its regions out contiguously, each region as the concatenation of the ranks' records in rank
//     order (= the order of a single-GPU sort of the concatenated buffers): regionBegin / regionCount, and one copy
//     segment per (owned region, source rank) for k_pull.
//   * activeRegions: the owned non-empty regions, LARGEST FIRST, so the longest fits start in the first wave.
// (replaces SampleCollector::getSortedData's offset computation, src/SampleCollector.cpp:100-126)
struct GSegment { uint32_t src, srcOff, dstOff, len; };
struct GPlanSummary {
    uint32_t numActive, numOwned, numSegments, localValid;
    unsigned long long ownedSamples, totalSamples;
    uint32_t sendStart[B200PT_MAX_RANKS + 1];    // slice of rank d inside this rank's sorted buffer
    uint32_t recvCount[B200PT_MAX_RANKS];        // records of rank s that belong to regions owned here
    uint32_t stageStart[B200PT_MAX_RANKS];       // NCCL mode: where rank s's slice is staged
};

__global__ void __launch_bounds__(1024) k_plan(const uint32_t *__restrict__ allCounts, int N, int me, uint32_t R, uint32_t stride, int peerMode,
                                               uint32_t *srcStart, uint32_t *regionBegin, uint32_t *regionCount, uint32_t *regionOffset,
                                               uint32_t *activeRegions, uint32_t *totalAll, uint8_t *ownerOut, uint32_t *regionSlot, GSegment *segments, GPlanSummary *summary) {
    __shared__ uint32_t total[G_PLAN_MAX_REGIONS];
    __shared__ uint16_t order[G_PLAN_MAX_REGIONS], binOf[G_PLAN_MAX_REGIONS], regionOfBin[G_PLAN_MAX_REGIONS];
    __shared__ uint8_t own[G_PLAN_MAX_REGIONS];
    __shared__ uint32_t firstBin[B200PT_MAX_RANKS + 1], sliceStart[B200PT_MAX_RANKS], stageStart[B200PT_MAX_RANKS], recvCount[B200PT_MAX_RANKS];
    const unsigned t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    // (per-region work: thread t takes regions t, t + 1024, ...)
    for (uint32_t g = t; g < R; g += 1024u) {
        uint32_t v = 0;
        for (int s = 0; s < N; s++) v += allCounts[size_t(s) * stride + g];
        total[g] = v;
    }
    __syncthreads();
    for (uint32_t g = t; g < R; g += 1024u) {     // position in the order (total descending, region id ascending)
        const uint32_t v = total[g];
        uint32_t rk = 0;
        for (uint32_t u = 0; u < R; u++) rk += (total[u] > v) || (total[u] == v && u < g);
        order[rk] = uint16_t(g);
        totalAll[g] = v;
    }
    __syncthreads();
    if (t == 0) {
        unsigned long long loads[B200PT_MAX_RANKS];
        for (int d = 0; d < N; d++) loads[d] = 0ull;
        uint32_t numActive = 0;
        unsigned long long tot = 0ull;
        for (uint32_t i = 0; i < R; i++) {
            const uint32_t g = order[i];
            if (total[g] == 0u) { own[g] = uint8_t(g % uint32_t(N)); continue; }
            int d = 0;
            for (int k = 1; k < N; k++) if (loads[k] < loads[d]) d = k;
            own[g] = uint8_t(d); loads[d] += total[g]; tot += total[g];
            if (d == me) activeRegions[numActive++] = g;
        }
        summary->numActive = numActive; summary->ownedSamples = loads[me]; summary->totalSamples = tot;
    }
    __syncthreads();
    for (uint32_t g = t; g < R; g += 1024u) {     // bins: regions ordered by (owner, region id)
        uint32_t b = 0;
        for (uint32_t u = 0; u < R; u++) b += (own[u] < own[g]) || (own[u] == own[g] && u < g);
        binOf[g] = uint16_t(b); regionOfBin[b] = uint16_t(g);
        ownerOut[g] = own[g];
        regionCount[g] = own[g] == me ? total[g] : 0u;
    }
    if (t <= unsigned(N)) { uint32_t c = 0; for (uint32_t u = 0; u < R; u++) c += own[u] < t; firstBin[t] = c; }
    __syncthreads();
    for (uint32_t g = t; g < R; g += 1024u) regionSlot[g] = own[g] == me ? uint32_t(binOf[g]) - firstBin[me] : 0u;      // position among this rank's regions
    for (int s = int(warp); s < N; s += 32) {      // every rank's sorted layout: exclusive scan of its counts over the bins
        uint32_t run = 0;
        for (uint32_t base = 0; base < R; base += 32) {
            const uint32_t b = base + lane;
            const uint32_t g = b < R ? regionOfBin[b] : 0u;
            const uint32_t c = b < R ? allCounts[size_t(s) * stride + g] : 0u;
            uint32_t incl = c;
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (int(lane) >= o) incl += u; }
            if (b < R) srcStart[size_t(s) * stride + g] = run + incl - c;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (s == me && lane == 0) summary->localValid = run;
    }
    if (t >= 32u * 16u && t < 32u * 16u + unsigned(N)) {        // my slice inside rank s's buffer
        const int s = int(t) - 32 * 16;
        uint32_t before = 0, mine = 0;
        for (uint32_t g = 0; g < R; g++) { const uint32_t c = allCounts[size_t(s) * stride + g]; before += own[g] < me ? c : 0u; mine += own[g] == me ? c : 0u; }
        sliceStart[s] = before; recvCount[s] = mine; summary->recvCount[s] = mine;
    }
    if (t >= 32u * 17u && t <= 32u * 17u + unsigned(N)) {       // rank d's slice inside my buffer
        const unsigned d = t - 32u * 17u;
        uint32_t before = 0;
        for (uint32_t g = 0; g < R; g++) before += own[g] < d ? allCounts[size_t(me) * stride + g] : 0u;
        summary->sendStart[d] = before;
    }
    if (warp == 18) {                              // my fit layout: owned regions in bin order, each one contiguous
        uint32_t run = 0;
        for (uint32_t base = firstBin[me]; base < firstBin[me + 1]; base += 32) {
            const uint32_t b = base + lane;
            const bool in = b < firstBin[me + 1];
            const uint32_t g = in ? regionOfBin[b] : 0u;
            const uint32_t c = in ? total[g] : 0u;
            uint32_t incl = c;
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (int(lane) >= o) incl += u; }
            if (in) regionBegin[g] = run + incl - c;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    if (t == 0) {
        uint32_t run = 0;
        for (int s = 0; s < N; s++) { stageStart[s] = run; summary->stageStart[s] = run; if (s != me) run += recvCount[s]; }
        summary->numOwned = firstBin[me + 1] - firstBin[me];
        summary->numSegments = (firstBin[me + 1] - firstBin[me]) * uint32_t(N);
    }
    __syncthreads();
    for (uint32_t g = t; g < R; g += 1024u) {
        if (own[g] != me) continue;
        const uint32_t idx = binOf[g] - firstBin[me];
        uint32_t dst = regionBegin[g];
        for (int s = 0; s < N; s++) {
            const uint32_t len = allCounts[size_t(s) * stride + g];
            uint32_t off = srcStart[size_t(s) * stride + g];
            if (!peerMode && s != me) off = stageStart[s] + (off - sliceStart[s]);
            segments[size_t(idx) * N + s] = GSegment{uint32_t(s), off, dst, len};
            dst += len;
        }
    }
    if (N == 1) {                                  // region-order offsets of the sorted buffer (parity hook b200pt_guiding_get_sorted)
        for (uint32_t g = t; g < R; g += 1024u) regionOffset[g] = regionBegin[g];
        if (t == 0) regionOffset[R] = uint32_t(summary->totalSamples);
    }
}

// the exchange: every block copies (part of) one segment — the records of one owned region that rank `src` holds — from
// that rank's sorted buffer (a CUDA-IPC mapping of its HBM, read over NVLink; or the local staging buffer in NCCL mode)
// to the region's place in the fit buffer.  Four independent 16-byte + 8-byte loads in flight per thread.
struct GPeerPtrs { const float4 *dirw[B200PT_MAX_RANKS]; const float2 *pd[B200PT_MAX_RANKS]; };
__global__ void __launch_bounds__(256) k_pull(const GSegment *__restrict__ segs, GPeerPtrs peers, float4 *__restrict__ dstDirw, float2 *__restrict__ dstPd) {
    const GSegment sg = segs[blockIdx.x];
    const float4 *__restrict__ sd = peers.dirw[sg.src] + sg.srcOff;
    const float2 *__restrict__ sp = peers.pd[sg.src] + sg.srcOff;
    float4 *dd = dstDirw + sg.dstOff;
    float2 *dp = dstPd + sg.dstOff;
    const uint32_t step = gridDim.y * 256u;
    uint32_t i = blockIdx.y * 256u + threadIdx.x;
    for (; i + 3u * step < sg.len; i += 4u * step) {
        const float4 a0 = sd[i], a1 = sd[i + step], a2 = sd[i + 2u * step], a3 = sd[i + 3u * step];
        const float2 b0 = sp[i], b1 = sp[i + step], b2 = sp[i + 2u * step], b3 = sp[i + 3u * step];
        dd[i] = a0; dd[i + step] = a1; dd[i + 2u * step] = a2; dd[i + 3u * step] = a3;
        dp[i] = b0; dp[i + step] = b1; dp[i + 2u * step] = b2; dp[i + 3u * step] = b3;
    }
    for (; i < sg.len; i += step) { dd[i] = sd[i]; dp[i] = sp[i]; }
}

// after the fit: every rank holds all ranks' mixture arrays (all-gather); a region that received samples takes its
// owner's result
__global__ void __launch_bounds__(128) k_pick_results(GMix *mixes, b200pt_vmm_theta *vmms, const GMix *__restrict__ gatherMix,
                                                       const b200pt_vmm_theta *__restrict__ gatherVmm, const uint8_t *__restrict__ owner,
                                                       const uint32_t *__restrict__ totalAll, uint32_t stride, int me) {
    const uint32_t g = blockIdx.x;
    if (totalAll[g] == 0u || owner[g] == me) return;
    const uint32_t *ms = reinterpret_cast<const uint32_t *>(gatherMix + size_t(owner[g]) * stride + g); uint32_t *md = reinterpret_cast<uint32_t *>(mixes + g);
    for (uint32_t i = threadIdx.x; i < sizeof(GMix) / 4; i += blockDim.x) md[i] = ms[i];
    const uint32_t *vs = reinterpret_cast<const uint32_t *>(gatherVmm + size_t(owner[g]) * stride + g); uint32_t *vd = reinterpret_cast<uint32_t *>(vmms + g);
    for (uint32_t i = threadIdx.x; i < sizeof(b200pt_vmm_theta) / 4; i += blockDim.x) vd[i] = vs[i];
}

__global__ void __launch_bounds__(SORT_WARPS * 32) k_sort_scatter(const b200pt_directional_data *__restrict__ recs, uint64_t n, uint32_t R,
                                                                  const uint32_t *__restrict__ tileOffsets, uint32_t numTiles,
                                                                  const uint32_t *__restrict__ regionStart, const b200pt_aabb *__restrict__ aabbs,
                                                                  int parallax, float4 *__restrict__ dirw, float2 *__restrict__ pdfDist,
                                                                  uint32_t *__restrict__ srcIndex) {
    extern __shared__ uint32_t hist[];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x * (blockDim.x >> 5) + warp;      // 8 warps per block, fewer when R histograms of 8 warps exceed 48 KB
    uint32_t *h = hist + warp * R;
    if (tile < numTiles) for (uint32_t i = lane; i < R; i += 32) h[i] = regionStart[i] + tileOffsets[uint64_t(tile) * R + i];
    __syncwarp();
    if (tile >= numTiles) return;
    const uint64_t begin = uint64_t(tile) * SORT_TILE;
    for (uint32_t k = 0; k < SORT_TILE; k += 32) {
        const uint64_t i = begin + k + lane;
        b200pt_directional_data d;
        d.flags = 0xffffffffu;
        if (i < n) d = recs[i];
        const bool valid = d.flags < R;
        const uint32_t dst = warpRank(h, d.flags, valid, lane);
        if (valid) {
            if (parallax) {     // PathGuiding::preFit: parallaxMean = aabb.min + 0.5 * (aabb.max - aabb.min)
                const b200pt_aabb bb = aabbs[d.flags];
                float mean[3];
                for (int a = 0; a < 3; a++) mean[a] = bb.min[a] + 0.5f * (bb.max[a] - bb.min[a]);
                gPrefitSample(d.position, d.direction, d.distance, mean);
            }
            dirw[dst] = make_float4(d.direction[0], d.direction[1], d.direction[2], d.weight);
            pdfDist[dst] = make_float2(d.pdf, d.distance);
            srcIndex[dst] = uint32_t(i);
        }
    }
}

// ---- per-region fit ----------------------------------------------------------------------------------------------------
struct BlockShared {
    GMix mix;
    GPacked packed;
    GFrames frames;
    GFitState fitState;
    union { EmAcc em; StatAcc stat; DistAcc dist; float raw[G_STATACC_FLOATS]; } acc;
    float metric[G_MAXK * (G_MAXK - 1) / 2];
    float red[G_WARPS][G_STATACC_FLOATS];
    float staged[G_STATACC_FLOATS];
    int bc;
};

// block-wide sum of NV per-thread values (registers) -> out[0..NV), deterministic order
template <int NV>
__device__ __forceinline__ void blockReduce(const float (&v)[NV], float (*red)[G_STATACC_FLOATS], float *out) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        float x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) red[warp][i] = x;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NV; i += blockDim.x) {
        float s = 0.0f;
        for (int w = 0; w < G_WARPS; w++) s += red[w][i];
        out[i] = s;
    }
    __syncthreads();
}

// One region is processed by a thread-block CLUSTER: every CTA of the cluster keeps its own copy of the mixture, loops
// over its slice of the region's samples, and after the block-level reduction the partial sums are combined through
// distributed shared memory in fixed rank order — every CTA reads all partials, so all of them hold the bit-identical
// totals and run the (tiny, deterministic) per-component logic redundantly; no broadcast of the mixture is needed.
struct BlockExec {
    BlockShared &sh;
    const float4 *dirw;
    const float2 *pdfDist;
    uint32_t N;
    uint32_t crank, csize;      // rank in the cluster / cluster size (1 = plain block)

    // combine sh.staged[0..nv) over the cluster; result back in sh.staged of every CTA
    __device__ void clusterCombine(int nv) {
        if (csize == 1) return;
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();
        float total = 0.0f;
        if (int(threadIdx.x) < nv)
            for (uint32_t r = 0; r < csize; r++) total += cluster.map_shared_rank(sh.staged, r)[threadIdx.x];
        cluster.sync();          // everybody has read every partial before any staged[] is overwritten
        if (int(threadIdx.x) < nv) sh.staged[threadIdx.x] = total;
        __syncthreads();
    }

    __device__ bool leader() const { return threadIdx.x == 0; }
    __device__ int bcast(int v) {
        __syncthreads();
        if (threadIdx.x == 0) sh.bc = v;
        __syncthreads();
        return sh.bc;
    }
    __device__ EmAcc &em() { return sh.acc.em; }
    __device__ StatAcc &stat() { return sh.acc.stat; }
    __device__ DistAcc &dst() { return sh.acc.dist; }
    __device__ GFrames &frames() { return sh.frames; }
    __device__ float *metric() { return sh.metric; }
    __device__ GFitState &fit() { return sh.fitState; }

    __device__ void publish(const GMix &m) {     // leader's mixture -> packed broadcast copy, visible to the block
        __syncthreads();
        if (threadIdx.x < G_MAXK) {
            const int c = threadIdx.x;
            sh.packed.a[c].mx = m.mux[c]; sh.packed.a[c].my = m.muy[c]; sh.packed.a[c].mz = m.muz[c]; sh.packed.a[c].kappa = m.kappa[c];
            sh.packed.b[c].norm = m.norm[c]; sh.packed.b[c].w = m.w[c];
        }
        __syncthreads();
    }

    template <int KPAD> __device__ void emPassT(EmAcc &out) {
        EmAcc a;   // register view: only the first KPAD slots of each array are touched
#pragma unroll
        for (int c = 0; c < KPAD; c++) { a.W[c] = 0.0f; a.Rx[c] = 0.0f; a.Ry[c] = 0.0f; a.Rz[c] = 0.0f; }
        a.sumWeight = 0.0f; a.logLikelihood = 0.0f;
        for (uint32_t i = crank * G_BLOCK + threadIdx.x; i < N; i += csize * G_BLOCK) {
            G_NO_HOIST();
            const float4 s = dirw[i];
            gEmSample<KPAD>(sh.packed, s.x, s.y, s.z, s.w, a);
        }
        float v[4 * KPAD + 2];
#pragma unroll
        for (int c = 0; c < KPAD; c++) { v[c] = a.W[c]; v[KPAD + c] = a.Rx[c]; v[2 * KPAD + c] = a.Ry[c]; v[3 * KPAD + c] = a.Rz[c]; }
        v[4 * KPAD] = a.sumWeight; v[4 * KPAD + 1] = a.logLikelihood;
        float *staged = sh.staged;   // reduce into a staging row, then unpack into the EmAcc layout
        blockReduce<4 * KPAD + 2>(v, sh.red, staged);
        clusterCombine(4 * KPAD + 2);
        if (threadIdx.x < KPAD) {
            const int c = threadIdx.x;
            out.W[c] = staged[c]; out.Rx[c] = staged[KPAD + c]; out.Ry[c] = staged[2 * KPAD + c]; out.Rz[c] = staged[3 * KPAD + c];
        }
        if (threadIdx.x == 0) { out.sumWeight = staged[4 * KPAD]; out.logLikelihood = staged[4 * KPAD + 1]; }
        __syncthreads();
    }
    __device__ void emPass(const GMix &m, EmAcc &out) {
        publish(m);
        switch (gKpad(m.K)) {
            case 4: emPassT<4>(out); break;
            case 8: emPassT<8>(out); break;
            case 12: emPassT<12>(out); break;
            default: emPassT<16>(out); break;
        }
    }

    template <int KPAD> __device__ void statPassT(const GFrames &f, StatAcc &out) {
        StatAcc a;
#pragma unroll
        for (int c = 0; c < KPAD; c++) { a.chi[c] = 0.0f; a.covW[c] = 0.0f; a.covXX[c] = 0.0f; a.covYY[c] = 0.0f; a.covXY[c] = 0.0f; }
        for (uint32_t i = crank * G_BLOCK + threadIdx.x; i < N; i += csize * G_BLOCK) {
            const float4 s = dirw[i];
            const float2 pd = pdfDist[i];
            G_NO_HOIST();
            gStatSample<KPAD>(sh.packed, f, s.x, s.y, s.z, s.w, pd.x, a);
        }
        float v[5 * KPAD];
#pragma unroll
        for (int c = 0; c < KPAD; c++) { v[c] = a.chi[c]; v[KPAD + c] = a.covW[c]; v[2 * KPAD + c] = a.covXX[c]; v[3 * KPAD + c] = a.covYY[c]; v[4 * KPAD + c] = a.covXY[c]; }
        float *staged = sh.staged;
        blockReduce<5 * KPAD>(v, sh.red, staged);
        clusterCombine(5 * KPAD);
        if (threadIdx.x < KPAD) {
            const int c = threadIdx.x;
            out.chi[c] = staged[c]; out.covW[c] = staged[KPAD + c]; out.covXX[c] = staged[2 * KPAD + c]; out.covYY[c] = staged[3 * KPAD + c]; out.covXY[c] = staged[4 * KPAD + c];
        }
        __syncthreads();
    }
    __device__ void statPass(const GMix &m, const GFrames &f, StatAcc &out) {
        publish(m);
        switch (gKpad(m.K)) {
            case 4: statPassT<4>(f, out); break;
            case 8: statPassT<8>(f, out); break;
            case 12: statPassT<12>(f, out); break;
            default: statPassT<16>(f, out); break;
        }
    }

    template <int KPAD> __device__ void distPassT(DistAcc &out) {
        DistAcc a;
#pragma unroll
        for (int c = 0; c < KPAD; c++) { a.w[c] = 0.0f; a.wd[c] = 0.0f; }
        for (uint32_t i = crank * G_BLOCK + threadIdx.x; i < N; i += csize * G_BLOCK) {
            const float4 s = dirw[i];
            const float2 pd = pdfDist[i];
            G_NO_HOIST();
            gDistSample<KPAD>(sh.packed, s.x, s.y, s.z, s.w, pd.y, a);
        }
        float v[2 * KPAD];
#pragma unroll
        for (int c = 0; c < KPAD; c++) { v[c] = a.w[c]; v[KPAD + c] = a.wd[c]; }
        float *staged = sh.staged;
        blockReduce<2 * KPAD>(v, sh.red, staged);
        clusterCombine(2 * KPAD);
        if (threadIdx.x < KPAD) { const int c = threadIdx.x; out.w[c] = staged[c]; out.wd[c] = staged[KPAD + c]; }
        __syncthreads();
    }
    __device__ void distPass(const GMix &m, DistAcc &out) {
        publish(m);
        switch (gKpad(m.K)) {
            case 4: distPassT<4>(out); break;
            case 8: distPassT<8>(out); break;
            case 12: distPassT<12>(out); break;
            default: distPassT<16>(out); break;
        }
    }

    __device__ void metricPass(const GMix &m, float *metric) {
        __syncthreads();
        const int K = m.K, numPairs = K * (K - 1) / 2;
        for (int p = threadIdx.x; p < numPairs; p += G_BLOCK) {
            int idx = p, a = 0;
            while (idx >= K - 1 - a) { idx -= K - 1 - a; a++; }
            metric[p] = gMergeMetricPair(m, a, a + 1 + idx);
        }
        __syncthreads();
    }
};

__global__ void __launch_bounds__(G_BLOCK, G_BLOCKS_PER_SM) k_guiding_update(GMix *mixes, b200pt_vmm_theta *vmms, const b200pt_aabb *__restrict__ aabbs,
                                                              const uint32_t *__restrict__ activeRegions, const uint32_t *__restrict__ numActive,
                                                              const uint32_t *__restrict__ regionBegin, const uint32_t *__restrict__ regionCount,
                                                              const float4 *__restrict__ dirw, const float2 *__restrict__ pdfDist,
                                                              b200pt_guiding_params gp, int firstFit, unsigned long long *emSampleIterations) {
    __shared__ BlockShared sh;
    const long long t0 = clock64();
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t csize = cluster.num_blocks(), crank = cluster.block_rank();
    if (blockIdx.x / csize >= *numActive) return;      // the grid is sized for every region; the whole cluster leaves together
    const uint32_t region = activeRegions[blockIdx.x / csize];
    const uint32_t begin = regionBegin[region], end = begin + regionCount[region];
    {   // mixture -> shared
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&mixes[region]);
        uint32_t *dstw = reinterpret_cast<uint32_t *>(&sh.mix);
        for (uint32_t i = threadIdx.x; i < sizeof(GMix) / 4; i += G_BLOCK) dstw[i] = src[i];
    }
    __syncthreads();
    BlockExec x{sh, dirw + begin, pdfDist + begin, end - begin, crank, csize};
    const b200pt_aabb bb = aabbs[region];
    float mean[3];
    for (int a = 0; a < 3; a++) mean[a] = bb.min[a] + 0.5f * (bb.max[a] - bb.min[a]);
    uint64_t iters = 0;
    gUpdateRegion(x, sh.mix, gp, end - begin, firstFit != 0, mean, &iters);
    if (threadIdx.x == 0) sh.mix.lastUpdateKCycles = uint32_t((clock64() - t0) >> 10);
    __syncthreads();
    if (csize > 1) cluster.sync();      // no CTA may exit while a peer could still read its shared memory
    if (crank != 0) return;             // every CTA holds the same result; rank 0 stores it
    {
        uint32_t *dstw = reinterpret_cast<uint32_t *>(&mixes[region]);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&sh.mix);
        for (uint32_t i = threadIdx.x; i < sizeof(GMix) / 4; i += G_BLOCK) dstw[i] = src[i];
    }
    if (threadIdx.x == 0) {
        gPackTheta(sh.mix, gp.useParallaxCompensation != 0, vmms[region]);
        atomicAdd(emSampleIterations, (unsigned long long)iters);
    }
}

// ---- work-sharing executor (block-parallel sums, the fast path) --------------------------------------------------------
// One thread block per region ("owner") runs the updateRegion sequence; every pass over the region's samples is cut
// into fixed chunks of GS_CHUNK samples that are handed out through one atomic word — to the owner itself and to every
// block that has no region of its own left ("helper").  The whole machine therefore works on whatever pass is open:
// a region that needs 100 EM iterations, or that holds a third of all samples, no longer sets the kernel time.
//   * a sample is evaluated by FOUR threads, thread q taking the components q, q+4, q+8, q+12 — lightpmm's SSE lanes:
//     the lane sums of the mixture pdf are combined as (l0+l2)+(l1+l3) with two shuffles, the values are the reference's;
//     18 accumulators per thread instead of 66 (registers for 3-4 resident blocks per SM instead of 1);
//   * partial sums: thread -> warp (shuffles) -> block (shared memory, warp order) -> one row per chunk in global memory;
//     the owner adds the rows in chunk order.  Sample -> thread, chunk boundaries and all summation orders are fixed, so
//     the result does not depend on which block computed which chunk (bit-identical from run to run and across ranks).
#define GS_BLOCK 256
#define GS_QUADS (GS_BLOCK / 4)
#define GS_WARPS (GS_BLOCK / 32)
#define GS_CHUNK_MAX 2048            // 32 samples per quad
#define GS_ROWS_PER_REGION 66        // a region's passes have at most max(65, N_r / 2048 + 1) chunks
enum { GS_PASS_EM = 0, GS_PASS_STAT = 1, GS_PASS_DIST = 2 };
// Chunk size of a region's passes: a function of the region's sample count only (so the summation order is fixed):
// about 64 chunks per pass for mid-size regions — enough for the whole machine to finish the last regions' iterations
// together — between 4 and 32 samples per quad.
__host__ __device__ __forceinline__ uint32_t gsChunkSamples(uint32_t count) {
    uint32_t perQuad = (count + 4095u) / 4096u;
    perQuad = perQuad < 4u ? 4u : (perQuad > 32u ? 32u : perQuad);
    return perQuad * 64u;
}

struct alignas(16) GPass {           // one per region, global memory
    unsigned long long word;         // number of chunks << 32 | next chunk; atomicAdd(word, 1) hands out a chunk
    uint32_t done;                   // chunks finished
    uint32_t type;                   // GS_PASS_* | KPAD << 8
    uint32_t begin, count, chunkBase, passCount;   // passCount: passes this region has opened in this update (helpers prefer long runners)
    GPacked packed;
    GFrames frames;
};

struct SharedShared {
    GMix mix;
    GPacked packed;                  // lobes of the pass being processed (own region's or a helped one's)
    GFrames frames;
    GFitState fitState;
    union { EmAcc em; StatAcc stat; DistAcc dist; float raw[G_STATACC_FLOATS]; } acc;
    float metric[G_MAXK * (G_MAXK - 1) / 2];
    float staged[G_STATACC_FLOATS];
    float red[GS_WARPS][G_STATACC_FLOATS];
    int bc;
    unsigned long long grab;
    uint32_t hType, hBegin, hCount, hChunkBase, hSlot;
};

__device__ __forceinline__ uint32_t ldAcquire(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// (l0 + l2) + (l1 + l3) over the four threads of a quad; every thread gets the same float (addition commutes)
__device__ __forceinline__ float quadLaneSum(float lane) {
    const float a = lane + __shfl_xor_sync(0xffffffffu, lane, 2);
    return a + __shfl_xor_sync(0xffffffffu, a, 1);
}
// weighted component pdfs of this thread's components (gMixturePdf restricted to lane q) and the mixture pdf
template <int KQ, bool WANT_PDF>
__device__ __forceinline__ float quadMixturePdf(const GPacked &m, unsigned q, float dx, float dy, float dz, float *wpdf, float *pdf) {
    float lane = 0.0f;
#pragma unroll
    for (int k = 0; k < KQ; k++) {
        const GPacked::A a = m.a[4 * k + q];
        const GPacked::B b = m.b[4 * k + q];
        const float cosTheta = a.mx * dx + a.my * dy + a.mz * dz;
        const float t = gSseMin(cosTheta - 1.0f, 0.0f);
        const float p = b.norm * gFastExp(a.kappa * t);
        if (WANT_PDF) pdf[k] = p;
        wpdf[k] = b.w * p;
        lane += wpdf[k];
    }
    return quadLaneSum(lane);
}

// sum NA per-thread accumulators over the block: threads with the same quad lane q hold the same components.
// value index of accumulator j of lane q: idx(j, q).  Result rows -> out[0..NV) (global), fixed order.
template <int NA, class Idx>
__device__ __forceinline__ void quadBlockReduce(float (&v)[NA], Idx idx, int NV, float (*red)[G_STATACC_FLOATS], float *out) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, q = threadIdx.x & 3u;
#pragma unroll
    for (int j = 0; j < NA; j++) {
        float x = v[j];
        x += __shfl_xor_sync(0xffffffffu, x, 4);
        x += __shfl_xor_sync(0xffffffffu, x, 8);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        const int i = idx(j, q);
        if (lane < 4 && i >= 0) red[warp][i] = x;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NV; i += GS_BLOCK) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < GS_WARPS; w++) s += red[w][i];
        out[i] = s;
    }
}

template <int KQ>
__device__ void chunkEm(const GPacked &pk, const float4 *__restrict__ dirw, uint32_t first, uint32_t end, float (*red)[G_STATACC_FLOATS], float *out) {
    constexpr int KPAD = 4 * KQ;
    const unsigned q = threadIdx.x & 3u, quad = threadIdx.x >> 2;
    float acc[4 * KQ + 2];
#pragma unroll
    for (int j = 0; j < 4 * KQ + 2; j++) acc[j] = 0.0f;
    // the log-likelihood term needs one logf per SAMPLE, not per thread: lane (it & 3) of the quad keeps the sample's
    // (weight, mixture pdf) and every fourth iteration each lane takes the logarithm of the one it holds
    float pendW = 0.0f, pendPdf = 1.0f;
    const float4 idle = make_float4(0.0f, 0.0f, 1.0f, 0.0f);
    uint32_t i = first + quad, it = 0;
    bool valid = i < end;
    float4 s = valid ? dirw[i] : idle;
    for (uint32_t base = first; base < end; base += GS_QUADS, it++) {
        G_NO_HOIST();
        const uint32_t in = i + GS_QUADS;
        const bool validNext = in < end;
        const float4 sn = validNext ? dirw[in] : idle;          // next sample: in flight while this one is evaluated
        float sw[KQ];
        const float mixturePDF = quadMixturePdf<KQ, false>(pk, q, s.x, s.y, s.z, sw, (float *)0);
        const bool ok = valid && mixturePDF > G_PMM_EPSILON;
        if (ok) {
            const float inv = 1.0f / mixturePDF;
#pragma unroll
            for (int k = 0; k < KQ; k++) {
                const float w = (sw[k] * inv) * s.w;
                acc[KQ + k] += s.x * w; acc[2 * KQ + k] += s.y * w; acc[3 * KQ + k] += s.z * w;
                acc[k] += w;
            }
            acc[4 * KQ] += s.w;
        }
        if ((it & 3u) == q) { pendW = ok ? s.w : 0.0f; pendPdf = ok ? mixturePDF : 1.0f; }
        if ((it & 3u) == 3u) { acc[4 * KQ + 1] += pendW * logf(pendPdf); pendW = 0.0f; pendPdf = 1.0f; }
        s = sn; valid = validNext; i = in;
    }
    float ll = acc[4 * KQ + 1] + pendW * logf(pendPdf);
    ll += __shfl_xor_sync(0xffffffffu, ll, 1);
    ll += __shfl_xor_sync(0xffffffffu, ll, 2);
    acc[4 * KQ + 1] = ll;
    // EmAcc layout of the partial row: W[c] | Rx[c] | Ry[c] | Rz[c] | sumWeight | logLikelihood, c = 4k + q
    quadBlockReduce<4 * KQ + 2>(acc, [](int j, unsigned q) -> int {
        if (j < 4 * KQ) return (j / KQ) * KPAD + 4 * (j % KQ) + int(q);
        return q == 0 ? 4 * KPAD + (j - 4 * KQ) : -1;
    }, 4 * KPAD + 2, red, out);
}

template <int KQ>
__device__ void chunkStat(const GPacked &pk, const GFrames &f, const float4 *__restrict__ dirw, const float2 *__restrict__ pdfDist, uint32_t first, uint32_t end,
                          float (*red)[G_STATACC_FLOATS], float *out) {
    constexpr int KPAD = 4 * KQ;
    const unsigned q = threadIdx.x & 3u, quad = threadIdx.x >> 2;
    float acc[5 * KQ];
#pragma unroll
    for (int j = 0; j < 5 * KQ; j++) acc[j] = 0.0f;
    const float4 idle = make_float4(0.0f, 0.0f, 1.0f, 0.0f);
    uint32_t i = first + quad;
    bool valid = i < end;
    float4 s = valid ? dirw[i] : idle;
    float2 pd = valid ? pdfDist[i] : make_float2(1.0f, 0.0f);
    for (uint32_t base = first; base < end; base += GS_QUADS) {
        G_NO_HOIST();
        const uint32_t in = i + GS_QUADS;
        const bool validNext = in < end;
        const float4 sn = validNext ? dirw[in] : idle;
        const float2 pdn = validNext ? pdfDist[in] : make_float2(1.0f, 0.0f);
        float wpdf[KQ], pdf[KQ];
        const float mixturePDF = quadMixturePdf<KQ, true>(pk, q, s.x, s.y, s.z, wpdf, pdf);
        if (valid && mixturePDF > G_PMM_EPSILON) {
            const float mixturePDFSqr = mixturePDF * mixturePDF;
            const float ideal = s.w * s.w * pd.x / mixturePDFSqr;
            const float inv = 1.0f / mixturePDF;
#pragma unroll
            for (int k = 0; k < KQ; k++) {
                const int c = 4 * k + int(q);
                acc[k] += pdf[k] * ideal;
                const float ws = s.w * (wpdf[k] * inv);
                acc[KQ + k] += ws;
                const float lx = f.sx[c] * s.x + f.sy[c] * s.y + f.sz[c] * s.z;
                const float ly = f.tx[c] * s.x + f.ty[c] * s.y + f.tz[c] * s.z;
                acc[2 * KQ + k] += lx * lx * ws;
                acc[3 * KQ + k] += ly * ly * ws;
                acc[4 * KQ + k] += lx * ly * ws;
            }
        }
        s = sn; pd = pdn; valid = validNext; i = in;
    }
    quadBlockReduce<5 * KQ>(acc, [](int j, unsigned q) -> int { return (j / KQ) * KPAD + 4 * (j % KQ) + int(q); }, 5 * KPAD, red, out);
}

template <int KQ>
__device__ void chunkDist(const GPacked &pk, const float4 *__restrict__ dirw, const float2 *__restrict__ pdfDist, uint32_t first, uint32_t end,
                          float (*red)[G_STATACC_FLOATS], float *out) {
    constexpr int KPAD = 4 * KQ;
    const unsigned q = threadIdx.x & 3u, quad = threadIdx.x >> 2;
    float acc[2 * KQ];
#pragma unroll
    for (int j = 0; j < 2 * KQ; j++) acc[j] = 0.0f;
    const float4 idle = make_float4(0.0f, 0.0f, 1.0f, 0.0f);
    uint32_t i = first + quad;
    bool valid = i < end;
    float4 s = valid ? dirw[i] : idle;
    float2 pd = valid ? pdfDist[i] : make_float2(1.0f, 0.0f);
    for (uint32_t base = first; base < end; base += GS_QUADS) {
        G_NO_HOIST();
        const uint32_t in = i + GS_QUADS;
        const bool validNext = in < end;
        const float4 sn = validNext ? dirw[in] : idle;
        const float2 pdn = validNext ? pdfDist[in] : make_float2(1.0f, 0.0f);
        float wpdf[KQ], pdf[KQ];
        const float mixturePDF = quadMixturePdf<KQ, true>(pk, q, s.x, s.y, s.z, wpdf, pdf);
        if (valid && pd.y > 0.0f && mixturePDF > G_PMM_EPSILON) {
            const float sw = s.w / mixturePDF;
#pragma unroll
            for (int k = 0; k < KQ; k++) {
                const float v = wpdf[k] * pdf[k] * sw;
                acc[k] += v;
                acc[KQ + k] += v / pd.y;
            }
        }
        s = sn; pd = pdn; valid = validNext; i = in;
    }
    quadBlockReduce<2 * KQ>(acc, [](int j, unsigned q) -> int { return (j / KQ) * KPAD + 4 * (j % KQ) + int(q); }, 2 * KPAD, red, out);
}

// one chunk of one pass; lobes (and frames) are in shared memory
__device__ void processChunk(SharedShared &sh, uint32_t type, uint32_t begin, uint32_t count, uint32_t chunk, const float4 *__restrict__ dirw,
                             const float2 *__restrict__ pdfDist, float *out) {
    const uint32_t cs = gsChunkSamples(count);
    const uint32_t first = begin + chunk * cs, end = begin + min(count, (chunk + 1) * cs);
    const uint32_t kq = (type >> 8) / 4, pass = type & 0xffu;
#define GS_DISPATCH(FN, ...)                                                   \
    switch (kq) {                                                              \
        case 1: FN<1>(__VA_ARGS__); break;                                     \
        case 2: FN<2>(__VA_ARGS__); break;                                     \
        case 3: FN<3>(__VA_ARGS__); break;                                     \
        default: FN<4>(__VA_ARGS__); break;                                    \
    }
    if (pass == GS_PASS_EM) { GS_DISPATCH(chunkEm, sh.packed, dirw, first, end, sh.red, out) }
    else if (pass == GS_PASS_STAT) { GS_DISPATCH(chunkStat, sh.packed, sh.frames, dirw, pdfDist, first, end, sh.red, out) }
    else { GS_DISPATCH(chunkDist, sh.packed, dirw, pdfDist, first, end, sh.red, out) }
#undef GS_DISPATCH
}

struct SharedExec {
    SharedShared &sh;
    GPass *pass;                // this region's pass descriptor (global)
    float *partials;            // [chunks][G_STATACC_FLOATS] (global)
    const float4 *dirw;         // whole sorted buffer
    const float2 *pdfDist;
    uint32_t begin, N, chunkBase;
    uint32_t passCount;         // thread 0 only

    __device__ bool leader() const { return threadIdx.x == 0; }
    __device__ int bcast(int v) {
        __syncthreads();
        if (threadIdx.x == 0) sh.bc = v;
        __syncthreads();
        return sh.bc;
    }
    __device__ EmAcc &em() { return sh.acc.em; }
    __device__ StatAcc &stat() { return sh.acc.stat; }
    __device__ DistAcc &dst() { return sh.acc.dist; }
    __device__ GFrames &frames() { return sh.frames; }
    __device__ float *metric() { return sh.metric; }
    __device__ GFitState &fit() { return sh.fitState; }

    // open a pass: lobes (+ frames) to shared AND global memory, then the chunk counter; work on it until no chunk is
    // left; wait for the helpers' chunks; add the partial rows in chunk order -> sh.staged[0..nv)
    __device__ void runPass(const GMix &m, uint32_t passType, int nv) {
        __syncthreads();
        const uint32_t kpad = uint32_t(gKpad(m.K)), type = passType | (kpad << 8);
        if (threadIdx.x < G_MAXK) {
            const int c = threadIdx.x;
            GPacked::A a; a.mx = m.mux[c]; a.my = m.muy[c]; a.mz = m.muz[c]; a.kappa = m.kappa[c];
            GPacked::B b; b.norm = m.norm[c]; b.w = m.w[c];
            sh.packed.a[c] = a; sh.packed.b[c] = b;
            pass->packed.a[c] = a; pass->packed.b[c] = b;
            if (passType == GS_PASS_STAT) {
                pass->frames.sx[c] = sh.frames.sx[c]; pass->frames.sy[c] = sh.frames.sy[c]; pass->frames.sz[c] = sh.frames.sz[c];
                pass->frames.tx[c] = sh.frames.tx[c]; pass->frames.ty[c] = sh.frames.ty[c]; pass->frames.tz[c] = sh.frames.tz[c];
            }
            __threadfence();
        }
        const uint32_t cs = gsChunkSamples(N), numChunks = (N + cs - 1) / cs;
        __syncthreads();
        if (threadIdx.x == 0) {
            pass->type = type; pass->done = 0u; pass->passCount = ++passCount;
            __threadfence();
            atomicExch(&pass->word, (unsigned long long)numChunks << 32);
        }
        for (;;) {
            if (threadIdx.x == 0) sh.grab = atomicAdd(&pass->word, 1ull);
            __syncthreads();
            const unsigned long long g = sh.grab;
            const uint32_t c = uint32_t(g & 0xffffffffull);
            if (c >= uint32_t(g >> 32)) break;
            processChunk(sh, type, begin, N, c, dirw, pdfDist, partials + size_t(chunkBase + c) * G_STATACC_FLOATS);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(&pass->done, 1u);
        }
        if (threadIdx.x == 0) while (ldAcquire(&pass->done) < numChunks) __nanosleep(100);
        __syncthreads();
        if (int(threadIdx.x) < nv) {      // rows in chunk order; eight loads in flight
            float sum = 0.0f;
            const float *row = partials + size_t(chunkBase) * G_STATACC_FLOATS + threadIdx.x;
            uint32_t c = 0;
            for (; c + 8 <= numChunks; c += 8) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = __ldcg(row + size_t(c + j) * G_STATACC_FLOATS);
#pragma unroll
                for (int j = 0; j < 8; j++) sum += v[j];
            }
            for (; c < numChunks; c++) sum += __ldcg(row + size_t(c) * G_STATACC_FLOATS);
            sh.staged[threadIdx.x] = sum;
        }
        __syncthreads();
    }

    __device__ void emPass(const GMix &m, EmAcc &out) {
        const int KPAD = gKpad(m.K);
        runPass(m, GS_PASS_EM, 4 * KPAD + 2);
        const float *staged = sh.staged;
        if (int(threadIdx.x) < KPAD) {
            const int c = threadIdx.x;
            out.W[c] = staged[c]; out.Rx[c] = staged[KPAD + c]; out.Ry[c] = staged[2 * KPAD + c]; out.Rz[c] = staged[3 * KPAD + c];
        }
        if (threadIdx.x == 0) { out.sumWeight = staged[4 * KPAD]; out.logLikelihood = staged[4 * KPAD + 1]; }
        __syncthreads();
    }
    __device__ void statPass(const GMix &m, const GFrames &f, StatAcc &out) {
        (void)f;    // == sh.frames
        const int KPAD = gKpad(m.K);
        runPass(m, GS_PASS_STAT, 5 * KPAD);
        const float *staged = sh.staged;
        if (int(threadIdx.x) < KPAD) {
            const int c = threadIdx.x;
            out.chi[c] = staged[c]; out.covW[c] = staged[KPAD + c]; out.covXX[c] = staged[2 * KPAD + c]; out.covYY[c] = staged[3 * KPAD + c]; out.covXY[c] = staged[4 * KPAD + c];
        }
        __syncthreads();
    }
    __device__ void distPass(const GMix &m, DistAcc &out) {
        const int KPAD = gKpad(m.K);
        runPass(m, GS_PASS_DIST, 2 * KPAD);
        const float *staged = sh.staged;
        if (int(threadIdx.x) < KPAD) { const int c = threadIdx.x; out.w[c] = staged[c]; out.wd[c] = staged[KPAD + c]; }
        __syncthreads();
    }
    __device__ void metricPass(const GMix &m, float *metric) {
        __syncthreads();
        const int K = m.K, numPairs = K * (K - 1) / 2;
        for (int p = threadIdx.x; p < numPairs; p += GS_BLOCK) {
            int idx = p, a = 0;
            while (idx >= K - 1 - a) { idx -= K - 1 - a; a++; }
            metric[p] = gMergeMetricPair(m, a, a + 1 + idx);
        }
        __syncthreads();
    }
};

// persistent: every block first takes regions to own (largest first), then helps until all regions are done.
// control[0] = next region to hand out, control[1] = regions finished
#ifndef GS_MIN_BLOCKS
#define GS_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(GS_BLOCK, GS_MIN_BLOCKS) k_guiding_update_shared(GMix *mixes, b200pt_vmm_theta *vmms, const b200pt_aabb *__restrict__ aabbs,
                                                                     const uint32_t *__restrict__ activeRegions, const uint32_t *__restrict__ numActivePtr,
                                                                     const uint32_t *__restrict__ regionBegin, const uint32_t *__restrict__ regionCount,
                                                                     const uint32_t *__restrict__ regionSlot, const float4 *__restrict__ dirw,
                                                                     const float2 *__restrict__ pdfDist, GPass *passes, float *partials, uint32_t *control,
                                                                     b200pt_guiding_params gp, int firstFit, unsigned long long *emSampleIterations) {
    __shared__ SharedShared sh;
    const uint32_t numActive = *numActivePtr;
    for (;;) {      // ---- owner phase
        if (threadIdx.x == 0) sh.hSlot = atomicAdd(&control[0], 1u);
        __syncthreads();
        const uint32_t slot = sh.hSlot;
        __syncthreads();
        if (slot >= numActive) break;
        const long long t0 = clock64();
        const uint32_t region = activeRegions[slot];
        const uint32_t begin = regionBegin[region], count = regionCount[region];
        {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&mixes[region]);
            uint32_t *dstw = reinterpret_cast<uint32_t *>(&sh.mix);
            for (uint32_t i = threadIdx.x; i < sizeof(GMix) / 4; i += GS_BLOCK) dstw[i] = src[i];
        }
        __syncthreads();
        GPass *pass = passes + region;
        const uint32_t chunkBase = begin / GS_CHUNK_MAX + regionSlot[region] * GS_ROWS_PER_REGION;
        if (threadIdx.x == 0) { pass->begin = begin; pass->count = count; pass->chunkBase = chunkBase; }
        SharedExec x{sh, pass, partials, dirw, pdfDist, begin, count, chunkBase, 0u};
        const b200pt_aabb bb = aabbs[region];
        float mean[3];
        for (int a = 0; a < 3; a++) mean[a] = bb.min[a] + 0.5f * (bb.max[a] - bb.min[a]);
        uint64_t iters = 0;
        gUpdateRegion(x, sh.mix, gp, count, firstFit != 0, mean, &iters);
        if (threadIdx.x == 0) sh.mix.lastUpdateKCycles = uint32_t((clock64() - t0) >> 10);
        __syncthreads();
        {
            uint32_t *dstw = reinterpret_cast<uint32_t *>(&mixes[region]);
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&sh.mix);
            for (uint32_t i = threadIdx.x; i < sizeof(GMix) / 4; i += GS_BLOCK) dstw[i] = src[i];
        }
        if (threadIdx.x == 0) {
            gPackTheta(sh.mix, gp.useParallaxCompensation != 0, vmms[region]);
            atomicAdd(emSampleIterations, (unsigned long long)iters);
            __threadfence();
            atomicAdd(&control[1], 1u);
        }
        __syncthreads();
    }
    // ---- helper phase: look for the open pass with the most chunks left, take one chunk, repeat
    for (;;) {
        if (threadIdx.x < 32) {      // one warp looks at the open passes; the others wait at the barrier (no issue slots)
            uint32_t best = 0u, bestSlot = 0xffffffffu;
            for (uint32_t sl = threadIdx.x; sl < numActive; sl += 32) {
                const GPass *ps = passes + activeRegions[sl];
                const unsigned long long w = __ldcg(&ps->word);
                const uint32_t next = uint32_t(w & 0xffffffffull), n = uint32_t(w >> 32);
                if (next >= n) continue;
                // prefer the region that has run the most passes (it is the one the kernel will end up waiting for), then the fullest pass
                const uint32_t key32 = (min(__ldcg(&ps->passCount), 0xffffu) << 16) | min(n - next, 0xffffu);
                if (key32 > best) { best = key32; bestSlot = sl; }
            }
            unsigned long long key = (unsigned long long)best << 32 | (0xffffffffu - bestSlot);
            for (int o = 16; o > 0; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o); key = other > key ? other : key; }
            if (threadIdx.x == 0) {
                sh.hSlot = (key >> 32) ? 0xffffffffu - uint32_t(key & 0xffffffffull) : 0xffffffffu;
                sh.grab = ~0ull;
                if (sh.hSlot != 0xffffffffu) sh.grab = atomicAdd(&passes[activeRegions[sh.hSlot]].word, 1ull);
                else if (ldAcquire(&control[1]) >= numActive) sh.hSlot = 0xfffffffeu;      // nothing open and every region finished: leave
                else __nanosleep(1000);
            }
        }
        __syncthreads();
        const uint32_t sl = sh.hSlot;
        const unsigned long long g = sh.grab;
        const uint32_t c = uint32_t(g & 0xffffffffull);
        if (sl == 0xfffffffeu) break;
        if (sl == 0xffffffffu || c >= uint32_t(g >> 32)) { __syncthreads(); continue; }
        // a chunk of a foreign pass is ours: its parameters were published before the counter was opened
        GPass *pass = passes + activeRegions[sl];
        __threadfence();
        if (threadIdx.x < G_MAXK) {
            const int cc = threadIdx.x;
            const float4 a = __ldcg(reinterpret_cast<const float4 *>(&pass->packed.a[cc]));
            const float2 b = __ldcg(reinterpret_cast<const float2 *>(&pass->packed.b[cc]));
            sh.packed.a[cc].mx = a.x; sh.packed.a[cc].my = a.y; sh.packed.a[cc].mz = a.z; sh.packed.a[cc].kappa = a.w;
            sh.packed.b[cc].norm = b.x; sh.packed.b[cc].w = b.y;
            sh.frames.sx[cc] = __ldcg(&pass->frames.sx[cc]); sh.frames.sy[cc] = __ldcg(&pass->frames.sy[cc]); sh.frames.sz[cc] = __ldcg(&pass->frames.sz[cc]);
            sh.frames.tx[cc] = __ldcg(&pass->frames.tx[cc]); sh.frames.ty[cc] = __ldcg(&pass->frames.ty[cc]); sh.frames.tz[cc] = __ldcg(&pass->frames.tz[cc]);
        }
        if (threadIdx.x == 32) { sh.hType = __ldcg(&pass->type); sh.hBegin = __ldcg(&pass->begin); sh.hCount = __ldcg(&pass->count); sh.hChunkBase = __ldcg(&pass->chunkBase); }
        __syncthreads();
        processChunk(sh, sh.hType, sh.hBegin, sh.hCount, c, dirw, pdfDist, partials + size_t(sh.hChunkBase + c) * G_STATACC_FLOATS);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(&pass->done, 1u);
    }
}

// ---- strict-order executor ---------------------------------------------------------------------------------------------
// lightpmm sums the sufficient statistics sample after sample in float (VMMFactory.h:497-536): float addition does not
// commute with regrouping, so any parallel partition of a region's samples changes the last bits of the sums, and the EM
// iteration amplifies that (weights ~1e-4, kappa ~1e-3, now and then one EM iteration more or less).  This executor keeps
// the reference's order exactly and is still parallel — over the ACCUMULATORS instead of the samples:
//   * producer warps evaluate the per-sample terms (soft assignment, 4K+2 values per sample for an EM pass) for a tile of
//     GC_TILE consecutive samples, four threads per sample (lightpmm's lanes, as in the fast path: with one thread per
//     sample five warps per region could not hide their own latency — 73 ms for config 1, profiles/r02_ncu_guiding_strict.txt),
//     and store them value-major in shared memory;
//   * consumer thread v owns running sum v and adds its row of the tile in sample order: a dependent FADD chain
//     (4 cycles per sample), 4K+2 chains side by side, fed by LDS.128 (row stride = 41 x 16 B: conflict-free);
//   * the two tile buffers alternate: the consumers add tile j while the producers compute tile j + 1.
// A skipped sample (mixture pdf <= 1e-8) contributes +0 to every sum, which leaves a float sum unchanged.  The result
// is the serial sum, bit for bit the one the host build of guiding_math.cuh (tests/harness) produces.
#define GC_CONSUMERS 96
#define GC_TILE 160
#define GC_PRODUCERS (4 * GC_TILE)            // four threads per sample, like the fast path (thread q: components q, q+4, q+8, q+12)
#define GC_BLOCK (GC_CONSUMERS + GC_PRODUCERS)
#define GC_STRIDE (GC_TILE + 4)
#define GC_ROWS G_STATACC_FLOATS
#define GC_SMEM_BYTES (2 * GC_ROWS * GC_STRIDE * sizeof(float))

struct ChainShared {
    GMix mix;
    GPacked packed;
    GFrames frames;
    GFitState fitState;
    union { EmAcc em; StatAcc stat; DistAcc dist; float raw[G_STATACC_FLOATS]; } acc;
    float metric[G_MAXK * (G_MAXK - 1) / 2];
    float staged[G_STATACC_FLOATS];
    int bc;
};

struct ChainExec {
    ChainShared &sh;
    float *tiles;               // dynamic shared memory: [2][GC_ROWS][GC_STRIDE]
    const float4 *dirw;
    const float2 *pdfDist;
    uint32_t N;

    __device__ bool leader() const { return threadIdx.x == 0; }
    __device__ int bcast(int v) {
        __syncthreads();
        if (threadIdx.x == 0) sh.bc = v;
        __syncthreads();
        return sh.bc;
    }
    __device__ EmAcc &em() { return sh.acc.em; }
    __device__ StatAcc &stat() { return sh.acc.stat; }
    __device__ DistAcc &dst() { return sh.acc.dist; }
    __device__ GFrames &frames() { return sh.frames; }
    __device__ float *metric() { return sh.metric; }
    __device__ GFitState &fit() { return sh.fitState; }

    __device__ void publish(const GMix &m) {
        __syncthreads();
        if (threadIdx.x < G_MAXK) {
            const int c = threadIdx.x;
            sh.packed.a[c].mx = m.mux[c]; sh.packed.a[c].my = m.muy[c]; sh.packed.a[c].mz = m.muz[c]; sh.packed.a[c].kappa = m.kappa[c];
            sh.packed.b[c].norm = m.norm[c]; sh.packed.b[c].w = m.w[c];
        }
        __syncthreads();
    }

    // One pass over the region's samples.  produce(i, valid, q, col): evaluation of sample i by the four threads q = 0..3 of
    // a quad (all of them call it, also for i >= N: the mixture pdf is combined with warp shuffles); thread q stores the values
    // of ITS components at col[row * GC_STRIDE].  Afterwards sh.staged[0..NV) holds the sequential sums.
    struct Smp { float4 s; float2 pd; };
    template <bool PD>
    __device__ __forceinline__ Smp loadSample(uint32_t i) const {
        Smp r;
        const bool valid = i < N;
        r.s = valid ? dirw[i] : make_float4(0.0f, 0.0f, 1.0f, 0.0f);
        r.pd = (PD && valid) ? pdfDist[i] : make_float2(1.0f, 0.0f);
        return r;
    }
    template <int NV, bool PD, class Produce>
    __device__ __forceinline__ void chainPass(Produce produce) {
        const unsigned tid = threadIdx.x;
        const bool consumer = tid < GC_CONSUMERS;
        const unsigned p = tid - GC_CONSUMERS, quad = p >> 2, q = p & 3u;
        const uint32_t numTiles = (N + GC_TILE - 1) / GC_TILE;
        float acc = 0.0f;
        Smp next;
        if (!consumer) {
            const Smp cur = loadSample<PD>(quad);
            next = loadSample<PD>(GC_TILE + quad);            // the following tile's sample is in flight while this one is evaluated
            produce(cur, quad < N, q, tiles + quad);
        }
        __syncthreads();
        for (uint32_t j = 0; j < numTiles; j++) {
            if (consumer) {
                if (tid < unsigned(NV)) {
                    const float *row = tiles + (j & 1u) * (GC_ROWS * GC_STRIDE) + tid * GC_STRIDE;
                    const uint32_t count = min(uint32_t(GC_TILE), N - j * GC_TILE);
                    uint32_t sIdx = 0;
                    if (count == GC_TILE) {
#pragma unroll 8
                        for (; sIdx < GC_TILE; sIdx += 4) {
                            const float4 t4 = *reinterpret_cast<const float4 *>(row + sIdx);
                            acc += t4.x; acc += t4.y; acc += t4.z; acc += t4.w;
                        }
                    } else {
                        for (; sIdx < count; sIdx++) acc += row[sIdx];
                    }
                }
            } else if (j + 1 < numTiles) {
                const uint32_t i = (j + 1) * GC_TILE + quad;
                const Smp cur = next;
                next = loadSample<PD>(i + GC_TILE);
                produce(cur, i < N, q, tiles + ((j + 1) & 1u) * (GC_ROWS * GC_STRIDE) + quad);
            }
            __syncthreads();
        }
        if (tid < unsigned(NV)) sh.staged[tid] = acc;
        __syncthreads();
    }

    template <int KPAD> __device__ void emPassT(EmAcc &out) {
        const GPacked &pk = sh.packed;
        const float4 *d = dirw;
        constexpr int KQ = KPAD / 4;
        chainPass<4 * KPAD + 2, false>([&](const Smp &smp, bool valid, unsigned q, float *col) {
            G_NO_HOIST();
            const float4 s = smp.s;
            float sw[KQ];
            const float mixturePDF = quadMixturePdf<KQ, false>(pk, q, s.x, s.y, s.z, sw, (float *)0);
            if (!valid) return;
            const bool ok = mixturePDF > G_PMM_EPSILON;
            const float inv = 1.0f / mixturePDF;
#pragma unroll
            for (int k = 0; k < KQ; k++) {
                const int c = 4 * k + int(q);
                const float w = ok ? (sw[k] * inv) * s.w : 0.0f;
                col[c * GC_STRIDE] = w;
                col[(KPAD + c) * GC_STRIDE] = s.x * w;
                col[(2 * KPAD + c) * GC_STRIDE] = s.y * w;
                col[(3 * KPAD + c) * GC_STRIDE] = s.z * w;
            }
            if (q == 0u) col[(4 * KPAD) * GC_STRIDE] = ok ? s.w : 0.0f;
            if (q == 1u) col[(4 * KPAD + 1) * GC_STRIDE] = ok ? s.w * logf(mixturePDF) : 0.0f;
        });
        const float *staged = sh.staged;
        if (threadIdx.x < KPAD) {
            const int c = threadIdx.x;
            out.W[c] = staged[c]; out.Rx[c] = staged[KPAD + c]; out.Ry[c] = staged[2 * KPAD + c]; out.Rz[c] = staged[3 * KPAD + c];
        }
        if (threadIdx.x == 0) { out.sumWeight = staged[4 * KPAD]; out.logLikelihood = staged[4 * KPAD + 1]; }
        __syncthreads();
    }
    __device__ void emPass(const GMix &m, EmAcc &out) {
        publish(m);
        switch (gKpad(m.K)) {
            case 4: emPassT<4>(out); break;
            case 8: emPassT<8>(out); break;
            case 12: emPassT<12>(out); break;
            default: emPassT<16>(out); break;
        }
    }

    template <int KPAD> __device__ void statPassT(const GFrames &f, StatAcc &out) {
        const GPacked &pk = sh.packed;
        const float4 *d = dirw;
        const float2 *pd2 = pdfDist;
        constexpr int KQ = KPAD / 4;
        chainPass<5 * KPAD, true>([&](const Smp &smp, bool valid, unsigned q, float *col) {
            const float4 s = smp.s;
            const float2 pd = smp.pd;
            G_NO_HOIST();
            float wpdf[KQ], pdf[KQ];
            const float mixturePDF = quadMixturePdf<KQ, true>(pk, q, s.x, s.y, s.z, wpdf, pdf);
            if (!valid) return;
            const bool ok = mixturePDF > G_PMM_EPSILON;
            const float mixturePDFSqr = mixturePDF * mixturePDF;
            const float ideal = s.w * s.w * pd.x / mixturePDFSqr;
            const float inv = 1.0f / mixturePDF;
#pragma unroll
            for (int k = 0; k < KQ; k++) {
                const int c = 4 * k + int(q);
                const float ws = s.w * (wpdf[k] * inv);
                const float lx = f.sx[c] * s.x + f.sy[c] * s.y + f.sz[c] * s.z;
                const float ly = f.tx[c] * s.x + f.ty[c] * s.y + f.tz[c] * s.z;
                col[c * GC_STRIDE] = ok ? pdf[k] * ideal : 0.0f;
                col[(KPAD + c) * GC_STRIDE] = ok ? ws : 0.0f;
                col[(2 * KPAD + c) * GC_STRIDE] = ok ? lx * lx * ws : 0.0f;
                col[(3 * KPAD + c) * GC_STRIDE] = ok ? ly * ly * ws : 0.0f;
                col[(4 * KPAD + c) * GC_STRIDE] = ok ? lx * ly * ws : 0.0f;
            }
        });
        const float *staged = sh.staged;
        if (threadIdx.x < KPAD) {
            const int c = threadIdx.x;
            out.chi[c] = staged[c]; out.covW[c] = staged[KPAD + c]; out.covXX[c] = staged[2 * KPAD + c]; out.covYY[c] = staged[3 * KPAD + c]; out.covXY[c] = staged[4 * KPAD + c];
        }
        __syncthreads();
    }
    __device__ void statPass(const GMix &m, const GFrames &f, StatAcc &out) {
        publish(m);
        switch (gKpad(m.K)) {
            case 4: statPassT<4>(f, out); break;
            case 8: statPassT<8>(f, out); break;
            case 12: statPassT<12>(f, out); break;
            default: statPassT<16>(f, out); break;
        }
    }

    template <int KPAD> __device__ void distPassT(DistAcc &out) {
        const GPacked &pk = sh.packed;
        const float4 *d = dirw;
        const float2 *pd2 = pdfDist;
        constexpr int KQ = KPAD / 4;
        chainPass<2 * KPAD, true>([&](const Smp &smp, bool valid, unsigned q, float *col) {
            const float4 s = smp.s;
            const float2 pd = smp.pd;
            G_NO_HOIST();
            float wpdf[KQ], pdf[KQ];
            const float mixturePDF = quadMixturePdf<KQ, true>(pk, q, s.x, s.y, s.z, wpdf, pdf);
            if (!valid) return;
            const bool ok = pd.y > 0.0f && mixturePDF > G_PMM_EPSILON;
            const float sw = s.w / mixturePDF;
#pragma unroll
            for (int k = 0; k < KQ; k++) {
                const int c = 4 * k + int(q);
                const float v = wpdf[k] * pdf[k] * sw;
                col[c * GC_STRIDE] = ok ? v : 0.0f;
                col[(KPAD + c) * GC_STRIDE] = ok ? v / pd.y : 0.0f;
            }
        });
        const float *staged = sh.staged;
        if (threadIdx.x < KPAD) { const int c = threadIdx.x; out.w[c] = staged[c]; out.wd[c] = staged[KPAD + c]; }
        __syncthreads();
    }
    __device__ void distPass(const GMix &m, DistAcc &out) {
        publish(m);
        switch (gKpad(m.K)) {
            case 4: distPassT<4>(out); break;
            case 8: distPassT<8>(out); break;
            case 12: distPassT<12>(out); break;
            default: distPassT<16>(out); break;
        }
    }

    __device__ void metricPass(const GMix &m, float *metric) {
        __syncthreads();
        const int K = m.K, numPairs = K * (K - 1) / 2;
        for (int p = threadIdx.x; p < numPairs; p += GC_BLOCK) {
            int idx = p, a = 0;
            while (idx >= K - 1 - a) { idx -= K - 1 - a; a++; }
            metric[p] = gMergeMetricPair(m, a, a + 1 + idx);
        }
        __syncthreads();
    }
};

__global__ void __launch_bounds__(GC_BLOCK, 1) k_guiding_update_strict(GMix *mixes, b200pt_vmm_theta *vmms, const b200pt_aabb *__restrict__ aabbs,
                                                                     const uint32_t *__restrict__ activeRegions, const uint32_t *__restrict__ numActive,
                                                                     const uint32_t *__restrict__ regionBegin, const uint32_t *__restrict__ regionCount,
                                                                     const float4 *__restrict__ dirw, const float2 *__restrict__ pdfDist,
                                                                     b200pt_guiding_params gp, int firstFit, unsigned long long *emSampleIterations) {
    __shared__ ChainShared sh;
    extern __shared__ __align__(16) float gcTiles[];
    if (blockIdx.x >= *numActive) return;
    const long long t0 = clock64();
    const uint32_t region = activeRegions[blockIdx.x];
    const uint32_t begin = regionBegin[region], count = regionCount[region];
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&mixes[region]);
        uint32_t *dstw = reinterpret_cast<uint32_t *>(&sh.mix);
        for (uint32_t i = threadIdx.x; i < sizeof(GMix) / 4; i += GC_BLOCK) dstw[i] = src[i];
    }
    __syncthreads();
    ChainExec x{sh, gcTiles, dirw + begin, pdfDist + begin, count};
    const b200pt_aabb bb = aabbs[region];
    float mean[3];
    for (int a = 0; a < 3; a++) mean[a] = bb.min[a] + 0.5f * (bb.max[a] - bb.min[a]);
    uint64_t iters = 0;
    gUpdateRegion(x, sh.mix, gp, count, firstFit != 0, mean, &iters);
    if (threadIdx.x == 0) sh.mix.lastUpdateKCycles = uint32_t((clock64() - t0) >> 10);
    __syncthreads();
    {
        uint32_t *dstw = reinterpret_cast<uint32_t *>(&mixes[region]);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&sh.mix);
        for (uint32_t i = threadIdx.x; i < sizeof(GMix) / 4; i += GC_BLOCK) dstw[i] = src[i];
    }
    if (threadIdx.x == 0) {
        gPackTheta(sh.mix, gp.useParallaxCompensation != 0, vmms[region]);
        atomicAdd(emSampleIterations, (unsigned long long)iters);
    }
}

#define G_TRY(expr)                                                                              \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) { error = std::string(#expr) + ": " + cudaGetErrorString(_e); return B200PT_E_CUDA; } \
    } while (0)

// PathGuiding::createPMMs + syncPMMsToVMM_Thetas at construction (src/PathGuiding.cpp:42-51, 101-103).  The initial
// mixture is evaluated on the HOST (VMMFactory::initialize uses acos / sin / cos: the C library's floats are the
// reference's, CUDA's differ in the last bit, and a kappa = 50000 lobe turns one ulp of mu into 0.5 % of pdf) and
// replicated to every region on the device.
__global__ void k_guiding_init(GMix *mixes, b200pt_vmm_theta *vmms, uint32_t R, const GMix *__restrict__ mix0, const b200pt_vmm_theta *__restrict__ vmm0) {
    const uint32_t r = blockIdx.x;
    if (r >= R) return;
    const uint32_t *ms = reinterpret_cast<const uint32_t *>(mix0); uint32_t *md = reinterpret_cast<uint32_t *>(mixes + r);
    for (uint32_t i = threadIdx.x; i < sizeof(GMix) / 4; i += blockDim.x) md[i] = ms[i];
    const uint32_t *vs = reinterpret_cast<const uint32_t *>(vmm0); uint32_t *vd = reinterpret_cast<uint32_t *>(vmms + r);
    for (uint32_t i = threadIdx.x; i < sizeof(b200pt_vmm_theta) / 4; i += blockDim.x) vd[i] = vs[i];
}
static int launchGuidingInit(GMix *mixes, b200pt_vmm_theta *vmms, int regions, const b200pt_guiding_params &gp, cudaStream_t stream, std::string &error) {
    GMix m;
    b200pt_vmm_theta v;
    gInitialize(m, gp);
    gPackTheta(m, gp.useParallaxCompensation != 0, v);
    char *d = nullptr;
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&d), sizeof(GMix) + sizeof(b200pt_vmm_theta)));
    G_TRY(cudaMemcpyAsync(d, &m, sizeof(GMix), cudaMemcpyHostToDevice, stream));
    G_TRY(cudaMemcpyAsync(d + sizeof(GMix), &v, sizeof(v), cudaMemcpyHostToDevice, stream));
    k_guiding_init<<<unsigned(regions), 128, 0, stream>>>(mixes, vmms, uint32_t(regions), reinterpret_cast<const GMix *>(d), reinterpret_cast<const b200pt_vmm_theta *>(d + sizeof(GMix)));
    G_TRY(cudaGetLastError());
    G_TRY(cudaStreamSynchronize(stream));
    cudaFree(d);
    return B200PT_OK;
}

// known-answer hook: the device build of lightpmm's approximate exp on an array
__global__ void k_fastexp(const float *__restrict__ in, float *__restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = gFastExp(in[i]);
}
// exhaustive check of gFastExpDiv against the IEEE division it replaces: every float bit pattern in [lo, hi]
__global__ void k_fastexp_div_check(uint32_t loBits, uint32_t count, unsigned long long *mismatches) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float d = __uint_as_float(loBits + i);
    if (__float_as_uint(gFastExpDiv(d)) != __float_as_uint(__fdiv_rn(27.7280233f, d))) atomicAdd(mismatches, 1ull);
}
int guidingDivisionSelfTest(float lo, float hi, unsigned long long *mismatchesOut, unsigned long long *testedOut, cudaStream_t stream, std::string &error) {
    uint32_t a, b;
    memcpy(&a, &lo, 4); memcpy(&b, &hi, 4);
    if (!(lo > 0.0f) || !(hi >= lo)) { error = "bad range"; return B200PT_E_INVALID; }
    const uint32_t count = b - a + 1;
    unsigned long long *d = nullptr, h = 0;
    if (cudaMalloc(reinterpret_cast<void **>(&d), sizeof(h)) != cudaSuccess) { error = "cudaMalloc failed"; return B200PT_E_CUDA; }
    cudaMemsetAsync(d, 0, sizeof(h), stream);
    k_fastexp_div_check<<<(count + 255) / 256, 256, 0, stream>>>(a, count, d);
    cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, stream);
    cudaError_t e = cudaStreamSynchronize(stream);
    cudaFree(d);
    if (e != cudaSuccess) { error = cudaGetErrorString(e); return B200PT_E_CUDA; }
    *mismatchesOut = h; *testedOut = count;
    return B200PT_OK;
}
int guidingFastExp(const float *hostIn, float *hostOut, int n, cudaStream_t stream, std::string &error) {
    if (n <= 0) return B200PT_OK;
    float *d = nullptr;
    if (cudaMalloc(reinterpret_cast<void **>(&d), size_t(n) * 2 * sizeof(float)) != cudaSuccess) { error = "cudaMalloc failed"; return B200PT_E_CUDA; }
    cudaMemcpyAsync(d, hostIn, size_t(n) * sizeof(float), cudaMemcpyHostToDevice, stream);
    k_fastexp<<<(n + 255) / 256, 256, 0, stream>>>(d, d + n, n);
    cudaMemcpyAsync(hostOut, d + n, size_t(n) * sizeof(float), cudaMemcpyDeviceToHost, stream);
    cudaError_t e = cudaStreamSynchronize(stream);
    cudaFree(d);
    if (e != cudaSuccess) { error = cudaGetErrorString(e); return B200PT_E_CUDA; }
    return B200PT_OK;
}

// ---- host side ---------------------------------------------------------------------------------------------------------
// PathGuiding::splitRegion (src/PathGuiding.cpp:328-348), device half: decay the split mixture's sample counters, then the
// new region starts as a copy of it (mixture, extra statistics, packed VMM_Theta)
__global__ void __launch_bounds__(128) k_guiding_split_regions(GMix *mixes, b200pt_vmm_theta *vmms, const int2 *pairs) {
    const int2 pr = pairs[blockIdx.x];
    if (threadIdx.x == 0) {
        const float decayTerm = 0.25f;
        mixes[pr.x].numSamples *= decayTerm;
        mixes[pr.x].sampleWeight *= decayTerm;
    }
    __syncthreads();
    const uint32_t *ms = reinterpret_cast<const uint32_t *>(&mixes[pr.x]); uint32_t *md = reinterpret_cast<uint32_t *>(&mixes[pr.y]);
    for (uint32_t i = threadIdx.x; i < sizeof(GMix) / 4; i += blockDim.x) md[i] = ms[i];
    const uint32_t *vs = reinterpret_cast<const uint32_t *>(&vmms[pr.x]); uint32_t *vd = reinterpret_cast<uint32_t *>(&vmms[pr.y]);
    for (uint32_t i = threadIdx.x; i < sizeof(b200pt_vmm_theta) / 4; i += blockDim.x) vd[i] = vs[i];
}

int GuidingState::init(int splits, const float sceneMin[3], const float sceneMax[3], cudaStream_t stream) {
    release();
    regionCount = 1 << splits;
    // capacity of the per-region buffers: 1024 up to GUIDING_SPLITS 9 (512 initial regions), G_PLAN_MAX_REGIONS beyond (b200pt_create limits
    // splits to B200PT_MAX_GUIDING_SPLITS = 11, 2048 initial regions) — adaptive refinement always has room to double the initial count
    maxRegions = regionCount <= 512 ? 1024 : G_PLAN_MAX_REGIONS;
    if (2 * regionCount > maxRegions) { error = "too many guiding regions (GUIDING_SPLITS <= 11)"; return B200PT_E_INVALID; }
    b200pt_aabb scene;
    for (int a = 0; a < 3; a++) {   // Aabb::addEpsilon, src/Shapes.h:48-53
        float extent = sceneMax[a] - sceneMin[a];
        float center = sceneMin[a] + 0.5f * extent;
        scene.min[a] = center - 0.50001f * extent;
        scene.max[a] = center + 0.50001f * extent;
    }
    hostAabbs.assign(1, scene);
    std::vector<b200pt_aabb> allLevels(1, scene);      // every level of the halving tree, for the tracer's region lookup
    for (int i = 0; i < splits; i++) {
        std::vector<b200pt_aabb> next;
        next.reserve(hostAabbs.size() * 2);
        for (const auto &b : hostAabbs) { b200pt_aabb l, r; splitAabb(b, l, r); next.push_back(l); next.push_back(r); }
        hostAabbs.swap(next);
        allLevels.insert(allLevels.end(), hostAabbs.begin(), hostAabbs.end());
    }
    this->splits = splits;
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&levelAabbs), allLevels.size() * sizeof(b200pt_aabb)));
    G_TRY(cudaMemcpyAsync(levelAabbs, allLevels.data(), allLevels.size() * sizeof(b200pt_aabb), cudaMemcpyHostToDevice, stream));
    {   // what the tracer's greedy descent reads per level (guiding_device.cuh getGuidingRegion): the children differ from their parent
        // on the split axis only (splitAabb), so "child contains p" is one comparison there once p is known to be inside the parent
        const size_t inner = (size_t(1) << splits) - 1;
        std::vector<float4> recs(std::max<size_t>(inner, 1), make_float4(0.0f, 0.0f, 0.0f, 0.0f));
        for (size_t i = 0; i < inner; i++) {
            const b200pt_aabb &b = allLevels[i];
            const float size[3] = {b.max[0] - b.min[0], b.max[1] - b.min[1], b.max[2] - b.min[2]};
            const int axis = size[0] > size[1] ? (size[0] > size[2] ? 0 : 2) : (size[1] > size[2] ? 1 : 2);      // splitAabb's choice
            const b200pt_aabb &l = allLevels[2 * i + 1], &r = allLevels[2 * i + 2];
            recs[i] = make_float4(__int_as_float_host(axis), l.max[axis], r.min[axis], 0.0f);
        }
        G_TRY(cudaMalloc(reinterpret_cast<void **>(&levelSplits), recs.size() * sizeof(float4)));
        G_TRY(cudaMemcpyAsync(levelSplits, recs.data(), recs.size() * sizeof(float4), cudaMemcpyHostToDevice, stream));
    }
    G_TRY(cudaStreamSynchronize(stream));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&aabbs), size_t(maxRegions) * sizeof(b200pt_aabb)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&vmms), size_t(maxRegions) * sizeof(b200pt_vmm_theta)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&mixes), size_t(maxRegions) * sizeof(GMix)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&regionTotal), size_t(maxRegions) * sizeof(uint32_t)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&regionOffset), size_t(maxRegions + 1) * sizeof(uint32_t)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&spawnFirst), size_t(maxRegions) * sizeof(int32_t)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&spawnNext), size_t(maxRegions) * sizeof(int32_t)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&splitPairs), size_t(maxRegions) * sizeof(int2)));
    hostSpawnFirst.assign(size_t(maxRegions), -1); hostSpawnNext.assign(size_t(maxRegions), -1); hasSpawns = false;
    G_TRY(cudaMemsetAsync(spawnFirst, 0xff, size_t(maxRegions) * sizeof(int32_t), stream));
    G_TRY(cudaMemsetAsync(spawnNext, 0xff, size_t(maxRegions) * sizeof(int32_t), stream));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&activeRegions), size_t(maxRegions) * sizeof(uint32_t)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&regionBegin), size_t(maxRegions) * sizeof(uint32_t)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&regionLen), size_t(maxRegions) * sizeof(uint32_t)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&totalAll), size_t(maxRegions) * sizeof(uint32_t)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&owner), size_t(maxRegions)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&regionSlot), size_t(maxRegions) * sizeof(uint32_t)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&passes), size_t(maxRegions) * sizeof(GPass)));
    G_TRY(cudaMemsetAsync(passes, 0, size_t(maxRegions) * sizeof(GPass), stream));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&fitControl), 2 * sizeof(uint32_t)));
    {
        int dev = 0, sms = 0, occ = 0;
        G_TRY(cudaGetDevice(&dev));
        G_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        G_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_guiding_update_shared, GS_BLOCK, 0));
        sharedGrid = sms * std::max(1, occ);
    }
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&planDev), sizeof(GPlanSummary)));
    G_TRY(cudaMemsetAsync(planDev, 0, sizeof(GPlanSummary), stream));
    G_TRY(cudaMallocHost(reinterpret_cast<void **>(&planHost), sizeof(GPlanSummary)));
    memset(planHost, 0, sizeof(GPlanSummary));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&barrierWord), sizeof(float)));
    G_TRY(cudaMemsetAsync(barrierWord, 0, sizeof(float), stream));
    for (auto &e : ev) G_TRY(cudaEventCreate(&e));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&devScalars), 4 * sizeof(unsigned long long)));
    G_TRY(cudaMallocHost(reinterpret_cast<void **>(&hostScalars), 4 * sizeof(unsigned long long)));
    G_TRY(cudaMemcpyAsync(aabbs, hostAabbs.data(), size_t(regionCount) * sizeof(b200pt_aabb), cudaMemcpyHostToDevice, stream));
    G_TRY(cudaMemsetAsync(devScalars, 0, 4 * sizeof(unsigned long long), stream));
    b200pt_default_guiding_params(&lastParams);
    { int rci = launchGuidingInit(mixes, vmms, regionCount, lastParams, stream, error); if (rci != B200PT_OK) return rci; }
    firstFit = true;
    ready = true;
    return B200PT_OK;
}

int GuidingState::reset(const b200pt_guiding_params &params, cudaStream_t stream) {
    if (!ready) { error = "guiding state not initialised"; return B200PT_E_STATE; }
    lastParams = params;
    { int rci = launchGuidingInit(mixes, vmms, regionCount, params, stream, error); if (rci != B200PT_OK) return rci; }
    firstFit = true;
    return B200PT_OK;
}

int GuidingState::ensureCapacity(int64_t numSamples) {
    if (numSamples <= capacity) return B200PT_OK;
    if (dirw) cudaFree(dirw);
    if (pdfDist) cudaFree(pdfDist);
    if (srcIndex) cudaFree(srcIndex);
    if (tileCounts) cudaFree(tileCounts);
    dirw = nullptr; pdfDist = nullptr; srcIndex = nullptr; tileCounts = nullptr; capacity = 0;
    const size_t n = size_t(std::max<int64_t>(numSamples, 1));
    const size_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&dirw), n * sizeof(float4)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&pdfDist), n * sizeof(float2)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&srcIndex), n * sizeof(uint32_t)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&tileCounts), tiles * size_t(maxRegions) * sizeof(uint32_t)));
    capacity = numSamples;
    return B200PT_OK;
}

int GuidingState::ensurePlan(int ranks) {
    if (ranks <= planRanks) return B200PT_OK;
    if (allCounts) cudaFree(allCounts);
    if (srcStart) cudaFree(srcStart);
    if (segments) cudaFree(segments);
    if (gatherMix) cudaFree(gatherMix);
    if (gatherVmm) cudaFree(gatherVmm);
    allCounts = nullptr; srcStart = nullptr; segments = nullptr; gatherMix = nullptr; gatherVmm = nullptr; planRanks = 0;
    const size_t n = size_t(ranks) * size_t(maxRegions);
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&allCounts), n * sizeof(uint32_t)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&srcStart), n * sizeof(uint32_t)));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&segments), n * sizeof(GSegment)));
    if (ranks > 1) {
        G_TRY(cudaMalloc(reinterpret_cast<void **>(&gatherMix), n * sizeof(GMix)));
        G_TRY(cudaMalloc(reinterpret_cast<void **>(&gatherVmm), n * sizeof(b200pt_vmm_theta)));
    }
    planRanks = ranks;
    return B200PT_OK;
}

// parity hook of the multi-rank plan: runs k_plan on caller-supplied per-rank counts as rank `me` of `N` and returns what it
// decided (no communicator, no samples involved) — lets a single-GPU test check ownership and layouts for any rank count
int GuidingState::planDebug(const uint32_t *counts, int N, int me, int peerMode, uint8_t *ownerOut, uint32_t *regionBeginOut, uint32_t *regionLenOut,
                            uint32_t *srcStartOut, uint32_t *activeOut, uint32_t *summaryOut /* numActive, numOwned, numSegments, localValid, ownedSamples, totalSamples */,
                            uint32_t *segmentsOut /* numOwned * N x {src, srcOff, dstOff, len} */, cudaStream_t stream) {
    if (!ready) { error = "guiding state not initialised"; return B200PT_E_STATE; }
    if (N < 1 || N > B200PT_MAX_RANKS || me < 0 || me >= N) { error = "bad rank arguments"; return B200PT_E_INVALID; }
    int rcode = ensurePlan(N);
    if (rcode != B200PT_OK) return rcode;
    const uint32_t R = uint32_t(regionCount), stride = uint32_t(maxRegions);
    for (int s = 0; s < N; s++) G_TRY(cudaMemcpyAsync(allCounts + size_t(s) * stride, counts + size_t(s) * R, R * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    k_plan<<<1, 1024, 0, stream>>>(allCounts, N, me, R, stride, peerMode, srcStart, regionBegin, regionLen, regionOffset, activeRegions, totalAll, owner, regionSlot, segments, planDev);
    G_TRY(cudaGetLastError());
    G_TRY(cudaMemcpyAsync(planHost, planDev, sizeof(GPlanSummary), cudaMemcpyDeviceToHost, stream));
    G_TRY(cudaMemcpyAsync(ownerOut, owner, R, cudaMemcpyDeviceToHost, stream));
    G_TRY(cudaMemcpyAsync(regionBeginOut, regionBegin, R * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    G_TRY(cudaMemcpyAsync(regionLenOut, regionLen, R * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    for (int s = 0; s < N; s++) G_TRY(cudaMemcpyAsync(srcStartOut + size_t(s) * R, srcStart + size_t(s) * stride, R * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    G_TRY(cudaMemcpyAsync(activeOut, activeRegions, R * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    G_TRY(cudaStreamSynchronize(stream));
    summaryOut[0] = planHost->numActive; summaryOut[1] = planHost->numOwned; summaryOut[2] = planHost->numSegments; summaryOut[3] = planHost->localValid;
    summaryOut[4] = uint32_t(planHost->ownedSamples); summaryOut[5] = uint32_t(planHost->totalSamples);
    if (planHost->numSegments) G_TRY(cudaMemcpy(segmentsOut, segments, size_t(planHost->numSegments) * sizeof(GSegment), cudaMemcpyDeviceToHost));
    return B200PT_OK;
}

#define G_NCCL(expr)                                                                             \
    do {                                                                                         \
        ncclResult_t _r = (expr);                                                                \
        if (_r != ncclSuccess) { error = std::string(#expr) + ": " + g_nccl.GetErrorString(_r); return B200PT_E_CUDA; } \
    } while (0)

// Map every rank's sorted-sample buffers (dirw, pdfDist) into this process with CUDA IPC: the handles travel through the
// communicator, and the decision to use them is taken collectively (a rank that cannot open a peer's handle makes all
// ranks fall back to ncclSend / ncclRecv).  B200PT_EXCHANGE=nccl forces the fallback.
int GuidingState::setupPeers(RankComm &rc, cudaStream_t stream) {
    closePeers();
    peerSelf = rc.rank;
    peerTried = true; peerCapacity = capacity; rc.peerMode = false;
    struct Handles { cudaIpcMemHandle_t dirw, pd; };
    Handles mine{};
    float ok = 1.0f;
    if (const char *e = getenv("B200PT_EXCHANGE")) if (!strcmp(e, "nccl")) ok = 0.0f;
    if (ok != 0.0f && (cudaIpcGetMemHandle(&mine.dirw, dirw) != cudaSuccess || cudaIpcGetMemHandle(&mine.pd, pdfDist) != cudaSuccess)) { cudaGetLastError(); ok = 0.0f; }
    Handles *devAll = nullptr;
    std::vector<Handles> all(size_t(rc.nranks));
    G_TRY(cudaMalloc(reinterpret_cast<void **>(&devAll), sizeof(Handles) * size_t(rc.nranks)));
    G_TRY(cudaMemcpyAsync(devAll + rc.rank, &mine, sizeof(Handles), cudaMemcpyHostToDevice, stream));
    G_NCCL(g_nccl.AllGather(devAll + rc.rank, devAll, sizeof(Handles), ncclChar, rc.comm, stream));
    G_TRY(cudaMemcpyAsync(all.data(), devAll, sizeof(Handles) * size_t(rc.nranks), cudaMemcpyDeviceToHost, stream));
    G_TRY(cudaStreamSynchronize(stream));
    if (ok != 0.0f)
        for (int s = 0; s < rc.nranks; s++) {
            if (s == rc.rank) { peerDirw[s] = dirw; peerPdfDist[s] = pdfDist; continue; }
            void *a = nullptr, *b = nullptr;
            if (cudaIpcOpenMemHandle(&a, all[size_t(s)].dirw, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                cudaIpcOpenMemHandle(&b, all[size_t(s)].pd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError(); if (a) cudaIpcCloseMemHandle(a); ok = 0.0f; break;
            }
            peerDirw[s] = static_cast<const float4 *>(a); peerPdfDist[s] = static_cast<const float2 *>(b);
        }
    // collective decision: min over ranks
    G_TRY(cudaMemcpyAsync(barrierWord, &ok, sizeof(float), cudaMemcpyHostToDevice, stream));
    G_NCCL(g_nccl.AllReduce(barrierWord, barrierWord, 1, ncclFloat, ncclMin, rc.comm, stream));
    G_TRY(cudaMemcpyAsync(&ok, barrierWord, sizeof(float), cudaMemcpyDeviceToHost, stream));
    G_TRY(cudaStreamSynchronize(stream));
    cudaFree(devAll);
    if (ok == 0.0f) closePeers();
    rc.peerMode = ok != 0.0f;
    if (getenv("B200PT_COMM_DEBUG")) fprintf(stderr, "[b200pt rank %d] sample exchange: %s\n", rc.rank, rc.peerMode ? "CUDA-IPC peer reads over NVLink" : "ncclSend/ncclRecv");
    return B200PT_OK;
}

void GuidingState::closePeers() {
    const int me = peerSelf;
    for (int s = 0; s < B200PT_MAX_RANKS; s++) {
        if (s != me && peerDirw[s]) cudaIpcCloseMemHandle(const_cast<float4 *>(peerDirw[s]));
        if (s != me && peerPdfDist[s]) cudaIpcCloseMemHandle(const_cast<float2 *>(peerPdfDist[s]));
        peerDirw[s] = nullptr; peerPdfDist[s] = nullptr;
    }
}

int GuidingState::update(b200pt_directional_data *samples, int64_t numSamples, const b200pt_guiding_params &params, cudaStream_t stream,
                         b200pt_stats *stats, RankComm *rc) {
    if (!ready) { error = "guiding state not initialised"; return B200PT_E_STATE; }
    if (regionCount > G_PLAN_MAX_REGIONS) { error = "the device sort supports at most 4096 guiding regions"; return B200PT_E_INVALID; }
    if (numSamples < 0 || numSamples > int64_t(0xfffffff0u)) { error = "bad sample count"; return B200PT_E_INVALID; }
    if (params.numInitialComponents < 1 || params.numInitialComponents > G_MAXK || params.maxItr < 0 || params.minItr < 0) {
        error = "bad guiding parameters"; return B200PT_E_INVALID;
    }
    const int N = (rc && rc->comm) ? rc->nranks : 1, me = N > 1 ? rc->rank : 0;
    if (N > B200PT_MAX_RANKS) { error = "too many ranks"; return B200PT_E_INVALID; }
    if (firstFit && memcmp(&params, &lastParams, sizeof(params)) != 0) {   // factory properties changed before the first fit: re-initialise
        int r0 = reset(params, stream);
        if (r0 != B200PT_OK) return r0;
    }
    lastParams = params;
    int rcode = ensureCapacity(numSamples);
    if (rcode != B200PT_OK) return rcode;
    rcode = ensurePlan(N);
    if (rcode != B200PT_OK) return rcode;
    if (N > 1 && (!peerTried || peerCapacity != capacity)) {               // (re)map the peers' buffers: collective, same condition on every rank
        rcode = setupPeers(*rc, stream);
        if (rcode != B200PT_OK) return rcode;
    }
    const bool peerMode = N > 1 && rc->peerMode;
    const uint32_t R = uint32_t(regionCount), stride = uint32_t(maxRegions);
    const uint32_t numTiles = uint32_t((uint64_t(numSamples) + SORT_TILE - 1) / SORT_TILE);
    // one histogram of R bins per warp in shared memory: 8 warps per block while that fits the 48 KB every kernel may use, fewer beyond
    uint32_t sortWarps = SORT_WARPS;
    while (sortWarps > 1u && size_t(sortWarps) * R * sizeof(uint32_t) > 48u * 1024u) sortWarps >>= 1;
    const uint32_t sortBlocks = (numTiles + sortWarps - 1) / sortWarps;
    const size_t sortSmem = size_t(sortWarps) * R * sizeof(uint32_t);
    unsigned launches = 0;
    G_TRY(cudaEventRecord(ev[0], stream));
    G_TRY(cudaMemsetAsync(devScalars, 0, 2 * sizeof(unsigned long long), stream));
    // ---- count (local), exchange the counts, plan
    uint32_t *myCounts = allCounts + size_t(me) * stride;
    if (numTiles) {
        k_sort_count<<<sortBlocks, sortWarps * 32, sortSmem, stream>>>(samples, uint64_t(numSamples), R, tileCounts, numTiles);
        k_sort_scan_tiles<<<R, 256, 0, stream>>>(tileCounts, numTiles, R, myCounts);
        launches += 2;
    } else {
        G_TRY(cudaMemsetAsync(myCounts, 0, R * sizeof(uint32_t), stream));
    }
    if (N > 1) G_NCCL(g_nccl.AllGather(myCounts, allCounts, size_t(stride) * sizeof(uint32_t), ncclChar, rc->comm, stream));
    k_plan<<<1, 1024, 0, stream>>>(allCounts, N, me, R, stride, peerMode ? 1 : 0, srcStart, regionBegin, regionLen, regionOffset, activeRegions, totalAll,
                                   owner, regionSlot, segments, planDev);
    launches++;
    G_TRY(cudaGetLastError());
    G_TRY(cudaMemcpyAsync(planHost, planDev, sizeof(GPlanSummary), cudaMemcpyDeviceToHost, stream));
    G_TRY(cudaEventRecord(ev[2], stream));
    // ---- scatter (sorted by (owner, region); preFit applied), while the host waits for the plan
    if (numTiles) {
        k_sort_scatter<<<sortBlocks, sortWarps * 32, sortSmem, stream>>>(samples, uint64_t(numSamples), R, tileCounts, numTiles, srcStart + size_t(me) * stride, aabbs,
                                                                          params.useParallaxCompensation, dirw, pdfDist, srcIndex);
        launches++;
    }
    G_TRY(cudaEventRecord(ev[1], stream));
    const float4 *fitD = dirw;
    const float2 *fitP = pdfDist;
    uint32_t fitRegions = R;                     // single GPU: the grid covers every region, empty ones leave at once (no host round trip)
    uint64_t bytesReceived = 0;
    if (N > 1) {
        G_TRY(cudaEventSynchronize(ev[2]));      // the only host wait inside the update; the scatter is running meanwhile
        const GPlanSummary &ps = *planHost;
        fitRegions = ps.numActive;
        if (int64_t(ps.ownedSamples) > fitCapacity) {
            if (fitDirw) cudaFree(fitDirw);
            if (fitPdfDist) cudaFree(fitPdfDist);
            fitDirw = nullptr; fitPdfDist = nullptr; fitCapacity = 0;
            const size_t n = size_t(ps.ownedSamples) + size_t(ps.ownedSamples) / 8 + 1024;      // head room: the shares move a little from update to update
            G_TRY(cudaMalloc(reinterpret_cast<void **>(&fitDirw), n * sizeof(float4)));
            G_TRY(cudaMalloc(reinterpret_cast<void **>(&fitPdfDist), n * sizeof(float2)));
            fitCapacity = int64_t(n);
        }
        GPeerPtrs peers{};
        uint64_t fromPeers = 0;
        for (int s = 0; s < N; s++) if (s != me) fromPeers += ps.recvCount[s];
        bytesReceived = fromPeers * (sizeof(float4) + sizeof(float2));
        if (peerMode) {
            for (int s = 0; s < N; s++) { peers.dirw[s] = peerDirw[s]; peers.pd[s] = peerPdfDist[s]; }
            // every rank's scatter must have finished before anybody reads its buffer: a one-word all-reduce is the barrier
            G_NCCL(g_nccl.AllReduce(barrierWord, barrierWord, 1, ncclFloat, ncclMin, rc->comm, stream));
        } else {
            if (int64_t(fromPeers) > stageCapacity) {
                if (stageDirw) cudaFree(stageDirw);
                if (stagePdfDist) cudaFree(stagePdfDist);
                stageDirw = nullptr; stagePdfDist = nullptr; stageCapacity = 0;
                const size_t n = size_t(fromPeers) + size_t(fromPeers) / 8 + 1024;
                G_TRY(cudaMalloc(reinterpret_cast<void **>(&stageDirw), n * sizeof(float4)));
                G_TRY(cudaMalloc(reinterpret_cast<void **>(&stagePdfDist), n * sizeof(float2)));
                stageCapacity = int64_t(n);
            }
            for (int s = 0; s < N; s++) { peers.dirw[s] = s == me ? dirw : stageDirw; peers.pd[s] = s == me ? pdfDist : stagePdfDist; }
            G_NCCL(g_nccl.GroupStart());
            for (int d = 0; d < N; d++) {
                const size_t cnt = size_t(ps.sendStart[d + 1] - ps.sendStart[d]);
                if (d == me || !cnt) continue;
                G_NCCL(g_nccl.Send(dirw + ps.sendStart[d], cnt * sizeof(float4), ncclChar, d, rc->comm, stream));
                G_NCCL(g_nccl.Send(pdfDist + ps.sendStart[d], cnt * sizeof(float2), ncclChar, d, rc->comm, stream));
            }
            for (int s = 0; s < N; s++) {
                const size_t cnt = size_t(ps.recvCount[s]);
                if (s == me || !cnt) continue;
                G_NCCL(g_nccl.Recv(stageDirw + ps.stageStart[s], cnt * sizeof(float4), ncclChar, s, rc->comm, stream));
                G_NCCL(g_nccl.Recv(stagePdfDist + ps.stageStart[s], cnt * sizeof(float2), ncclChar, s, rc->comm, stream));
            }
            G_NCCL(g_nccl.GroupEnd());
        }
        if (ps.numSegments) {
            // blocks per segment: enough to keep ~8 blocks per SM busy on the average segment
            const uint64_t avg = ps.ownedSamples / ps.numSegments + 1;
            const unsigned perSeg = unsigned(std::max<uint64_t>(1, std::min<uint64_t>(64, avg / 2048 + 1)));
            k_pull<<<dim3(ps.numSegments, perSeg, 1), 256, 0, stream>>>(segments, peers, fitDirw, fitPdfDist);
            launches++;
        }
        fitD = fitDirw; fitP = fitPdfDist;
    }
    G_TRY(cudaEventRecord(ev[3], stream));
    int order = summationOrder;
    if (const char *e = getenv("B200PT_GUIDING_ORDER")) order = !strcmp(e, "reordered") ? 1 : 0;
    if (fitRegions && order == 0) {
        // strict order: one CTA per region, 2 CTAs per SM, regions largest first
        static bool attrSet = false;
        if (!attrSet) { G_TRY(cudaFuncSetAttribute(k_guiding_update_strict, cudaFuncAttributeMaxDynamicSharedMemorySize, int(GC_SMEM_BYTES))); attrSet = true; }
        k_guiding_update_strict<<<fitRegions, GC_BLOCK, GC_SMEM_BYTES, stream>>>(mixes, vmms, aabbs, activeRegions, &planDev->numActive, regionBegin, regionLen, fitD, fitP,
                                                                                 params, firstFit ? 1 : 0, devScalars);
        launches++;
    } else if (fitRegions && !getenv("B200PT_GUIDING_CLUSTER")) {
        // block-parallel sums (a different float summation order than the reference's), work shared between all blocks
        const int64_t rows = (fitD == dirw ? capacity : fitCapacity) / GS_CHUNK_MAX + int64_t(maxRegions) * GS_ROWS_PER_REGION + 2;
        if (rows > partialRows) {
            if (partials) cudaFree(partials);
            partials = nullptr; partialRows = 0;
            G_TRY(cudaMalloc(reinterpret_cast<void **>(&partials), size_t(rows) * G_STATACC_FLOATS * sizeof(float)));
            partialRows = rows;
        }
        G_TRY(cudaMemsetAsync(fitControl, 0, 2 * sizeof(uint32_t), stream));
        k_guiding_update_shared<<<sharedGrid, GS_BLOCK, 0, stream>>>(mixes, vmms, aabbs, activeRegions, &planDev->numActive, regionBegin, regionLen, regionSlot, fitD, fitP,
                                                                     passes, partials, fitControl, params, firstFit ? 1 : 0, devScalars);
        launches++;
    } else if (fitRegions) {
        // the round-1 kernel, kept for A/B (B200PT_GUIDING_CLUSTER=<size>): one cluster of CTAs per region
        // per region (portable maximum 8); B200PT_GUIDING_CLUSTER overrides
        int clusterSize = 4;
        if (const char *e = getenv("B200PT_GUIDING_CLUSTER")) clusterSize = std::max(1, std::min(8, atoi(e)));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(fitRegions * clusterSize, 1, 1);
        cfg.blockDim = dim3(G_BLOCK, 1, 1);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = clusterSize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const uint32_t *activePtr = activeRegions, *numActivePtr = &planDev->numActive, *beginPtr = regionBegin, *countPtr = regionLen;
        const b200pt_aabb *aabbPtr = aabbs;
        G_TRY(cudaLaunchKernelEx(&cfg, k_guiding_update, mixes, vmms, aabbPtr, activePtr, numActivePtr, beginPtr, countPtr, fitD, fitP, params, firstFit ? 1 : 0, devScalars));
        launches++;
    }
    G_TRY(cudaGetLastError());
    G_TRY(cudaEventRecord(ev[4], stream));
    if (N > 1) {
        // results: all-gather of the mixture arrays, then every region takes its owner's copy.  This also is the barrier
        // that keeps a rank from overwriting its sorted buffer (next update) while a peer still reads it.
        G_NCCL(g_nccl.GroupStart());
        G_NCCL(g_nccl.AllGather(mixes, gatherMix, size_t(R) * sizeof(GMix), ncclChar, rc->comm, stream));      // only the regions in use
        G_NCCL(g_nccl.AllGather(vmms, gatherVmm, size_t(R) * sizeof(b200pt_vmm_theta), ncclChar, rc->comm, stream));
        G_NCCL(g_nccl.GroupEnd());
        k_pick_results<<<R, 128, 0, stream>>>(mixes, vmms, gatherMix, gatherVmm, owner, totalAll, R, me);
        launches++;
    }
    G_TRY(cudaEventRecord(ev[5], stream));
    G_TRY(cudaMemcpyAsync(hostScalars, devScalars, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    if (N == 1) G_TRY(cudaMemcpyAsync(planHost, planDev, sizeof(GPlanSummary), cudaMemcpyDeviceToHost, stream));
    G_TRY(cudaStreamSynchronize(stream));        // the one synchronisation of a single-GPU update
    float msSort = 0, msExchange = 0, msFit = 0, msGather = 0;
    cudaEventElapsedTime(&msSort, ev[0], ev[1]);
    cudaEventElapsedTime(&msExchange, ev[1], ev[3]);
    cudaEventElapsedTime(&msFit, ev[3], ev[4]);
    cudaEventElapsedTime(&msGather, ev[4], ev[5]);
    firstFit = false;
    lastValidSamples = planHost->localValid;
    if (params.splitRegions) { int rcs = splitRegions(params, stream); if (rcs != B200PT_OK) return rcs; }
    if (getenv("B200PT_GUIDING_PROFILE")) {   // per-region cost distribution (development aid)
        std::vector<GMix> hm;
        hm.resize(size_t(regionCount));
        cudaMemcpy(hm.data(), mixes, hm.size() * sizeof(GMix), cudaMemcpyDeviceToHost);
        std::vector<uint32_t> kc, it;
        for (const GMix &m : hm) { kc.push_back(m.lastUpdateKCycles); it.push_back(m.numEMIterations); }
        std::sort(kc.begin(), kc.end()); std::sort(it.begin(), it.end());
        double sum = 0; for (uint32_t v : kc) sum += v;
        fprintf(stderr, "[guiding profile] regions %d fit %.2f ms | per-region Mcycles: min %.2f median %.2f p90 %.2f max %.2f sum %.1f | EM iterations (cumulative): min %u median %u max %u\n",
                regionCount, msFit, kc.front() / 1024.0, kc[kc.size() / 2] / 1024.0, kc[kc.size() * 9 / 10] / 1024.0, kc.back() / 1024.0, sum / 1024.0,
                it.front(), it[it.size() / 2], it.back());
    }
    if (stats) {
        stats->guiding_samples += planHost->ownedSamples;
        stats->guiding_samples_all_ranks += planHost->totalSamples;
        stats->guiding_bytes_received += bytesReceived;
        stats->guiding_em_sample_iterations += hostScalars[0];
        stats->guiding_regions_fit += planHost->numActive;
        stats->ms_guiding_sort += msSort;
        stats->ms_guiding_exchange += msExchange;
        stats->ms_guiding_fit += msFit;
        stats->ms_guiding_gather += msGather;
        stats->kernel_launches += launches;
        stats->launches_guiding += launches;
    }
    return B200PT_OK;
}

// "Check for regions to split" of PathGuiding::update (src/PathGuiding.cpp:291-300): every region whose mixture has
// seen more than samplesForRegionSplit samples is halved along its longest axis; it keeps the left half, the right
// half is appended as a new region that starts from a copy of the mixture (splitRegion, :328-348)
int GuidingState::splitRegions(const b200pt_guiding_params &params, cudaStream_t stream) {
    std::vector<float> numSamples;
    numSamples.resize(size_t(regionCount));
    G_TRY(cudaMemcpy2DAsync(numSamples.data(), sizeof(float), &mixes[0].numSamples, sizeof(GMix), sizeof(float), size_t(regionCount), cudaMemcpyDeviceToHost, stream));
    G_TRY(cudaStreamSynchronize(stream));
    std::vector<int2> pairs;
    const int currentRegionCount = regionCount;
    for (int r = 0; r < currentRegionCount; r++) {
        if (!(numSamples[size_t(r)] > params.samplesForRegionSplit)) continue;
        if (regionCount >= maxRegions) break;          // the per-region buffers are full: refinement stops there
        b200pt_aabb l, rr;
        splitAabb(hostAabbs[size_t(r)], l, rr);
        hostAabbs[size_t(r)] = l;
        hostAabbs.push_back(rr);
        const int nr = regionCount++;
        hostSpawnNext[size_t(nr)] = hostSpawnFirst[size_t(r)];
        hostSpawnFirst[size_t(r)] = nr;
        pairs.push_back(make_int2(r, nr));
    }
    if (pairs.empty()) return B200PT_OK;
    G_TRY(cudaMemcpyAsync(splitPairs, pairs.data(), pairs.size() * sizeof(int2), cudaMemcpyHostToDevice, stream));
    k_guiding_split_regions<<<unsigned(pairs.size()), 128, 0, stream>>>(mixes, vmms, splitPairs);
    G_TRY(cudaGetLastError());
    G_TRY(cudaMemcpyAsync(aabbs, hostAabbs.data(), size_t(regionCount) * sizeof(b200pt_aabb), cudaMemcpyHostToDevice, stream));
    G_TRY(cudaMemcpyAsync(spawnFirst, hostSpawnFirst.data(), size_t(maxRegions) * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    G_TRY(cudaMemcpyAsync(spawnNext, hostSpawnNext.data(), size_t(maxRegions) * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    G_TRY(cudaStreamSynchronize(stream));
    hasSpawns = true;
    return B200PT_OK;
}

// checkpoint of the guiding state (b200pt_save_state / b200pt_load_state)
int GuidingState::save(FILE *f, cudaStream_t stream) {
    if (!ready) { error = "guiding state not initialised"; return B200PT_E_STATE; }
    const int32_t hdr[6] = {splits, regionCount, maxRegions, firstFit ? 1 : 0, hasSpawns ? 1 : 0, int32_t(sizeof(GMix))};
    std::vector<char> mixBuf(size_t(regionCount) * sizeof(GMix));
    std::vector<b200pt_vmm_theta> vmmBuf;
    vmmBuf.resize(size_t(regionCount));
    G_TRY(cudaMemcpyAsync(mixBuf.data(), mixes, mixBuf.size(), cudaMemcpyDeviceToHost, stream));
    G_TRY(cudaMemcpyAsync(vmmBuf.data(), vmms, vmmBuf.size() * sizeof(b200pt_vmm_theta), cudaMemcpyDeviceToHost, stream));
    G_TRY(cudaStreamSynchronize(stream));
    bool ok = fwrite(hdr, sizeof(hdr), 1, f) == 1 && fwrite(&lastParams, sizeof(lastParams), 1, f) == 1 &&
              fwrite(hostAabbs.data(), sizeof(b200pt_aabb), size_t(regionCount), f) == size_t(regionCount) &&
              fwrite(hostSpawnFirst.data(), sizeof(int32_t), size_t(regionCount), f) == size_t(regionCount) &&
              fwrite(hostSpawnNext.data(), sizeof(int32_t), size_t(regionCount), f) == size_t(regionCount) &&
              fwrite(mixBuf.data(), 1, mixBuf.size(), f) == mixBuf.size() &&
              fwrite(vmmBuf.data(), sizeof(b200pt_vmm_theta), vmmBuf.size(), f) == vmmBuf.size();
    if (!ok) { error = "write failed"; return B200PT_E_IO; }
    return B200PT_OK;
}
// checkpoint, part 1: read everything into host memory and validate it (nothing is touched on failure)
int GuidingState::readCheckpoint(FILE *f, GuidingCheckpoint &ck) {
    if (!ready) { error = "guiding state not initialised"; return B200PT_E_STATE; }
    int32_t hdr[6];
    if (fread(hdr, sizeof(hdr), 1, f) != 1) { error = "truncated checkpoint"; return B200PT_E_IO; }
    if (hdr[0] != splits || hdr[2] != maxRegions || hdr[5] != int32_t(sizeof(GMix)) || hdr[1] < (1 << splits) || hdr[1] > maxRegions) {
        error = "checkpoint belongs to a different guiding configuration"; return B200PT_E_INVALID;
    }
    const int n = hdr[1];
    ck.regionCount = n; ck.firstFit = hdr[3] != 0; ck.hasSpawns = hdr[4] != 0;
    ck.aabbs.resize(size_t(n));
    ck.spawnFirst.assign(size_t(maxRegions), -1); ck.spawnNext.assign(size_t(maxRegions), -1);
    ck.mixes.resize(size_t(n) * sizeof(GMix));
    ck.vmms.resize(size_t(n));
    bool ok = fread(&ck.params, sizeof(ck.params), 1, f) == 1 && fread(ck.aabbs.data(), sizeof(b200pt_aabb), size_t(n), f) == size_t(n) &&
              fread(ck.spawnFirst.data(), sizeof(int32_t), size_t(n), f) == size_t(n) && fread(ck.spawnNext.data(), sizeof(int32_t), size_t(n), f) == size_t(n) &&
              fread(ck.mixes.data(), 1, ck.mixes.size(), f) == ck.mixes.size() && fread(ck.vmms.data(), sizeof(b200pt_vmm_theta), ck.vmms.size(), f) == ck.vmms.size();
    if (!ok) { error = "truncated checkpoint"; return B200PT_E_IO; }
    // what the device code indexes with: spawn links, component counts
    for (int r = 0; r < n; r++) {
        const int32_t a = ck.spawnFirst[size_t(r)], b = ck.spawnNext[size_t(r)];
        if (a < -1 || a >= n || b < -1 || b >= n) { error = "checkpoint: region spawn link out of range"; return B200PT_E_INVALID; }
        GMix m;
        memcpy(&m, ck.mixes.data() + size_t(r) * sizeof(GMix), sizeof(GMix));
        if (m.K < 1 || m.K > G_MAXK) { error = "checkpoint: mixture component count out of range"; return B200PT_E_INVALID; }
        const int used = ck.vmms[size_t(r)].usedDistributions;
        if (used < 0 || used > G_MAXK) { error = "checkpoint: VMM_Theta component count out of range"; return B200PT_E_INVALID; }
    }
    if (ck.params.numInitialComponents < 1 || ck.params.numInitialComponents > G_MAXK) { error = "checkpoint: bad guiding parameters"; return B200PT_E_INVALID; }
    return B200PT_OK;
}
// checkpoint, part 2: upload validated data
int GuidingState::applyCheckpoint(const GuidingCheckpoint &ck, cudaStream_t stream) {
    const int n = ck.regionCount;
    regionCount = n; firstFit = ck.firstFit; hasSpawns = ck.hasSpawns; lastParams = ck.params;
    hostAabbs = ck.aabbs; hostSpawnFirst = ck.spawnFirst; hostSpawnNext = ck.spawnNext;
    G_TRY(cudaMemcpyAsync(aabbs, hostAabbs.data(), size_t(n) * sizeof(b200pt_aabb), cudaMemcpyHostToDevice, stream));
    G_TRY(cudaMemcpyAsync(spawnFirst, hostSpawnFirst.data(), size_t(maxRegions) * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    G_TRY(cudaMemcpyAsync(spawnNext, hostSpawnNext.data(), size_t(maxRegions) * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    G_TRY(cudaMemcpyAsync(mixes, ck.mixes.data(), ck.mixes.size(), cudaMemcpyHostToDevice, stream));
    G_TRY(cudaMemcpyAsync(vmms, ck.vmms.data(), ck.vmms.size() * sizeof(b200pt_vmm_theta), cudaMemcpyHostToDevice, stream));
    G_TRY(cudaStreamSynchronize(stream));
    return B200PT_OK;
}

int GuidingState::getState(int region, float scalars5[5], float perComponent[14 * 16], cudaStream_t stream) {
    if (!ready || region < 0 || region >= regionCount) { error = "bad region"; return B200PT_E_INVALID; }
    GMix m;
    G_TRY(cudaMemcpyAsync(&m, mixes + region, sizeof(GMix), cudaMemcpyDeviceToHost, stream));
    G_TRY(cudaStreamSynchronize(stream));
    scalars5[0] = float(m.K); scalars5[1] = m.sampleWeight; scalars5[2] = m.numSamples; scalars5[3] = float(m.totalNumSamples); scalars5[4] = float(m.numEMIterations);
    const float *src[14] = {m.w, m.kappa, m.r, m.mux, m.muy, m.muz, m.dist, m.distSumW, m.chi, m.chiN, m.covxx, m.covyy, m.covxy, m.covSumW};
    for (int f = 0; f < 14; f++) memcpy(perComponent + f * 16, src[f], 16 * sizeof(float));
    return B200PT_OK;
}

int GuidingState::getSorted(b200pt_directional_data *out, uint32_t *offsets, const b200pt_directional_data *rawDevice, cudaStream_t stream) {
    if (!ready) { error = "guiding state not initialised"; return B200PT_E_STATE; }
    const uint32_t n = lastValidSamples;
    if (offsets) G_TRY(cudaMemcpyAsync(offsets, regionOffset, size_t(regionCount + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    if (out && n) {
        std::vector<float4> dw(n); std::vector<float2> pd(n); std::vector<uint32_t> off(size_t(regionCount) + 1);
        G_TRY(cudaMemcpyAsync(dw.data(), dirw, n * sizeof(float4), cudaMemcpyDeviceToHost, stream));
        G_TRY(cudaMemcpyAsync(pd.data(), pdfDist, n * sizeof(float2), cudaMemcpyDeviceToHost, stream));
        G_TRY(cudaMemcpyAsync(off.data(), regionOffset, off.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        G_TRY(cudaStreamSynchronize(stream));
        uint32_t r = 0;
        for (uint32_t i = 0; i < n; i++) {
            while (r + 1 < uint32_t(regionCount) && i >= off[r + 1]) r++;
            b200pt_directional_data &d = out[i];
            const b200pt_aabb &bb = hostAabbs[r];
            for (int a = 0; a < 3; a++) d.position[a] = bb.min[a] + 0.5f * (bb.max[a] - bb.min[a]);
            d.direction[0] = dw[i].x; d.direction[1] = dw[i].y; d.direction[2] = dw[i].z; d.weight = dw[i].w;
            d.pdf = pd[i].x; d.distance = pd[i].y; d.flags = r;
        }
    }
    G_TRY(cudaStreamSynchronize(stream));
    (void)rawDevice;
    return B200PT_OK;
}

void GuidingState::release() {
    closePeers();
    peerTried = false; peerCapacity = -1;
    for (void *p : {(void *)allCounts, (void *)srcStart, (void *)segments, (void *)gatherMix, (void *)gatherVmm, (void *)regionBegin, (void *)regionLen,
                    (void *)totalAll, (void *)owner, (void *)regionSlot, (void *)passes, (void *)partials, (void *)fitControl, (void *)planDev, (void *)fitDirw, (void *)fitPdfDist, (void *)stageDirw, (void *)stagePdfDist, (void *)barrierWord})
        if (p) cudaFree(p);
    allCounts = nullptr; srcStart = nullptr; segments = nullptr; gatherMix = nullptr; gatherVmm = nullptr; regionBegin = nullptr; regionLen = nullptr;
    totalAll = nullptr; owner = nullptr; regionSlot = nullptr; passes = nullptr; partials = nullptr; fitControl = nullptr; partialRows = 0; planDev = nullptr; fitDirw = nullptr; fitPdfDist = nullptr; stageDirw = nullptr; stagePdfDist = nullptr; barrierWord = nullptr;
    fitCapacity = 0; stageCapacity = 0; planRanks = 0;
    if (planHost) cudaFreeHost(planHost);
    planHost = nullptr;
    for (auto &e : ev) { if (e) cudaEventDestroy(e); e = nullptr; }
    if (aabbs) cudaFree(aabbs);
    if (levelAabbs) cudaFree(levelAabbs);
    levelAabbs = nullptr;
    if (levelSplits) cudaFree(levelSplits);
    levelSplits = nullptr;
    if (spawnFirst) cudaFree(spawnFirst);
    if (spawnNext) cudaFree(spawnNext);
    if (splitPairs) cudaFree(splitPairs);
    spawnFirst = nullptr; spawnNext = nullptr; splitPairs = nullptr; hasSpawns = false; maxRegions = 0;
    hostSpawnFirst.clear(); hostSpawnNext.clear();
    if (vmms) cudaFree(vmms);
    if (mixes) cudaFree(mixes);
    if (regionTotal) cudaFree(regionTotal);
    if (regionOffset) cudaFree(regionOffset);
    if (activeRegions) cudaFree(activeRegions);
    if (devScalars) cudaFree(devScalars);
    if (hostScalars) cudaFreeHost(hostScalars);
    if (dirw) cudaFree(dirw);
    if (pdfDist) cudaFree(pdfDist);
    if (srcIndex) cudaFree(srcIndex);
    if (tileCounts) cudaFree(tileCounts);
    aabbs = nullptr; vmms = nullptr; mixes = nullptr; regionTotal = nullptr; regionOffset = nullptr; activeRegions = nullptr;
    devScalars = nullptr; hostScalars = nullptr; dirw = nullptr; pdfDist = nullptr; srcIndex = nullptr; tileCounts = nullptr;
    capacity = 0; ready = false; regionCount = 0; firstFit = true; lastValidSamples = 0;
    hostAabbs.clear();
}

}  // namespace b200pt
