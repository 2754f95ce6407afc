// Kernels around the irradiance cache: frame-start snapshot + lookup grid, ordered pixel compaction, and the cache
// build (update before the frame's paths, create after them) — see ic_device.cuh for the frame semantic.
#pragma once
#include "wavefront.cuh"

namespace b200pt {

// snapshot header (device): what the frame's lookups see
enum { ICH_COUNT = 0, ICH_MAX = 1, ICH_NEXT_UPDATE = 2, ICH_GRID_TOTAL = 3, ICH_LIST_COUNT = 4, ICH_NEXT_CACHE = 5, ICH_NUM = 8 };

struct ICBuffers {      // raw device pointers of one context (host keeps ownership, api.cu)
    b200pt_cache_header *header;      // live, binding 13 header
    b200pt_cache_data *data;          // live, binding 13
    b200pt_sphere *spheres;           // live, binding 12
    float4 *snapSphere, *snapNormalR, *snapColor, *snapRot, *snapTrans;
    uint2 *ranges;                    // per snapshot entry: packed cell range (lo xyz bytes, hi xyz bytes)
    uint32_t *cellCount, *cellStart, *cellItems;
    float4 *cellSpheres;              // per cell-list item: center.xyz, radius
    uint32_t *snapHdr;                // ICH_*
    uint32_t *blockCounts;            // ordered compaction scratch
    uint32_t *list;                   // compacted pixel ids (pixel order)
    int32_t *updSlot;                 // per list entry: cache index to update or -1
    uint32_t *validFlags, *validOffsets;   // create: per (list entry, k) 1 when the entry is to be appended; exclusive scan
    float4 *pending;                  // per list entry (update) / per (list entry, k) (create): 3 float4
    int icSize;
    int numCells;
};

// ---- snapshot + grid ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ic_snapshot(ICBuffers b, ICView grid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t next = b.header->nextCacheSlot, mx = b.header->maxCaches;
    const uint32_t count = min(min(next, mx), uint32_t(b.icSize));
    if (i == 0) { b.snapHdr[ICH_COUNT] = count; b.snapHdr[ICH_MAX] = mx; b.snapHdr[ICH_NEXT_UPDATE] = b.header->nextUpdateSlot; b.snapHdr[ICH_GRID_TOTAL] = 0; b.snapHdr[ICH_LIST_COUNT] = 0; b.snapHdr[ICH_NEXT_CACHE] = next; }
    if (uint32_t(i) >= count) return;
    const b200pt_sphere s = b.spheres[i];
    const b200pt_cache_data d = b.data[i];
    b.snapSphere[i] = make_float4(s.center[0], s.center[1], s.center[2], s.radius);
    b.snapNormalR[i] = make_float4(d.normal[0], d.normal[1], d.normal[2], d.harmonicR);
    b.snapColor[i] = make_float4(d.color[0], d.color[1], d.color[2], __uint_as_float(d.numUpdates));
    b.snapRot[i] = make_float4(d.rotGrad[0], d.rotGrad[1], d.rotGrad[2], 0.0f);
    b.snapTrans[i] = make_float4(d.transGrad[0], d.transGrad[1], d.transGrad[2], 0.0f);
    // conservative cell range: the containment test is length(p - c) <= r in floats, so pad the radius slightly
    const float rp = s.radius * 1.0001f + 1e-6f;
    uint32_t lo = 0, hi = 0;
    for (int a = 0; a < 3; a++) {
        lo |= uint32_t(icCellCoord(grid, a, s.center[a] - rp)) << (8 * a);
        hi |= uint32_t(icCellCoord(grid, a, s.center[a] + rp)) << (8 * a);
    }
    b.ranges[i] = make_uint2(lo, hi);
}

// one thread per cell; the entries are streamed through shared memory in index order, so every cell list is ascending
template <bool FILL>
__global__ void __launch_bounds__(256) k_ic_cells(ICBuffers b, ICView grid) {
    __shared__ uint2 tile[1024];
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t count = b.snapHdr[ICH_COUNT];
    const int cx = cell % grid.dim[0], cy = (cell / grid.dim[0]) % grid.dim[1], cz = cell / (grid.dim[0] * grid.dim[1]);
    const uint32_t cp = uint32_t(cx) | (uint32_t(cy) << 8) | (uint32_t(cz) << 16);
    const bool valid = cell < b.numCells;
    uint32_t n = 0;
    uint32_t out = (FILL && valid) ? b.cellStart[cell] : 0u;
    for (uint32_t base = 0; base < count; base += 1024) {
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < 1024 && base + k < count; k += blockDim.x) tile[k] = b.ranges[base + k];
        __syncthreads();
        const uint32_t m = min(1024u, count - base);
        if (valid) {
            for (uint32_t k = 0; k < m; k++) {
                const uint2 r = tile[k];
                if ((__vcmpgeu4(cp, r.x) & __vcmpleu4(cp, r.y)) == 0xffffffffu) {
                    if (FILL) { b.cellItems[out] = base + k; b.cellSpheres[out] = b.snapSphere[base + k]; out++; } else n++;
                }
            }
        }
    }
    if (!FILL && valid) b.cellCount[cell] = n;
}

// exclusive scan of `n` counters by one block; out[n] = total (also stored to *total when given)
__global__ void __launch_bounds__(1024) k_scan_single_block(const uint32_t *in, uint32_t *out, int n, uint32_t *total) {   // in == out is allowed
    __shared__ uint32_t partial[1024];
    const int per = (n + 1023) / 1024;
    const int b0 = threadIdx.x * per, b1 = min(n, b0 + per);
    uint32_t s = 0;
    for (int i = b0; i < b1; i++) s += in[i];
    partial[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {       // Hillis–Steele inclusive scan
        uint32_t v = threadIdx.x >= off ? partial[threadIdx.x - off] : 0u;
        __syncthreads();
        partial[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = threadIdx.x ? partial[threadIdx.x - 1] : 0u;
    for (int i = b0; i < b1; i++) { const uint32_t c = in[i]; out[i] = run; run += c; }
    if (threadIdx.x == 1023) { out[n] = partial[1023]; if (total) *total = partial[1023]; }
}

// ---- ordered compaction of pixels --------------------------------------------------------------------------------
struct PredICUpdate {      // rgen:1639: `pushC.useIrradianceCache && rnd() < pushC.irradianceUpdateProb`, first draw of the pixel's stream
    uint32_t randomUInt; float prob;
    __device__ __forceinline__ bool operator()(int p) const { uint32_t s = tea(uint32_t(p), randomUInt); return rnd(s) < prob; }
};
struct PredICCreate {      // the pixel queued new cache entries during the frame
    const uint32_t *newCount;
    __device__ __forceinline__ bool operator()(int p) const { return newCount[p] != 0u; }
};

template <typename Pred>
__global__ void __launch_bounds__(256) k_flag_count(Pred pred, int numPixels, uint32_t *blockCounts) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = __syncthreads_count(p < numPixels && pred(p));
    if (threadIdx.x == 0) blockCounts[blockIdx.x] = uint32_t(c);
}
template <typename Pred>
__global__ void __launch_bounds__(256) k_flag_fill(Pred pred, int numPixels, const uint32_t *__restrict__ blockOffsets, uint32_t *list) {
    __shared__ uint32_t warpCount[8];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool f = p < numPixels && pred(p);
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) warpCount[warp] = __popc(bal);
    __syncthreads();
    uint32_t off = blockOffsets[blockIdx.x];
    for (int w = 0; w < warp; w++) off += warpCount[w];
    if (f) list[off + __popc(bal & ((1u << lane) - 1u))] = uint32_t(p);
}

// ---- cache update (before the frame's paths): rgen:1334-1381 ------------------------------------------------------
// update slots in pixel order (`header.nextUpdateSlot++`, reset when it runs past the last entry)
__global__ void k_ic_update_assign(ICBuffers b) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t n = b.snapHdr[ICH_LIST_COUNT];
    const uint32_t count = b.snapHdr[ICH_COUNT];           // = header.nextCacheSlot clipped to the buffer
    const uint32_t nextCacheSlot = b.header->nextCacheSlot;
    uint32_t slot = b.header->nextUpdateSlot;
    for (uint32_t k = 0; k < n; k++) {
        int32_t assigned = -1;
        const uint32_t cacheIndex = slot++;
        if (cacheIndex >= nextCacheSlot) slot = 0;
        else if (cacheIndex < count && __float_as_uint(b.snapColor[cacheIndex].w) >= 1u) assigned = int32_t(cacheIndex);
        b.updSlot[k] = assigned;
    }
    b.header->nextUpdateSlot = slot;
}

// Launch shape of the two build kernels (buildLaunchShape, api.cu): `entryStride` threads per entry slot, of which the first
// `lanes` work on the entry.  lanes = 1: the entry's 200 paths run on one lane (entryStride 32 = a warp of its own ... 1 = every
// lane an entry, for the long lists of the first prepare frames).  lanes = 8 (entryStride 32 / 16 / 8): eight lanes run the SAME
// entry in lockstep — same seed, same control flow, every ray traced by the group (traceRayGroup) — and lane 0 of the group stores.
__global__ void __launch_bounds__(128) k_ic_update(FrameParams fp, DeviceScene sc, Wavefront wf, ICBuffers b, int entryStride, int lanes) {
    __shared__ uint2 stack[PT_STACK_SMEM * 128];
    const uint32_t n = b.snapHdr[ICH_LIST_COUNT];
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid % uint32_t(entryStride) >= uint32_t(lanes)) return;
    const bool group = lanes == 8, leader = gtid % uint32_t(entryStride) == 0u;
    for (uint32_t e = gtid / uint32_t(entryStride); e < n; e += gridDim.x * blockDim.x / uint32_t(entryStride)) {
        const int pid = int(b.list[e]);
        uint32_t seed = tea(uint32_t(pid), fp.pc.randomUInt);
        rnd(seed);                                              // the draw that selected this pixel
        const int32_t cacheIndex = b.updSlot[e];
        if (cacheIndex >= 0) {
            InlineTracer tr(fp.pc, sc, wf.ic, wf.guide, stack + threadIdx.x, blockDim.x, pid, seed, group);
            vec3 color, rotGrad, transGrad;
            const float harmonicR = tr.calculateCacheData(make_vec3(b.snapSphere[cacheIndex]), make_vec3(b.snapNormalR[cacheIndex]), color, rotGrad, transGrad);
            seed = tr.seed;
            if (leader) {
                b.pending[e * 3 + 0] = make_f4(color, harmonicR);
                b.pending[e * 3 + 1] = make_f4(rotGrad, 0.0f);
                b.pending[e * 3 + 2] = make_f4(transGrad, 0.0f);
                atomicAdd(&wf.dstats[DST_EXTEND], (unsigned long long)tr.extendRays);
                atomicAdd(&wf.dstats[DST_SHADOW], (unsigned long long)tr.shadowRays);
                atomicAdd(&wf.dstats[DST_VERTICES], (unsigned long long)tr.vertices);
            }
        }
        if (leader) wf.seed[pid] = seed;                        // k_generate continues this pixel's stream from here
    }
}

// blend into the live arrays in pixel order (rgen:1351-1380; the old values are the frame-start snapshot)
__global__ void k_ic_update_commit(FrameParams fp, ICBuffers b) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t n = b.snapHdr[ICH_LIST_COUNT];
    for (uint32_t e = 0; e < n; e++) {
        const int32_t ci = b.updSlot[e];
        if (ci < 0) continue;
        const float4 p0 = b.pending[e * 3 + 0];
        const uint32_t numUpdates = __float_as_uint(b.snapColor[ci].w);
        const float a = fminf(float(numUpdates) / float(numUpdates + 1u), 0.95f);
        const vec3 color = mix(make_vec3(p0), make_vec3(b.snapColor[ci]), a);
        vec3 rotGrad = mix(make_vec3(b.pending[e * 3 + 1]), make_vec3(b.snapRot[ci]), a);
        vec3 transGrad = mix(make_vec3(b.pending[e * 3 + 2]), make_vec3(b.snapTrans[ci]), a);
        clampGradients(fp.pc.irradianceGradientsMaxLength, rotGrad, transGrad);
        const float oldR = b.snapNormalR[ci].w;
        float harmonicR = p0.w;
        if (harmonicR > 0.0f) harmonicR = mixf(harmonicR, oldR, a); else harmonicR = oldR;
        harmonicR = fmaxf(harmonicR, fp.pc.irradianceCacheMinRadius);
        b200pt_cache_data &cd = b.data[ci];
        cd.color[0] = color.x; cd.color[1] = color.y; cd.color[2] = color.z;
        cd.rotGrad[0] = rotGrad.x; cd.rotGrad[1] = rotGrad.y; cd.rotGrad[2] = rotGrad.z;
        cd.transGrad[0] = transGrad.x; cd.transGrad[1] = transGrad.y; cd.transGrad[2] = transGrad.z;
        cd.harmonicR = harmonicR;
        b.spheres[ci].radius = fp.pc.irradianceA * harmonicR;
        cd.numUpdates = numUpdates + 1u;
    }
}

// ---- cache creation (after the frame's last path): rgen:1821-1827 + :1383-1421 ------------------------------------
__global__ void __launch_bounds__(128) k_ic_create(FrameParams fp, DeviceScene sc, Wavefront wf, ICBuffers b, int entryStride, int lanes) {
    __shared__ uint2 stack[PT_STACK_SMEM * 128];
    const uint32_t n = b.snapHdr[ICH_LIST_COUNT];
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid % uint32_t(entryStride) >= uint32_t(lanes)) return;
    const bool group = lanes == 8, leader = gtid % uint32_t(entryStride) == 0u;
    const bool full = b.header->nextCacheSlot > b.header->maxCaches;   // rgen:1384 (the live header only changes in k_ic_create_commit)
    for (uint32_t e = gtid / uint32_t(entryStride); e < n; e += gridDim.x * blockDim.x / uint32_t(entryStride)) {
        const int pid = int(b.list[e]);
        const uint32_t cnt = min(wf.ic.newCount[pid], uint32_t(IC_MAX_NEW));
        const uint32_t seed0 = wf.seed[pid];
        if (group) __syncwarp(0xffu << (threadIdx.x & 24u));    // every lane of the group has read the seed before the leader stores the new one
        InlineTracer tr(fp.pc, sc, wf.ic, wf.guide, stack + threadIdx.x, blockDim.x, pid, seed0, group);
        for (uint32_t k = 0; k < IC_MAX_NEW; k++) {
            float4 *pe = b.pending + (size_t(e) * IC_MAX_NEW + k) * 3;
            if (k >= cnt || full) { if (leader) b.validFlags[size_t(e) * IC_MAX_NEW + k] = 0u; continue; }
            const float4 o = wf.ic.newEntries[(size_t(pid) * IC_MAX_NEW + k) * 2 + 0], nn = wf.ic.newEntries[(size_t(pid) * IC_MAX_NEW + k) * 2 + 1];
            vec3 color, rotGrad, transGrad;
            const float harmonicR = tr.calculateCacheData(make_vec3(o), make_vec3(nn), color, rotGrad, transGrad);
            if (!leader) continue;
            pe[0] = make_f4(color, harmonicR);
            pe[1] = make_f4(rotGrad, 0.0f);
            pe[2] = make_f4(transGrad, 0.0f);
            b.validFlags[size_t(e) * IC_MAX_NEW + k] = harmonicR < 0.0f ? 0u : 1u;      // rgen:1391-1393
        }
        if (!leader) continue;
        wf.seed[pid] = tr.seed;
        atomicAdd(&wf.dstats[DST_EXTEND], (unsigned long long)tr.extendRays);
        atomicAdd(&wf.dstats[DST_SHADOW], (unsigned long long)tr.shadowRays);
        atomicAdd(&wf.dstats[DST_VERTICES], (unsigned long long)tr.vertices);
    }
}

// cache slots in pixel order (`header.nextCacheSlot++`): the j-th valid entry of the ordered list gets slot start + j
// (validOffsets = exclusive scan of validFlags), including the reference's off-by-one at the end of the buffer
// (rgen:1384,1397: the slot counter may reach maxCaches + 1, the entry at index maxCaches is dropped — quirk 11)
__global__ void __launch_bounds__(256) k_ic_create_commit(FrameParams fp, Wavefront wf, ICBuffers b) {
    const uint32_t n = b.snapHdr[ICH_LIST_COUNT] * IC_MAX_NEW;
    const uint32_t start = b.snapHdr[ICH_NEXT_CACHE], maxCaches = b.snapHdr[ICH_MAX];
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0 && start <= maxCaches) b.header->nextCacheSlot = min(start + b.validOffsets[n], maxCaches + 1u);
    if (idx >= n || start > maxCaches || !b.validFlags[idx]) return;
    const uint32_t ci = start + b.validOffsets[idx];
    if (ci > maxCaches || ci >= uint32_t(b.icSize)) return;
    const uint32_t e = idx / IC_MAX_NEW, k = idx % IC_MAX_NEW;
    const int pid = int(b.list[e]);
    const float4 *pe = b.pending + size_t(idx) * 3;
    const float harmonicR = fmaxf(pe[0].w, fp.pc.irradianceCacheMinRadius);
    vec3 rotGrad = make_vec3(pe[1]), transGrad = make_vec3(pe[2]);
    clampGradients(fp.pc.irradianceGradientsMaxLength, rotGrad, transGrad);
    const float4 o = wf.ic.newEntries[(size_t(pid) * IC_MAX_NEW + k) * 2 + 0], nn = wf.ic.newEntries[(size_t(pid) * IC_MAX_NEW + k) * 2 + 1];
    b200pt_cache_data &cd = b.data[ci];
    cd.normal[0] = nn.x; cd.normal[1] = nn.y; cd.normal[2] = nn.z;
    cd.color[0] = pe[0].x; cd.color[1] = pe[0].y; cd.color[2] = pe[0].z;
    cd.harmonicR = harmonicR;
    cd.rotGrad[0] = rotGrad.x; cd.rotGrad[1] = rotGrad.y; cd.rotGrad[2] = rotGrad.z;
    cd.transGrad[0] = transGrad.x; cd.transGrad[1] = transGrad.y; cd.transGrad[2] = transGrad.z;
    cd.numUpdates = 1u;
    b200pt_sphere &s = b.spheres[ci];
    s.center[0] = o.x; s.center[1] = o.y; s.center[2] = o.z;
    s.radius = fp.pc.irradianceA * harmonicR;
}

}  // namespace b200pt
