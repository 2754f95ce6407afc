// Software ray traversal of the compressed BVH8 (host/bvh.h) — replaces the RT-core traceRayEXT calls of the
// reference (shaders/raytrace.rgen:1011-1022 closest hit, :626-637 shadow) and the analytic sphere intersection
// shader (shaders/raytrace.sphere.rint:11-29).
//
// Semantics (ours; the reference leaves them to the driver, SURVEY.md §7 "hard parts"):
//   - triangles: Möller–Trumbore with every product/sum individually rounded (no FMA contraction) so the CPU oracle
//     reproduces t,u,v bit for bit; a hit needs tmin < t < tmax (exclusive, Vulkan triangle rule)
//   - spheres: both roots reported far-then-near like the .rint shader; accepted when tmin <= t <= tmax
//   - closest hit = lexicographic minimum of (t, primitive id); primitive ids: triangles of instance 0,1,.. then spheres
//   - node boxes are dilated on the host (Bvh8::pad) so culling can compare against the current best t exactly
#pragma once
#include "device_math.cuh"

namespace b200pt {

// what the stochastic alpha test of textured triangles needs (raytrace.rahit:22-46); nullptr in TraceScene when no
// texture of the scene has a texel with alpha < 1 (then no triangle carries the alpha flag either)
struct AlphaTexture { const float4 *texels; int width, height; };
struct AlphaScene {
    const int4 *primVerts;            // per global triangle: global vertex ids v0,v1,v2 + instance
    const float4 *vertices;           // b200pt_vertex as 3 float4: pos|-, normal|-, texCoord.xy materialIndex -
    const int *materialTexture;       // per material: textureIdDiffuse
    const AlphaTexture *textures;
};

struct TraceScene {
    const float4 *nodes;      // 5 float4 per Bvh8Node
    const float4 *tris;       // 3 float4 per PackedTri (e1.w != 0: triangle of an alpha-tested material)
    const float4 *spheres;    // center.xyz, radius
    const AlphaScene *alpha;
    uint32_t numTris;
    uint32_t numSpheres;
    uint32_t alphaSeed;       // pushC.randomUInt of the frame (the any-hit shader seeds its own stream with it)
};

struct HitRec { float t; uint32_t prim; float u, v; };
#define PT_MISS 0xFFFFFFFFu

#define PT_STACK_SMEM 8       // entries per thread kept in shared memory
#define PT_STACK_LOCAL 24     // overflow entries in local memory
#define PT_TRACE_BLOCK 128
#ifndef PT_TRACE_MIN_BLOCKS
#define PT_TRACE_MIN_BLOCKS 7   // __launch_bounds__ minimum CTAs per SM of the trace kernels (register cap)
#endif

// exact (non-contracted) Möller–Trumbore; returns true and t,u,v when the ray hits the triangle's plane inside it
__device__ __forceinline__ bool intersectTriExact(const float4 a, const float4 b, const float4 c, const vec3 o, const vec3 d,
                                                  float &t, float &u, float &v) {
    // a = v0 (w = prim), b = e1, c = e2
    const float px = __fsub_rn(__fmul_rn(d.y, c.z), __fmul_rn(d.z, c.y));
    const float py = __fsub_rn(__fmul_rn(d.z, c.x), __fmul_rn(d.x, c.z));
    const float pz = __fsub_rn(__fmul_rn(d.x, c.y), __fmul_rn(d.y, c.x));
    const float det = __fadd_rn(__fadd_rn(__fmul_rn(b.x, px), __fmul_rn(b.y, py)), __fmul_rn(b.z, pz));
    if (det == 0.0f) return false;
    const float inv = __fdiv_rn(1.0f, det);
    const float tx = __fsub_rn(o.x, a.x), ty = __fsub_rn(o.y, a.y), tz = __fsub_rn(o.z, a.z);
    u = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(tx, px), __fmul_rn(ty, py)), __fmul_rn(tz, pz)), inv);
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    const float qx = __fsub_rn(__fmul_rn(ty, b.z), __fmul_rn(tz, b.y));
    const float qy = __fsub_rn(__fmul_rn(tz, b.x), __fmul_rn(tx, b.z));
    const float qz = __fsub_rn(__fmul_rn(tx, b.y), __fmul_rn(ty, b.x));
    v = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(d.x, qx), __fmul_rn(d.y, qy)), __fmul_rn(d.z, qz)), inv);
    if (!(v >= 0.0f && __fadd_rn(u, v) <= 1.0f)) return false;
    t = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(c.x, qx), __fmul_rn(c.y, qy)), __fmul_rn(c.z, qz)), inv);
    return true;
}

// raytrace.sphere.rint:13-28 with individually rounded operations; returns the two roots (far first)
__device__ __forceinline__ bool intersectSphereExact(const float4 s, const vec3 o, const vec3 d, float &t1, float &t2) {
    const float ox = __fsub_rn(o.x, s.x), oy = __fsub_rn(o.y, s.y), oz = __fsub_rn(o.z, s.z);
    const float dotDOC = __fadd_rn(__fadd_rn(__fmul_rn(d.x, ox), __fmul_rn(d.y, oy)), __fmul_rn(d.z, oz));
    const float ococ = __fadd_rn(__fadd_rn(__fmul_rn(ox, ox), __fmul_rn(oy, oy)), __fmul_rn(oz, oz));
    const float rootTerm = __fadd_rn(__fsub_rn(__fmul_rn(dotDOC, dotDOC), ococ), __fmul_rn(s.w, s.w));
    if (rootTerm < 0.0f) return false;
    const float root = __fsqrt_rn(rootTerm);
    t1 = __fadd_rn(-dotDOC, root);
    t2 = __fsub_rn(-dotDOC, root);
    return true;
}

__device__ __forceinline__ uint32_t byteOf(uint32_t x, int i) { return (x >> (8 * i)) & 0xffu; }

// Slab arithmetic of the node step: bit-identical variants, measured in profiles/r01e_slab_variants.txt — both lose.
// ncu names ALU the busiest pipe of k_trace (48.6 %), so the I2F.U8 conversions (conversion pipe) and the scalar FFMAs
// (FMA pipe) are already off the critical pipe; FFMA2 halves the FMA issue slots but costs registers (28 B of spills at 72).
//   PT_SLAB_F2  1 = the six fmaf of a child as three packed fma.rn.f32x2 (Blackwell FFMA2): trace -2.4 %
//   PT_SLAB_CVT 0 = I2F.U8 with byte select, 1 = PRMT into 0x4B0000qq (= 2^23 + q) and an exact FADD of -2^23: trace -4.9 %
#ifndef PT_SLAB_F2
#define PT_SLAB_F2 0
#endif
#ifndef PT_SLAB_CVT
#define PT_SLAB_CVT 0
#endif
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    return ((unsigned long long)__float_as_uint(hi) << 32) | (unsigned long long)__float_as_uint(lo);
}
__device__ __forceinline__ void ffma2(float &rlo, float &rhi, unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    rlo = __uint_as_float(uint32_t(d)); rhi = __uint_as_float(uint32_t(d >> 32));
}
__device__ __forceinline__ float byteToFloat(uint32_t x, int i) {
#if PT_SLAB_CVT == 1
    return __fadd_rn(__uint_as_float(__byte_perm(x, 0x4B000000u, 0x7540u + uint32_t(i))), -8388608.0f);
#else
    return float(byteOf(x, i));
#endif
}

// raytrace.rahit:22-46 — any-hit shader of the (non-opaque) triangle geometry: a hit on a textured material is ignored
// when rnd(seed) > alpha, seed = tea(uint(uv.x * 1e8 + rayOrigin.x * t), pushC.randomUInt).  Runs for closest-hit and
// shadow rays alike.  Out of line: only scenes with a non-opaque texel ever get here.
__device__ __forceinline__ bool alphaRejectsInline(const TraceScene &sc, uint32_t prim, float bu, float bv, float ox, float t) {
    const AlphaScene &A = *sc.alpha;
    const int4 pv = __ldg(&A.primVerts[prim]);
    const float4 t0 = __ldg(&A.vertices[3 * pv.x + 2]), t1 = __ldg(&A.vertices[3 * pv.y + 2]), t2 = __ldg(&A.vertices[3 * pv.z + 2]);
    const int tex = __ldg(&A.materialTexture[__float_as_int(t0.z)]);      // material of v0 (quirk 4)
    if (tex == -1) return false;
    const float bx = 1.0f - bu - bv;
    const float tu = t0.x * bx + t1.x * bu + t2.x * bv;
    const float tv = t0.y * bx + t1.y * bu + t2.y * bv;
    uint32_t seed = tea(uint32_t(tu * 100000000.0f + ox * t), sc.alphaSeed);   // float -> uint saturates (negative, NaN -> 0), like the oracle
    // sampler2D: linear filter, repeat (the alpha channel only)
    const AlphaTexture T = A.textures[tex];
    const float x = tu * float(T.width) - 0.5f, y = tv * float(T.height) - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float ax = x - fx, ay = y - fy;
    int x0 = int(fx) % T.width, y0 = int(fy) % T.height;
    if (x0 < 0) x0 += T.width;
    if (y0 < 0) y0 += T.height;
    const int x1 = x0 + 1 == T.width ? 0 : x0 + 1, y1 = y0 + 1 == T.height ? 0 : y0 + 1;
    const float a = __ldg(&T.texels[y0 * T.width + x0]).w, b = __ldg(&T.texels[y0 * T.width + x1]).w;
    const float c = __ldg(&T.texels[y1 * T.width + x0]).w, d = __ldg(&T.texels[y1 * T.width + x1]).w;
    const float alpha = (a * (1 - ax) + b * ax) * (1 - ay) + (c * (1 - ax) + d * ax) * ay;
    return rnd(seed) > alpha;
}
__device__ __noinline__ bool alphaRejects(const TraceScene &sc, uint32_t prim, float bu, float bv, float ox, float t) {
    return alphaRejectsInline(sc, prim, bu, bv, ox, t);
}

// order-preserving map float -> uint32 (and back); -0 is folded onto +0 first
__device__ __forceinline__ uint32_t orderedBitsOf(float f) {
    const uint32_t b = __float_as_uint(__fadd_rn(f, 0.0f));
    return b ^ (uint32_t(int32_t(b) >> 31) | 0x80000000u);
}
__device__ __forceinline__ float orderedBitsToFloat(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

// Traversal is written as an explicit per-lane state machine (init / step) so that the persistent trace kernel can
// hand a lane a NEW ray as soon as its current one is finished (dynamic ray fetch), instead of idling until the slowest
// ray of the warp is done.  One step = pop one entry: visit a BVH8 node (5 x 128-bit loads, 8 slab tests) or test one
// group of leaf triangles (3 x 128-bit loads each).
struct Trav {
    vec3 o, d;
    float idx, idy, idz;
    float tmin, tmax, best;       // triangles need t < best  (or == best with a lower id once something was hit)
    uint32_t octInv;
    HitRec hit;
    uint2 cur;                    // node group: (base index, hit bits | imask) or triangle group (base, bits)
    int sp;
};

// returns true when the ray is already finished (no triangles, or ANY and a sphere was hit)
__device__ __forceinline__ bool travInit(Trav &s, const TraceScene &sc, const vec3 o, const vec3 d, const float tmin, const float tmax, const bool ANY) {
    s.o = o; s.d = d; s.tmin = tmin; s.tmax = tmax;
    s.hit.prim = PT_MISS; s.hit.t = tmax; s.hit.u = 0.0f; s.hit.v = 0.0f;
    s.best = tmax;
    // analytic spheres first (few per scene; kept out of the BVH)
    for (uint32_t i = 0; i < sc.numSpheres; i++) {
        float t1, t2;
        if (!intersectSphereExact(__ldg(&sc.spheres[i]), o, d, t1, t2)) continue;
        // reportIntersectionEXT(t1) then (t2): each accepted if inside [tmin, current tmax]
        const uint32_t id = sc.numTris + i;
        if (t1 >= tmin && (t1 < s.best || (t1 == s.best && (s.hit.prim == PT_MISS || id < s.hit.prim)))) { s.best = t1; s.hit.t = t1; s.hit.prim = id; }
        if (t2 >= tmin && (t2 < s.best || (t2 == s.best && (s.hit.prim == PT_MISS || id < s.hit.prim)))) { s.best = t2; s.hit.t = t2; s.hit.prim = id; }
        if (ANY && s.hit.prim != PT_MISS) return true;
    }
    if (sc.numTris == 0) return true;
    const float eps = 1e-20f;
    s.idx = 1.0f / (fabsf(d.x) > eps ? d.x : copysignf(eps, d.x));
    s.idy = 1.0f / (fabsf(d.y) > eps ? d.y : copysignf(eps, d.y));
    s.idz = 1.0f / (fabsf(d.z) > eps ? d.z : copysignf(eps, d.z));
    s.octInv = (d.x < 0.0f ? 0u : 4u) | (d.y < 0.0f ? 0u : 2u) | (d.z < 0.0f ? 0u : 1u);
    s.cur = make_uint2(0u, 0x80000000u);
    s.sp = 0;
    return false;
}

// one traversal step; returns true when the ray is finished.  The stack is split: the first PT_STACK_SMEM entries
// live in shared memory (stride = blockDim.x, conflict-free), the overflow in a per-thread local array.
// triCap bounds the triangle tests of one step: a node can hand a lane up to 24 leaf triangles while its neighbours get
// none, and the whole warp would sit through that lane's loop; the surplus goes back on the stack as a triangle group.
#ifndef PT_TRI_CAP
#define PT_TRI_CAP 0          // compile-time: 0 = no cap (test every triangle of the group in this step)
#endif
// node half of a step: pops the next child of the current node group (pushing the rest of the group), fetches that
// node, slab-tests its 8 children and returns the new node group in `cur` and the hit leaf triangles in `triGroup`
__device__ __forceinline__ void travNode(Trav &s, const TraceScene &sc, uint2 &cur, uint2 &triGroup, uint2 *smemStack, uint2 *localStack, const int stride) {
    const vec3 o = s.o, d = s.d;
    const uint32_t hits = cur.y;
    const int bit = 31 - __clz(hits);
    cur.y &= ~(1u << bit);
    if (cur.y & 0xff000000u) {   // push the rest of the group
        if (s.sp < PT_STACK_SMEM) smemStack[s.sp * stride] = cur;
        else localStack[s.sp - PT_STACK_SMEM] = cur;
        s.sp++;
    }
    const uint32_t octInv = s.octInv;
    const uint32_t octInv4 = octInv * 0x01010101u;
    const uint32_t slot = (uint32_t(bit) - 24u) ^ octInv;
    const uint32_t rel = __popc(hits & ~(0xffffffffu << slot));   // low byte of hits = imask
    const uint32_t nodeIdx = cur.x + rel;

    const float4 n0 = __ldg(&sc.nodes[nodeIdx * 5 + 0]);
    const float4 n1 = __ldg(&sc.nodes[nodeIdx * 5 + 1]);
    const float4 n2 = __ldg(&sc.nodes[nodeIdx * 5 + 2]);
    const float4 n3 = __ldg(&sc.nodes[nodeIdx * 5 + 3]);
    const float4 n4 = __ldg(&sc.nodes[nodeIdx * 5 + 4]);
    const uint32_t e = __float_as_uint(n0.w);
    const float ax = __uint_as_float((e & 0xffu) << 23) * s.idx;
    const float ay = __uint_as_float(((e >> 8) & 0xffu) << 23) * s.idy;
    const float az = __uint_as_float(((e >> 16) & 0xffu) << 23) * s.idz;
    const float ox = (n0.x - o.x) * s.idx, oy = (n0.y - o.y) * s.idy, oz = (n0.z - o.z) * s.idz;
    const float best = s.best;
#if PT_SLAB_F2
    const unsigned long long axy = pack2(ax, ay), oxy = pack2(ox, oy), azz = pack2(az, az), ozz = pack2(oz, oz);
#endif
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t meta4 = __float_as_uint(half ? n1.w : n1.z);
        const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t innerMask4 = (isInner4 >> 4) * 0xffu;
        const uint32_t bitIndex4 = (meta4 ^ (octInv4 & innerMask4)) & 0x1f1f1f1fu;
        const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t qlox = __float_as_uint(half ? n2.y : n2.x), qloy = __float_as_uint(half ? n2.w : n2.z);
        const uint32_t qloz = __float_as_uint(half ? n3.y : n3.x), qhix = __float_as_uint(half ? n3.w : n3.z);
        const uint32_t qhiy = __float_as_uint(half ? n4.y : n4.x), qhiz = __float_as_uint(half ? n4.w : n4.z);
        const uint32_t xn = d.x < 0.0f ? qhix : qlox, xf = d.x < 0.0f ? qlox : qhix;
        const uint32_t yn = d.y < 0.0f ? qhiy : qloy, yf = d.y < 0.0f ? qloy : qhiy;
        const uint32_t zn = d.z < 0.0f ? qhiz : qloz, zf = d.z < 0.0f ? qloz : qhiz;
#pragma unroll
        for (int j = 0; j < 4; j++) {
#if PT_SLAB_F2
            float t0x, t0y, t1x, t1y, t0z, t1z;
            ffma2(t0x, t0y, pack2(byteToFloat(xn, j), byteToFloat(yn, j)), axy, oxy);
            ffma2(t1x, t1y, pack2(byteToFloat(xf, j), byteToFloat(yf, j)), axy, oxy);
            ffma2(t0z, t1z, pack2(byteToFloat(zn, j), byteToFloat(zf, j)), azz, ozz);
#else
            const float t0x = fmaf(byteToFloat(xn, j), ax, ox), t1x = fmaf(byteToFloat(xf, j), ax, ox);
            const float t0y = fmaf(byteToFloat(yn, j), ay, oy), t1y = fmaf(byteToFloat(yf, j), ay, oy);
            const float t0z = fmaf(byteToFloat(zn, j), az, oz), t1z = fmaf(byteToFloat(zf, j), az, oz);
#endif
            const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
            const float tf = fminf(fminf(t1x, t1y), fminf(t1z, best));
            if (tn <= tf) hitmask |= byteOf(childBits4, j) << byteOf(bitIndex4, j);
        }
    }
    cur.x = __float_as_uint(n1.x);
    cur.y = (hitmask & 0xff000000u) | (e >> 24);
    triGroup.x = __float_as_uint(n1.y);
    triGroup.y = hitmask & 0x00ffffffu;
}

// ALPHA compiles the alpha test of textured triangles in (scenes without a transparent texel run the plain kernel):
// 0 = none, 1 = out-of-line call (the trace kernels: keeps their register count), 2 = in line (callers that must not
// contain an ABI call: a CALL anywhere in the shade kernel costs it 7 % even when it is never executed)
template <int ALPHA>
__device__ __forceinline__ bool travStep(Trav &s, const TraceScene &sc, uint2 *smemStack, uint2 *localStack, const int stride, const bool ANY) {
    const vec3 o = s.o, d = s.d;
    uint2 cur = s.cur;
    uint2 triGroup;
    if (cur.y & 0xff000000u) {
        travNode(s, sc, cur, triGroup, smemStack, localStack, stride);
    } else {
        triGroup = cur;
        cur = make_uint2(0u, 0u);
    }

#if PT_TRI_CAP > 0
#pragma unroll
    for (int it = 0; it < PT_TRI_CAP; it++) {
        if (!triGroup.y) break;
#else
    while (triGroup.y) {
#endif
        const int ti = __ffs(triGroup.y) - 1;
        triGroup.y &= triGroup.y - 1;
        const uint32_t base = (triGroup.x + uint32_t(ti)) * 3u;
        const float4 a = __ldg(&sc.tris[base + 0]);
        const float4 b = __ldg(&sc.tris[base + 1]);
        const float4 c = __ldg(&sc.tris[base + 2]);
        float t, u, v;
        if (intersectTriExact(a, b, c, o, d, t, u, v)) {
            const uint32_t id = __float_as_uint(a.w);
            if (t > s.tmin && t < s.tmax && (t < s.best || (t == s.best && id < s.hit.prim))) {
                if (!ALPHA || __float_as_uint(b.w) == 0u || !(ALPHA == 2 ? alphaRejectsInline(sc, id, u, v, o.x, t) : alphaRejects(sc, id, u, v, o.x, t))) {
                    s.best = t; s.hit.t = t; s.hit.prim = id; s.hit.u = u; s.hit.v = v;
                    if (ANY) return true;
                }
            }
        }
    }

#if PT_TRI_CAP > 0
    if (triGroup.y) {            // budget exhausted: continue with these triangles in the next step
        if (cur.y & 0xff000000u) {
            if (s.sp < PT_STACK_SMEM) smemStack[s.sp * stride] = cur;
            else localStack[s.sp - PT_STACK_SMEM] = cur;
            s.sp++;
        }
        s.cur = triGroup;
        return false;
    }
#endif
    if ((cur.y & 0xff000000u) == 0) {
        if (s.sp == 0) return true;
        s.sp--;
        if (s.sp < PT_STACK_SMEM) cur = smemStack[s.sp * stride];
        else cur = localStack[s.sp - PT_STACK_SMEM];
    }
    s.cur = cur;
    return false;
}

// whole-ray convenience wrapper (used for the rare in-line visibility rays of the shade kernel)
template <bool ANY, int ALPHA>
__device__ __forceinline__ void traceRayT(const TraceScene &sc, const vec3 o, const vec3 d, const float tmin, const float tmax,
                                          HitRec &hit, uint2 *smemStack, const int stride) {
    Trav s;
    uint2 localStack[PT_STACK_LOCAL];
    bool done = travInit(s, sc, o, d, tmin, tmax, ANY);
    while (!done) done = travStep<ALPHA>(s, sc, smemStack, localStack, stride, ANY);
    hit = s.hit;
}
// out of line on purpose: the callers (the rare in-line visibility ray of the shade kernel, the serial cache-build
// tracer) keep their registers for their own work and share one copy of the traversal code
template <bool ANY>
__device__ __noinline__ void traceRay(const TraceScene &sc, const vec3 o, const vec3 d, const float tmin, const float tmax,
                                      HitRec &hit, uint2 *smemStack, const int stride) {
    if (sc.alpha) traceRayT<ANY, 2>(sc, o, d, tmin, tmax, hit, smemStack, stride);
    else traceRayT<ANY, 0>(sc, o, d, tmin, tmax, hit, smemStack, stride);
}
// the same, in line (no ABI call in the caller)
template <bool ANY>
__device__ __forceinline__ void traceRayInline(const TraceScene &sc, const vec3 o, const vec3 d, const float tmin, const float tmax,
                                               HitRec &hit, uint2 *smemStack, const int stride) {
    if (sc.alpha) traceRayT<ANY, 2>(sc, o, d, tmin, tmax, hit, smemStack, stride);
    else traceRayT<ANY, 0>(sc, o, d, tmin, tmax, hit, smemStack, stride);
}

// ---------------------------------------------------------------------------------------------------------------
// One ray on EIGHT lanes (the serial cache build, ic_device.cuh InlineTracer): the eight lanes of a group hold the same ray
// and the same traversal state; in a node step lane c slab-tests child c only and the hit masks are OR-reduced over the
// group, in a leaf the set triangles are dealt round-robin and the candidates merged by the closest-hit rule (minimum of
// (t, id), the same key the cooperative triangle phase uses).  Per child / per triangle the arithmetic is travNode's /
// travStep's, so the result is the per-lane traversal's bit for bit; what shrinks is the dependent instruction chain of a
// step (~230 -> ~60 instructions), and that chain IS the run time of a kernel whose parallelism is one entry per warp.
// All eight lanes must call with identical arguments, converged; `gmask` = the group's lanes.
__device__ __forceinline__ void travNodeGroup(Trav &s, const TraceScene &sc, uint2 &cur, uint2 &triGroup, uint2 *smemStack, uint2 *localStack, const int stride,
                                              const unsigned sub, const unsigned gmask) {
    const vec3 o = s.o, d = s.d;
    const uint32_t hits = cur.y;
    const int bit = 31 - __clz(hits);
    cur.y &= ~(1u << bit);
    if (cur.y & 0xff000000u) {   // push the rest of the group
        if (s.sp < PT_STACK_SMEM) smemStack[s.sp * stride] = cur;
        else localStack[s.sp - PT_STACK_SMEM] = cur;
        s.sp++;
    }
    const uint32_t octInv = s.octInv;
    const uint32_t octInv4 = octInv * 0x01010101u;
    const uint32_t slot = (uint32_t(bit) - 24u) ^ octInv;
    const uint32_t rel = __popc(hits & ~(0xffffffffu << slot));
    const uint32_t nodeIdx = cur.x + rel;
    const float4 n0 = __ldg(&sc.nodes[nodeIdx * 5 + 0]);
    const float4 n1 = __ldg(&sc.nodes[nodeIdx * 5 + 1]);
    const float4 n2 = __ldg(&sc.nodes[nodeIdx * 5 + 2]);
    const float4 n3 = __ldg(&sc.nodes[nodeIdx * 5 + 3]);
    const float4 n4 = __ldg(&sc.nodes[nodeIdx * 5 + 4]);
    const uint32_t e = __float_as_uint(n0.w);
    const float ax = __uint_as_float((e & 0xffu) << 23) * s.idx;
    const float ay = __uint_as_float(((e >> 8) & 0xffu) << 23) * s.idy;
    const float az = __uint_as_float(((e >> 16) & 0xffu) << 23) * s.idz;
    const float ox = (n0.x - o.x) * s.idx, oy = (n0.y - o.y) * s.idy, oz = (n0.z - o.z) * s.idz;
    const bool half = (sub & 4u) != 0u;
    const int j = int(sub & 3u);
    const uint32_t meta4 = __float_as_uint(half ? n1.w : n1.z);
    const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t innerMask4 = (isInner4 >> 4) * 0xffu;
    const uint32_t bitIndex4 = (meta4 ^ (octInv4 & innerMask4)) & 0x1f1f1f1fu;
    const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;
    const uint32_t qlox = __float_as_uint(half ? n2.y : n2.x), qloy = __float_as_uint(half ? n2.w : n2.z);
    const uint32_t qloz = __float_as_uint(half ? n3.y : n3.x), qhix = __float_as_uint(half ? n3.w : n3.z);
    const uint32_t qhiy = __float_as_uint(half ? n4.y : n4.x), qhiz = __float_as_uint(half ? n4.w : n4.z);
    const uint32_t xn = d.x < 0.0f ? qhix : qlox, xf = d.x < 0.0f ? qlox : qhix;
    const uint32_t yn = d.y < 0.0f ? qhiy : qloy, yf = d.y < 0.0f ? qloy : qhiy;
    const uint32_t zn = d.z < 0.0f ? qhiz : qloz, zf = d.z < 0.0f ? qloz : qhiz;
    const float t0x = fmaf(float(byteOf(xn, j)), ax, ox), t1x = fmaf(float(byteOf(xf, j)), ax, ox);
    const float t0y = fmaf(float(byteOf(yn, j)), ay, oy), t1y = fmaf(float(byteOf(yf, j)), ay, oy);
    const float t0z = fmaf(float(byteOf(zn, j)), az, oz), t1z = fmaf(float(byteOf(zf, j)), az, oz);
    const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
    const float tf = fminf(fminf(t1x, t1y), fminf(t1z, s.best));
    const uint32_t mine = tn <= tf ? byteOf(childBits4, j) << byteOf(bitIndex4, j) : 0u;
    const uint32_t hitmask = __reduce_or_sync(gmask, mine);
    cur.x = __float_as_uint(n1.x);
    cur.y = (hitmask & 0xff000000u) | (e >> 24);
    triGroup.x = __float_as_uint(n1.y);
    triGroup.y = hitmask & 0x00ffffffu;
}

template <bool ANY, int ALPHA>
__device__ __forceinline__ void traceRayGroupT(const TraceScene &sc, const vec3 o, const vec3 d, const float tmin, const float tmax,
                                               HitRec &hit, uint2 *smemStack, const int stride) {
    const unsigned lane = threadIdx.x & 31u, sub = lane & 7u;
    const unsigned gmask = 0xffu << (lane & 24u);
    Trav s;
    uint2 localStack[PT_STACK_LOCAL];
    bool done = travInit(s, sc, o, d, tmin, tmax, ANY);
    while (!done) {
        uint2 cur = s.cur;
        uint2 triGroup;
        if (cur.y & 0xff000000u) travNodeGroup(s, sc, cur, triGroup, smemStack, localStack, stride, sub, gmask);
        else { triGroup = cur; cur = make_uint2(0u, 0u); }
        if (triGroup.y) {
            // the k-th set triangle goes to lane k mod 8; every lane keeps the best candidate of its share
            const unsigned long long curKey = ((unsigned long long)orderedBitsOf(s.best) << 32) | (s.hit.prim == PT_MISS ? 0xffffffffu : s.hit.prim);
            unsigned long long key = curKey;
            float bu = 0.0f, bv = 0.0f;
            uint32_t m = triGroup.y, k = 0u;
            while (m) {
                const int ti = __ffs(m) - 1;
                m &= m - 1u;
                if ((k++ & 7u) != sub) continue;
                const uint32_t base = (triGroup.x + uint32_t(ti)) * 3u;
                const float4 a = __ldg(&sc.tris[base + 0]);
                const float4 b = __ldg(&sc.tris[base + 1]);
                const float4 c = __ldg(&sc.tris[base + 2]);
                float t, u, v;
                if (intersectTriExact(a, b, c, o, d, t, u, v) && t > s.tmin && t < s.tmax) {
                    const uint32_t id = __float_as_uint(a.w);
                    const unsigned long long cand = ((unsigned long long)orderedBitsOf(t) << 32) | id;
                    // travStep's rule: t < best, or t == best with a lower id (a miss so far: nothing to tie with, tmax is exclusive)
                    if (cand < key && (t < s.best || s.hit.prim != PT_MISS)) {
                        if (!ALPHA || __float_as_uint(b.w) == 0u || !(ALPHA == 2 ? alphaRejectsInline(sc, id, u, v, o.x, t) : alphaRejects(sc, id, u, v, o.x, t))) { key = cand; bu = u; bv = v; }
                    }
                }
            }
            unsigned long long best = key;
#pragma unroll
            for (int dl = 1; dl < 8; dl <<= 1) {
                const unsigned long long other = __shfl_xor_sync(gmask, best, dl);
                best = other < best ? other : best;
            }
            if (best != curKey) {
                const int src = __ffs(__ballot_sync(gmask, key == best) & gmask) - 1;
                s.hit.u = __shfl_sync(gmask, bu, src);
                s.hit.v = __shfl_sync(gmask, bv, src);
                s.best = s.hit.t = orderedBitsToFloat(uint32_t(best >> 32));
                s.hit.prim = uint32_t(best);
                if (ANY) break;
            }
        }
        if ((cur.y & 0xff000000u) == 0u) {
            if (s.sp == 0) break;
            s.sp--;
            if (s.sp < PT_STACK_SMEM) cur = smemStack[s.sp * stride];
            else cur = localStack[s.sp - PT_STACK_SMEM];
        }
        s.cur = cur;
    }
    hit = s.hit;
}
template <bool ANY>
__device__ __noinline__ void traceRayGroup(const TraceScene &sc, const vec3 o, const vec3 d, const float tmin, const float tmax,
                                           HitRec &hit, uint2 *smemStack, const int stride) {
    if (sc.alpha) traceRayGroupT<ANY, 2>(sc, o, d, tmin, tmax, hit, smemStack, stride);
    else traceRayGroupT<ANY, 0>(sc, o, d, tmin, tmax, hit, smemStack, stride);
}

// Persistent-warp ray loop with dynamic fetch.  `total` rays are numbered 0..total-1; warps take chunks of `chunk`
// consecutive rays from the global work counter (one atomic per chunk) and hand them to lanes as lanes fall idle: a
// lane whose ray is finished gets the warp's next ray once at least `refillMin` lanes are idle (or the whole warp is).
// IO::load(idx, o, d, tmin, tmax, any) reads ray idx, IO::store(idx, hit, any) consumes the result; results are keyed
// by ray index, so the output does not depend on which lane traced which ray.
template <int ALPHA, typename IO>
__device__ __forceinline__ void tracePersistent(const TraceScene &sc, const IO &io, const uint32_t total, uint32_t *workCounter,
                                                const uint32_t chunk, const int refillMin, uint2 *smemStack) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned ltMask = (1u << lane) - 1u;
    const int stride = blockDim.x;
    uint2 localStack[PT_STACK_LOCAL];
    Trav s;
    uint32_t rayIdx = 0, wBase = 0, wEnd = 0;
    bool active = false, any = false, exhausted = false;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle == 0xffffffffu && exhausted) break;
        if (!exhausted && (idle == 0xffffffffu || __popc(idle) >= refillMin)) {
            if (wBase >= wEnd) {
                uint32_t b = 0;
                if (lane == 0) b = atomicAdd(workCounter, chunk);
                b = __shfl_sync(0xffffffffu, b, 0);
                if (b >= total) { exhausted = true; wBase = wEnd = 0; }
                else { wBase = b; wEnd = min(b + chunk, total); }
            }
            if (!active) {
                const uint32_t idx = wBase + __popc(idle & ltMask);
                if (idx < wEnd) {
                    vec3 o, d; float tmin, tmax;
                    io.load(idx, o, d, tmin, tmax, any);
                    rayIdx = idx;
                    active = !travInit(s, sc, o, d, tmin, tmax, any);
                    if (!active) io.store(idx, s.hit, any);
                }
            }
            wBase = min(wEnd, wBase + uint32_t(__popc(idle)));
        }
        if (active) {
            if (travStep<ALPHA>(s, sc, smemStack, localStack, stride, any)) {
                io.store(rayIdx, s.hit, any);
                active = false;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Cooperative triangle phase.  In the per-lane loop above a node step hands a lane 0..24 leaf triangles (0.8 on
// average) and the warp sits through the longest lane's loop: on cornell-dielectric the triangle loop ran 4.4
// iterations per step with 6 -> 2 of 32 lanes active and took 55 % of the kernel's issue slots
// (profiles/r01c_ncu_trace_sass.txt).  Here the lanes of a warp pool the triangles their node steps produced into a
// shared-memory work list and every lane — including lanes without a ray — tests list entry `lane`, `lane + 32`, ...
// against the OWNER's ray (kept in shared memory).  Candidates are merged per owner with a 64-bit atomicMin on
// (ordered bits of t) << 32 | primitive id, which IS the closest-hit rule (lexicographic minimum of (t, id)), so the
// result does not depend on which lane tested what; the winner's u,v follow after a warp barrier.
#ifndef PT_COOP
#define PT_COOP 1
#endif
#ifndef PT_PREFETCH
#define PT_PREFETCH 0         // L1 prefetch of the next node before the pooled triangle phase: measured -5.5 % (trace 4585 -> 4333 Mrays/s), the kernel is issue bound, not waiting for nodes
#endif
#ifndef PT_COOP_CAP
#define PT_COOP_CAP 128       // work-list entries per warp and round (a warp can produce up to 32 x 24 per node step)
#endif
#ifndef PT_COOP_STEPS
#define PT_COOP_STEPS 2       // node steps per pooled triangle phase, 1..4 (more: fewer phases, but `best` shrinks later)
#endif
#ifndef PT_COOP_LIST
#define PT_COOP_LIST 2        // pooled triangle work: 0 = list in shared memory built by one merged per-lane loop, 1 = one loop per group,
                              // 2 = no list, every lane finds the (owner, triangle) of its slot itself (round 2b: the list loops ran as long as the
                              // lane with the most triangles, 11 % of the kernel's instructions at 8 of 32 lanes on cornell-dielectric, 15 % at 4 lanes
                              // on sponzaXML; trace-alone 4 562 -> 4 602 Mrays/s on config 2, sponzaXML plain frames 1 781 -> 1 857 Mrays/s)
#endif
struct CoopSmem {
    float4 rayO[PT_TRACE_BLOCK];                 // origin, tmin of the lane's current ray
    float4 rayD[PT_TRACE_BLOCK];                 // direction
    unsigned long long key[PT_TRACE_BLOCK];      // best (t, id) of the lane's ray during a triangle phase
    float2 uv[PT_TRACE_BLOCK];
    uint32_t base[PT_COOP_STEPS][PT_TRACE_BLOCK];   // first packed triangle of the lane's triangle group(s)
    uint16_t items[PT_TRACE_BLOCK / 32][PT_COOP_CAP];   // (owner lane << 2 | group) << 5 | triangle offset (one step: lane << 5 | offset)
};

// order-preserving map float -> uint32 (and back); -0 is folded onto +0 first
__device__ __forceinline__ uint32_t orderedBits(float f) {
    const uint32_t b = __float_as_uint(__fadd_rn(f, 0.0f));
    return b ^ (uint32_t(int32_t(b) >> 31) | 0x80000000u);
}
__device__ __forceinline__ float fromOrderedBits(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

// must be called by all 32 lanes; `tg[g].y` = 0 for lanes that bring no triangles in group g
template <int ALPHA>
__device__ __forceinline__ void coopTriangles(Trav &s, const TraceScene &sc, const uint2 (&tg)[PT_COOP_STEPS], CoopSmem &sm, const unsigned lane, const unsigned tid) {
    uint32_t bits[PT_COOP_STEPS];
    uint32_t anyBits = 0u;
#pragma unroll
    for (int g = 0; g < PT_COOP_STEPS; g++) { bits[g] = tg[g].y; anyBits |= bits[g]; }
    if (!__any_sync(0xffffffffu, anyBits != 0u)) return;
    const unsigned wl = tid & ~31u, warp = tid >> 5;
    const bool mine = anyBits != 0u;
    unsigned long long initKey = 0ull;
    if (mine) {
        initKey = ((unsigned long long)orderedBits(s.best) << 32) | (s.hit.prim == PT_MISS ? 0u : s.hit.prim);
        sm.key[tid] = initKey;
#pragma unroll
        for (int g = 0; g < PT_COOP_STEPS; g++) sm.base[g][tid] = tg[g].x;
    }
    volatile unsigned long long *vkey = sm.key;
#if PT_COOP_LIST == 2
    // No list in memory: slot i of the pooled work belongs to the lane whose prefix-sum interval contains i (binary search over
    // the lanes' inclusive sums by shuffles) and, inside that lane's bits, to the (i - first)-th set triangle (popcount search).
    // Every lane does this for its own slot, so building the work costs ~50 instructions per 32 slots at full lanes instead of a
    // loop that runs as long as the lane with the most triangles.
    static_assert(PT_COOP_STEPS <= 2, "PT_COOP_LIST 2 is written for one or two triangle groups per lane");
    {
        const uint32_t b0own = bits[0], b1own = PT_COOP_STEPS > 1 ? bits[PT_COOP_STEPS - 1] : 0u;
        const uint32_t cnt = uint32_t(__popc(b0own)) + (PT_COOP_STEPS > 1 ? uint32_t(__popc(b1own)) : 0u);
        uint32_t incl = cnt;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
            const uint32_t nb = __shfl_up_sync(0xffffffffu, incl, dlt);
            if (int(lane) >= dlt) incl += nb;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t excl = incl - cnt;
        __syncwarp();            // sm.key / sm.base of the owners are visible
        for (uint32_t i0 = 0; i0 < total; i0 += 32u) {
            const uint32_t i = min(i0 + lane, total - 1u);
            const bool live = i0 + lane < total;
            uint32_t lo = 0u, hi = 31u;          // first lane whose inclusive sum exceeds i
#pragma unroll
            for (int step = 0; step < 5; step++) {
                const uint32_t mid = (lo + hi) >> 1;
                const uint32_t v = __shfl_sync(0xffffffffu, incl, int(mid));
                if (v > i) hi = mid; else lo = mid + 1u;
            }
            const uint32_t owner = lo;
            uint32_t r = i - __shfl_sync(0xffffffffu, excl, int(owner));
            const uint32_t ob0 = __shfl_sync(0xffffffffu, b0own, int(owner));
            const uint32_t ob1 = __shfl_sync(0xffffffffu, b1own, int(owner));
            const uint32_t c0 = uint32_t(__popc(ob0));
            const uint32_t g = (PT_COOP_STEPS > 1 && r >= c0) ? 1u : 0u;
            uint32_t bb = g ? ob1 : ob0;
            if (g) r -= c0;
            // position of the r-th set bit of the 24-bit mask bb (it has more than r bits set)
            uint32_t ti = 0u, c;
            c = uint32_t(__popc(bb & 0xfffu)); if (r >= c) { r -= c; ti = 12u; }
            c = uint32_t(__popc((bb >> ti) & 0x3fu)); if (r >= c) { r -= c; ti += 6u; }
            c = uint32_t(__popc((bb >> ti) & 0x7u)); if (r >= c) { r -= c; ti += 3u; }
            const uint32_t t3 = (bb >> ti) & 0x7u;
            // r-th set bit of three: t3 = 1:0 2:1 3:0,1 4:2 5:0,2 6:1,2 7:0,1,2  (two bits per (t3, r), r = 0..2)
            ti += uint32_t((0x909202101000ull >> ((t3 * 3u + r) * 2u)) & 3ull);
            bool won = false;
            unsigned long long cand = 0ull;
            float u = 0.0f, v = 0.0f;
            const uint32_t ol = wl + owner;
            if (live) {
                const uint32_t tb = (sm.base[g][ol] + ti) * 3u;
                const float4 ro = sm.rayO[ol], rd = sm.rayD[ol];
                const float4 a = __ldg(&sc.tris[tb + 0]);
                const float4 b = __ldg(&sc.tris[tb + 1]);
                const float4 c4 = __ldg(&sc.tris[tb + 2]);
                float t;
                if (intersectTriExact(a, b, c4, make_vec3(ro), make_vec3(rd), t, u, v) && t > ro.w) {
                    const uint32_t id = __float_as_uint(a.w);
                    cand = ((unsigned long long)orderedBits(t) << 32) | id;
                    if (cand < vkey[ol]) {
                        if (!ALPHA || __float_as_uint(b.w) == 0u || !(ALPHA == 2 ? alphaRejectsInline(sc, id, u, v, ro.x, t) : alphaRejects(sc, id, u, v, ro.x, t)))
                            won = cand < atomicMin(&sm.key[ol], cand);
                    }
                }
            }
            __syncwarp();
            if (won && vkey[ol] == cand) sm.uv[ol] = make_float2(u, v);
        }
    }
#else
    for (;;) {
        uint32_t cnt = 0u;
#pragma unroll
        for (int g = 0; g < PT_COOP_STEPS; g++) cnt += __popc(bits[g]);
        uint32_t incl = cnt;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
            const uint32_t nb = __shfl_up_sync(0xffffffffu, incl, dlt);
            if (int(lane) >= dlt) incl += nb;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t pos = incl - cnt;
        if (PT_COOP_STEPS == 1) {
            while (bits[0] && pos < PT_COOP_CAP) {
                const uint32_t ti = uint32_t(__ffs(bits[0]) - 1);
                bits[0] &= bits[0] - 1u;
                sm.items[warp][pos++] = uint16_t((lane << 5) | ti);
            }
        } else if (PT_COOP_LIST == 1) {
            // one plain loop per group: half the instructions per entry of the merged loop below
#pragma unroll
            for (int g = 0; g < PT_COOP_STEPS; g++) {
                uint32_t b = bits[g];
                const uint32_t tag = ((lane << 2) | uint32_t(g)) << 5;
                while (b && pos < PT_COOP_CAP) {
                    const uint32_t ti = uint32_t(__ffs(b) - 1);
                    b &= b - 1u;
                    sm.items[warp][pos++] = uint16_t(tag | ti);
                }
                bits[g] = b;
            }
        } else {
            while (pos < PT_COOP_CAP) {
                uint32_t g = 0u, b = bits[0];
#pragma unroll
                for (int k = 1; k < PT_COOP_STEPS; k++) if (b == 0u) { g = uint32_t(k); b = bits[k]; }
                if (b == 0u) break;
                const uint32_t ti = uint32_t(__ffs(b) - 1);
                b &= b - 1u;
#pragma unroll
                for (int k = 0; k < PT_COOP_STEPS; k++) if (g == uint32_t(k)) bits[k] = b;
                sm.items[warp][pos++] = uint16_t((((lane << 2) | g) << 5) | ti);
            }
        }
        __syncwarp();
        const uint32_t n = min(total, uint32_t(PT_COOP_CAP));
        for (uint32_t i0 = 0; i0 < n; i0 += 32u) {
            const uint32_t i = i0 + lane;
            bool won = false;
            unsigned long long cand = 0ull;
            float u = 0.0f, v = 0.0f;
            uint32_t ol = 0u;
            if (i < n) {
                const uint32_t it = sm.items[warp][i];
                uint32_t tb;
                if (PT_COOP_STEPS == 1) { ol = wl + (it >> 5); tb = (sm.base[0][ol] + (it & 31u)) * 3u; }
                else { ol = wl + (it >> 7); tb = (sm.base[(it >> 5) & 3u][ol] + (it & 31u)) * 3u; }
                const float4 ro = sm.rayO[ol], rd = sm.rayD[ol];
                const float4 a = __ldg(&sc.tris[tb + 0]);
                const float4 b = __ldg(&sc.tris[tb + 1]);
                const float4 c = __ldg(&sc.tris[tb + 2]);
                float t;
                if (intersectTriExact(a, b, c, make_vec3(ro), make_vec3(rd), t, u, v) && t > ro.w) {
                    const uint32_t id = __float_as_uint(a.w);
                    cand = ((unsigned long long)orderedBits(t) << 32) | id;
                    if (cand < vkey[ol]) {
                        if (!ALPHA || __float_as_uint(b.w) == 0u || !(ALPHA == 2 ? alphaRejectsInline(sc, id, u, v, ro.x, t) : alphaRejects(sc, id, u, v, ro.x, t)))
                            won = cand < atomicMin(&sm.key[ol], cand);
                    }
                }
            }
            __syncwarp();
            if (won && vkey[ol] == cand) sm.uv[ol] = make_float2(u, v);
        }
        if (total <= uint32_t(PT_COOP_CAP)) break;
        __syncwarp();            // the next round rewrites the work list
    }
#endif
    __syncwarp();
    if (mine) {
        const unsigned long long k = sm.key[tid];
        if (k != initKey) {
            const float2 uv = sm.uv[tid];
            s.best = s.hit.t = fromOrderedBits(uint32_t(k >> 32));
            s.hit.prim = uint32_t(k);
            s.hit.u = uv.x; s.hit.v = uv.y;
        }
    }
}

// tracePersistent with the cooperative triangle phase: a step = node half for every lane with a ray, then ONE pooled
// triangle phase for the warp, then the stack pop.
template <int ALPHA, typename IO>
__device__ __forceinline__ void tracePersistentCoop(const TraceScene &sc, const IO &io, const uint32_t total, uint32_t *workCounter,
                                                    const uint32_t chunk, const int refillMin, uint2 *smemStack, CoopSmem &sm) {
    const unsigned tid = threadIdx.x;
    const unsigned lane = tid & 31u;
    const unsigned ltMask = (1u << lane) - 1u;
    const int stride = blockDim.x;
    uint2 localStack[PT_STACK_LOCAL];
    Trav s;
    uint32_t rayIdx = 0, wBase = 0, wEnd = 0;
    bool active = false, any = false, exhausted = false;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle == 0xffffffffu && exhausted) break;
        if (!exhausted && (idle == 0xffffffffu || __popc(idle) >= refillMin)) {
            if (wBase >= wEnd) {
                uint32_t b = 0;
                if (lane == 0) b = atomicAdd(workCounter, chunk);
                b = __shfl_sync(0xffffffffu, b, 0);
                if (b >= total) { exhausted = true; wBase = wEnd = 0; }
                else { wBase = b; wEnd = min(b + chunk, total); }
            }
            if (!active) {
                const uint32_t idx = wBase + __popc(idle & ltMask);
                if (idx < wEnd) {
                    vec3 o, d; float tmin, tmax;
                    io.load(idx, o, d, tmin, tmax, any);
                    rayIdx = idx;
                    active = !travInit(s, sc, o, d, tmin, tmax, any);
                    if (!active) io.store(idx, s.hit, any);
                    else { sm.rayO[tid] = make_float4(o.x, o.y, o.z, tmin); sm.rayD[tid] = make_float4(d.x, d.y, d.z, 0.0f); }
                }
            }
            wBase = min(wEnd, wBase + uint32_t(__popc(idle)));
        }
        // PT_COOP_STEPS node steps (each followed by its stack pop), then one pooled triangle phase for the warp
        uint2 tg[PT_COOP_STEPS];
        bool stackEmpty = false;
        uint2 cur = make_uint2(0u, 0u);
        if (active) cur = s.cur;
#pragma unroll
        for (int g = 0; g < PT_COOP_STEPS; g++) {
            tg[g] = make_uint2(0u, 0u);
            if (active && !stackEmpty) {
                travNode(s, sc, cur, tg[g], smemStack, localStack, stride);
                if ((cur.y & 0xff000000u) == 0u) {
                    if (s.sp == 0) stackEmpty = true;
                    else {
                        s.sp--;
                        if (s.sp < PT_STACK_SMEM) cur = smemStack[s.sp * stride];
                        else cur = localStack[s.sp - PT_STACK_SMEM];
                    }
                }
            }
        }
#if PT_PREFETCH
        // the node the next step will visit is known now (unless the triangle phase below culls it): start its two cache
        // lines towards L1 while the warp is busy with the pooled triangle tests
        if (active && !stackEmpty) {
            const int bit = 31 - __clz(cur.y);
            const uint32_t slot = (uint32_t(bit) - 24u) ^ s.octInv;
            const uint32_t nodeIdx = cur.x + __popc(cur.y & ~(0xffffffffu << slot));
            const float4 *np = sc.nodes + size_t(nodeIdx) * 5;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(np));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(np + 4));
        }
#endif
        coopTriangles<ALPHA>(s, sc, tg, sm, lane, tid);
        if (active) {
            s.cur = cur;
            if (stackEmpty || (any && s.hit.prim != PT_MISS)) { io.store(rayIdx, s.hit, any); active = false; }
        }
    }
}

}  // namespace b200pt
