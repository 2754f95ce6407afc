// Host-built acceleration structure: binned-SAH BVH2 collapsed into compressed 8-wide nodes (80 B, quantised child
// boxes, after Ylitie et al. 2017).  Replaces the driver BLAS/TLAS build of the reference
// (src/SceneLoader.cpp:1101-1193, external/nvvkpp/raytraceKHR_vkpp.hpp): the reference contains no BVH code of its
// own, so layout and tie-breaking rules are ours and are validated against the brute-force oracle.
#pragma once
#include <cstdint>
#include <vector>

namespace b200pt {

struct Bvh8Node {            // 80 bytes = 5 x 16 B, fetched with 128-bit loads
    float p[3];              // quantisation origin (node box min)
    uint8_t e[3];            // per-axis biased exponent: scale = 2^(e-127)
    uint8_t imask;           // which of the 8 slots hold inner nodes
    uint32_t childBase;      // index of the first inner child; inner children are contiguous in slot order
    uint32_t triBase;        // index of the first triangle of this node's leaf children
    uint8_t meta[8];         // inner: 0b001_11sss (24+slot); leaf: unary tri count << 5 | first tri offset; empty: 0
    uint8_t qlo[3][8];       // quantised child box minima  (x[8], y[8], z[8])
    uint8_t qhi[3][8];       // quantised child box maxima
};
static_assert(sizeof(Bvh8Node) == 80, "Bvh8Node must be 80 bytes");

struct PackedTri {           // 48 bytes = 3 x float4: v0 | e1 = v1-v0 | e2 = v2-v0, world space
    float v0[3]; uint32_t prim;
    float e1[3]; float pad0;
    float e2[3]; float pad1;
};
static_assert(sizeof(PackedTri) == 48, "PackedTri must be 48 bytes");

struct Bvh8 {
    std::vector<Bvh8Node> nodes;
    std::vector<PackedTri> tris;      // leaf order
    int maxDepth = 0;
    float pad = 0;                    // world-space dilation applied to every box
};

// verts: numTris * 9 floats (world space), prim ids are 0..numTris-1
void buildBvh8(const float *verts, uint32_t numTris, Bvh8 &out);

// Structural check of a built tree against the triangles it was built from (host only, no traversal): what the device's
// traversal relies on.  Every field counts violations and must be 0, except the statistics at the end.
struct Bvh8Report {
    uint32_t missingPrims = 0;        // primitives that appear in no leaf
    uint32_t duplicatePrims = 0;      // primitives that appear in more than one leaf slot
    uint32_t outsideBox = 0;          // leaf triangles with a vertex outside the dequantised box of their slot or of an ancestor's slot
    uint32_t badMeta = 0;             // meta / imask / child index / triangle range inconsistencies, triangles that are not v0, v1 - v0, v2 - v0
    uint32_t depthMismatch = 0;       // 1 when Bvh8::maxDepth is not the depth of the deepest node
    uint32_t unreachableNodes = 0;    // nodes no parent points to
    uint32_t numNodes = 0, numTris = 0, maxDepth = 0, innerChildren = 0, leafChildren = 0;
};
void validateBvh8(const Bvh8 &bvh, const float *verts, uint32_t numTris, Bvh8Report &report);

}  // namespace b200pt
