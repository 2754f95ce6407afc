#include "bvh.h"
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>

namespace b200pt {
namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; a++) { lo[a] = std::numeric_limits<float>::infinity(); hi[a] = -lo[a]; } }
    void grow(const float p[3]) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void grow(const Box &b) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float halfArea() const {
        float d[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
        if (d[0] < 0) return 0;
        return d[0] * d[1] + d[1] * d[2] + d[2] * d[0];
    }
};

struct Node2 {
    Box box;
    int left = -1, right = -1;   // children (inner) or -1
    uint32_t first = 0, count = 0;
};

struct Builder {
    const float *verts;
    std::vector<Box> triBox;
    std::vector<float> centroid;   // 3 per tri
    std::vector<uint32_t> order;
    std::vector<Node2> nodes;
    uint32_t maxLeaf = 3;
    float travCost = 0.5f;
    float nodeCost = 2.0f;         // cost of visiting one 8-wide node, in triangle tests (the collapse below)

    // Collapse of the binary tree into 8-wide nodes by dynamic programming over the SAH cost (Ylitie, Karras, Laine
    // 2017, section 3.1): cost[n][i-1] = cheapest way to represent the subtree of n as at most i roots, each root a
    // leaf (<= maxLeaf triangles) or an 8-wide node whose children are again such a forest of the two subtrees.
    // The greedy "open the largest child" collapse left half of the nodes of cornell-dielectric with two children
    // (4.0 children per node on average); every node visit costs the same 5 x 128-bit loads and 8 slab tests.
    std::vector<float> cost;       // 7 per node
    std::vector<float> costInner, costLeaf;
    std::vector<uint8_t> splitK;   // 9 per node: roots given to the left subtree when j roots are distributed (j = 2..8)
    float distribute(int n, int j, int &bestK) const {
        const int l = nodes[n].left, r = nodes[n].right;
        float best = std::numeric_limits<float>::infinity(); bestK = 1;
        for (int k = std::max(1, j - 7); k <= std::min(7, j - 1); k++) {
            const float c = cost[size_t(l) * 7 + (k - 1)] + cost[size_t(r) * 7 + (j - k - 1)];
            if (c < best) { best = c; bestK = k; }
        }
        return best;
    }
    void solve(int n) {
        const Node2 &nd = nodes[n];
        const float area = nd.box.halfArea();
        costLeaf[n] = nd.count <= maxLeaf ? area * float(nd.count) : std::numeric_limits<float>::infinity();
        if (nd.left < 0) {
            costInner[n] = std::numeric_limits<float>::infinity();
            for (int i = 0; i < 7; i++) cost[size_t(n) * 7 + i] = costLeaf[n];
            return;
        }
        solve(nd.left); solve(nd.right);
        float dist[9];
        for (int j = 2; j <= 8; j++) { int k; dist[j] = distribute(n, j, k); splitK[size_t(n) * 9 + j] = uint8_t(k); }
        costInner[n] = dist[8] + area * nodeCost;
        cost[size_t(n) * 7] = std::min(costLeaf[n], costInner[n]);
        for (int i = 2; i <= 7; i++) cost[size_t(n) * 7 + (i - 1)] = std::min(dist[i], cost[size_t(n) * 7 + (i - 2)]);
    }
    struct Kid { int node; bool leaf; };
    // the (at most i) roots that represent the subtree of n
    void collect(int n, int i, std::vector<Kid> &out) const {
        const Node2 &nd = nodes[n];
        if (nd.left < 0) { out.push_back({n, true}); return; }
        if (i == 1) { out.push_back({n, costLeaf[n] <= costInner[n]}); return; }
        int k; const float d = distribute(n, i, k);
        if (cost[size_t(n) * 7 + (i - 2)] <= d) { collect(n, i - 1, out); return; }
        collect(nd.left, k, out); collect(nd.right, i - k, out);
    }
    // children of the 8-wide node that n becomes
    void children(int n, std::vector<Kid> &out) const {
        const Node2 &nd = nodes[n];
        if (nd.left < 0) { out.push_back({n, true}); return; }
        const int k = splitK[size_t(n) * 9 + 8];
        collect(nd.left, k, out); collect(nd.right, 8 - k, out);
    }

    int build(uint32_t first, uint32_t count) {
        int idx = int(nodes.size());
        nodes.emplace_back();
        Box box, cbox;
        box.reset(); cbox.reset();
        for (uint32_t i = first; i < first + count; i++) { box.grow(triBox[order[i]]); cbox.grow(&centroid[3 * order[i]]); }
        nodes[idx].box = box;
        nodes[idx].first = first; nodes[idx].count = count;
        if (count <= 1) return idx;

        const int NB = 16;
        float bestCost = std::numeric_limits<float>::infinity();
        int bestAxis = -1, bestBin = -1;
        for (int axis = 0; axis < 3; axis++) {
            float ext = cbox.hi[axis] - cbox.lo[axis];
            if (!(ext > 0)) continue;
            Box bb[NB]; uint32_t bc[NB];
            for (int b = 0; b < NB; b++) { bb[b].reset(); bc[b] = 0; }
            float scale = NB / ext;
            for (uint32_t i = first; i < first + count; i++) {
                uint32_t t = order[i];
                int b = std::min(NB - 1, std::max(0, int((centroid[3 * t + axis] - cbox.lo[axis]) * scale)));
                bb[b].grow(triBox[t]); bc[b]++;
            }
            float rightArea[NB]; uint32_t rightCount[NB];
            Box acc; acc.reset(); uint32_t n = 0;
            for (int b = NB - 1; b > 0; b--) { acc.grow(bb[b]); n += bc[b]; rightArea[b] = acc.halfArea(); rightCount[b] = n; }
            acc.reset(); n = 0;
            for (int b = 0; b < NB - 1; b++) {
                acc.grow(bb[b]); n += bc[b];
                if (n == 0 || rightCount[b + 1] == 0) continue;
                float cost = acc.halfArea() * n + rightArea[b + 1] * rightCount[b + 1];
                if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestBin = b; }
            }
        }
        // leaves hold at most 3 triangles (unary count in 3 bits of the node meta byte); travCost = cost of one more
        // node visit in units of a triangle test (B200PT_BVH_MAX_LEAF / B200PT_BVH_TRAV_COST override for tuning)
        if (count <= maxLeaf) {
            float leafCost = box.halfArea() * count;
            if (bestAxis < 0 || bestCost + box.halfArea() * travCost >= leafCost) return idx;
        }
        uint32_t mid;
        if (bestAxis >= 0) {
            float ext = cbox.hi[bestAxis] - cbox.lo[bestAxis];
            float scale = NB / ext;
            auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](uint32_t t) {
                int b = std::min(NB - 1, std::max(0, int((centroid[3 * t + bestAxis] - cbox.lo[bestAxis]) * scale)));
                return b <= bestBin;
            });
            mid = uint32_t(it - order.begin());
        } else {
            mid = first + count / 2;   // all centroids coincide: split by index
        }
        if (mid == first || mid == first + count) mid = first + count / 2;
        int l = build(first, mid - first);
        int r = build(mid, first + count - mid);
        nodes[idx].left = l; nodes[idx].right = r;
        return idx;
    }
};

}  // namespace

void buildBvh8(const float *verts, uint32_t numTris, Bvh8 &out) {
    out.nodes.clear(); out.tris.clear(); out.maxDepth = 0;
    // world-space dilation: ~100 ulp of the largest coordinate, so that float rounding in the slab test can never
    // cull a box whose triangle the (exactly specified) triangle test would accept
    float maxAbs = 0;
    for (size_t i = 0; i < size_t(numTris) * 9; i++) maxAbs = std::max(maxAbs, std::fabs(verts[i]));
    float pad = std::max(maxAbs, 1e-3f) * 1.2e-5f;
    out.pad = pad;

    Builder b;
    if (const char *e = getenv("B200PT_BVH_MAX_LEAF")) b.maxLeaf = uint32_t(std::max(1, std::min(3, atoi(e))));
    if (const char *e = getenv("B200PT_BVH_TRAV_COST")) b.travCost = float(atof(e));
    b.verts = verts;
    b.triBox.resize(numTris); b.centroid.resize(size_t(numTris) * 3); b.order.resize(numTris);
    std::iota(b.order.begin(), b.order.end(), 0u);
    for (uint32_t t = 0; t < numTris; t++) {
        Box bx; bx.reset();
        for (int k = 0; k < 3; k++) bx.grow(verts + size_t(t) * 9 + k * 3);
        for (int a = 0; a < 3; a++) { b.centroid[3 * t + a] = 0.5f * (bx.lo[a] + bx.hi[a]); bx.lo[a] -= pad; bx.hi[a] += pad; }
        b.triBox[t] = bx;
    }
    if (numTris == 0) {
        Bvh8Node n; memset(&n, 0, sizeof(n));
        n.e[0] = n.e[1] = n.e[2] = 1;
        out.nodes.push_back(n);
        return;
    }
    b.nodes.reserve(size_t(numTris) * 2);
    int root2 = b.build(0, numTris);
    if (const char *e = getenv("B200PT_BVH_NODE_COST")) b.nodeCost = float(atof(e));
    const bool greedy = getenv("B200PT_BVH_GREEDY") != nullptr;
    std::vector<Builder::Kid> kidList;
    if (!greedy) {
        b.cost.assign(b.nodes.size() * 7, 0.0f); b.costInner.assign(b.nodes.size(), 0.0f); b.costLeaf.assign(b.nodes.size(), 0.0f);
        b.splitK.assign(b.nodes.size() * 9, 0);
        b.solve(root2);
    }

    struct Work { int node2; uint32_t node8; int depth; };
    std::vector<Work> queue;
    out.nodes.emplace_back();
    queue.push_back({root2, 0, 1});
    for (size_t qi = 0; qi < queue.size(); qi++) {
        Work w = queue[qi];
        out.maxDepth = std::max(out.maxDepth, w.depth);
        const Node2 &n2 = b.nodes[w.node2];
        // children of this 8-wide node: the DP's choice (B200PT_BVH_GREEDY=1: open the largest inner child until 8)
        int kids[8]; bool kidLeaf[8]; int nk = 0;
        if (!greedy) {
            kidList.clear();
            b.children(w.node2, kidList);
            for (const Builder::Kid &k : kidList) { kids[nk] = k.node; kidLeaf[nk] = k.leaf; nk++; }
        } else {
            if (n2.left < 0) kids[nk++] = w.node2;   // the whole (sub)tree is a single leaf
            else { kids[nk++] = n2.left; kids[nk++] = n2.right; }
            for (;;) {
                if (nk == 8) break;
                int best = -1; float bestArea = -1;
                for (int k = 0; k < nk; k++) {
                    const Node2 &c = b.nodes[kids[k]];
                    if (c.left < 0) continue;
                    float a = c.box.halfArea();
                    if (a > bestArea) { bestArea = a; best = k; }
                }
                if (best < 0) break;
                const Node2 &c = b.nodes[kids[best]];
                kids[best] = c.left; kids[nk++] = c.right;
            }
            for (int k = 0; k < nk; k++) kidLeaf[k] = b.nodes[kids[k]].left < 0;
        }
        // slot assignment: greedy matching of children to octant directions (slot bit set = high side of the axis)
        float nc[3];
        for (int a = 0; a < 3; a++) nc[a] = 0.5f * (n2.box.lo[a] + n2.box.hi[a]);
        int slotOf[8]; bool slotUsed[8] = {false, false, false, false, false, false, false, false};
        bool kidDone[8] = {false, false, false, false, false, false, false, false};
        float cost[8][8];
        for (int k = 0; k < nk; k++) {
            const Box &cb = b.nodes[kids[k]].box;
            for (int s = 0; s < 8; s++) {
                float c = 0;
                for (int a = 0; a < 3; a++) {
                    float d = 0.5f * (cb.lo[a] + cb.hi[a]) - nc[a];
                    c += ((s >> (2 - a)) & 1) ? d : -d;
                }
                cost[k][s] = c;
            }
        }
        for (int it = 0; it < nk; it++) {
            int bk = -1, bs = -1; float bc = -std::numeric_limits<float>::infinity();
            for (int k = 0; k < nk; k++) if (!kidDone[k])
                for (int s = 0; s < 8; s++) if (!slotUsed[s] && cost[k][s] > bc) { bc = cost[k][s]; bk = k; bs = s; }
            kidDone[bk] = true; slotUsed[bs] = true; slotOf[bk] = bs;
        }
        int kidInSlot[8]; bool leafInSlot[8];
        for (int s = 0; s < 8; s++) { kidInSlot[s] = -1; leafInSlot[s] = false; }
        for (int k = 0; k < nk; k++) { kidInSlot[slotOf[k]] = kids[k]; leafInSlot[slotOf[k]] = kidLeaf[k]; }

        Bvh8Node n; memset(&n, 0, sizeof(n));
        for (int a = 0; a < 3; a++) {
            n.p[a] = n2.box.lo[a];
            double ext = double(n2.box.hi[a]) - double(n2.box.lo[a]);
            int e = (ext > 0) ? int(std::ceil(std::log2(ext / 255.0))) : -126;
            // make sure 255 * 2^e really covers the extent (log2 rounding)
            while (e < 127 && std::ldexp(255.0, e) < ext) e++;
            e = std::max(-126, std::min(127, e));
            n.e[a] = uint8_t(e + 127);
        }
        n.childBase = uint32_t(out.nodes.size());
        n.triBase = uint32_t(out.tris.size());
        uint32_t triOffset = 0;
        for (int s = 0; s < 8; s++) {
            int k2 = kidInSlot[s];
            if (k2 < 0) continue;
            const Node2 &c = b.nodes[k2];
            for (int a = 0; a < 3; a++) {
                double scale = std::ldexp(1.0, int(n.e[a]) - 127);
                double lo = std::floor((double(c.box.lo[a]) - double(n.p[a])) / scale);
                double hi = std::ceil((double(c.box.hi[a]) - double(n.p[a])) / scale);
                n.qlo[a][s] = uint8_t(std::max(0.0, std::min(255.0, lo)));
                n.qhi[a][s] = uint8_t(std::max(0.0, std::min(255.0, hi)));
            }
            if (!leafInSlot[s]) {
                n.imask |= uint8_t(1u << s);
                n.meta[s] = uint8_t((1u << 5) | (24u + uint32_t(s)));
                out.nodes.emplace_back();
                queue.push_back({k2, uint32_t(out.nodes.size() - 1), w.depth + 1});
            } else {
                uint32_t unary = c.count == 1 ? 1u : c.count == 2 ? 3u : 7u;
                n.meta[s] = uint8_t((unary << 5) | triOffset);
                for (uint32_t i = c.first; i < c.first + c.count; i++) {
                    uint32_t t = b.order[i];
                    const float *v = verts + size_t(t) * 9;
                    PackedTri pt;
                    for (int a = 0; a < 3; a++) { pt.v0[a] = v[a]; pt.e1[a] = v[3 + a] - v[a]; pt.e2[a] = v[6 + a] - v[a]; }
                    pt.prim = t; pt.pad0 = 0; pt.pad1 = 0;
                    out.tris.push_back(pt);
                }
                triOffset += c.count;
            }
        }
        out.nodes[w.node8] = n;
    }
    if (getenv("B200PT_BVH_STATS")) {     // tuning aid: fill of the 8-wide nodes and the SAH cost of the collapsed tree
        size_t inner = 0, leafKids = 0, leafTris = 0, hist[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (const Bvh8Node &n : out.nodes) {
            int kids = 0;
            for (int s = 0; s < 8; s++) {
                if (n.meta[s] == 0) continue;
                kids++;
                if (n.imask & (1u << s)) inner++;
                else { leafKids++; uint32_t u = n.meta[s] >> 5; leafTris += u == 1 ? 1 : u == 3 ? 2 : 3; }
            }
            hist[kids]++;
        }
        fprintf(stderr, "[bvh8] tris %u nodes %zu depth %d inner children %zu leaf children %zu (%.2f tris each) children/node %.2f  hist",
                numTris, out.nodes.size(), out.maxDepth, inner, leafKids, leafKids ? double(leafTris) / leafKids : 0.0,
                double(inner + leafKids) / out.nodes.size());
        for (int k = 1; k <= 8; k++) fprintf(stderr, " %d:%zu", k, hist[k]);
        fprintf(stderr, "\n");
    }
}

void validateBvh8(const Bvh8 &bvh, const float *verts, uint32_t numTris, Bvh8Report &rep) {
    rep = Bvh8Report();
    rep.numNodes = uint32_t(bvh.nodes.size()); rep.numTris = uint32_t(bvh.tris.size()); rep.maxDepth = uint32_t(bvh.maxDepth);
    if (bvh.nodes.empty()) { rep.badMeta = 1; return; }
    std::vector<uint32_t> uses(numTris, 0);
    std::vector<uint8_t> visited(bvh.nodes.size(), 0);
    struct Item { uint32_t node; int depth; double lo[3], hi[3]; };        // lo / hi: intersection of the ancestors' slot boxes
    std::vector<Item> stack;
    Item root; root.node = 0; root.depth = 1;
    for (int a = 0; a < 3; a++) { root.lo[a] = -std::numeric_limits<double>::infinity(); root.hi[a] = std::numeric_limits<double>::infinity(); }
    stack.push_back(root);
    int deepest = 0;
    while (!stack.empty()) {
        const Item it = stack.back(); stack.pop_back();
        if (it.node >= bvh.nodes.size() || visited[it.node]) { rep.badMeta++; continue; }
        visited[it.node] = 1;
        deepest = std::max(deepest, it.depth);
        const Bvh8Node &n = bvh.nodes[it.node];
        uint32_t innerRank = 0, triOffset = 0;
        for (int s = 0; s < 8; s++) {
            const bool inner = (n.imask >> s) & 1;
            if (n.meta[s] == 0) { if (inner) rep.badMeta++; continue; }
            // the box the device's slab test sees: p + q * 2^(e - 127), exact in double
            double lo[3], hi[3];
            for (int a = 0; a < 3; a++) {
                const double scale = std::ldexp(1.0, int(n.e[a]) - 127);
                lo[a] = std::max(it.lo[a], double(n.p[a]) + double(n.qlo[a][s]) * scale);
                hi[a] = std::min(it.hi[a], double(n.p[a]) + double(n.qhi[a][s]) * scale);
            }
            if (inner) {
                if (n.meta[s] != uint8_t((1u << 5) | (24u + uint32_t(s)))) rep.badMeta++;
                Item c; c.node = n.childBase + innerRank; c.depth = it.depth + 1;       // inner children are contiguous in slot order
                for (int a = 0; a < 3; a++) { c.lo[a] = lo[a]; c.hi[a] = hi[a]; }
                stack.push_back(c);
                innerRank++; rep.innerChildren++;
                continue;
            }
            rep.leafChildren++;
            const uint32_t unary = n.meta[s] >> 5, offset = n.meta[s] & 31u;
            const uint32_t count = unary == 1 ? 1u : unary == 3 ? 2u : unary == 7 ? 3u : 0u;
            if (count == 0 || offset != triOffset || offset + count > 24u || size_t(n.triBase) + offset + count > bvh.tris.size()) { rep.badMeta++; continue; }
            triOffset += count;
            for (uint32_t k = 0; k < count; k++) {
                const PackedTri &pt = bvh.tris[size_t(n.triBase) + offset + k];
                if (pt.prim >= numTris) { rep.badMeta++; continue; }
                if (uses[pt.prim]++) rep.duplicatePrims++;
                const float *v = verts + size_t(pt.prim) * 9;
                bool same = true, inside = true;
                for (int a = 0; a < 3; a++) {
                    same = same && pt.v0[a] == v[a] && pt.e1[a] == v[3 + a] - v[a] && pt.e2[a] == v[6 + a] - v[a];
                    for (int c = 0; c < 3; c++) inside = inside && double(v[3 * c + a]) >= lo[a] && double(v[3 * c + a]) <= hi[a];
                }
                if (!same) rep.badMeta++;
                if (!inside) rep.outsideBox++;
            }
        }
    }
    for (uint32_t t = 0; t < numTris; t++) if (!uses[t]) rep.missingPrims++;
    for (uint8_t v : visited) if (!v) rep.unreachableNodes++;
    if (numTris && deepest != bvh.maxDepth) rep.depthMismatch = 1;       // (no triangles: one empty node, depth 0)
}

}  // namespace b200pt
