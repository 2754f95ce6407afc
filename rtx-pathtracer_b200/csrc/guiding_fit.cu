// Guiding region tree and mixture fit (device).  Region creation follows PathGuiding::createRegions
// (src/PathGuiding.cpp:81-104) with Aabb::addEpsilon / splitAabb (src/Shapes.h:32-53).
#include "guiding_fit.cuh"
#include <cstring>
#include <tuple>

namespace b200pt {

static void splitAabb(const b200pt_aabb &b, b200pt_aabb &l, b200pt_aabb &r) {   // src/Shapes.h:32-46, axis = largest
    float size[3] = {b.max[0] - b.min[0], b.max[1] - b.min[1], b.max[2] - b.min[2]};
    int axis = size[0] > size[1] ? (size[0] > size[2] ? 0 : 2) : (size[1] > size[2] ? 1 : 2);
    l = b; r = b;
    l.max[axis] -= 0.5f * size[axis];
    r.min[axis] += 0.5f * size[axis];
}

int GuidingState::init(int splits, const float sceneMin[3], const float sceneMax[3], cudaStream_t stream) {
    release();
    regionCount = 1 << splits;
    b200pt_aabb scene;
    for (int a = 0; a < 3; a++) {   // Aabb::addEpsilon, src/Shapes.h:48-53
        float extent = sceneMax[a] - sceneMin[a];
        float center = sceneMin[a] + 0.5f * extent;
        scene.min[a] = center - 0.50001f * extent;
        scene.max[a] = center + 0.50001f * extent;
    }
    hostAabbs.assign(1, scene);
    for (int i = 0; i < splits; i++) {
        std::vector<b200pt_aabb> next;
        next.reserve(hostAabbs.size() * 2);
        for (const auto &b : hostAabbs) { b200pt_aabb l, r; splitAabb(b, l, r); next.push_back(l); next.push_back(r); }
        hostAabbs.swap(next);
    }
    if (cudaMalloc(reinterpret_cast<void **>(&aabbs), size_t(regionCount) * sizeof(b200pt_aabb)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void **>(&vmms), size_t(regionCount) * sizeof(b200pt_vmm_theta)) != cudaSuccess) {
        error = "cudaMalloc failed";
        return B200PT_E_CUDA;
    }
    cudaMemcpyAsync(aabbs, hostAabbs.data(), size_t(regionCount) * sizeof(b200pt_aabb), cudaMemcpyHostToDevice, stream);
    cudaMemsetAsync(vmms, 0, size_t(regionCount) * sizeof(b200pt_vmm_theta), stream);
    if (cudaStreamSynchronize(stream) != cudaSuccess) { error = "upload failed"; return B200PT_E_CUDA; }
    ready = true;
    return B200PT_OK;
}

int GuidingState::update(b200pt_directional_data *, int64_t, const b200pt_guiding_params &, cudaStream_t, b200pt_stats *) {
    error = "guiding fit kernel not built yet";
    return B200PT_E_STATE;
}

void GuidingState::release() {
    if (aabbs) cudaFree(aabbs);
    if (vmms) cudaFree(vmms);
    aabbs = nullptr; vmms = nullptr; ready = false; regionCount = 0;
    hostAabbs.clear();
}

}  // namespace b200pt
