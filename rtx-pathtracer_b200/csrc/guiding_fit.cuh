// Guiding state on the device: region tree + per-region vMF mixtures (PathGuiding, src/PathGuiding.{h,cpp}).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "../../include/b200pt.h"

namespace b200pt {

struct GuidingState {
    bool ready = false;
    int regionCount = 0;
    std::vector<b200pt_aabb> hostAabbs;
    b200pt_aabb *aabbs = nullptr;          // device, binding 15
    b200pt_vmm_theta *vmms = nullptr;      // device, binding 16
    std::string error;

    int init(int splits, const float sceneMin[3], const float sceneMax[3], cudaStream_t stream);
    int update(b200pt_directional_data *samples, int64_t numSamples, const b200pt_guiding_params &params, cudaStream_t stream, b200pt_stats *stats);
    void release();
};

}  // namespace b200pt
