// TEST INFRASTRUCTURE ONLY.  C entry point around the reference's OWN WeightedSampler (src/WeightedSampler.cpp, compiled where
// it lies under /root/reference by oracle/Makefile into oracle/_ref/libhost_ref.so; nothing is copied into this repository).
// SceneLoader::getLightSamplingVector / getFaceSamplingVector (src/SceneLoader.cpp:861-944) draw the 10 000-entry light and
// face tables of binding 6 from it; tests/test_scene.py feeds it the weights our loader reports and compares the tables.
#include <vector>
#include "WeightedSampler.h"

extern "C" int host_ref_weighted_samples(const float *values, int n, int count, int *samples_out, float *probs_out, float *total_out) {
    std::vector<float> v(values, values + n);
    WeightedSampler sampler(v);                        // a fresh sampler per table, like the reference (default-seeded mt19937)
    std::vector<float> probs = sampler.getProbabilities();
    for (int i = 0; i < n; i++) probs_out[i] = probs[size_t(i)];
    *total_out = sampler.getTotal();
    for (int i = 0; i < count; i++) samples_out[i] = sampler.sample();
    return 0;
}
