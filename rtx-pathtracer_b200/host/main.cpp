// Headless command line of the path tracer.  Keeps the reference's positional interface
//   rtx_raytracer WIDTH HEIGHT IC_SIZE GUIDING_SPLITS scenes...      (src/main.cpp:8-35)
// and replaces the ImGui panels (src/RayTracingApp.cpp:932-1118) with --<RtPushConstant field>=value overrides.
// Instead of presenting a window it runs the reference's "collect N samples" evaluation (src/RayTracingApp.cpp:159-186):
// frames of `--samplesPerPixel` spp through the frame driver (b200pt_app_*: IC prepare frames, the ADRRS estimate
// frame, the guiding-optimisation limit) until `--frames` x spp samples are in the image, then writes an EXR whose name
// follows RayTracingApp::getModeString (src/RayTracingApp.cpp:288-316).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>
#include "../../include/b200pt.h"

struct Field { const char *name; size_t offset; bool isFloat; };
#define F_I(n) {#n, offsetof(b200pt_push_constants, n), false}
#define F_F(n) {#n, offsetof(b200pt_push_constants, n), true}
static const Field kFields[] = {
    F_I(maxDepth), F_I(maxFollowDiscrete), F_I(samplesPerPixel), F_I(enableRR), F_I(enableNEE), F_I(numNEE), F_I(enableAverageInsteadOfMix),
    F_I(enableMIS), F_I(usePowerHeuristic), F_I(storeEstimate), F_I(visualizeMode), F_I(useIrradianceCache), F_F(irradianceA),
    F_F(irradianceUpdateProb), F_F(irradianceCreateProb), F_I(useIrradianceGradients), F_I(useIrradianceCacheOnGlossy),
    F_F(irradianceGradientsMaxLength), F_I(irradianceNumNEE), F_F(irradianceCacheMinRadius), F_I(irradianceCachePerformVisibilityCheck),
    F_I(useVisibleSphereSampling), F_I(useADRRS), F_F(adrrsS), F_I(adrrsSplit), F_I(splitOnFirst), F_I(useGuiding), F_F(guidingProb),
    F_I(updateGuiding), F_I(useParallaxCompensation)};

static void printHelp() {
    printf("Required parameters: WIDTH HEIGHT IC_SIZE GUIDING_SPLITS Scenes ...\n");
    printf("Options: --frames=N | --seconds=S, --seed=S --out=FILE.exr --device=D --prepareFrames=N --numGuidingOptimizations=N\n");
    printf("         --splitRegions=0|1 --samplesForRegionSplit=N --splitAndMerge=0|1 (guiding refit)\n");
    printf("         and --<pushConstantField>=value, e.g. --samplesPerPixel=16 --enableMIS=1 --useADRRS=1\n");
}

// frame seed stream: the reference draws randomUInt from glm::linearRand (quirk 10); we take tea(frame, seed)
static uint32_t tea(uint32_t v0, uint32_t v1) {
    uint32_t s0 = 0;
    for (int n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}

int main(int argc, char **argv) {
    std::vector<std::string> positional, options;
    for (int i = 1; i < argc; i++) (strncmp(argv[i], "--", 2) == 0 ? options : positional).push_back(argv[i]);
    if (positional.size() < 5) { printHelp(); return EXIT_FAILURE; }
    int width = std::stoi(positional[0]), height = std::stoi(positional[1]);
    int icSize = std::stoi(positional[2]), guidingSplits = std::stoi(positional[3]);

    b200pt_app app;
    b200pt_app_init(&app);
    app.accumulateResults = 1;      // the evaluation modes switch it on (src/RayTracingApp.cpp:93,98)
    b200pt_push_constants &pc = app.settings;
    b200pt_guiding_params gparams;      // PathGuiding's knobs (the "Guiding" ImGui panel, src/RayTracingApp.cpp:1060-1090)
    b200pt_default_guiding_params(&gparams);
    int frames = 1, device = 0;
    double seconds = 0.0;           // > 0: the reference's "collect for N seconds" evaluation (src/RayTracingApp.cpp:188-218)
    uint32_t seed = 0xC0FFEEu;
    std::string out;
    for (const std::string &o : options) {
        size_t eq = o.find('=');
        std::string key = o.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
        std::string val = eq == std::string::npos ? "1" : o.substr(eq + 1);
        if (key == "frames") frames = std::stoi(val);
        else if (key == "seed") seed = uint32_t(std::stoul(val, nullptr, 0));
        else if (key == "out") out = val;
        else if (key == "device") device = std::stoi(val);
        else if (key == "seconds") seconds = std::stod(val);
        else if (key == "splitRegions") gparams.splitRegions = std::stoi(val);
        else if (key == "samplesForRegionSplit") gparams.samplesForRegionSplit = std::stof(val);
        else if (key == "splitAndMerge") gparams.splitAndMerge = std::stoi(val);
        else if (key == "prepareFrames") app.irradianceCachePrepareFrames = std::stoi(val);
        else if (key == "numGuidingOptimizations") app.numGuidingOptimizations = std::stoi(val);
        else {
            bool found = false;
            for (const Field &f : kFields)
                if (key == f.name) {
                    char *dst = reinterpret_cast<char *>(&pc) + f.offset;
                    if (f.isFloat) { float v = std::stof(val); memcpy(dst, &v, 4); } else { int v = std::stoi(val); memcpy(dst, &v, 4); }
                    found = true;
                }
            if (!found) { fprintf(stderr, "Unknown option %s\n", o.c_str()); printHelp(); return EXIT_FAILURE; }
        }
    }

    b200pt_ctx *ctx = nullptr;
    if (b200pt_create(device, width, height, icSize, guidingSplits, &ctx) != B200PT_OK) { fprintf(stderr, "%s\n", b200pt_last_error()); return EXIT_FAILURE; }
    int status = EXIT_SUCCESS;
    for (size_t si = 4; si < positional.size(); si++) {
        const std::string &scenePath = positional[si];
        auto t0 = std::chrono::high_resolution_clock::now();
        b200pt_scene *scene = nullptr;
        if (b200pt_scene_load(scenePath.c_str(), &scene) != B200PT_OK) { fprintf(stderr, "%s\n", b200pt_last_error()); status = EXIT_FAILURE; continue; }
        b200pt_scene_desc desc;
        b200pt_scene_get_desc(scene, &desc);
        float origin[3], target[3], up[3], vfov, view[16], proj[16];
        b200pt_scene_get_camera(scene, origin, target, up, &vfov);
        b200pt_camera_matrices(origin, target, up, vfov, float(width) / float(height), view, proj);
        if (b200pt_set_scene(ctx, &desc) != B200PT_OK || b200pt_set_camera(ctx, view, proj) != B200PT_OK) {
            fprintf(stderr, "%s\n", b200pt_last_error()); b200pt_scene_free(scene); status = EXIT_FAILURE; continue;
        }
        auto t1 = std::chrono::high_resolution_clock::now();
        printf("Startup time: %lld milliseconds\n", (long long)std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count());
        b200pt_stats_reset(ctx);
        bool ok = true;
        const b200pt_push_constants userSettings = pc;
        b200pt_app_scene_switched(&app);
        pc.previousFrames = 0xFFFFFFFFu;
        app.evalCurrentSamples = 0;
        const long long wanted = (long long)frames * pc.samplesPerPixel;
        int drawn = 0;
        for (uint32_t f = 0; ok; f++, drawn++) {
            if (seconds > 0.0) { if (std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t1).count() >= seconds) break; }
            else if (app.evalCurrentSamples >= wanted) break;
            gparams.useParallaxCompensation = pc.useParallaxCompensation;
            if (b200pt_app_draw_frame(&app, ctx, tea(f, seed), &gparams) != B200PT_OK) { fprintf(stderr, "%s\n", b200pt_last_error()); ok = false; }
        }
        auto t2 = std::chrono::high_resolution_clock::now();
        if (ok) {
            b200pt_stats st;
            b200pt_stats_get(ctx, &st);
            long long ms = (long long)std::chrono::duration_cast<std::chrono::milliseconds>(t2 - t1).count();
            int spp = seconds > 0.0 ? int(app.evalCurrentSamples) : frames * userSettings.samplesPerPixel;
            if (seconds > 0.0) printf("Collected %d in %g\n", spp, seconds);
            printf("Collecting %d samples took %lld milliseconds with %d samples per pixel per frame (%d frames drawn)\n", spp, ms, userSettings.samplesPerPixel, drawn);
            printf("rays: %llu extend + %llu shadow, %.1f Mrays/s (device time %.1f ms), %.2f spp/s\n", (unsigned long long)st.extend_rays,
                   (unsigned long long)st.shadow_rays, double(st.extend_rays + st.shadow_rays) / (double(st.ms_total) * 1e3), st.ms_total,
                   spp / (double(st.ms_total) * 1e-3));
            std::vector<float> img(size_t(width) * height * 4);
            b200pt_read_image(ctx, B200PT_IMAGE_OUTPUT, img.data());
            std::string file = out;
            if (file.empty()) {   // <scene><mode>_<N>samples.exr, mode string as RayTracingApp::getModeString
                std::string base = scenePath.substr(scenePath.find_last_of('/') + 1);
                base = base.substr(0, base.find_last_of('.'));
                std::string mode;
                const b200pt_push_constants &us = userSettings;     // (the driver may be in the middle of a prepare / estimate frame)
                if (us.enableNEE) { mode += "_NEE"; if (us.enableMIS) mode += "_MIS"; }
                if (us.useIrradianceCache) mode += "_IC";
                if (us.useADRRS) mode += "_ADRRS";
                if (us.useGuiding) { mode += "_Guiding"; if (us.useParallaxCompensation) mode += "_Parallax"; }
                file = base + mode + "_" + std::to_string(spp) + "samples.exr";
            }
            if (b200pt_write_exr(file.c_str(), img.data(), width, height) != B200PT_OK) { fprintf(stderr, "%s\n", b200pt_last_error()); status = EXIT_FAILURE; }
            else printf("Wrote file %s\n", file.c_str());
        } else status = EXIT_FAILURE;
        pc = userSettings;      // the next scene starts from the user's settings again
        b200pt_scene_free(scene);
    }
    b200pt_destroy(ctx);
    return status;
}
