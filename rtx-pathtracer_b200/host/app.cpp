// Frame driver: the per-frame host logic of the reference application, without the window.
//   b200pt_app_begin_frame  = RayTracingApp::raytrace up to the push (src/RayTracingApp.cpp:1120-1173)
//   b200pt_app_end_frame    = the bookkeeping of RayTracingApp::drawCallback after the frame (:120-163)
// State only — the device work happens in b200pt_render_frame / b200pt_guiding_update.
#include <cstring>
#include "../../include/b200pt.h"

extern "C" {

void b200pt_app_init(b200pt_app *app) {
    if (!app) return;
    memset(app, 0, sizeof(*app));
    b200pt_default_push_constants(&app->settings);     // previousFrames = -1 (RayTracingApp.h:118)
    app->accumulateResults = 0;
    app->irradianceCachePrepareFrames = 50;
    app->numGuidingOptimizations = 6;
    app->currentGuidingOptimizations = -1;
}

void b200pt_app_scene_switched(b200pt_app *app) {     // src/RayTracingApp.cpp:327-367
    if (!app) return;
    app->currentPrepareFrames = 0;
    app->hasInputChanged = 1;
}

static void setEstimateRTSettings(b200pt_push_constants &pc) {   // src/RayTracingApp.cpp:1206-1217
    pc.samplesPerPixel = 16;
    pc.useIrradianceCache = 1;
    pc.useIrradianceCacheOnGlossy = 1;
    pc.enableNEE = 1;
    pc.visualizeMode = 0;
    pc.maxDepth = 1;
    pc.maxFollowDiscrete = 10;
    pc.showIrradianceCacheOnly = 0;
    pc.numNEE = 5;
    pc.useADRRS = 0;
}

void b200pt_app_begin_frame(b200pt_app *app, uint32_t frame_seed, b200pt_push_constants *frame_pc) {
    if (!app) return;
    b200pt_push_constants &pc = app->settings;
    pc.randomUInt = frame_seed;     // the reference draws it from glm::linearRand (quirk 10); here the caller owns the seed stream
    if (app->accumulateResults && !app->hasInputChanged) pc.previousFrames += 1u;
    else pc.previousFrames = 0u;

    // the irradiance cache needs a few frames to create its first entries
    if (pc.useIrradianceCache || pc.useADRRS) {
        if (app->currentPrepareFrames < app->irradianceCachePrepareFrames) {
            if (pc.useADRRS) {      // ADRRS cannot run during the prepare frames, but must afterwards
                pc.useIrradianceCacheOnGlossy = 1;
                pc.useIrradianceCache = 1;
                pc.useADRRS = 0;
                app->activateADRRSAfterPrepareFrames = 1;
            }
            pc.isIrradiancePrepareFrame = 1;
            pc.previousFrames = 0xFFFFFFFFu;
            app->currentPrepareFrames++;
        } else if (app->activateADRRSAfterPrepareFrames && app->currentPrepareFrames == app->irradianceCachePrepareFrames) {
            pc.storeEstimate = 1;   // after the prepare frames: one frame that stores the estimate image
            pc.useIrradianceCache = 0;
            pc.isIrradiancePrepareFrame = 0;
            pc.useADRRS = 1;
            app->activateADRRSAfterPrepareFrames = 0;
            app->currentPrepareFrames++;
        } else pc.isIrradiancePrepareFrame = 0;
    } else pc.isIrradiancePrepareFrame = 0;

    if (pc.storeEstimate) {
        if (app->loadBackupNextIteration) {      // the frame after the estimate frame: the user's settings come back
            pc = app->backupPushConstant;        // (sic: including the estimate frame's randomUInt and previousFrames)
            pc.storeEstimate = 0;
            app->loadBackupNextIteration = 0;
        } else {
            app->backupPushConstant = pc;
            setEstimateRTSettings(pc);
            app->loadBackupNextIteration = 1;
        }
    }
    app->hasInputChanged = 0;
    if (frame_pc) *frame_pc = pc;
}

int b200pt_app_end_frame(b200pt_app *app) {
    if (!app) return 0;
    b200pt_push_constants &pc = app->settings;
    if (pc.updateGuiding) {       // limit the number of guiding optimisations
        if (app->currentGuidingOptimizations == -1) app->currentGuidingOptimizations = 0;
        else app->currentGuidingOptimizations++;
        if (app->numGuidingOptimizations != -1 && app->currentGuidingOptimizations >= app->numGuidingOptimizations) {
            pc.updateGuiding = 0;
            app->currentGuidingOptimizations = -1;
        }
    } else app->currentGuidingOptimizations = -1;
    const int runUpdate = pc.updateGuiding > 0;
    // prepare frames only fill the irradiance cache; they do not count as image samples
    if (!(pc.useIrradianceCache && app->currentPrepareFrames < app->irradianceCachePrepareFrames)) app->evalCurrentSamples += pc.samplesPerPixel;
    return runUpdate;
}

int b200pt_app_draw_frame(b200pt_app *app, b200pt_ctx *ctx, uint32_t frame_seed, const b200pt_guiding_params *guiding_params) {
    if (!app || !ctx) return B200PT_E_INVALID;
    b200pt_push_constants pc;
    b200pt_app_begin_frame(app, frame_seed, &pc);
    int rc = b200pt_render_frame(ctx, &pc);
    if (rc != B200PT_OK) return rc;
    if (b200pt_app_end_frame(app)) {
        b200pt_guiding_params gp;
        if (!guiding_params) { b200pt_default_guiding_params(&gp); gp.useParallaxCompensation = pc.useParallaxCompensation; guiding_params = &gp; }
        // spp-sharded training run (b200pt_comm_init was called): the ranks' samples are refitted together, region-sharded
        rc = b200pt_comm_exchange_mode(ctx) > 0 ? b200pt_guiding_update_all_ranks(ctx, guiding_params) : b200pt_guiding_update(ctx, guiding_params);
    }
    return rc;
}

}  // extern "C"
