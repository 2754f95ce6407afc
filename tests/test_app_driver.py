"""Host frame driver (b200pt_app_*) against the state machine of RayTracingApp::raytrace / drawCallback
(src/RayTracingApp.cpp:1120-1170, :120-163).  Pure host logic: runs without a GPU."""
import helpers


def _frames(app, n, seed0=100):
    out = []
    for f in range(n):
        pc = app.begin_frame(seed0 + f)
        out.append({k: getattr(pc, k) for k in ("randomUInt", "previousFrames", "samplesPerPixel", "maxDepth", "maxFollowDiscrete", "numNEE", "enableNEE",
                                                  "useIrradianceCache", "useIrradianceCacheOnGlossy", "useADRRS", "storeEstimate", "isIrradiancePrepareFrame")})
        out[-1]["update"] = app.end_frame()
        out[-1]["counted"] = app.state.evalCurrentSamples
    return out


def test_plain_frames_accumulate_a_running_mean():
    P = helpers.pt()
    app = P.App(accumulate=True, samplesPerPixel=4)
    fr = _frames(app, 4)
    assert [f["previousFrames"] for f in fr] == [0, 1, 2, 3]
    assert [f["randomUInt"] for f in fr] == [100, 101, 102, 103]
    assert [f["counted"] for f in fr] == [4, 8, 12, 16]
    assert not any(f["isIrradiancePrepareFrame"] or f["storeEstimate"] for f in fr)
    app.settings.maxDepth = 5
    app.input_changed()                       # an edited setting restarts the mean
    assert app.begin_frame(1).previousFrames == 0
    app2 = P.App(accumulate=False)
    assert [f["previousFrames"] for f in _frames(app2, 3)] == [0, 0, 0]


def test_adrrs_schedule_prepare_estimate_restore():
    P = helpers.pt()
    app = P.App(accumulate=True, samplesPerPixel=2, maxDepth=30, numNEE=1, useADRRS=1, useIrradianceCache=0)
    app.state.irradianceCachePrepareFrames = 3
    fr = _frames(app, 8)
    for f in fr[:3]:       # prepare frames: IC on (also on glossy), ADRRS postponed, 1 spp enforced by the shader flag
        assert f["isIrradiancePrepareFrame"] == 1 and f["useIrradianceCache"] == 1 and f["useIrradianceCacheOnGlossy"] == 1
        assert f["useADRRS"] == 0 and f["previousFrames"] == 0xFFFFFFFF and f["storeEstimate"] == 0
    e = fr[3]              # the estimate frame: setEstimateRTSettings
    assert (e["storeEstimate"], e["samplesPerPixel"], e["maxDepth"], e["maxFollowDiscrete"], e["numNEE"], e["enableNEE"]) == (1, 16, 1, 10, 5, 1)
    assert e["useIrradianceCache"] == 1 and e["useADRRS"] == 0 and e["isIrradiancePrepareFrame"] == 0 and e["previousFrames"] == 0
    for i, f in enumerate(fr[4:]):     # the user's settings are back, ADRRS on, IC as adjoint only
        assert (f["useADRRS"], f["useIrradianceCache"], f["storeEstimate"], f["isIrradiancePrepareFrame"]) == (1, 0, 0, 0)
        assert (f["samplesPerPixel"], f["maxDepth"], f["numNEE"]) == (2, 30, 1)
        assert f["previousFrames"] == i
    assert fr[4]["randomUInt"] == fr[3]["randomUInt"]      # sic: the backup restores the estimate frame's seed once
    assert fr[5]["randomUInt"] == 105
    # sample accounting: prepare frames do not count, except the last one (sic, :160); the estimate frame counts its 16
    assert [f["counted"] for f in fr] == [0, 0, 2, 18, 20, 22, 24, 26]


def test_ic_only_schedule_and_scene_switch():
    P = helpers.pt()
    app = P.App(accumulate=True, useIrradianceCache=1)
    app.state.irradianceCachePrepareFrames = 2
    fr = _frames(app, 4)
    assert [f["isIrradiancePrepareFrame"] for f in fr] == [1, 1, 0, 0]
    assert [f["previousFrames"] for f in fr] == [0xFFFFFFFF, 0xFFFFFFFF, 0, 1]
    assert all(f["useIrradianceCache"] == 1 and f["storeEstimate"] == 0 for f in fr)
    app.scene_switched()
    assert [f["isIrradiancePrepareFrame"] for f in _frames(app, 3)] == [1, 1, 0]


def test_guiding_optimisation_limit():
    P = helpers.pt()
    app = P.App(accumulate=True, updateGuiding=1)
    app.state.numGuidingOptimizations = 3
    fr = _frames(app, 6)
    assert [f["update"] for f in fr] == [True, True, True, False, False, False]
    assert app.settings.updateGuiding == 0 and app.state.currentGuidingOptimizations == -1
    app = P.App(accumulate=True, updateGuiding=1)
    app.state.numGuidingOptimizations = -1
    assert all(f["update"] for f in _frames(app, 5))
