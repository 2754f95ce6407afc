"""The headless command line (rtx-pathtracer_b200/b200pt): the reference's `WIDTH HEIGHT IC_SIZE GUIDING_SPLITS scenes...`
interface (src/main.cpp:8-35), frames driven by the frame driver, EXR out.  Compared with the same frames rendered
through the Python binding."""
import os
import subprocess

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
CLI = os.path.join(helpers.PKG_DIR, "b200pt")


def test_cli_usage_message_without_arguments():
    p = subprocess.run([CLI], capture_output=True, text=True)
    assert p.returncode != 0 and "WIDTH HEIGHT IC_SIZE GUIDING_SPLITS" in p.stdout


def test_cli_renders_the_same_image_as_the_binding(tmp_path):
    P = helpers.pt()
    out = str(tmp_path / "cornell.exr")
    p = subprocess.run([CLI, "96", "54", "0", "0", helpers.scene_path("cornell-dielectric"), "--frames=3", "--samplesPerPixel=2", "--enableMIS=1",
                        "--seed=77", "--out=" + out], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "Collecting 6 samples took" in p.stdout and "Wrote file" in p.stdout
    img = P.read_exr(out)
    scene = P.Scene(helpers.scene_path("cornell-dielectric"))
    view, proj = scene.camera_matrices(96 / 54)
    r = P.Renderer(96, 54, 0, 0)
    r.set_scene(scene)
    r.set_camera(view, proj)
    for f in range(3):
        r.render_frame(P.default_push_constants(randomUInt=P.tea(f, 77), previousFrames=f, samplesPerPixel=2, enableMIS=1))
    ref = r.read_image()
    assert img.shape[:2] == (54, 96)
    assert np.allclose(img[..., :3], ref[..., :3], rtol=2e-3, atol=1e-4)      # EXR stores half floats


def test_cli_adrrs_schedule_and_mode_string(tmp_path):
    p = subprocess.run([CLI, "64", "36", "256", "0", helpers.scene_path("irradianceCache"), "--frames=2", "--samplesPerPixel=1", "--useADRRS=1",
                        "--prepareFrames=4", "--irradianceCreateProb=0.02"], capture_output=True, text=True, cwd=str(tmp_path))
    assert p.returncode == 0, p.stdout + p.stderr
    # 4 prepare frames (the last one counts, sic) + the 16-spp estimate frame already exceed the 2 requested samples
    assert "(5 frames drawn)" in p.stdout, p.stdout
    assert os.path.exists(str(tmp_path / "irradianceCache_NEE_ADRRS_2samples.exr")), os.listdir(str(tmp_path))
    img = helpers.pt().read_exr(str(tmp_path / "irradianceCache_NEE_ADRRS_2samples.exr"))
    assert np.isfinite(img[..., :3]).all() and img[..., :3].mean() > 0


def test_cli_reports_errors(tmp_path):
    p = subprocess.run([CLI, "64", "36", "0", "0", str(tmp_path / "nope.xml")], capture_output=True, text=True)
    assert p.returncode != 0 and "nope.xml" in p.stderr
    p = subprocess.run([CLI, "64", "36", "0", "0", helpers.scene_path("veachMIS"), "--useADRRS=1"], capture_output=True, text=True, cwd=str(tmp_path))
    assert p.returncode != 0 and "ic_size" in p.stderr
