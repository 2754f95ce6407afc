"""CPU: the C-ABI library loads and exports every symbol include/b200pt.h declares; layouts have the reference sizes."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers


def header_symbols():
    text = open(os.path.join(helpers.ROOT, "include", "b200pt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200pt_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    P = helpers.pt()
    L = P.lib()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), "libb200pt.so does not export %s" % s
    assert sorted(P.EXPORTS) == syms


def test_struct_sizes_match_reference_layouts():
    P = helpers.pt()
    # SURVEY.md §8(a) layout table
    assert C.sizeof(P.Vertex) == 48 and P.Vertex.normal.offset == 16 and P.Vertex.texCoord.offset == 32 and P.Vertex.materialIndex.offset == 40
    assert C.sizeof(P.Material) == 96
    for name, off in dict(lightColor=0, diffuse=16, specular=32, specularHighlight=44, transparency=48, refractionIndex=52,
                          refractionIndexInv=56, eta=60, k=64, roughness=68, textureIdDiffuse=72, textureIdSpecular=76, type=80).items():
        assert getattr(P.Material, name).offset == off, name
    assert C.sizeof(P.Instance) == 144 and P.Instance.normalTransform.offset == 64 and P.Instance.modelIndex.offset == 128 and P.Instance.iLight.offset == 132
    assert C.sizeof(P.Light) == 40 and P.Light.pos.offset == 12 and P.Light.instanceIndex.offset == 24 and P.Light.type.offset == 36
    assert C.sizeof(P.FaceSample) == 12 and C.sizeof(P.Sphere) == 24 and C.sizeof(P.Aabb) == 24 and C.sizeof(P.CacheHeader) == 12
    assert C.sizeof(P.PushConstants) == 192
    assert P.DIRECTIONAL_DATA_DTYPE.itemsize == 40 and P.VMM_THETA_DTYPE.itemsize == 720 and P.CACHE_DATA_DTYPE.itemsize == 56
    assert P.VMM_THETA_DTYPE.fields["pi"][1] == 640 and P.VMM_THETA_DTYPE.fields["meanPosition"][1] == 704 and P.VMM_THETA_DTYPE.fields["usedDistributions"][1] == 716


def test_default_push_constants_are_the_reference_defaults():
    pc = helpers.pt().default_push_constants()   # src/RayTracingApp.h:116-165
    assert pc.previousFrames == 0xFFFFFFFF and pc.maxDepth == 30 and pc.maxFollowDiscrete == 3 and pc.samplesPerPixel == 1
    assert pc.enableRR == 0 and pc.enableNEE == 1 and pc.numNEE == 1 and pc.enableMIS == 0 and pc.usePowerHeuristic == 1
    assert pc.irradianceA == pytest.approx(0.2) and pc.irradianceUpdateProb == pytest.approx(1e-5) and pc.irradianceCreateProb == pytest.approx(5e-4)
    assert pc.irradianceGradientsMaxLength == 5 and pc.irradianceCacheMinRadius == pytest.approx(0.1) and pc.adrrsS == 5 and pc.adrrsSplit == 1
    assert pc.guidingProb == 0.5 and pc.useParallaxCompensation == 1 and pc.useGuiding == 0 and pc.updateGuiding == 0


def test_no_cpu_fallback_without_device():
    P = helpers.pt()
    if helpers.has_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(P.B200ptError, match="no CUDA device"):
        P.Renderer(16, 16)


def test_exr_round_trip(tmp_path):
    P = helpers.pt()
    rng = np.random.default_rng(0)
    img = rng.uniform(0, 4, (9, 13, 4)).astype(np.float32)
    path = str(tmp_path / "t.exr")
    P.write_exr(path, img)
    back = P.read_exr(path)
    # the reader mirrors CommonOps::readEXR, which goes through half-precision Imf::Rgba
    assert back.shape == img.shape
    assert np.allclose(back[..., :3], img[..., :3], rtol=1e-3)
    assert np.all(back[..., 3] == 1.0)
    import cv2  # independent OpenEXR implementation
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cvimg = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if cvimg is not None:
        assert np.array_equal(cvimg[..., ::-1], img[..., :3])


def test_cli_usage_and_no_cpu_fallback():
    """the headless command line (src/main.cpp:8-35): usage without arguments; without a CUDA device it fails loudly instead of rendering
    on the CPU"""
    import os
    import subprocess
    cli = os.path.join(helpers.PKG_DIR, "b200pt")
    p = subprocess.run([cli], capture_output=True, text=True)
    assert p.returncode != 0 and "WIDTH HEIGHT IC_SIZE GUIDING_SPLITS" in p.stdout
    if helpers.has_gpu():
        return
    p = subprocess.run([cli, "64", "36", "0", "0", helpers.scene_path("cornell-dielectric"), "--frames=1"], capture_output=True, text=True)
    assert p.returncode != 0 and "no CUDA device" in (p.stdout + p.stderr)
