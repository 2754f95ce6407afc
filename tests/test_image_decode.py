"""Bitmap texture decoding (host/image_decode.cpp) against the reference's decoder: stb_image.h compiled from the
reference tree (oracle/_ref/libstb_ref.so, when present) and the committed digests of its output.  CPU only."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import helpers

DIGESTS = json.load(open(os.path.join(helpers.ROOT, "tests", "golden", "image_digests.json")))
STB = os.path.join(helpers.ROOT, "oracle", "_ref", "libstb_ref.so")


@pytest.mark.parametrize("rel", sorted(DIGESTS))
def test_decoder_matches_reference_digest(rel):
    img = helpers.pt().read_image_file(os.path.join(helpers.ROOT, rel))
    want = DIGESTS[rel]
    assert img.shape == (want["height"], want["width"], 4) and img.dtype == np.uint8
    assert hashlib.sha256(img.tobytes()).hexdigest() == want["sha256"]      # byte-exact


def test_decoder_matches_stb_image_live():
    if not os.path.exists(STB):
        pytest.skip("oracle/_ref/libstb_ref.so not built (needs the reference tree)")
    S = C.CDLL(STB)
    S.stb_ref_load.restype = C.POINTER(C.c_uint8)
    S.stb_ref_load.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    S.stb_ref_free.argtypes = [C.c_void_p]
    for rel in sorted(DIGESTS):
        path = os.path.join(helpers.ROOT, rel)
        w, h = C.c_int(), C.c_int()
        p = S.stb_ref_load(path.encode(), C.byref(w), C.byref(h))
        ref = np.ctypeslib.as_array(p, shape=(h.value, w.value, 4)).copy()
        S.stb_ref_free(p)
        assert np.array_equal(helpers.pt().read_image_file(path), ref), rel


def test_png_alpha_channel_is_kept():
    img = helpers.pt().read_image_file(os.path.join(helpers.ROOT, "tests", "golden", "images", "leaf.png"))
    assert (img[..., 3] == 0).any() and (img[..., 3] == 255).any()


def test_bad_files_fail_loudly(tmp_path):
    P = helpers.pt()
    with pytest.raises(P.B200ptError):
        P.read_image_file(str(tmp_path / "missing.jpg"))
    bad = tmp_path / "bad.jpg"
    bad.write_bytes(b"\xff\xd8\xff\xc2" + b"\x00" * 64)        # progressive frame header
    with pytest.raises(P.B200ptError):
        P.read_image_file(str(bad))
    junk = tmp_path / "junk.png"
    junk.write_bytes(b"not an image at all")
    with pytest.raises(P.B200ptError):
        P.read_image_file(str(junk))


def test_sponza_scene_loads_with_textures():
    P = helpers.pt()
    scene = P.Scene(helpers.scene_path("sponzaXML"))
    d = scene.desc
    assert scene.num_triangles == 66445 and d.num_spheres == 1
    assert d.num_textures == 11                                            # slot 0 (env map) + 10 JPGs
    used = {d.materials[i].textureIdDiffuse for i in range(d.num_materials)}
    assert used - {-1} == set(range(1, 11))


EXR_DIGESTS = json.load(open(os.path.join(helpers.ROOT, "tests", "golden", "exr_digests.json")))


@pytest.mark.parametrize("rel", sorted(EXR_DIGESTS))
def test_exr_reader_matches_openexr_digest(rel):
    """host/exr.cpp against OpenEXR's own decode (digests made with tests/golden/make_exr_digests.py): the reference's
    PIZ-compressed environment map and the uncompressed file our writer produces."""
    img = helpers.pt().read_exr(os.path.join(helpers.ROOT, rel))
    want = EXR_DIGESTS[rel]
    assert img.shape == (want["height"], want["width"], 4)
    assert hashlib.sha256(np.ascontiguousarray(img[..., :3]).tobytes()).hexdigest() == want["sha256"]
    assert {EXR_DIGESTS[k]["compression"] for k in EXR_DIGESTS} >= {0, 4}
