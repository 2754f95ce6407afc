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
    bad.write_bytes(b"\xff\xd8\xff\xc2" + b"\x00" * 64)        # frame header without a body
    with pytest.raises(P.B200ptError):
        P.read_image_file(str(bad))
    arith = tmp_path / "arith.jpg"
    arith.write_bytes(b"\xff\xd8\xff\xc9\x00\x0b\x08\x00\x08\x00\x08\x01\x01\x11\x00" + b"\x00" * 32)   # SOF9: arithmetic coding
    with pytest.raises(P.B200ptError):
        P.read_image_file(str(arith))
    junk = tmp_path / "junk.png"
    junk.write_bytes(b"not an image at all")
    with pytest.raises(P.B200ptError):
        P.read_image_file(str(junk))


def test_interlaced_png_equals_plain_png(tmp_path):
    """Adam7 (RFC 2083 section 2.6) only reorders the pixels: the interlaced and the plain file of the same samples must decode to
    the same RGBA bytes — every colour type and bit depth, sizes that leave some of the seven passes empty, all five filters."""
    import sys
    sys.path.insert(0, os.path.join(helpers.ROOT, "tests", "golden"))
    import make_interlaced_png as M
    P = helpers.pt()
    for (w, h) in ((1, 1), (2, 3), (5, 2), (8, 8), (9, 17), (33, 20)):
        for name, colour, depth in M.cases():
            out = []
            for interlace in (0, 1):
                path = str(tmp_path / ("%s_%dx%d_%d.png" % (name, w, h, interlace)))
                M.make(path, name, colour, depth, w, h, interlace, seed=w * 100 + h)
                out.append(P.read_image_file(path))
            assert out[0].shape == (h, w, 4) and np.array_equal(out[0], out[1]), (name, w, h)


def test_progressive_jpeg_fixture_kinds():
    """the committed fixtures do cover what their names say (SOF2 frames, restart markers, several scans)"""
    img_dir = os.path.join(helpers.ROOT, "tests", "golden", "images")
    for name, sof, rst in (("prog_444.jpg", 0xC2, False), ("prog_422.jpg", 0xC2, False), ("prog_420.jpg", 0xC2, False), ("prog_gray.jpg", 0xC2, False),
                           ("prog_rst.jpg", 0xC2, True), ("base_rst.jpg", 0xC0, True)):
        d = open(os.path.join(img_dir, name), "rb").read()
        markers = [d[i + 1] for i in range(len(d) - 1) if d[i] == 0xFF and d[i + 1] not in (0x00, 0xFF)]
        assert sof in markers and (0xD0 in markers) == rst, name
        assert markers.count(0xDA) >= (2 if sof == 0xC2 else 1), name
        assert "tests/golden/images/" + name in DIGESTS


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
    PIZ-compressed environment map, the uncompressed file our writer produces, and one small HALF and FLOAT file per further
    compression (tests/golden/exr, written by make_exr_fixtures.py)."""
    img = helpers.pt().read_exr(os.path.join(helpers.ROOT, rel))
    want = EXR_DIGESTS[rel]
    assert img.shape == (want["height"], want["width"], 4)
    assert hashlib.sha256(np.ascontiguousarray(img[..., :3]).tobytes()).hexdigest() == want["sha256"]
    assert {EXR_DIGESTS[k]["compression"] for k in EXR_DIGESTS} >= {0, 1, 2, 3, 4, 5, 6, 7}        # every compression but DWAA / DWAB


def _expect_loader_error(fn, path):
    P = helpers.pt()
    with pytest.raises(P.B200ptError):
        fn(path)


def test_corrupt_exr_files_raise_instead_of_corrupting_memory(tmp_path):
    """A malformed .exr in a scene directory must end in the loader's error, not in out-of-bounds reads or writes
    (header attribute sizes, channel list, chunk-offset table, chunk y and size are all untrusted)."""
    import struct
    P = helpers.pt()
    rng = np.random.default_rng(3)
    good = str(tmp_path / "good.exr")
    P.write_exr(good, rng.uniform(0, 2, (7, 9, 4)).astype(np.float32))
    data = bytearray(open(good, "rb").read())
    assert P.read_exr(good).shape == (7, 9, 4)

    def variant(name, mutate):
        d = bytearray(data)
        mutate(d)
        p = str(tmp_path / name)
        open(p, "wb").write(d)
        _expect_loader_error(P.read_exr, p)

    for cut in (6, 20, 60, len(data) // 2, len(data) - 5):           # truncations everywhere
        variant("cut%d.exr" % cut, lambda d, cut=cut: d.__delitem__(slice(cut, None)))
    # attribute size of the first attribute blown up / negative
    first = data.index(b"\x00", 8)                                    # end of the first attribute name
    tpos = data.index(b"\x00", first + 1) + 1                         # after its type string: the int32 size
    variant("bigattr.exr", lambda d: d.__setitem__(slice(tpos, tpos + 4), struct.pack("<i", 0x7fffff00)))
    variant("negattr.exr", lambda d: d.__setitem__(slice(tpos, tpos + 4), struct.pack("<i", -8)))
    # data window turned inside out / huge
    dwp = data.index(b"dataWindow\x00box2i\x00") + len(b"dataWindow\x00box2i\x00") + 4
    variant("dwneg.exr", lambda d: d.__setitem__(slice(dwp, dwp + 16), struct.pack("<4i", 5, 5, 0, 0)))
    variant("dwhuge.exr", lambda d: d.__setitem__(slice(dwp, dwp + 16), struct.pack("<4i", 0, 0, 2000000000, 2000000000)))
    # the chunk table: an offset beyond the file, and a chunk whose y lies outside the data window / whose size is bogus
    hdr_end = data.index(b"\x00\x00", dwp) if False else None
    # locate the offset table: it follows the header's terminating zero byte; find it by parsing like the reader does
    i = 8
    while data[i] != 0:
        i = data.index(b"\x00", i) + 1
        i = data.index(b"\x00", i) + 1
        sz = struct.unpack_from("<i", data, i)[0]
        i += 4 + sz
    table = i + 1
    off0 = struct.unpack_from("<Q", data, table)[0]
    variant("badoff.exr", lambda d: struct.pack_into("<Q", d, table, len(data) + 1000))
    variant("bady.exr", lambda d: struct.pack_into("<i", d, off0, 123456))
    variant("badsize.exr", lambda d: struct.pack_into("<i", d, off0 + 4, 0x7ffffff0))
    variant("negsize.exr", lambda d: struct.pack_into("<i", d, off0 + 4, -4))


def test_corrupt_jpeg_segments_raise(tmp_path):
    P = helpers.pt()
    src = os.path.join(helpers.ROOT, "tests", "golden", "images", "cube.jpg")
    data = bytearray(open(src, "rb").read())
    assert P.read_image_file(src).ndim == 3
    raised = 0
    for marker in (b"\xff\xdb", b"\xff\xc0", b"\xff\xc4", b"\xff\xda"):
        pos = data.index(marker)
        d = bytearray(data)
        d[pos + 2:pos + 4] = (2).to_bytes(2, "big")                  # segment length shrunk to its minimum: the body is gone
        p = str(tmp_path / ("seg%02x.jpg" % marker[1]))
        open(p, "wb").write(d)
        try:                                                          # (the byte pair may also sit inside an APPn payload: then the file still decodes)
            P.read_image_file(p)
        except P.B200ptError:
            raised += 1
    assert raised >= 2
    for cut in (3, 30, 200):                                          # inside the header segments
        p = str(tmp_path / ("cut%d.jpg" % cut))
        open(p, "wb").write(data[:cut])
        _expect_loader_error(P.read_image_file, p)
    p = str(tmp_path / "cutscan.jpg")                                 # inside the entropy-coded data: like stb_image the decoder pads
    open(p, "wb").write(data[:len(data) // 3])                        # the missing bits with zeros — it must not read past the buffer
    try:
        assert P.read_image_file(p).ndim == 3
    except P.B200ptError:
        pass


def test_corrupted_bitmaps_raise_or_decode_but_never_hang(tmp_path):
    """Seeded byte flips in the header region and truncations of every fixture (progressive / restart JPEGs, Adam7 PNGs): the
    decoders either return an image or raise the loader's error; a corrupt size field is refused before it becomes an
    allocation of gigabytes (the whole loop runs in a second or two)."""
    import glob
    import time
    P = helpers.pt()
    rng = np.random.default_rng(2)
    files = sorted(glob.glob(os.path.join(helpers.ROOT, "tests", "golden", "images", "*")))
    t0 = time.time()
    outcomes = {"ok": 0, "raised": 0}
    for it in range(600):
        f = files[it % len(files)]
        d = bytearray(open(f, "rb").read())
        if rng.integers(0, 3) == 0:
            d = d[:rng.integers(1, len(d))]
        else:
            for _ in range(rng.integers(1, 6)):
                d[rng.integers(2, min(len(d), 400))] = rng.integers(0, 256)
        p = str(tmp_path / ("fuzz" + os.path.splitext(f)[1]))
        open(p, "wb").write(d)
        try:
            img = P.read_image_file(p)
            assert img.ndim == 3 and img.shape[2] == 4
            outcomes["ok"] += 1
        except P.B200ptError:
            outcomes["raised"] += 1
    assert outcomes["ok"] > 20 and outcomes["raised"] > 200, outcomes
    assert time.time() - t0 < 60


def test_corrupted_exr_chunks_of_every_compression_raise_or_decode(tmp_path):
    """Seeded byte flips and truncations of the RLE / ZIPS / ZIP / PXR24 / B44 / B44A fixtures: runs, byte planes and 4x4 blocks
    that point past their chunk end in the loader's error, never in an out-of-bounds access."""
    import glob
    P = helpers.pt()
    rng = np.random.default_rng(4)
    files = sorted(glob.glob(os.path.join(helpers.ROOT, "tests", "golden", "exr", "*.exr")))
    assert len(files) == 12
    outcomes = {"ok": 0, "raised": 0}
    for it in range(600):
        d = bytearray(open(files[it % len(files)], "rb").read())
        if rng.integers(0, 3) == 0:
            d = d[:rng.integers(1, len(d))]
        else:
            for _ in range(rng.integers(1, 6)):
                d[rng.integers(8, len(d))] = rng.integers(0, 256)
        p = str(tmp_path / "fuzz.exr")
        open(p, "wb").write(d)
        try:
            assert P.read_exr(p).shape[2] == 4
            outcomes["ok"] += 1
        except P.B200ptError:
            outcomes["raised"] += 1
    assert outcomes["ok"] > 50 and outcomes["raised"] > 200, outcomes
