"""Writes the small EXR fixtures of tests/golden/exr (one per compression the host reader handles beyond the scene files'
PIZ / NONE: RLE, ZIPS, ZIP, PXR24, B44, B44A; HALF and FLOAT) with OpenEXR through OpenCV — build container only.  Afterwards
run make_exr_digests.py, which stores what OpenEXR itself decodes from them."""
import os

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2  # noqa: E402
import numpy as np  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "exr")
os.makedirs(OUT, exist_ok=True)
rng = np.random.default_rng(3)
w, h = 37, 41                 # not a multiple of the 4x4 blocks, of the 16- and of the 32-line chunks
y, x = np.mgrid[0:h, 0:w]
img = np.stack([np.sin(x * 0.1) + 1.5 + 0.1 * rng.random((h, w)), (x + y) % 7 * 0.25, np.where((x // 8 + y // 8) % 2, 3.0, 0.0)], -1).astype(np.float32)
img[h // 2:, :w // 3] = 0.75  # a flat area: B44A's 3-byte blocks, RLE runs
img[0, 1] = -2.5
img[1, 2] = 1e4
COMP = {"rle": cv2.IMWRITE_EXR_COMPRESSION_RLE, "zips": cv2.IMWRITE_EXR_COMPRESSION_ZIPS, "zip": cv2.IMWRITE_EXR_COMPRESSION_ZIP,
        "pxr24": cv2.IMWRITE_EXR_COMPRESSION_PXR24, "b44": cv2.IMWRITE_EXR_COMPRESSION_B44, "b44a": cv2.IMWRITE_EXR_COMPRESSION_B44A}
for name, comp in COMP.items():
    for tn, typ in (("half", cv2.IMWRITE_EXR_TYPE_HALF), ("float", cv2.IMWRITE_EXR_TYPE_FLOAT)):
        assert cv2.imwrite(os.path.join(OUT, "%s_%s.exr" % (name, tn)), img[..., ::-1], [cv2.IMWRITE_EXR_COMPRESSION, comp, cv2.IMWRITE_EXR_TYPE, typ])
print(sorted(os.listdir(OUT)))
