"""Regenerates tests/golden/exr_digests.json: SHA-256 of the float32 RGB pixels that OpenEXR itself (through OpenCV's
imread, build container only) decodes from the EXR files kept in this repository — the known answers for the host's own
EXR reader (host/exr.cpp: NONE / RLE / ZIPS / ZIP / PIZ / PXR24 / B44 / B44A)."""
import glob
import hashlib
import json
import os

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2  # noqa: E402
import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = {}
for f in sorted(glob.glob(os.path.join(ROOT, "scenes", "**", "*.exr"), recursive=True) + glob.glob(os.path.join(ROOT, "tests", "golden", "exr", "*.exr"))):
    img = cv2.imread(f, cv2.IMREAD_UNCHANGED)
    # CommonOps::readEXR goes through Imf::RgbaInputFile: every value passes through half precision (a no-op for HALF files)
    rgb = np.ascontiguousarray(img[..., 2::-1].astype(np.float32).astype(np.float16).astype(np.float32))
    comp = open(f, "rb").read(4096)
    i = comp.find(b"compression\x00compression\x00")
    out[os.path.relpath(f, ROOT)] = {"width": int(rgb.shape[1]), "height": int(rgb.shape[0]), "compression": int(comp[i + 28]),
                                     "sha256": hashlib.sha256(rgb.tobytes()).hexdigest()}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "exr_digests.json"), "w"), indent=1, sort_keys=True)
print(out)
