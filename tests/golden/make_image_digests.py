"""Regenerates tests/golden/image_digests.json: SHA-256 of the RGBA8 bytes the REFERENCE's decoder (stb_image.h, compiled
from /root/reference into oracle/_ref/libstb_ref.so by oracle/Makefile) produces for the bitmap textures kept in this
repository.  Run in the build container (needs /root/reference); the digests travel, the reference does not."""
import ctypes as C
import glob
import hashlib
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
S = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libstb_ref.so"))
S.stb_ref_load.restype = C.POINTER(C.c_uint8)
S.stb_ref_load.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
S.stb_ref_free.argtypes = [C.c_void_p]

out = {}
files = sorted(glob.glob(os.path.join(ROOT, "scenes", "sponzaXML", "textures", "*")) + glob.glob(os.path.join(ROOT, "tests", "golden", "images", "*")))
for f in files:
    w, h = C.c_int(), C.c_int()
    p = S.stb_ref_load(f.encode(), C.byref(w), C.byref(h))
    assert p, f
    a = np.ctypeslib.as_array(p, shape=(h.value, w.value, 4)).copy()
    S.stb_ref_free(p)
    out[os.path.relpath(f, ROOT)] = {"width": w.value, "height": h.value, "sha256": hashlib.sha256(a.tobytes()).hexdigest()}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "image_digests.json"), "w"), indent=1, sort_keys=True)
print(len(out), "digests written")
