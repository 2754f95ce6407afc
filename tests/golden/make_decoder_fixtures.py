"""Writes the small progressive / restart-interval JPEG and Adam7 PNG fixtures of tests/golden/images (run once in the
build container: needs PIL and OpenCV for the JPEG encoders; the PNGs come from make_interlaced_png.py).  Afterwards run
make_image_digests.py, which stores what the reference's stb_image decodes from them."""
import os
import sys

import cv2
import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_interlaced_png as M

OUT = os.path.join(HERE, "images")
rng = np.random.default_rng(11)


def picture(w, h):
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([(x * 7 + y * 3) % 256, (x * y) % 256, (x * 2 + y * 5) % 256], -1)
    return (base + rng.integers(-25, 25, base.shape)).clip(0, 255).astype(np.uint8)


Image.fromarray(picture(33, 47), "RGB").save(os.path.join(OUT, "prog_444.jpg"), quality=85, progressive=True, optimize=True, subsampling=0)
Image.fromarray(picture(50, 21), "RGB").save(os.path.join(OUT, "prog_422.jpg"), quality=60, progressive=True, optimize=True, subsampling=1)
Image.fromarray(picture(41, 30), "RGB").save(os.path.join(OUT, "prog_420.jpg"), quality=40, progressive=True, optimize=True, subsampling=2)
Image.fromarray(picture(19, 19)[..., 0], "L").save(os.path.join(OUT, "prog_gray.jpg"), quality=75, progressive=True)
cv2.imwrite(os.path.join(OUT, "prog_rst.jpg"), picture(100, 37), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1, cv2.IMWRITE_JPEG_RST_INTERVAL, 3, cv2.IMWRITE_JPEG_QUALITY, 75])
cv2.imwrite(os.path.join(OUT, "base_rst.jpg"), picture(70, 45), [cv2.IMWRITE_JPEG_PROGRESSIVE, 0, cv2.IMWRITE_JPEG_RST_INTERVAL, 2, cv2.IMWRITE_JPEG_QUALITY, 80])
for name, colour, depth in M.cases():
    M.make(os.path.join(OUT, "adam7_%s.png" % name), name, colour, depth, 13, 11, 1, seed=depth * 10 + colour)
print(len(os.listdir(OUT)), "files in", OUT)
