#!/usr/bin/env python3
"""Generates the committed golden fixtures from the reference's own converged images (run in the build container,
where /root/reference exists):  1280x720 float EXRs next to the scenes (Mitsuba renders) are box-filtered
down to 160x90 and stored as float16 .npy, small enough to commit.  Usage: python tests/golden/make_golden.py"""
import os
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2
import numpy as np

REF = "/root/reference/scenes"
HERE = os.path.dirname(os.path.abspath(__file__))
IMAGES = {"cornell-dielectric": "cornell-dielectric/cornell-dielectric.exr", "veachMIS": "veachMIS/veachMIS.exr", "miPhong": "miPhong/miPhong.exr",
          "envMap": "envMap/envMap.exr", "testSpheres": "testSpheres/testSpheres.exr", "irradianceCache": "irradianceCache/irradianceCache.exr",
          "sponzaXML": "sponzaXML/sponzaXML_360spp_10m40.exr"}
# (roughConductor/roughConductor.exr is not usable: the rough-conductor object of that scene is shell.obj, a missing blob of the checkout)

for name, rel in IMAGES.items():
    img = cv2.imread(os.path.join(REF, rel), cv2.IMREAD_UNCHANGED)[..., :3][..., ::-1].astype(np.float64)
    h, w, _ = img.shape
    assert (w, h) == (1280, 720)
    small = img.reshape(90, 8, 160, 8, 3).mean(axis=(1, 3))
    np.save(os.path.join(HERE, name + "_160x90.npy"), small.astype(np.float16))
    print(name, small.mean(axis=(0, 1)))
