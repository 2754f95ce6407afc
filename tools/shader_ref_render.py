#!/usr/bin/env python
"""Render a scene with the REFERENCE'S OWN SHADER SOURCE compiled as C++ (oracle/_ref/libshader_ref.so, built from /root/reference by
oracle/Makefile) and with the oracle's restatement, and compare the two frames.  CPU only.

  python tools/shader_ref_render.py scenes/veachMIS/veachMIS.xml [--width 160 --height 90 --spp 4 --frames 2 --out ref.exr] [--<pushConstantField>=v ...]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import test_shader_ref as T  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene")
    ap.add_argument("--width", type=int, default=160)
    ap.add_argument("--height", type=int, default=90)
    ap.add_argument("--spp", type=int, default=4)
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--out", default=None)
    args, extra = ap.parse_known_args()
    kw = {}
    for e in extra:                      # --enableMIS=1 --maxDepth=8 ... like the command line of the product
        k, v = e.lstrip("-").split("=")
        kw[k] = float(v) if "." in v else int(v)
    P, O = helpers.pt(), helpers.oracle()
    T.W, T.H = args.width, args.height
    scene = P.Scene(args.scene)
    view, proj = scene.camera_matrices(args.width / args.height)
    o = O.TracerOracle(args.width, args.height, 0, accel=True)
    o.set_scene(scene.desc)
    o.set_camera(view, proj, P.mat4_inverse(view), P.mat4_inverse(proj))
    ref = T.ShaderRef(scene, view, proj, o)
    t_ref = t_ora = 0.0
    for f in range(args.frames):
        pc = P.default_push_constants(randomUInt=P.tea(f, 0xC0FFEE), previousFrames=f, samplesPerPixel=args.spp, **kw)
        t0 = time.time(); ref.render(pc); t_ref += time.time() - t0
        t0 = time.time(); o.render_region(pc, threads=os.cpu_count() or 1); t_ora += time.time() - t0
    a, b = ref.image(), o.image()
    print("compiled reference shaders: %.2f s (1 thread)   oracle: %.2f s (%d threads)" % (t_ref, t_ora, os.cpu_count() or 1))
    print("image mean %.6f / %.6f   pixels bit-equal: %.4f   within 1e-4: %.4f" % (
        a[..., :3].mean(), b[..., :3].mean(), float((a.view(np.uint32) == b.view(np.uint32)).all(-1).mean()),
        float((np.abs(a[..., :3] - b[..., :3]) <= 1e-4 * np.maximum(np.abs(a[..., :3]), 1e-3)).all(-1).mean())))
    if args.out:
        P.write_exr(args.out, a)
        print("wrote", args.out)


if __name__ == "__main__":
    main()
