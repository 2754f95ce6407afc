"""Generates scenes/envSynthetic/envmap.exr: a small procedural lat-long environment (sky gradient + a warm sun lobe +
dark ground) written with the product's own EXR writer.  Stand-in for the reference's envMap/envmap.exr, which is
PIZ-compressed (the host EXR reader supports NONE / ZIPS / ZIP)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402

W, H = 128, 64
v, u = np.meshgrid((np.arange(H) + 0.5) / H, (np.arange(W) + 0.5) / W, indexing="ij")
theta, phi = v * np.pi, u * 2 * np.pi
d = np.stack([np.sin(theta) * np.sin(phi), np.cos(theta), -np.sin(theta) * np.cos(phi)], -1)
sun = np.array([0.4, 0.7, 0.59]); sun /= np.linalg.norm(sun)
sky = np.where(d[..., 1:2] > 0, np.array([0.35, 0.55, 0.9]) * (0.3 + 0.7 * d[..., 1:2]), np.array([0.08, 0.07, 0.06]))
lobe = np.exp(40.0 * (d @ sun - 1.0))[..., None] * np.array([30.0, 24.0, 16.0])
img = np.concatenate([sky + lobe, np.ones((H, W, 1))], -1).astype(np.float32)
os.makedirs(os.path.join(ROOT, "scenes", "envSynthetic"), exist_ok=True)
helpers.pt().write_exr(os.path.join(ROOT, "scenes", "envSynthetic", "envmap.exr"), img)
print("wrote", img.shape, float(img[..., :3].mean()))
