"""Shared test plumbing: loads the product binding (the package directory has a hyphen, so importlib by path) and the
oracle wrapper, and builds matching renderer / oracle pairs on the same scene, camera and push constants."""
import ctypes as C
import importlib.util
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "rtx-pathtracer_b200")
SCENES = os.path.join(ROOT, "scenes")


def _load(name, path):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def pt():
    return _load("b200pt_binding", os.path.join(PKG_DIR, "b200pt.py"))


def oracle():
    return _load("b200pt_oracle", os.path.join(ROOT, "oracle", "oracle.py"))


def ensure_built():
    if not os.path.exists(os.path.join(PKG_DIR, "libb200pt.so")) or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle_tracer.so")):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()


def has_gpu():
    try:
        return pt().lib().b200pt_device_count() > 0
    except Exception:
        return False


def scene_path(name):
    xml = os.path.join(SCENES, name, name + ".xml")
    return xml if os.path.exists(xml) else os.path.join(SCENES, name, name + ".json")     # the reference's two scene formats


def make_pair(scene_name, width, height, ic_size=0, guiding_splits=0, accel=True, gpu=True):
    """Returns (scene, renderer or None, oracle) set up identically."""
    P, O = pt(), oracle()
    scene = P.Scene(scene_path(scene_name))
    view, proj = scene.camera_matrices(width / height)
    vinv, pinv = P.mat4_inverse(view), P.mat4_inverse(proj)
    r = None
    if gpu:
        r = P.Renderer(width, height, ic_size, guiding_splits)
        r.set_scene(scene)
        r.set_camera(view, proj)
    o = O.TracerOracle(width, height, ic_size, accel=accel)
    o.set_scene(scene.desc)
    o.set_camera(view, proj, vinv, pinv)
    return scene, r, o


def camera_rays(scene, width, height, seed=1, jitter=True):
    """Primary rays through pixel centres (+ jitter), computed in float64 then rounded: inputs for traversal parity."""
    P = pt()
    view, proj = scene.camera_matrices(width / height)
    vinv = P.mat4_inverse(view).astype(np.float64).reshape(4, 4).T
    pinv = P.mat4_inverse(proj).astype(np.float64).reshape(4, 4).T
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:height, 0:width]
    jx = rng.uniform(-0.5, 0.5, xs.shape) if jitter else 0.0
    jy = rng.uniform(-0.5, 0.5, xs.shape) if jitter else 0.0
    u = (xs + 0.5 + jx) / width * 2 - 1
    v = (ys + 0.5 + jy) / height * 2 - 1
    clip = np.stack([u, v, np.ones_like(u), np.ones_like(u)], -1).reshape(-1, 4)
    t = clip @ pinv.T
    d = t[:, :3] / np.linalg.norm(t[:, :3], axis=1, keepdims=True)
    d = d @ vinv[:3, :3].T
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros(d.shape[0], dtype=P.RAY_DTYPE)
    rays["origin"] = vinv[:3, 3].astype(np.float32)
    rays["dir"] = d.astype(np.float32)
    rays["tmin"] = 1e-3
    rays["tmax"] = 1e6
    return rays


def secondary_rays(scene_oracle, primary, seed=2, shadow=False):
    """Incoherent rays: start at the primary hit points (oracle), uniformly random directions."""
    hits = scene_oracle.trace_rays(primary, threads=os.cpu_count() or 1)
    ok = hits["prim"] != 0xFFFFFFFF
    rng = np.random.default_rng(seed)
    o = primary["origin"][ok] + primary["dir"][ok] * hits["t"][ok, None]
    d = rng.normal(size=o.shape)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros(o.shape[0], dtype=primary.dtype)
    rays["origin"] = o.astype(np.float32)
    rays["dir"] = d.astype(np.float32)
    rays["tmin"] = 1e-3
    rays["tmax"] = rng.uniform(0.2, 3.0, o.shape[0]).astype(np.float32) if shadow else 1e6
    return rays


def scene_box(scene_name="cornell-dielectric"):
    """SceneLoader::calculateSceneSize box of a scene exactly as the loader computes it (float32)."""
    scene = pt().Scene(scene_path(scene_name))
    return [float(x) for x in scene.desc.scene_min[:]], [float(x) for x in scene.desc.scene_max[:]]
