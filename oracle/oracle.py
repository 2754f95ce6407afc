"""ORACLE — test infrastructure only.  ctypes wrapper of oracle/liboracle_tracer.so (tracer_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_tracer.so")
_lib = None

HIT_DTYPE = np.dtype([("t", "<f4"), ("prim", "<u4"), ("u", "<f4"), ("v", "<f4")])


def build():
    subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_scene.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_set_camera.argtypes = [C.c_void_p] + [C.POINTER(C.c_float)] * 4
        L.oracle_render_region.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int]
        L.oracle_image.restype = C.POINTER(C.c_float)
        L.oracle_image.argtypes = [C.c_void_p, C.c_int]
        L.oracle_get_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.oracle_reset_counters.argtypes = [C.c_void_p]
        L.oracle_set_guiding.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_samples.restype = C.c_void_p
        L.oracle_samples.argtypes = [C.c_void_p]
        L.oracle_ic_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_ic_put.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


class TracerOracle:
    def __init__(self, width, height, ic_size=0, accel=True):
        self.width, self.height, self.ic_size = width, height, ic_size
        self._h = lib().oracle_create(width, height, ic_size, int(accel))

    def set_scene(self, desc):
        lib().oracle_set_scene(self._h, C.addressof(desc))

    def set_camera(self, view, proj, view_inv, proj_inv):
        arrs = [np.ascontiguousarray(x, dtype=np.float32) for x in (view, proj, view_inv, proj_inv)]
        lib().oracle_set_camera(self._h, *[a.ctypes.data_as(C.POINTER(C.c_float)) for a in arrs])

    def render_region(self, pc, x0=0, y0=0, x1=None, y1=None, threads=1):
        x1 = self.width if x1 is None else x1
        y1 = self.height if y1 is None else y1
        lib().oracle_render_region(self._h, C.addressof(pc), x0, y0, x1, y1, threads)

    def trace_rays(self, rays, any_hit=False, threads=1):
        rays = np.ascontiguousarray(rays)
        hits = np.empty(rays.shape[0], dtype=HIT_DTYPE)
        lib().oracle_trace_rays(self._h, rays.ctypes.data, rays.shape[0], hits.ctypes.data, int(any_hit), threads)
        return hits

    def image(self, which=0):
        p = lib().oracle_image(self._h, which)
        return np.ctypeslib.as_array(p, shape=(self.height, self.width, 4)).copy()

    def counters(self):
        out = (C.c_uint64 * 3)()
        lib().oracle_get_counters(self._h, out)
        return {"extend_rays": out[0], "shadow_rays": out[1], "path_vertices": out[2]}

    def reset_counters(self):
        lib().oracle_reset_counters(self._h)

    def set_image(self, which, rgba):
        p = lib().oracle_image(self._h, which)
        np.ctypeslib.as_array(p, shape=(self.height, self.width, 4))[...] = np.asarray(rgba, dtype=np.float32).reshape(self.height, self.width, 4)

    def ic_get(self, P):
        hdr = P.CacheHeader()
        data = np.zeros(self.ic_size, dtype=P.CACHE_DATA_DTYPE)
        spheres = np.zeros(self.ic_size, dtype=P.SPHERE_DTYPE)
        lib().oracle_ic_get(self._h, C.addressof(hdr), data.ctypes.data, spheres.ctypes.data, self.ic_size)
        return hdr, data, spheres

    def ic_put(self, hdr, data, spheres):
        d = np.ascontiguousarray(data)
        s = np.ascontiguousarray(spheres)
        lib().oracle_ic_put(self._h, C.addressof(hdr), d.ctypes.data, s.ctypes.data, d.shape[0])

    def set_guiding(self, aabbs, vmms):
        a = np.ascontiguousarray(aabbs)
        v = np.ascontiguousarray(vmms)
        lib().oracle_set_guiding(self._h, a.ctypes.data, v.ctypes.data, a.shape[0])

    def samples(self, dtype):
        n = self.width * self.height * 16
        p = lib().oracle_samples(self._h)
        buf = (C.c_char * (n * dtype.itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dtype).copy()

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- guiding fit: the reference's own lightpmm / guiding headers compiled into oracle/_ref/libguiding_ref.so ----------
REF_GUIDING_PATH = os.path.join(_HERE, "_ref", "libguiding_ref.so")
_glib = None

DIRECTIONAL_DATA_DTYPE = np.dtype([("position", "<f4", 3), ("direction", "<f4", 3), ("weight", "<f4"), ("pdf", "<f4"),
                                   ("distance", "<f4"), ("flags", "<u4")])
VMF_THETA_DTYPE = np.dtype([("mu", "<f4", 3), ("k", "<f4"), ("norm", "<f4"), ("eMin2K", "<f4"), ("distance", "<f4"),
                            ("target", "<f4", 3)])
VMM_THETA_DTYPE = np.dtype([("thetas", VMF_THETA_DTYPE, 16), ("pi", "<f4", 16), ("meanPosition", "<f4", 3),
                            ("usedDistributions", "<i4")])
AABB_DTYPE = np.dtype([("min", "<f4", 3), ("max", "<f4", 3)])
STATE_FIELDS = ("weight", "kappa", "r", "mux", "muy", "muz", "distance", "distSumW", "chi", "chiN", "covxx", "covyy", "covxy", "covSumW")


def ref_guiding_available():
    return os.path.exists(REF_GUIDING_PATH)


def glib():
    """oracle/_ref/libguiding_ref.so — built by oracle/Makefile only where /root/reference exists; the prebuilt file travels."""
    global _glib
    if _glib is None:
        if not os.path.exists(REF_GUIDING_PATH):
            build()
        if not os.path.exists(REF_GUIDING_PATH):
            raise RuntimeError("oracle/_ref/libguiding_ref.so is missing and /root/reference is not here to build it")
        L = C.CDLL(REF_GUIDING_PATH)
        L.refguiding_create.restype = C.c_void_p
        L.refguiding_create.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]
        L.refguiding_destroy.argtypes = [C.c_void_p]
        L.refguiding_region_count.argtypes = [C.c_void_p]
        L.refguiding_get_aabbs.argtypes = [C.c_void_p, C.c_void_p]
        L.refguiding_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.refguiding_get_vmms.argtypes = [C.c_void_p, C.c_void_p]
        L.refguiding_em_sample_iterations.restype = C.c_uint64
        L.refguiding_em_sample_iterations.argtypes = [C.c_void_p]
        L.refguiding_sorted_count.restype = C.c_int64
        L.refguiding_sorted_count.argtypes = [C.c_void_p]
        L.refguiding_get_sorted.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.refguiding_get_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.refguiding_fastexp.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _glib = L
    return _glib


class GuidingRef:
    """PathGuiding (region tree + per-region lightpmm mixtures) on the CPU: the reference's headers + restated glue."""

    def __init__(self, splits, scene_min, scene_max, params):
        self._params = params      # ctypes b200pt_guiding_params (same layout as include/b200pt.h)
        mn = (C.c_float * 3)(*scene_min)
        mx = (C.c_float * 3)(*scene_max)
        self._h = glib().refguiding_create(splits, mn, mx, C.addressof(params))
        self.region_count = glib().refguiding_region_count(self._h)

    def aabbs(self):
        out = np.empty(self.region_count, dtype=AABB_DTYPE)
        glib().refguiding_get_aabbs(self._h, out.ctypes.data)
        return out

    def update(self, samples, threads=1):
        a = np.ascontiguousarray(samples, dtype=DIRECTIONAL_DATA_DTYPE)
        glib().refguiding_update(self._h, a.ctypes.data, a.shape[0], threads)
        self.region_count = glib().refguiding_region_count(self._h)      # adaptive refinement may have split regions

    def vmms(self):
        out = np.empty(self.region_count, dtype=VMM_THETA_DTYPE)
        glib().refguiding_get_vmms(self._h, out.ctypes.data)
        return out

    def em_sample_iterations(self):
        return int(glib().refguiding_em_sample_iterations(self._h))

    def sorted_samples(self):
        n = glib().refguiding_sorted_count(self._h)
        out = np.empty(n, dtype=DIRECTIONAL_DATA_DTYPE)
        off = np.empty(self.region_count + 1, dtype=np.uint32)
        glib().refguiding_get_sorted(self._h, out.ctypes.data, off.ctypes.data)
        return out, off

    def state(self, region):
        sc = np.empty(5, dtype=np.float32)
        pc = np.empty((14, 16), dtype=np.float32)
        glib().refguiding_get_state(self._h, region, sc.ctypes.data, pc.ctypes.data)
        d = {"K": int(sc[0]), "sampleWeight": float(sc[1]), "numSamples": float(sc[2]), "totalNumSamples": int(sc[3]), "numEMIterations": int(sc[4])}
        for i, f in enumerate(STATE_FIELDS):
            d[f] = pc[i].copy()
        return d

    def close(self):
        if self._h:
            glib().refguiding_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ref_fastexp(x):
    a = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(a)
    glib().refguiding_fastexp(a.ctypes.data, out.ctypes.data, a.size)
    return out
