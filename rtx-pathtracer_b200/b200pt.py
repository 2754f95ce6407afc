"""ctypes binding of libb200pt.so — the thin Python face of the C ABI in include/b200pt.h.

Used by tests/, bench.py and __graft_entry__.py.  It mirrors the reference's host objects by name
(SceneLoader -> Scene, RayTracingApp -> Renderer, RtPushConstant -> PushConstants) and never computes anything
itself: every call goes straight into the shared library, which fails loudly when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200PT_LIB") or os.path.join(_HERE, "libb200pt.so")   # B200PT_LIB: A/B a second build of the same library

SIZE_LIGHT_RANDOM = 10000
SIZE_TRI_RANDOM = 10000
MAX_DISTRIBUTIONS = 16
MAX_DIRECTIONAL_DATA_PER_PIXEL = 16
INVALID_REGION = 0xFFFFFFFF
MISS = 0xFFFFFFFF
IMAGE_OUTPUT, IMAGE_ACCUM, IMAGE_ESTIMATE = 0, 1, 2


class B200ptError(RuntimeError):
    pass


class Vertex(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("_pad0", C.c_float), ("normal", C.c_float * 3), ("_pad1", C.c_float),
                ("texCoord", C.c_float * 2), ("materialIndex", C.c_int32), ("_pad2", C.c_int32)]


class Material(C.Structure):
    _fields_ = [("lightColor", C.c_float * 3), ("_pad0", C.c_float), ("diffuse", C.c_float * 3), ("_pad1", C.c_float),
                ("specular", C.c_float * 3), ("specularHighlight", C.c_float), ("transparency", C.c_float),
                ("refractionIndex", C.c_float), ("refractionIndexInv", C.c_float), ("eta", C.c_float), ("k", C.c_float),
                ("roughness", C.c_float), ("textureIdDiffuse", C.c_int32), ("textureIdSpecular", C.c_int32),
                ("type", C.c_int32), ("_pad2", C.c_int32 * 3)]


class Instance(C.Structure):
    _fields_ = [("transform", C.c_float * 16), ("normalTransform", C.c_float * 16), ("modelIndex", C.c_int32),
                ("iLight", C.c_int32), ("_pad", C.c_int32 * 2)]


class Light(C.Structure):
    _fields_ = [("color", C.c_float * 3), ("pos", C.c_float * 3), ("instanceIndex", C.c_uint32),
                ("sampleProb", C.c_float), ("area", C.c_float), ("type", C.c_int32)]


class FaceSample(C.Structure):
    _fields_ = [("index", C.c_int32), ("sampleProb", C.c_float), ("faceArea", C.c_float)]


class Sphere(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("radius", C.c_float), ("materialIndex", C.c_int32), ("iLight", C.c_int32)]


class Aabb(C.Structure):
    _fields_ = [("min", C.c_float * 3), ("max", C.c_float * 3)]


class CacheHeader(C.Structure):
    _fields_ = [("nextCacheSlot", C.c_uint32), ("maxCaches", C.c_uint32), ("nextUpdateSlot", C.c_uint32)]


class Texture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("format", C.c_int32), ("_pad", C.c_int32),
                ("pixels", C.c_void_p)]


class SceneDesc(C.Structure):
    _fields_ = [("num_models", C.c_int32), ("vertices", C.POINTER(C.POINTER(Vertex))), ("num_vertices", C.POINTER(C.c_int32)),
                ("indices", C.POINTER(C.POINTER(C.c_uint32))), ("num_indices", C.POINTER(C.c_int32)),
                ("num_materials", C.c_int32), ("materials", C.POINTER(Material)),
                ("num_instances", C.c_int32), ("instances", C.POINTER(Instance)),
                ("num_lights", C.c_int32), ("lights", C.POINTER(Light)),
                ("random_light_index", C.POINTER(C.c_int32)),
                ("num_face_tables", C.c_int32), ("random_tri_index", C.POINTER(FaceSample)),
                ("num_spheres", C.c_int32), ("spheres", C.POINTER(Sphere)),
                ("num_textures", C.c_int32), ("textures", C.POINTER(Texture)),
                ("scene_min", C.c_float * 3), ("scene_max", C.c_float * 3)]


_PC_FIELDS = [
    ("randomUInt", C.c_uint32), ("previousFrames", C.c_uint32), ("maxDepth", C.c_int32), ("maxFollowDiscrete", C.c_int32),
    ("samplesPerPixel", C.c_int32), ("enableRR", C.c_int32), ("enableNEE", C.c_int32), ("numNEE", C.c_int32),
    ("enableAverageInsteadOfMix", C.c_int32), ("enableMIS", C.c_int32), ("usePowerHeuristic", C.c_int32),
    ("storeEstimate", C.c_int32), ("visualizeMode", C.c_int32), ("showIrradianceCacheOnly", C.c_int32),
    ("showIrradianceGradients", C.c_int32), ("useIrradianceCache", C.c_int32), ("highlightIrradianceCacheColor", C.c_int32),
    ("irradianceA", C.c_float), ("irradianceUpdateProb", C.c_float), ("irradianceCreateProb", C.c_float),
    ("irradianceVisualizationScale", C.c_float), ("useIrradianceGradients", C.c_int32), ("useIrradianceCacheOnGlossy", C.c_int32),
    ("irradianceGradientsMaxLength", C.c_float), ("isIrradiancePrepareFrame", C.c_int32), ("irradianceNumNEE", C.c_int32),
    ("irradianceCacheMinRadius", C.c_float), ("irradianceCachePerformVisibilityCheck", C.c_int32),
    ("useVisibleSphereSampling", C.c_int32), ("useADRRS", C.c_int32), ("adrrsS", C.c_float), ("adrrsSplit", C.c_int32),
    ("splitOnFirst", C.c_int32), ("useGuiding", C.c_int32), ("guidingProb", C.c_float), ("guidingVisuScale", C.c_float),
    ("guidingVisuMax", C.c_float), ("guidingVisuIgnoreOcclusioon", C.c_int32), ("updateGuiding", C.c_int32),
    ("useParallaxCompensation", C.c_int32), ("time", C.c_float), ("guidingVisuMove", C.c_int32),
    ("guidingVisuPhiScale", C.c_float), ("guidingVisuThetaScale", C.c_float), ("numGuidingRegions", C.c_int32),
    ("guidingPiPHighlightRegion", C.c_int32), ("guidingPiPShowSpheres", C.c_int32), ("guidingPiPSize", C.c_float)]


class BvhReport(C.Structure):
    """b200pt_bvh_report: the first six fields count violations (0 for a valid tree), the rest are statistics."""
    _fields_ = [(n, C.c_uint32) for n in ("missing_prims", "duplicate_prims", "outside_box", "bad_meta", "depth_mismatch", "unreachable_nodes",
                                          "num_nodes", "num_tris", "max_depth", "inner_children", "leaf_children")]


class PushConstants(C.Structure):
    """RtPushConstant — src/RayTracingApp.h:116-165 (192 bytes)."""
    _fields_ = _PC_FIELDS


class GuidingParams(C.Structure):
    _fields_ = [("useParallaxCompensation", C.c_int32), ("splitAndMerge", C.c_int32), ("minSamplesForMerging", C.c_int32),
                ("minSamplesForSplitting", C.c_int32), ("minSamplesForPostSplitFitting", C.c_int32),
                ("splitMinDivergence", C.c_float), ("mergeMaxDivergence", C.c_float), ("numInitialComponents", C.c_int32),
                ("minItr", C.c_int32), ("maxItr", C.c_int32), ("relLogLikelihoodThreshold", C.c_float), ("initKappa", C.c_float),
                ("maxKappa", C.c_float), ("vPrior", C.c_float), ("rPrior", C.c_float), ("rPriorWeight", C.c_float),
                ("splitRegions", C.c_int32), ("samplesForRegionSplit", C.c_float)]


class AppState(C.Structure):
    """b200pt_app — the per-frame host state of RayTracingApp (src/RayTracingApp.h:147-189)."""
    _fields_ = [("settings", PushConstants), ("accumulateResults", C.c_int32), ("hasInputChanged", C.c_int32),
                ("irradianceCachePrepareFrames", C.c_int32), ("currentPrepareFrames", C.c_int32), ("numGuidingOptimizations", C.c_int32),
                ("currentGuidingOptimizations", C.c_int32), ("loadBackupNextIteration", C.c_int32), ("activateADRRSAfterPrepareFrames", C.c_int32),
                ("evalCurrentSamples", C.c_int64), ("backupPushConstant", PushConstants)]


class Stats(C.Structure):
    _fields_ = [("extend_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("path_vertices", C.c_uint64), ("samples", C.c_uint64),
                ("iterations", C.c_uint64), ("kernel_launches", C.c_uint64), ("launches_extend", C.c_uint64),
                ("launches_shadow", C.c_uint64), ("launches_shade", C.c_uint64), ("ms_extend", C.c_float), ("ms_shadow", C.c_float),
                ("ms_shade", C.c_float), ("ms_total", C.c_float),
                ("guiding_samples", C.c_uint64), ("guiding_em_sample_iterations", C.c_uint64), ("guiding_regions_fit", C.c_uint64),
                ("launches_guiding", C.c_uint64), ("ms_guiding_sort", C.c_float), ("ms_guiding_fit", C.c_float),
                ("guiding_samples_all_ranks", C.c_uint64), ("guiding_bytes_received", C.c_uint64), ("ms_guiding_exchange", C.c_float),
                ("ms_guiding_gather", C.c_float)]


GUIDING_ORDER_STRICT, GUIDING_ORDER_REORDERED = 0, 1      # b200pt_guiding_set_order
DM_FUNCTIONS = ("sin", "cos", "tan", "asin", "acos", "atan", "atan2", "pow", "log", "exp", "div3")      # B200PT_DM_*


def detmath_host(fn, a, b=None):
    """include/b200pt_detmath.h as compiled for the host (inside libb200pt.so): fn in DM_FUNCTIONS over float32 arrays"""
    a = np.ascontiguousarray(a, dtype=np.float32)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.float32)
    out = np.empty_like(a)
    _check(lib().b200pt_detmath_eval_host(DM_FUNCTIONS.index(fn), a.ctypes.data, None if bb is None else bb.ctypes.data, out.ctypes.data, a.size))
    return out

RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("tmin", "<f4"), ("dir", "<f4", 3), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("prim", "<u4"), ("u", "<f4"), ("v", "<f4")])
DIRECTIONAL_DATA_DTYPE = np.dtype([("position", "<f4", 3), ("direction", "<f4", 3), ("weight", "<f4"), ("pdf", "<f4"),
                                   ("distance", "<f4"), ("flags", "<u4")])
VMF_THETA_DTYPE = np.dtype([("mu", "<f4", 3), ("k", "<f4"), ("norm", "<f4"), ("eMin2K", "<f4"), ("distance", "<f4"),
                            ("target", "<f4", 3)])
VMM_THETA_DTYPE = np.dtype([("thetas", VMF_THETA_DTYPE, 16), ("pi", "<f4", 16), ("meanPosition", "<f4", 3),
                            ("usedDistributions", "<i4")])
CACHE_DATA_DTYPE = np.dtype([("color", "<f4", 3), ("normal", "<f4", 3), ("rotGrad", "<f4", 3), ("transGrad", "<f4", 3),
                             ("harmonicR", "<f4"), ("numUpdates", "<u4")])
SPHERE_DTYPE = np.dtype([("center", "<f4", 3), ("radius", "<f4"), ("materialIndex", "<i4"), ("iLight", "<i4")])
AABB_DTYPE = np.dtype([("min", "<f4", 3), ("max", "<f4", 3)])
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 16 and DIRECTIONAL_DATA_DTYPE.itemsize == 40
assert VMM_THETA_DTYPE.itemsize == 720 and CACHE_DATA_DTYPE.itemsize == 56 and C.sizeof(PushConstants) == 192
assert C.sizeof(Vertex) == 48 and C.sizeof(Material) == 96 and C.sizeof(Instance) == 144 and C.sizeof(Light) == 40

# every symbol include/b200pt.h declares (checked by tests/test_abi.py against the header text)
EXPORTS = [
    "b200pt_last_error", "b200pt_device_count", "b200pt_create", "b200pt_destroy", "b200pt_set_scene", "b200pt_set_camera",
    "b200pt_render_frame", "b200pt_render_frames", "b200pt_read_image", "b200pt_write_image", "b200pt_read_image_device", "b200pt_write_image_device",
    "b200pt_trace_rays", "b200pt_trace_rays_device", "b200pt_stats_get", "b200pt_set_stage_timing", "b200pt_stats_reset", "b200pt_synchronize",
    "b200pt_timer_start", "b200pt_timer_stop",
    "b200pt_default_guiding_params", "b200pt_guiding_update", "b200pt_guiding_region_count", "b200pt_guiding_get_aabbs",
    "b200pt_guiding_get_vmms", "b200pt_guiding_put_vmms", "b200pt_guiding_get_samples", "b200pt_guiding_put_samples",
    "b200pt_guiding_sample_capacity", "b200pt_guiding_get_samples_device", "b200pt_guiding_reset", "b200pt_guiding_update_host", "b200pt_guiding_update_device",
    "b200pt_guiding_set_order", "b200pt_guiding_sorted_count", "b200pt_guiding_get_sorted", "b200pt_guiding_get_state", "b200pt_guiding_fastexp", "b200pt_guiding_selftest_division", "b200pt_detmath_eval", "b200pt_detmath_eval_host", "b200pt_ic_get", "b200pt_ic_put", "b200pt_default_push_constants", "b200pt_scene_load",
    "b200pt_scene_free", "b200pt_scene_get_desc", "b200pt_scene_get_camera", "b200pt_camera_matrices", "b200pt_mat4_inverse", "b200pt_write_exr",
    "b200pt_read_exr", "b200pt_read_image_file", "b200pt_free",
    "b200pt_set_aovs", "b200pt_read_aovs", "b200pt_save_state", "b200pt_load_state", "b200pt_comm_unique_id", "b200pt_comm_init", "b200pt_comm_destroy", "b200pt_reduce_image", "b200pt_allgather_samples", "b200pt_guiding_update_all_ranks", "b200pt_guiding_update_all_ranks_device", "b200pt_guiding_plan_debug", "b200pt_comm_exchange_mode",
    "b200pt_scene_bvh_check",
    "b200pt_app_init", "b200pt_app_scene_switched", "b200pt_app_begin_frame", "b200pt_app_end_frame", "b200pt_app_draw_frame"]

_lib = None


def lib():
    """Load libb200pt.so (built in-tree by `make` / __graft_entry__.build()). No fallback: raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200ptError("libb200pt.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        L.b200pt_last_error.restype = C.c_char_p
        L.b200pt_guiding_sample_capacity.restype = C.c_int64
        L.b200pt_guiding_sample_capacity.argtypes = [C.c_void_p]
        L.b200pt_camera_matrices.restype = None
        L.b200pt_default_push_constants.restype = None
        L.b200pt_default_guiding_params.restype = None
        L.b200pt_free.restype = None
        L.b200pt_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.b200pt_destroy.argtypes = [C.c_void_p]
        L.b200pt_set_scene.argtypes = [C.c_void_p, C.POINTER(SceneDesc)]
        L.b200pt_scene_bvh_check.argtypes = [C.POINTER(SceneDesc), C.POINTER(BvhReport)]
        L.b200pt_set_camera.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.b200pt_render_frame.argtypes = [C.c_void_p, C.POINTER(PushConstants)]
        L.b200pt_render_frames.argtypes = [C.c_void_p, C.POINTER(PushConstants), C.c_int]
        L.b200pt_read_image.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.b200pt_write_image.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.b200pt_read_image_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.b200pt_write_image_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.b200pt_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.b200pt_trace_rays_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.b200pt_stats_get.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.b200pt_stats_reset.argtypes = [C.c_void_p]
        L.b200pt_set_stage_timing.argtypes = [C.c_void_p, C.c_int]
        L.b200pt_synchronize.argtypes = [C.c_void_p]
        L.b200pt_timer_start.argtypes = [C.c_void_p]
        L.b200pt_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.b200pt_guiding_update.argtypes = [C.c_void_p, C.POINTER(GuidingParams)]
        L.b200pt_guiding_region_count.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.b200pt_guiding_get_aabbs.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.b200pt_guiding_get_vmms.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.b200pt_guiding_put_vmms.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.b200pt_guiding_get_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.b200pt_guiding_put_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.b200pt_guiding_get_samples_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.b200pt_guiding_reset.argtypes = [C.c_void_p, C.POINTER(GuidingParams)]
        L.b200pt_guiding_update_host.argtypes = [C.c_void_p, C.POINTER(GuidingParams), C.c_void_p, C.c_int64]
        L.b200pt_guiding_update_device.argtypes = [C.c_void_p, C.POINTER(GuidingParams), C.c_void_p, C.c_int64]
        L.b200pt_guiding_sorted_count.restype = C.c_int64
        L.b200pt_guiding_sorted_count.argtypes = [C.c_void_p]
        L.b200pt_guiding_get_sorted.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.b200pt_guiding_get_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.b200pt_guiding_fastexp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.b200pt_ic_get.argtypes = [C.c_void_p, C.POINTER(CacheHeader), C.c_void_p, C.c_void_p, C.c_int]
        L.b200pt_ic_put.argtypes = [C.c_void_p, C.POINTER(CacheHeader), C.c_void_p, C.c_void_p, C.c_int]
        L.b200pt_scene_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.b200pt_scene_free.argtypes = [C.c_void_p]
        L.b200pt_scene_get_desc.argtypes = [C.c_void_p, C.POINTER(SceneDesc)]
        L.b200pt_scene_get_camera.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.b200pt_camera_matrices.argtypes = [C.POINTER(C.c_float)] * 3 + [C.c_float, C.c_float] + [C.POINTER(C.c_float)] * 2
        L.b200pt_mat4_inverse.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.b200pt_write_exr.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        L.b200pt_read_exr.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.b200pt_free.argtypes = [C.c_void_p]
        L.b200pt_read_image_file.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.b200pt_set_aovs.argtypes = [C.c_void_p, C.c_int]
        L.b200pt_read_aovs.argtypes = [C.c_void_p, C.c_void_p]
        L.b200pt_save_state.argtypes = [C.c_void_p, C.c_char_p]
        L.b200pt_load_state.argtypes = [C.c_void_p, C.c_char_p]
        L.b200pt_comm_unique_id.argtypes = [C.c_char_p]
        L.b200pt_comm_init.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        L.b200pt_comm_destroy.argtypes = [C.c_void_p]
        L.b200pt_reduce_image.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.b200pt_allgather_samples.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.b200pt_guiding_update_all_ranks.argtypes = [C.c_void_p, C.POINTER(GuidingParams)]
        L.b200pt_guiding_update_all_ranks_device.argtypes = [C.c_void_p, C.POINTER(GuidingParams), C.c_void_p, C.c_int64]
        L.b200pt_comm_exchange_mode.argtypes = [C.c_void_p]
        L.b200pt_guiding_plan_debug.argtypes = [C.c_void_p] * 2 + [C.c_int] * 3 + [C.c_void_p] * 7
        L.b200pt_guiding_set_order.argtypes = [C.c_void_p, C.c_int]
        L.b200pt_detmath_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.b200pt_detmath_eval_host.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.b200pt_guiding_selftest_division.argtypes = [C.c_void_p, C.c_float, C.c_float, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.b200pt_app_init.restype = None
        L.b200pt_app_init.argtypes = [C.POINTER(AppState)]
        L.b200pt_app_scene_switched.restype = None
        L.b200pt_app_scene_switched.argtypes = [C.POINTER(AppState)]
        L.b200pt_app_begin_frame.restype = None
        L.b200pt_app_begin_frame.argtypes = [C.POINTER(AppState), C.c_uint32, C.POINTER(PushConstants)]
        L.b200pt_app_end_frame.argtypes = [C.POINTER(AppState)]
        L.b200pt_app_draw_frame.argtypes = [C.POINTER(AppState), C.c_void_p, C.c_uint32, C.POINTER(GuidingParams)]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise B200ptError("%s (code %d)" % (lib().b200pt_last_error().decode(), rc))


def default_push_constants(**overrides):
    pc = PushConstants()
    lib().b200pt_default_push_constants(C.byref(pc))
    for k, v in overrides.items():
        if not hasattr(pc, k):
            raise AttributeError("no push constant named %r" % k)
        setattr(pc, k, v)
    return pc


COMM_ID_BYTES = 128


def comm_unique_id():
    """ncclGetUniqueId as 128 bytes: create on one rank, send to the others (any side channel), pass to Renderer.comm_init."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(lib().b200pt_comm_unique_id(buf))
    return buf.raw


def default_guiding_params(**overrides):
    gp = GuidingParams()
    lib().b200pt_default_guiding_params(C.byref(gp))
    for k, v in overrides.items():
        setattr(gp, k, v)
    return gp


def tea(val0, val1):
    """random.glsl:13-27 — used for the frame-seed stream tea(frame, seed)."""
    v0, v1, s0 = val0 & 0xFFFFFFFF, val1 & 0xFFFFFFFF, 0
    for _ in range(16):
        s0 = (s0 + 0x9E3779B9) & 0xFFFFFFFF
        v0 = (v0 + ((((v1 << 4) & 0xFFFFFFFF) + 0xA341316C) ^ (v1 + s0) ^ ((v1 >> 5) + 0xC8013EA4))) & 0xFFFFFFFF
        v1 = (v1 + ((((v0 << 4) & 0xFFFFFFFF) + 0xAD90777D) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7E95761E))) & 0xFFFFFFFF
    return v0


class Scene:
    """SceneLoader(filepath): parsed scene buffers living in host memory owned by the library."""

    def __init__(self, path):
        self.path = path
        self._h = C.c_void_p()
        _check(lib().b200pt_scene_load(os.fsencode(path), C.byref(self._h)))
        self.desc = SceneDesc()
        _check(lib().b200pt_scene_get_desc(self._h, C.byref(self.desc)))

    def camera(self):
        o, t, u = (C.c_float * 3)(), (C.c_float * 3)(), (C.c_float * 3)()
        f = C.c_float()
        _check(lib().b200pt_scene_get_camera(self._h, o, t, u, C.byref(f)))
        return list(o), list(t), list(u), f.value

    def camera_matrices(self, aspect):
        o, t, u, f = self.camera()
        return camera_matrices(o, t, u, f, aspect)

    def bvh_check(self):
        """Host-only structural check of the BVH8 that set_scene would build for this scene (b200pt_scene_bvh_check)."""
        rep = BvhReport()
        _check(lib().b200pt_scene_bvh_check(C.byref(self.desc), C.byref(rep)))
        return rep

    @property
    def num_triangles(self):
        d = self.desc
        return sum(d.num_indices[d.instances[i].modelIndex] // 3 for i in range(d.num_instances))

    def close(self):
        if self._h:
            lib().b200pt_scene_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def camera_matrices(origin, target, up, vfov, aspect):
    view, proj = (C.c_float * 16)(), (C.c_float * 16)()
    lib().b200pt_camera_matrices((C.c_float * 3)(*origin), (C.c_float * 3)(*target), (C.c_float * 3)(*up), vfov, aspect, view, proj)
    return np.array(view, dtype=np.float32), np.array(proj, dtype=np.float32)


def mat4_inverse(m):
    a = np.ascontiguousarray(m, dtype=np.float32)
    out = np.empty(16, dtype=np.float32)
    if not lib().b200pt_mat4_inverse(a.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float))):
        raise B200ptError("singular matrix")
    return out


class Renderer:
    """RayTracingApp(width, height, icSize, guidingSplits): one context on one GPU."""

    def __init__(self, width, height, ic_size=0, guiding_splits=0, device=0):
        self.width, self.height = width, height
        self._h = C.c_void_p()
        _check(lib().b200pt_create(device, width, height, ic_size, guiding_splits, C.byref(self._h)))
        self.ic_size = ic_size

    def set_scene(self, scene):
        desc = scene.desc if isinstance(scene, Scene) else scene
        _check(lib().b200pt_set_scene(self._h, C.byref(desc)))

    def set_camera(self, view, proj):
        v = np.ascontiguousarray(view, dtype=np.float32)
        p = np.ascontiguousarray(proj, dtype=np.float32)
        _check(lib().b200pt_set_camera(self._h, v.ctypes.data_as(C.POINTER(C.c_float)), p.ctypes.data_as(C.POINTER(C.c_float))))

    def render_frame(self, pc):
        _check(lib().b200pt_render_frame(self._h, C.byref(pc)))

    def render_frames(self, pcs):
        """`len(pcs)` consecutive frames in one call (b200pt_render_frames): same images as render_frame in a loop."""
        arr = (PushConstants * len(pcs))(*pcs)
        _check(lib().b200pt_render_frames(self._h, arr, len(pcs)))

    def read_image(self, which=IMAGE_OUTPUT, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.float32)
        _check(lib().b200pt_read_image(self._h, which, out.ctypes.data))
        return out

    def write_image(self, which, rgba):
        a = np.ascontiguousarray(rgba, dtype=np.float32)
        assert a.size == self.width * self.height * 4
        _check(lib().b200pt_write_image(self._h, which, a.ctypes.data))

    def read_image_device(self, which, device_ptr):
        _check(lib().b200pt_read_image_device(self._h, which, C.c_void_p(device_ptr)))

    def write_image_device(self, which, device_ptr):
        _check(lib().b200pt_write_image_device(self._h, which, C.c_void_p(device_ptr)))

    def trace_rays(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(rays.shape[0], dtype=HIT_DTYPE)
        _check(lib().b200pt_trace_rays(self._h, rays.ctypes.data, rays.shape[0], hits.ctypes.data, int(any_hit)))
        return hits

    def trace_rays_device(self, rays_ptr, n, hits_ptr, any_hit=False):
        _check(lib().b200pt_trace_rays_device(self._h, C.c_void_p(rays_ptr), n, C.c_void_p(hits_ptr), int(any_hit)))

    def synchronize(self):
        _check(lib().b200pt_synchronize(self._h))

    def timer_start(self):
        _check(lib().b200pt_timer_start(self._h))

    def timer_stop(self):
        """Device milliseconds since timer_start, measured with CUDA events on the library's stream."""
        ms = C.c_float()
        _check(lib().b200pt_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def stats(self):
        s = Stats()
        _check(lib().b200pt_stats_get(self._h, C.byref(s)))
        return s

    def set_stage_timing(self, enabled):
        _check(lib().b200pt_set_stage_timing(self._h, int(enabled)))

    def stats_reset(self):
        _check(lib().b200pt_stats_reset(self._h))

    # guiding
    def guiding_region_count(self):
        n = C.c_int()
        _check(lib().b200pt_guiding_region_count(self._h, C.byref(n)))
        return n.value

    def guiding_aabbs(self):
        out = np.empty(self.guiding_region_count(), dtype=AABB_DTYPE)
        _check(lib().b200pt_guiding_get_aabbs(self._h, out.ctypes.data, out.shape[0]))
        return out

    def guiding_get_vmms(self):
        out = np.empty(self.guiding_region_count(), dtype=VMM_THETA_DTYPE)
        _check(lib().b200pt_guiding_get_vmms(self._h, out.ctypes.data, out.shape[0]))
        return out

    def guiding_put_vmms(self, vmms):
        a = np.ascontiguousarray(vmms, dtype=VMM_THETA_DTYPE)
        _check(lib().b200pt_guiding_put_vmms(self._h, a.ctypes.data, a.shape[0]))

    def guiding_sample_capacity(self):
        return lib().b200pt_guiding_sample_capacity(self._h)

    def guiding_get_samples(self, n=None):
        n = self.guiding_sample_capacity() if n is None else n
        out = np.empty(n, dtype=DIRECTIONAL_DATA_DTYPE)
        _check(lib().b200pt_guiding_get_samples(self._h, out.ctypes.data, n))
        return out

    def guiding_put_samples(self, samples):
        a = np.ascontiguousarray(samples, dtype=DIRECTIONAL_DATA_DTYPE)
        _check(lib().b200pt_guiding_put_samples(self._h, a.ctypes.data, a.shape[0]))

    def guiding_update(self, params=None):
        params = params or default_guiding_params()
        _check(lib().b200pt_guiding_update(self._h, C.byref(params)))

    def guiding_get_samples_device(self, device_ptr, n=None):
        n = self.guiding_sample_capacity() if n is None else n
        _check(lib().b200pt_guiding_get_samples_device(self._h, C.c_void_p(device_ptr), n))

    def guiding_reset(self, params=None):
        params = params or default_guiding_params()
        _check(lib().b200pt_guiding_reset(self._h, C.byref(params)))

    def guiding_update_host(self, samples, params=None):
        """PathGuiding::update on caller-provided host records (copied to the device inside the call)."""
        params = params or default_guiding_params()
        a = np.ascontiguousarray(samples, dtype=DIRECTIONAL_DATA_DTYPE)
        _check(lib().b200pt_guiding_update_host(self._h, C.byref(params), a.ctypes.data, a.shape[0]))

    def guiding_update_device(self, device_ptr, n, params=None):
        params = params or default_guiding_params()
        _check(lib().b200pt_guiding_update_device(self._h, C.byref(params), C.c_void_p(device_ptr), n))

    def guiding_sorted(self):
        n = lib().b200pt_guiding_sorted_count(self._h)
        out = np.empty(n, dtype=DIRECTIONAL_DATA_DTYPE)
        off = np.empty(self.guiding_region_count() + 1, dtype=np.uint32)
        _check(lib().b200pt_guiding_get_sorted(self._h, out.ctypes.data, off.ctypes.data))
        return out, off

    GUIDING_STATE_FIELDS = ("weight", "kappa", "r", "mux", "muy", "muz", "distance", "distSumW", "chi", "chiN", "covxx", "covyy", "covxy", "covSumW")

    def detmath(self, fn, a, b=None):
        """the same functions evaluated by a kernel on this context's device"""
        a = np.ascontiguousarray(a, dtype=np.float32)
        bb = None if b is None else np.ascontiguousarray(b, dtype=np.float32)
        out = np.empty_like(a)
        _check(lib().b200pt_detmath_eval(self._h, DM_FUNCTIONS.index(fn), a.ctypes.data, None if bb is None else bb.ctypes.data, out.ctypes.data, a.size))
        return out

    def guiding_selftest_division(self, lo, hi):
        """(mismatches, tested) of the device's fastexp division vs IEEE division over every float in [lo, hi]"""
        m, t = C.c_uint64(), C.c_uint64()
        _check(lib().b200pt_guiding_selftest_division(self._h, lo, hi, C.byref(m), C.byref(t)))
        return m.value, t.value

    def guiding_set_order(self, order):
        """0 = strict (the reference's sequential float sums, default), 1 = reordered (block-parallel sums)"""
        _check(lib().b200pt_guiding_set_order(self._h, int(order)))

    def guiding_state(self, region):
        sc = np.empty(5, dtype=np.float32)
        pc = np.empty((14, 16), dtype=np.float32)
        _check(lib().b200pt_guiding_get_state(self._h, region, sc.ctypes.data, pc.ctypes.data))
        d = {"K": int(sc[0]), "sampleWeight": float(sc[1]), "numSamples": float(sc[2]), "totalNumSamples": int(sc[3]), "numEMIterations": int(sc[4])}
        for i, f in enumerate(self.GUIDING_STATE_FIELDS):
            d[f] = pc[i].copy()
        return d

    def guiding_fastexp(self, x):
        a = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(a)
        _check(lib().b200pt_guiding_fastexp(self._h, a.ctypes.data, out.ctypes.data, a.size))
        return out

    def set_aovs(self, enabled=True):
        _check(lib().b200pt_set_aovs(self._h, int(enabled)))

    def read_aovs(self):
        """(H, W, 4): maxReachedDepth, depthSum, depthsCounter, nextSplitSlot of the last frame (rgen:1653-1655, :1722-1741)."""
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        _check(lib().b200pt_read_aovs(self._h, out.ctypes.data))
        return out

    def save_state(self, path):
        _check(lib().b200pt_save_state(self._h, os.fsencode(path)))

    def load_state(self, path):
        _check(lib().b200pt_load_state(self._h, os.fsencode(path)))

    # multi-GPU (native NCCL path of the C ABI; sharding.py is the torch.distributed variant)
    def comm_init(self, unique_id, rank, nranks):
        _check(lib().b200pt_comm_init(self._h, unique_id, rank, nranks))

    def comm_destroy(self):
        _check(lib().b200pt_comm_destroy(self._h))

    def reduce_image(self, which=IMAGE_OUTPUT, frames_local=1):
        _check(lib().b200pt_reduce_image(self._h, which, frames_local))

    def allgather_samples(self):
        n = C.c_int64()
        _check(lib().b200pt_allgather_samples(self._h, C.byref(n)))
        return n.value

    def guiding_update_all_ranks(self, params=None):
        p = params if params is not None else default_guiding_params()
        _check(lib().b200pt_guiding_update_all_ranks(self._h, C.byref(p)))

    def guiding_update_all_ranks_device(self, device_ptr, n, params=None):
        p = params if params is not None else default_guiding_params()
        _check(lib().b200pt_guiding_update_all_ranks_device(self._h, C.byref(p), C.c_void_p(device_ptr), n))

    def guiding_plan_debug(self, counts, rank, peer_mode=1):
        """k_plan as rank `rank` on counts[nranks][regions] (b200pt_guiding_plan_debug); returns a dict of numpy arrays"""
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        n, R = counts.shape
        assert R == self.guiding_region_count()
        owner = np.empty(R, np.uint8); begin = np.empty(R, np.uint32); length = np.empty(R, np.uint32)
        src = np.empty((n, R), np.uint32); active = np.empty(R, np.uint32); summary = np.zeros(6, np.uint32); seg = np.zeros((R * n, 4), np.uint32)
        _check(lib().b200pt_guiding_plan_debug(self._h, counts.ctypes.data, n, rank, peer_mode, owner.ctypes.data, begin.ctypes.data, length.ctypes.data, src.ctypes.data,
                                               active.ctypes.data, summary.ctypes.data, seg.ctypes.data))
        return dict(owner=owner, region_begin=begin, region_len=length, src_start=src, active=active[:summary[0]], num_owned=int(summary[1]),
                    segments=seg[:summary[2]], local_valid=int(summary[3]), owned_samples=int(summary[4]), total_samples=int(summary[5]))

    def comm_exchange_mode(self):
        """0 no communicator, 1 ncclSend/ncclRecv, 2 CUDA-IPC peer reads (decided by the first all-ranks update)"""
        return lib().b200pt_comm_exchange_mode(self._h)

    # irradiance cache
    def ic_get(self):
        hdr = CacheHeader()
        data = np.empty(self.ic_size, dtype=CACHE_DATA_DTYPE)
        spheres = np.empty(self.ic_size, dtype=SPHERE_DTYPE)
        _check(lib().b200pt_ic_get(self._h, C.byref(hdr), data.ctypes.data, spheres.ctypes.data, self.ic_size))
        return hdr, data, spheres

    def ic_put(self, hdr, data, spheres):
        d = np.ascontiguousarray(data, dtype=CACHE_DATA_DTYPE)
        s = np.ascontiguousarray(spheres, dtype=SPHERE_DTYPE)
        _check(lib().b200pt_ic_put(self._h, C.byref(hdr), d.ctypes.data, s.ctypes.data, d.shape[0]))

    def close(self):
        if self._h:
            lib().b200pt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class App:
    """Frame driver: RayTracingApp::raytrace / drawCallback without the window (include/b200pt.h, b200pt_app_*).
    `settings` is the user's RtPushConstant (edit it, then call input_changed()); draw_frame renders one frame."""

    def __init__(self, renderer=None, accumulate=True, **settings):
        self.state = AppState()
        lib().b200pt_app_init(C.byref(self.state))
        self.state.accumulateResults = int(accumulate)
        self.renderer = renderer
        for k, v in settings.items():
            if not hasattr(self.state.settings, k):
                raise AttributeError("no push constant named %r" % k)
            setattr(self.state.settings, k, v)

    @property
    def settings(self):
        return self.state.settings

    def input_changed(self):
        self.state.hasInputChanged = 1

    def scene_switched(self):
        lib().b200pt_app_scene_switched(C.byref(self.state))

    def begin_frame(self, frame_seed):
        pc = PushConstants()
        lib().b200pt_app_begin_frame(C.byref(self.state), C.c_uint32(frame_seed), C.byref(pc))
        return pc

    def end_frame(self):
        return bool(lib().b200pt_app_end_frame(C.byref(self.state)))

    def draw_frame(self, frame_seed, guiding_params=None):
        gp = C.byref(guiding_params) if guiding_params is not None else None
        _check(lib().b200pt_app_draw_frame(C.byref(self.state), self.renderer._h, C.c_uint32(frame_seed), gp))


def read_image_file(path):
    """Baseline JPEG / PNG -> (H, W, 4) uint8, as the scene loader decodes bitmap textures."""
    p = C.POINTER(C.c_uint8)()
    w, h = C.c_int(), C.c_int()
    _check(lib().b200pt_read_image_file(os.fsencode(path), C.byref(p), C.byref(w), C.byref(h)))
    try:
        return np.ctypeslib.as_array(p, shape=(h.value, w.value, 4)).copy()
    finally:
        lib().b200pt_free(p)


def write_exr(path, rgba):
    a = np.ascontiguousarray(rgba, dtype=np.float32)
    h, w = a.shape[0], a.shape[1]
    _check(lib().b200pt_write_exr(os.fsencode(path), a.ctypes.data, w, h))


def read_exr(path):
    p = C.POINTER(C.c_float)()
    w, h = C.c_int(), C.c_int()
    _check(lib().b200pt_read_exr(os.fsencode(path), C.byref(p), C.byref(w), C.byref(h)))
    try:
        return np.ctypeslib.as_array(p, shape=(h.value, w.value, 4)).copy()
    finally:
        lib().b200pt_free(p)
