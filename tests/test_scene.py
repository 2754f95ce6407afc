"""CPU: the headless scene front-end against facts about the reference's scenes recorded in SURVEY.md."""
import os

import numpy as np
import pytest

import helpers


def test_cornell_dielectric_counts_and_aabb():
    P = helpers.pt()
    s = P.Scene(helpers.scene_path("cornell-dielectric"))
    d = s.desc
    assert d.num_models == 8 and d.num_instances == 8 and d.num_spheres == 0
    # 3950 triangles without the glass shell (5 walls x 2 + ball 3936 + emitter 4) + stand-in shell 3936
    assert s.num_triangles == 3950 + 3936
    # materials: 4 top-level bsdfs + the emitter copy (src/SceneLoader.cpp:437-453)
    assert d.num_materials == 5
    assert [d.materials[i].type for i in range(5)] == [0, 0, 0, 2, 3]
    assert d.materials[3].refractionIndex == pytest.approx(1.5) and d.materials[3].refractionIndexInv == pytest.approx(1 / 1.5)
    assert list(d.materials[4].lightColor) == [15, 15, 15] and list(d.materials[4].diffuse) == pytest.approx([0.3, 0.3, 0.3])
    # one area light on the emitter instance, uniform selection probability
    assert d.num_lights == 1 and d.lights[0].type == 0 and d.lights[0].instanceIndex == 7 and d.lights[0].sampleProb == 1.0
    assert d.instances[7].iLight == 0 and all(d.instances[i].iLight == -1 for i in range(7))
    # emitter: two quads of 1.51779^2 each
    assert d.lights[0].area == pytest.approx(2 * 1.51779 ** 2, rel=1e-4)
    # scene AABB as SURVEY §8(d) config 1 states it (the stand-in shell does not enlarge x/z; ymin of the ball mesh)
    assert list(d.scene_min)[0] == pytest.approx(-1.686433) and list(d.scene_max) == pytest.approx([1.686433, 3.386433, 1.686433])
    assert list(d.scene_min)[2] == pytest.approx(-1.686433)
    o, t, u, f = s.camera()
    assert o == [0, 2, 5] and t == [0, 1.5, 0] and u == [0, 1, 0] and f == 45
    assert d.num_textures == 1 and d.textures[0].width == 1   # 1x1 default env texture


def test_light_tables_follow_weighted_sampler():
    P = helpers.pt()
    s = P.Scene(helpers.scene_path("cornell-dielectric"))
    d = s.desc
    rl = np.ctypeslib.as_array(d.random_light_index, shape=(P.SIZE_LIGHT_RANDOM,))
    assert np.all(rl == 0)
    assert d.num_face_tables == 1
    ft = np.ctypeslib.as_array(C_cast(d.random_tri_index, P), shape=(P.SIZE_TRI_RANDOM, 3))
    idx = ft[:, 0].view(np.int32)
    # 4 emissive triangles (two quads), equal areas -> roughly uniform; the stream is std::mt19937(5489)
    assert set(np.unique(idx)) == {0, 1, 2, 3}
    counts = np.bincount(idx, minlength=4)
    assert counts.min() > 2300 and counts.max() < 2700
    assert np.allclose(ft[:, 1], 0.25, atol=1e-6)
    # first draws of libstdc++'s mt19937 + uniform_real_distribution<float>: u0 = 0.8147237 -> last quarter
    assert idx[0] == 3


def C_cast(ptr, P):
    import ctypes as C
    return C.cast(ptr, C.POINTER(C.c_float))


def test_veach_mis_scene():
    P = helpers.pt()
    s = P.Scene(helpers.scene_path("veachMIS"))
    d = s.desc
    assert d.num_spheres == 4 and d.num_instances == 5 and s.num_triangles == 12
    assert d.num_lights == 4 and all(d.lights[i].type == 2 for i in range(4))
    assert [round(d.spheres[i].radius, 5) for i in range(4)] == [0.1, 0.03333, 0.3, 0.9]
    assert d.lights[3].area == pytest.approx(4 * np.pi * 0.81, rel=1e-6)
    rough = [d.materials[i].roughness for i in range(d.num_materials) if d.materials[i].type == 6]
    assert rough == pytest.approx([0.02, 0.06, 0.1, 0.2])
    # OBJ files without normals get per-face normals (src/SceneLoader.cpp:299-308)
    v = d.vertices[4][0]
    assert np.linalg.norm(list(v.normal)) == pytest.approx(1.0, rel=1e-6)
    rl = np.ctypeslib.as_array(d.random_light_index, shape=(P.SIZE_LIGHT_RANDOM,))
    assert set(np.unique(rl)) == {0, 1, 2, 3}


def test_missing_obj_is_an_error(tmp_path):
    P = helpers.pt()
    xml = tmp_path / "s.xml"
    xml.write_text('<scene version="0.6.0"><bsdf type="diffuse" id="a"><rgb name="reflectance" value="0.5"/></bsdf>'
                   '<shape type="obj"><string name="filename" value="nope.obj"/><ref id="a"/></shape></scene>')
    with pytest.raises(P.B200ptError, match="nope.obj"):   # SceneLoader::readObjFile throws, src/SceneLoader.cpp:128-137
        P.Scene(str(xml))


def test_camera_matrices():
    P = helpers.pt()
    view, proj = P.camera_matrices([0, 2, 5], [0, 1.5, 0], [0, 1, 0], 45.0, 16 / 9)
    V = view.reshape(4, 4).T
    # rotation part orthonormal, camera origin maps to 0, view direction maps to -z
    assert np.allclose(V[:3, :3] @ V[:3, :3].T, np.eye(3), atol=1e-6)
    assert np.allclose(V @ np.array([0, 2, 5, 1]), [0, 0, 0, 1], atol=1e-5)
    fwd = np.array([0, -0.5, -5]); fwd /= np.linalg.norm(fwd)
    assert np.allclose(V[:3, :3] @ fwd, [0, 0, -1], atol=1e-6)
    Pm = proj.reshape(4, 4).T
    assert Pm[1, 1] == pytest.approx(-1 / np.tan(np.radians(45) / 2), rel=1e-6)   # proj[1][1] *= -1
    assert Pm[0, 0] == pytest.approx(1 / (16 / 9 * np.tan(np.radians(45) / 2)), rel=1e-6)
    inv = P.mat4_inverse(view).reshape(4, 4).T
    assert np.allclose(inv @ V, np.eye(4), atol=1e-5)


VIOLATIONS = ("missing_prims", "duplicate_prims", "outside_box", "bad_meta", "depth_mismatch", "unreachable_nodes")


@pytest.mark.parametrize("name", ["cornell-dielectric", "veachMIS", "sponzaXML", "test-scene", "irradianceCache", "alphaLeaf", "stackedCards", "envMap", "testSpheres"])
def test_bvh8_structure_is_valid_for_every_scene(name):
    """host/bvh.cpp on the CPU (b200pt_scene_bvh_check, no GPU): every primitive sits in exactly one leaf, every triangle lies
    inside the dequantised 8-bit box of its slot and of all its ancestors' slots (what the slab test sees), child and
    triangle indexing is consistent, and the recorded depth — which set_scene checks against the traversal stack — is exact."""
    P = helpers.pt()
    scene = P.Scene(helpers.scene_path(name))
    rep = scene.bvh_check()
    assert all(getattr(rep, f) == 0 for f in VIOLATIONS), {f: getattr(rep, f) for f in VIOLATIONS}
    assert rep.num_tris == scene.num_triangles
    if rep.num_tris > 24:
        assert rep.inner_children == rep.num_nodes - 1 and rep.max_depth >= 2
        assert (rep.inner_children + rep.leaf_children) / rep.num_nodes > 5.0        # the DP collapse fills the 8-wide nodes (greedy: ~4)
    assert rep.max_depth <= 24                                                         # PT_STACK_LOCAL


def test_bvh8_check_notices_damage(monkeypatch):
    P = helpers.pt()
    scene = P.Scene(helpers.scene_path("cornell-dielectric"))
    for kind, field in ((1, "outside_box"), (2, "duplicate_prims"), (3, "bad_meta"), (4, "depth_mismatch")):
        monkeypatch.setenv("B200PT_BVH_CHECK_MUTATE", str(kind))
        rep = scene.bvh_check()
        assert getattr(rep, field) > 0, (kind, field)
        if kind == 2:
            assert rep.missing_prims == 1
    monkeypatch.delenv("B200PT_BVH_CHECK_MUTATE")
    assert all(getattr(scene.bvh_check(), f) == 0 for f in VIOLATIONS)


def test_bvh8_of_degenerate_and_coincident_triangles():
    """zero-area triangles, all centroids equal, and one huge + many tiny triangles: the builder must still place every
    primitive and keep it inside its boxes (the quantisation grid is 255 steps of the node's extent)."""
    import ctypes as C
    P = helpers.pt()
    rng = np.random.default_rng(0)

    def check(tri_pos):
        n = len(tri_pos)
        verts = (P.Vertex * (3 * n))()
        for i, t in enumerate(tri_pos):
            for k in range(3):
                verts[3 * i + k].pos[0], verts[3 * i + k].pos[1], verts[3 * i + k].pos[2] = [float(x) for x in t[k]]
        idx = (C.c_uint32 * (3 * n))(*range(3 * n))
        d = P.SceneDesc()
        vp = (C.POINTER(P.Vertex) * 1)(C.cast(verts, C.POINTER(P.Vertex)))
        ip = (C.POINTER(C.c_uint32) * 1)(C.cast(idx, C.POINTER(C.c_uint32)))
        nv, ni = (C.c_int32 * 1)(3 * n), (C.c_int32 * 1)(3 * n)
        inst = (P.Instance * 1)()
        for k in range(4):
            inst[0].transform[5 * k] = 1.0
            inst[0].normalTransform[5 * k] = 1.0
        d.num_models, d.vertices, d.num_vertices, d.indices, d.num_indices = 1, vp, nv, ip, ni
        d.num_instances, d.instances = 1, inst
        rep = P.BvhReport()
        assert P.lib().b200pt_scene_bvh_check(C.byref(d), C.byref(rep)) == 0
        assert all(getattr(rep, f) == 0 for f in VIOLATIONS), {f: getattr(rep, f) for f in VIOLATIONS}
        assert rep.num_tris == n
        return rep

    check(np.zeros((40, 3, 3)))                                                       # 40 points at the origin
    same = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], float), (300, 1, 1))   # 300 coincident triangles (the tie test's stack)
    check(same)
    tiny = rng.random((500, 1, 3)) * 1e-3 + rng.random((500, 3, 3)) * 1e-6
    huge = np.array([[[-1e4, -1e4, 0], [1e4, -1e4, 0], [0, 1e4, 0]]])
    rep = check(np.concatenate([huge, tiny]))
    assert rep.max_depth <= 24
    check(rng.normal(size=(5000, 3, 3)) * np.array([1e3, 1.0, 1e-3]))                 # very anisotropic soup


HOST_REF = os.path.join(helpers.ROOT, "oracle", "_ref", "libhost_ref.so")
LIGHT_TABLE_DIGESTS = os.path.join(helpers.ROOT, "tests", "golden", "light_table_digests.json")


def _light_tables(P, name):
    """(random_light_index, lights, [face table per area light]) of a scene as our loader built them"""
    import ctypes as C
    s = P.Scene(helpers.scene_path(name))
    d = s.desc
    rl = np.ctypeslib.as_array(d.random_light_index, shape=(P.SIZE_LIGHT_RANDOM,)).copy()
    lights = [d.lights[i] for i in range(d.num_lights)]
    raw = np.ctypeslib.as_array(C.cast(d.random_tri_index, C.POINTER(C.c_uint32)), shape=(max(1, d.num_face_tables) * P.SIZE_TRI_RANDOM, 3)).copy()
    tables = [raw[k * P.SIZE_TRI_RANDOM:(k + 1) * P.SIZE_TRI_RANDOM] for k in range(d.num_face_tables)]
    keep = dict(rl=rl, tables=tables, probs=np.array([l.sampleProb for l in lights], np.float32), areas=np.array([l.area for l in lights], np.float32),
                types=[l.type for l in lights])
    return s, keep


@pytest.mark.parametrize("name", ["cornell-dielectric", "test-scene", "veachMIS", "sponzaXML"])
def test_light_and_face_tables_equal_the_reference_sampler(name):
    """Binding 6 (LightSamplerBuffer): the 10 000-entry light table and the 10 000-entry face table of every area light against
    the reference's OWN WeightedSampler (src/WeightedSampler.cpp compiled into oracle/_ref/libhost_ref.so) fed with the
    weights SceneLoader::getLightSamplingVector / getFaceSamplingVector give it (src/SceneLoader.cpp:861-944): every light
    weighs 1; an area light's faces weigh their areas, in face order.  Indices, probabilities, per-sample face areas and the
    light's total area must be bit-equal.  Where the reference tree is absent the committed digests of the same tables are checked."""
    import ctypes as C
    import hashlib
    import json
    P = helpers.pt()
    scene, t = _light_tables(P, name)
    digest = hashlib.sha256(t["rl"].tobytes() + b"".join(x.tobytes() for x in t["tables"]) + t["probs"].tobytes() + t["areas"].tobytes()).hexdigest()
    assert json.load(open(LIGHT_TABLE_DIGESTS))[name] == digest
    if not os.path.exists(HOST_REF):
        return
    R = C.CDLL(HOST_REF)
    R.host_ref_weighted_samples.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]

    def reference(values, count):
        v = np.ascontiguousarray(values, np.float32)
        samples, probs, total = np.zeros(count, np.int32), np.zeros(len(v), np.float32), np.zeros(1, np.float32)
        R.host_ref_weighted_samples(v.ctypes.data, len(v), count, samples.ctypes.data, probs.ctypes.data, total.ctypes.data)
        return samples, probs, total[0]

    n = len(t["types"])
    samples, probs, _ = reference(np.ones(n, np.float32), P.SIZE_LIGHT_RANDOM)
    assert np.array_equal(t["rl"], samples) and np.array_equal(t["probs"], probs)
    area_lights = [i for i in range(n) if t["types"][i] == 0]             # B200PT_LIGHT_AREA
    if not area_lights:                                                    # "Dummy data" (src/SceneLoader.cpp:909-918): one zeroed table
        assert len(t["tables"]) == 1 and not t["tables"][0].any()
        return
    assert len(area_lights) == len(t["tables"])
    for k, i in enumerate(area_lights):
        tab = t["tables"][k]
        idx, prob, area = tab[:, 0].view(np.int32), tab[:, 1].view(np.float32), tab[:, 2].view(np.float32)
        faces = np.unique(idx)                                             # the emissive faces of the light's model, in face order
        weights = np.array([area[idx == f][0] for f in faces], np.float32)
        assert abs(float(np.unique(prob).astype(np.float64).sum()) - 1.0) < 1e-4 or len(np.unique(prob)) < len(faces)    # every face was drawn at least once
        samples, probs, total = reference(weights, P.SIZE_TRI_RANDOM)
        assert np.array_equal(idx, faces[samples]), (name, i)
        assert np.array_equal(prob, probs[samples]) and np.array_equal(area, weights[samples])
        assert t["areas"][i] == total


def _reference_obj(path, mtl_dir, override):
    import ctypes as C
    R = C.CDLL(HOST_REF)
    R.host_ref_obj_load.restype = C.c_void_p
    R.host_ref_obj_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    R.host_ref_obj_error.restype = C.c_char_p
    R.host_ref_obj_error.argtypes = [C.c_void_p]
    R.host_ref_obj_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    R.host_ref_obj_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    R.host_ref_obj_free.argtypes = [C.c_void_p]
    h = R.host_ref_obj_load(path.encode(), mtl_dir.encode(), override)
    assert not R.host_ref_obj_error(h), R.host_ref_obj_error(h)
    nv, ni, nm = C.c_int(), C.c_int(), C.c_int()
    R.host_ref_obj_counts(h, C.byref(nv), C.byref(ni), C.byref(nm))
    verts, mats, idx = np.zeros((nv.value, 8), np.float32), np.zeros(nv.value, np.int32), np.zeros(ni.value, np.uint32)
    R.host_ref_obj_copy(h, verts.ctypes.data, mats.ctypes.data, idx.ctypes.data)
    R.host_ref_obj_free(h)
    return verts, mats, idx


def _model_arrays(P, d, m):
    import ctypes as C
    nv, ni = d.num_vertices[m], d.num_indices[m]
    raw = np.ctypeslib.as_array(C.cast(d.vertices[m], C.POINTER(C.c_float)), shape=(nv, 12)).copy()       # b200pt_vertex: 48 B
    verts = np.concatenate([raw[:, 0:3], raw[:, 4:7], raw[:, 8:10]], 1)
    mats = raw[:, 10].view(np.int32).copy()
    idx = np.ctypeslib.as_array(d.indices[m], shape=(ni,)).copy()
    return verts, mats, idx


def _obj_files_of(scene_file):
    """the OBJ file of every model, in model order: <shape type="obj"> elements of a Mitsuba XML file, the "models" list of a JSON scene"""
    base = os.path.dirname(scene_file)
    if scene_file.endswith(".json"):
        import json
        return [(os.path.join(base, "models", list(m.values())[0]), os.path.join(base, "materials") + "/", -1) for m in json.load(open(scene_file))["models"]]
    import xml.etree.ElementTree as ET
    out = []
    for shape in ET.parse(scene_file).getroot().iter("shape"):
        if shape.get("type") == "obj":
            name = [s.get("value") for s in shape.findall("string") if s.get("name") == "filename"][0]
            out.append((os.path.join(base, name), base + "/", 0))
    return out


@pytest.mark.parametrize("name", ["cornell-dielectric", "veachMIS", "miPhong", "irradianceCache", "envMap", "sponzaXML", "test-scene", "alphaLeaf", "stackedCards"])
def test_obj_front_end_equals_the_reference_parser(name):
    """Every model of every scene: the vertex and index buffers of the product's loader against the reference's own OBJ parser
    (external/tiny_obj_loader.h compiled into oracle/_ref/libhost_ref.so: triangulation, index resolution, number parsing) followed by
    the corner-merging of SceneLoader::converteObjData (restated in oracle/host_ref.cpp: v-flip, face normals for files without
    normals, identical corners merged in first-occurrence order).  Positions, normals, texture coordinates and indices bit-equal;
    material indices equal up to the per-file offset the scene adds."""
    if not os.path.exists(HOST_REF):
        pytest.skip("oracle/_ref/libhost_ref.so needs the reference tree")
    P = helpers.pt()
    scene_file = helpers.scene_path(name)
    scene = P.Scene(scene_file)
    files = _obj_files_of(scene_file)
    assert len(files) == scene.desc.num_models and len(files) > 0
    for m, (path, mtl_dir, override) in enumerate(files):
        ours_v, ours_m, ours_i = _model_arrays(P, scene.desc, m)
        ref_v, ref_m, ref_i = _reference_obj(path, mtl_dir, override)
        assert ours_v.shape == ref_v.shape and ours_i.shape == ref_i.shape, (path, ours_v.shape, ref_v.shape)
        assert np.array_equal(ours_i, ref_i), path
        assert np.array_equal(ours_v.view(np.uint32), ref_v.view(np.uint32)) or np.array_equal(ours_v, ref_v), path      # (-0 / +0 compare equal)
        off = ours_m - ref_m
        assert (off == off[0]).all(), path


def test_obj_line_endings_lf_crlf_and_cr_only(tmp_path):
    """scenes/test-scene/models/sphere.obj and spherePhong.obj end their lines with a lone carriage return (old Mac export).  tinyobjloader
    reads them (safeGetline); a getline-based reader sees one long line and loads an empty model — which is what this loader did until
    round 2c: three of the JSON test scene's fourteen instances were missing on both sides of every parity test."""
    P = helpers.pt()
    body = ["# quad", "v 0 0 0", "v 1 0 0", "v 1 1 0", "v 0 1 0", "vn 0 0 1", "vt 0 0", "vt 1 0", "vt 1 1", "vt 0 1", "f 1/1/1 2/2/1 3/3/1 4/4/1"]
    shapes = []
    for name, eol in (("lf", "\n"), ("crlf", "\r\n"), ("cr", "\r")):
        (tmp_path / (name + ".obj")).write_bytes(eol.join(body).encode() + eol.encode())
        shapes.append('<shape type="obj"><string name="filename" value="%s.obj"/><bsdf type="diffuse"><rgb name="reflectance" value="0.5,0.5,0.5"/></bsdf></shape>' % name)
    xml = tmp_path / "s.xml"
    xml.write_text('<scene version="0.6.0"><sensor type="perspective"><float name="fov" value="40"/><transform name="toWorld"><lookat origin="0,0,3" target="0,0,0" up="0,1,0"/></transform></sensor>%s</scene>' % "".join(shapes))
    quad_scene = P.Scene(str(xml))       # (keep it alive: the description points into its buffers)
    d = quad_scene.desc
    assert d.num_models == 3
    for m in range(3):
        assert d.num_vertices[m] == 4 and d.num_indices[m] == 6, m          # the quad, triangulated as a fan
    scene = P.Scene(helpers.scene_path("test-scene"))
    assert all(scene.desc.num_indices[m] > 0 for m in range(scene.desc.num_models))
    assert scene.num_triangles == 19046 + 3 * 320                             # 3 instances of the two 320-triangle spheres were empty before


def test_mtl_materials_equal_the_reference_parser():
    """The JSON scenes take their materials from .mtl files: every material of scenes/test-scene against what the reference's parser
    (tinyobjloader, compiled into oracle/_ref/libhost_ref.so) reads from the same files, mapped like SceneLoader::addMaterials
    (src/SceneLoader.cpp:139-182): Ke / Kd / Ks / Ns / Ni, 1 / Ni, illum -> material type, materials accumulated file after file."""
    import ctypes as C
    if not os.path.exists(HOST_REF):
        pytest.skip("oracle/_ref/libhost_ref.so needs the reference tree")
    P = helpers.pt()
    scene_file = helpers.scene_path("test-scene")
    scene = P.Scene(scene_file)
    R = C.CDLL(HOST_REF)
    R.host_ref_obj_load.restype = C.c_void_p
    R.host_ref_obj_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    R.host_ref_obj_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    R.host_ref_obj_material.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    R.host_ref_obj_free.argtypes = [C.c_void_p]
    illum_to_type = {0: 0, 1: 0, 2: 4, 3: 1, 4: 2, 7: 2, 11: 3}          # eDiffuse, ePhong, eSpecular, eDielectric, eLight
    k = 0
    for path, mtl_dir, _ in _obj_files_of(scene_file):
        h = R.host_ref_obj_load(path.encode(), mtl_dir.encode(), -1)
        nv, ni, nm = C.c_int(), C.c_int(), C.c_int()
        R.host_ref_obj_counts(h, C.byref(nv), C.byref(ni), C.byref(nm))
        assert nm.value >= 1, path
        for i in range(nm.value):
            vals, illum = np.zeros(11, np.float32), C.c_int()
            dtex, stex = C.create_string_buffer(512), C.create_string_buffer(512)
            R.host_ref_obj_material(h, i, vals.ctypes.data, C.byref(illum), dtex, stex, 512)
            m = scene.desc.materials[k]
            assert np.array_equal(np.array(m.lightColor[:], np.float32), vals[0:3]) and np.array_equal(np.array(m.diffuse[:], np.float32), vals[3:6]), (path, i)
            assert np.array_equal(np.array(m.specular[:], np.float32), vals[6:9]) and np.float32(m.specularHighlight) == vals[9], (path, i)
            assert np.float32(m.refractionIndex) == vals[10] and np.float32(m.refractionIndexInv) == np.float32(1.0) / vals[10], (path, i)
            assert m.type == illum_to_type[illum.value], (path, i, illum.value)
            assert (m.textureIdDiffuse != -1) == bool(dtex.value or stex.value), (path, i)       # (sic: a specular map lands in textureIdDiffuse too, :175-177)
            k += 1
        R.host_ref_obj_free(h)
    assert k == scene.desc.num_materials


def test_json_instance_transforms_follow_parse_instances():
    """SceneLoader::parseInstances (src/SceneLoader.cpp:706-758): transform = T * Rx * Ry * Rz * S with GLM's post-multiplying
    translate / rotate (degrees -> radians, axes x, y, z in that order) / scale, normalTransform = Rx * Ry * Rz * S^-1; area lights get an
    instance index, point lights follow.  Recomputed here in float64 from the JSON file and compared with the loader's column-major matrices."""
    import json
    P = helpers.pt()
    path = helpers.scene_path("test-scene")
    scene = P.Scene(path)
    j = json.load(open(path))
    names = [list(m.keys())[0] for m in j["models"]]

    def rot(deg, axis):
        c, s = np.cos(np.radians(deg)), np.sin(np.radians(deg))
        m = np.eye(4)
        a, b = [(1, 2), (2, 0), (0, 1)][axis]
        m[a, a], m[a, b], m[b, a], m[b, b] = c, -s, s, c
        return m

    assert scene.desc.num_instances == len(j["instances"])
    for i, inst in enumerate(j["instances"]):
        name, props = list(inst.items())[0]
        T, N = np.eye(4), np.eye(4)
        if "translate" in props:
            t = np.eye(4); t[:3, 3] = props["translate"]; T = T @ t
        if "rotate" in props:
            for axis in range(3):
                T = T @ rot(props["rotate"][axis], axis); N = N @ rot(props["rotate"][axis], axis)
        if "scale" in props:
            T = T @ np.diag(list(props["scale"]) + [1.0]); N = N @ np.diag([1.0 / x for x in props["scale"]] + [1.0])
        got = scene.desc.instances[i]
        assert got.modelIndex == names.index(name)
        assert np.allclose(np.array(got.transform[:]).reshape(4, 4).T, T, rtol=1e-6, atol=1e-6), (i, name)
        assert np.allclose(np.array(got.normalTransform[:]).reshape(4, 4).T, N, rtol=1e-6, atol=1e-6), (i, name)
    lights = [scene.desc.lights[k] for k in range(scene.desc.num_lights)]
    area = [l for l in lights if l.type == 0]
    assert [l.instanceIndex for l in area] == [i for i, inst in enumerate(j["instances"]) if list(inst.keys())[0] == "lightBox"]
    points = [l for l in lights if l.type == 1]
    assert len(points) == len(j["lights"]) and lights[-len(points):] == points or all(l.type == 1 for l in lights[len(area):])
    for l, jl in zip(points, j["lights"]):
        assert np.allclose(l.color[:], jl["color"]) and np.allclose(l.pos[:], jl["position"])


def _expected_xml_scene(scene_file):
    """SceneLoader::parseMitsubaSceneFile recomputed with ElementTree (src/SceneLoader.cpp:329-648, src/MitsubaXML.h): the material list —
    top-level <bsdf> in document order, then per shape an inline <bsdf> and, for a shape with an <emitter>, a light copy of its
    material — and the material index of every obj model and sphere."""
    import xml.etree.ElementTree as ET
    root = ET.parse(scene_file).getroot()

    def named(el, tag, name):
        for c in el.findall(tag):
            if c.get("name") == name:
                return c
        raise KeyError(name)

    def rgb(el, name):
        vals = [float(x) for x in named(el, "rgb", name).get("value").replace(",", " ").split()]
        return vals * 3 if len(vals) == 1 else vals

    textures = {}
    for i, t in enumerate(root.findall("texture")):
        textures[t.get("id")] = 1 + i                      # slot 0 is the environment map; scenes here do not repeat a file

    def bsdf(el):
        m = dict(type=None, diffuse=None, specular=None, exponent=None, ior=None, eta=None, k=None, alpha=None, tex=-1, light=None)
        t = el.get("type")
        if t == "phong":
            m.update(type=4, specular=rgb(el, "specularReflectance"), diffuse=rgb(el, "diffuseReflectance"), exponent=float(named(el, "float", "exponent").get("value")))
        elif t == "diffuse":
            m.update(type=0, specular=[0, 0, 0], exponent=0.0)
            ref = el.find("ref")
            if ref is not None and ref.get("name") == "reflectance":
                m.update(diffuse=[1, 1, 1], tex=textures[ref.get("id")])
            else:
                m.update(diffuse=rgb(el, "reflectance"))
        elif t == "dielectric":
            m.update(type=2, specular=[1, 1, 1], ior=np.float32(named(el, "float", "intIOR").get("value")) / np.float32(named(el, "float", "extIOR").get("value")))
        elif t == "conductor":
            mat = [s for s in el.findall("string") if s.get("name") == "material"]
            if mat and mat[0].get("value") == "none":
                m.update(type=1, specular=[1, 1, 1])
            else:
                m.update(type=5, eta=float(named(el, "spectrum", "eta").get("value")), k=float(named(el, "spectrum", "k").get("value")))
        elif t == "roughconductor":
            m.update(type=6, alpha=float(named(el, "float", "alpha").get("value")), eta=float(named(el, "spectrum", "eta").get("value")), k=float(named(el, "spectrum", "k").get("value")))
        return m

    mats, by_id = [], {}
    for el in root.findall("bsdf"):
        by_id[el.get("id")] = len(mats)
        mats.append(bsdf(el))
    models, spheres = [], []
    for sh in root.findall("shape"):
        idx = -1
        inline = sh.find("bsdf")
        if inline is not None:
            idx = len(mats); mats.append(bsdf(inline))
        if idx < 0 and sh.find("ref") is not None:
            idx = by_id[sh.find("ref").get("id")]
        em = sh.find("emitter")
        if em is not None:
            light = dict(mats[idx]); light.update(type=3, light=rgb(em, "radiance"))
            mats.append(light); idx = len(mats) - 1
        if sh.get("type") == "obj":
            models.append(idx)
        elif sh.get("type") == "sphere":
            p = named(sh, "point", "center")
            spheres.append((idx, float(named(sh, "float", "radius").get("value")), [float(p.get(a)) for a in "xyz"]))
    return mats, models, spheres


@pytest.mark.parametrize("name", ["cornell-dielectric", "veachMIS", "miPhong", "irradianceCache", "envMap", "sponzaXML", "alphaLeaf", "stackedCards", "testSpheres", "envSynthetic"])
def test_xml_materials_shapes_and_spheres_follow_the_reference_loader(name):
    P = helpers.pt()
    scene_file = helpers.scene_path(name)
    scene = P.Scene(scene_file)
    d = scene.desc
    mats, models, spheres = _expected_xml_scene(scene_file)
    assert d.num_materials == len(mats) and d.num_models == len(models) and d.num_spheres == len(spheres)
    f32 = lambda x: np.array(x, np.float32)
    for i, e in enumerate(mats):
        m = d.materials[i]
        assert m.type == e["type"], (i, m.type, e)
        if e["diffuse"] is not None:
            assert np.array_equal(f32(m.diffuse[:]), f32(e["diffuse"])), i
        if e["specular"] is not None:
            assert np.array_equal(f32(m.specular[:]), f32(e["specular"])), i
        if e["exponent"] is not None:
            assert np.float32(m.specularHighlight) == np.float32(e["exponent"]), i
        if e["ior"] is not None:
            assert np.float32(m.refractionIndex) == e["ior"] and np.float32(m.refractionIndexInv) == np.float32(1.0) / e["ior"], i
        if e["eta"] is not None:
            assert np.float32(m.eta) == np.float32(e["eta"]) and np.float32(m.k) == np.float32(e["k"]), i
        if e["alpha"] is not None:
            assert np.float32(m.roughness) == np.float32(e["alpha"]), i
        if e["light"] is not None:
            assert np.array_equal(f32(m.lightColor[:]), f32(e["light"])), i
        assert m.textureIdDiffuse == e["tex"], (i, m.textureIdDiffuse, e["tex"])
    for mi, want in enumerate(models):
        _, vm, _ = _model_arrays(P, d, mi)
        assert (vm == want).all(), (mi, want, np.unique(vm))
    for si, (want, radius, center) in enumerate(spheres):
        s = d.spheres[si]
        assert s.materialIndex == want and np.float32(s.radius) == np.float32(radius) and np.array_equal(f32(s.center[:]), f32(center)), si


def test_malformed_scene_files_raise_instead_of_crashing(tmp_path):
    """Scene files are untrusted input.  Found by fuzzing in round 2c: a face index of 0 / out of range / overflowing long made the OBJ
    reader index outside its arrays (the reference's parser does not look either); a truncated close tag sent the XML parser into endless
    recursion; a coordinate that overflows float (1e40) sent the BVH builder's binning out of bounds.  All are loader errors now."""
    import shutil
    P = helpers.pt()
    src = os.path.join(helpers.SCENES, "alphaLeaf")

    def variant(name, rel, mutate):
        dst = tmp_path / name
        shutil.copytree(src, dst)
        p = dst / rel
        p.write_bytes(mutate(p.read_bytes()))
        with pytest.raises(P.B200ptError):
            P.Scene(str(dst / "alphaLeaf.xml")).bvh_check()

    face = b"f 1/1/1 3/3/1 2/2/1"
    assert face in open(os.path.join(src, "floor.obj"), "rb").read()
    for k, bad in enumerate((b"f 1/1/1 -999999/3/1 2/2/1", b"f 1/2147483647/1 3/3/1 2/2/1", b"f 1/1/1 3/3/99999999999 2/2/1", b"f 0/1/1 3/3/1 2/2/1", b"f 1/1/1 3/3/1 77/2/1")):
        variant("idx%d" % k, "floor.obj", lambda d, bad=bad: d.replace(face, bad))
    variant("inf", "floor.obj", lambda d: d.replace(b"v ", b"v 1e40 ", 1))
    xml = open(os.path.join(src, "alphaLeaf.xml"), "rb").read()
    for k, cut in enumerate((len(xml) - 3, len(xml) - 12, len(xml) // 2, xml.rindex(b"</shape>") + 5)):
        variant("cut%d" % k, "alphaLeaf.xml", lambda d, cut=cut: d[:cut])
    variant("deep", "alphaLeaf.xml", lambda d: d.replace(b"<scene", b"<a>" * 400 + b"<scene", 1))
    # and a seeded sweep of byte flips / truncations over the three files: the loader returns or raises
    rng = np.random.default_rng(3)
    outcomes = {"ok": 0, "raised": 0}
    for it in range(120):
        dst = tmp_path / ("fz%d" % it)
        shutil.copytree(src, dst)
        p = dst / ["floor.obj", "leafquad.obj", "alphaLeaf.xml"][it % 3]
        d = bytearray(p.read_bytes())
        if it % 2:
            d = d[:rng.integers(1, len(d))]
        else:
            for _ in range(rng.integers(1, 8)):
                d[rng.integers(0, len(d))] = rng.integers(32, 127)
        p.write_bytes(bytes(d))
        try:
            P.Scene(str(dst / "alphaLeaf.xml")).bvh_check()
            outcomes["ok"] += 1
        except P.B200ptError:
            outcomes["raised"] += 1
        shutil.rmtree(dst)
    assert outcomes["ok"] > 5 and outcomes["raised"] > 20, outcomes
