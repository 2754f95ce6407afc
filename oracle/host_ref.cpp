// TEST INFRASTRUCTURE ONLY.  C entry point around the reference's OWN WeightedSampler (src/WeightedSampler.cpp, compiled where
// it lies under /root/reference by oracle/Makefile into oracle/_ref/libhost_ref.so; nothing is copied into this repository).
// SceneLoader::getLightSamplingVector / getFaceSamplingVector (src/SceneLoader.cpp:861-944) draw the 10 000-entry light and
// face tables of binding 6 from it; tests/test_scene.py feeds it the weights our loader reports and compares the tables.
#include <vector>
#include "WeightedSampler.h"

extern "C" int host_ref_weighted_samples(const float *values, int n, int count, int *samples_out, float *probs_out, float *total_out) {
    std::vector<float> v(values, values + n);
    WeightedSampler sampler(v);                        // a fresh sampler per table, like the reference (default-seeded mt19937)
    std::vector<float> probs = sampler.getProbabilities();
    for (int i = 0; i < n; i++) probs_out[i] = probs[size_t(i)];
    *total_out = sampler.getTotal();
    for (int i = 0; i < count; i++) samples_out[i] = sampler.sample();
    return 0;
}

// ---- OBJ front end: the reference's own parser (external/tiny_obj_loader.h, compiled here where it lies) followed by
// SceneLoader::converteObjData (src/SceneLoader.cpp:250-327; SceneLoader.cpp itself needs Vulkan, so this step is restated):
// positions / normals / texture coordinates per face corner (v flipped, face normal when the file has none), identical corners
// merged in first-occurrence order (Vertex::operator==, src/Model.h:29-31).  tests/test_scene.py compares the product's loader.
#define TINYOBJLOADER_IMPLEMENTATION
#include "tiny_obj_loader.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>

namespace {
struct RefVertex {
    float pos[3], normal[3], uv[2]; int mat;
    bool operator==(const RefVertex &o) const {
        for (int i = 0; i < 3; i++) if (!(pos[i] == o.pos[i] && normal[i] == o.normal[i])) return false;
        return uv[0] == o.uv[0] && uv[1] == o.uv[1] && mat == o.mat;
    }
};
struct RefVertexHash {
    size_t operator()(const RefVertex &v) const {       // any hash consistent with operator== (-0 == +0): the output order does not depend on it
        size_t h = 17;
        auto mix = [&](float f) { if (f == 0.0f) f = 0.0f; uint32_t b; memcpy(&b, &f, 4); h = h * 31 + b; };
        for (int i = 0; i < 3; i++) { mix(v.pos[i]); mix(v.normal[i]); }
        mix(v.uv[0]); mix(v.uv[1]);
        return h * 31 + size_t(v.mat);
    }
};
struct RefObj { std::vector<RefVertex> vertices; std::vector<uint32_t> indices; int numMaterials = 0; std::string error; std::vector<tinyobj::material_t> materials; };
}

extern "C" void *host_ref_obj_load(const char *path, const char *mtl_dir, int material_override) {
    RefObj *r = new RefObj();
    tinyobj::attrib_t attrib; std::vector<tinyobj::shape_t> shapes; std::vector<tinyobj::material_t> materials;
    std::string warn, err;
    if (!tinyobj::LoadObj(&attrib, &shapes, &materials, &warn, &err, path, mtl_dir)) { r->error = warn + err; return r; }
    r->numMaterials = int(materials.size());
    r->materials = materials;
    const bool hasNormals = !attrib.normals.empty(), hasTexCoords = !attrib.texcoords.empty();
    std::unordered_map<RefVertex, uint32_t, RefVertexHash> unique;
    for (const auto &shape : shapes) {
        const size_t numFaces = shape.mesh.indices.size() / 3;
        for (size_t f = 0; f < numFaces; f++) {
            const int mat = material_override >= 0 ? material_override : shape.mesh.material_ids[f];
            RefVertex fv[3];
            for (int i = 0; i < 3; i++) {
                const auto &idx = shape.mesh.indices[3 * f + i];
                RefVertex v; memset(&v, 0, sizeof(v));
                for (int a = 0; a < 3; a++) v.pos[a] = attrib.vertices[3 * idx.vertex_index + a];
                if (hasNormals) for (int a = 0; a < 3; a++) v.normal[a] = attrib.normals[3 * idx.normal_index + a];
                if (hasTexCoords) { v.uv[0] = attrib.texcoords[2 * idx.texcoord_index + 0]; v.uv[1] = 1.0f - attrib.texcoords[2 * idx.texcoord_index + 1]; }
                v.mat = mat;
                fv[i] = v;
            }
            if (!hasNormals) {       // glm::normalize(glm::cross(ab, ac)): x * inversesqrt(dot(x, x)) in GLM's scalar path
                float ab[3], ac[3], n[3];
                for (int a = 0; a < 3; a++) { ab[a] = fv[1].pos[a] - fv[0].pos[a]; ac[a] = fv[2].pos[a] - fv[0].pos[a]; }
                n[0] = ab[1] * ac[2] - ac[1] * ab[2]; n[1] = ab[2] * ac[0] - ac[2] * ab[0]; n[2] = ab[0] * ac[1] - ac[0] * ab[1];
                const float inv = 1.0f / std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
                for (int i = 0; i < 3; i++) for (int a = 0; a < 3; a++) fv[i].normal[a] = n[a] * inv;
            }
            for (int i = 0; i < 3; i++) {
                auto it = unique.find(fv[i]);
                if (it == unique.end()) { it = unique.emplace(fv[i], uint32_t(r->vertices.size())).first; r->vertices.push_back(fv[i]); }
                r->indices.push_back(it->second);
            }
        }
    }
    return r;
}
extern "C" const char *host_ref_obj_error(void *h) { return static_cast<RefObj *>(h)->error.c_str(); }
extern "C" void host_ref_obj_counts(void *h, int *num_vertices, int *num_indices, int *num_materials) {
    RefObj *r = static_cast<RefObj *>(h);
    *num_vertices = int(r->vertices.size()); *num_indices = int(r->indices.size()); *num_materials = r->numMaterials;
}
extern "C" void host_ref_obj_copy(void *h, float *verts8 /* pos, normal, uv */, int *mats, uint32_t *indices) {
    RefObj *r = static_cast<RefObj *>(h);
    for (size_t i = 0; i < r->vertices.size(); i++) { memcpy(verts8 + 8 * i, &r->vertices[i], 32); mats[i] = r->vertices[i].mat; }
    memcpy(indices, r->indices.data(), r->indices.size() * 4);
}
extern "C" void host_ref_obj_free(void *h) { delete static_cast<RefObj *>(h); }

// material i of the file's .mtl as the reference's parser read it: emission, diffuse, specular (3 each), shininess, ior, then illum and
// the two texture names SceneLoader::addMaterials looks at (src/SceneLoader.cpp:139-182)
extern "C" void host_ref_obj_material(void *h, int i, float out11[11], int *illum, char *diffuse_tex, char *specular_tex, int cap) {
    const tinyobj::material_t &m = static_cast<RefObj *>(h)->materials[size_t(i)];
    for (int a = 0; a < 3; a++) { out11[a] = m.emission[a]; out11[3 + a] = m.diffuse[a]; out11[6 + a] = m.specular[a]; }
    out11[9] = m.shininess; out11[10] = m.ior; *illum = m.illum;
    snprintf(diffuse_tex, size_t(cap), "%s", m.diffuse_texname.c_str()); snprintf(specular_tex, size_t(cap), "%s", m.specular_texname.c_str());
}
