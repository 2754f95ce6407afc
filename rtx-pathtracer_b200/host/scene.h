// Host-side scene representation: the headless replacement of the reference's SceneLoader / Model / Shapes /
// WeightedSampler classes (src/SceneLoader.{h,cpp}, src/Model.{h,cpp}, src/Shapes.h, src/WeightedSampler.{h,cpp}).
// No Vulkan objects: the output is the set of plain buffers the reference binds to descriptor set 1.
#pragma once
#include <string>
#include <vector>
#include <map>
#include <stdexcept>
#include "../../include/b200pt.h"

namespace b200pt {

struct Vec3 { float x, y, z; };

struct Model {                       // src/Model.h:42-66 without the vk:: members
    std::vector<b200pt_vertex> vertices;
    std::vector<uint32_t> indices;
    float faceArea(int iFace, const float transform[16]) const;     // Model::getFaceArea, src/Model.cpp:35-57
    void aabb(const float transform[16], float mn[3], float mx[3]) const;  // Model::getAabb, src/Model.cpp:66-78
};

struct TextureImage {
    int width = 0, height = 0, format = B200PT_TEX_RGBA8_SRGB;
    std::vector<uint8_t> rgba8;      // format RGBA8_SRGB
    std::vector<float> rgba32f;      // format RGBA32F
    std::string path;
};

// JPEG (baseline) / PNG file -> RGBA8, row 0 first (host/image_decode.cpp; stands in for stbi_load(path, .., 4))
void decodeImageFile(const std::string &path, int &width, int &height, std::vector<uint8_t> &rgba);

// CDF-walk sampler with a default-seeded std::mt19937 — src/WeightedSampler.{h,cpp}
class WeightedSampler {
public:
    explicit WeightedSampler(const std::vector<float> &values);
    int sample();
    std::vector<float> probabilities() const;
    float total() const { return total_; }
private:
    std::vector<float> values_;
    float total_;
    struct Impl; Impl *impl_;
public:
    ~WeightedSampler();
    WeightedSampler(const WeightedSampler &) = delete;
};

struct Scene {                       // src/SceneLoader.h:98-150
    std::vector<Model> models;
    std::vector<std::vector<int>> emissiveFacesPerModel;
    std::vector<b200pt_material> materials;
    std::vector<b200pt_instance> instances;
    std::vector<b200pt_light> lights;
    std::vector<b200pt_sphere> spheres;
    std::vector<TextureImage> textures;          // [0] = env map slot
    std::vector<int32_t> randomLightIndex;       // SIZE_LIGHT_RANDOM
    std::vector<b200pt_face_sample> randomTriIndex; // nMeshLights * SIZE_TRI_RANDOM
    int numFaceTables = 0;
    float sceneMin[3], sceneMax[3];
    // camera defaults: src/SceneLoader.h:147-150
    float origin[3] = {0, -10, 4}, target[3] = {0, 0, 4}, upDir[3] = {0, 0, 1};
    float vfov = 28.0f;

    // filled by finalize(): pointer tables for b200pt_scene_desc
    std::vector<const b200pt_vertex *> vertexPtrs;
    std::vector<const uint32_t *> indexPtrs;
    std::vector<int32_t> numVertices, numIndices;
    std::vector<b200pt_texture> textureDescs;

    void loadFile(const std::string &path);       // SceneLoader::SceneLoader, src/SceneLoader.cpp:29-65
    void finalize();                              // calculateSceneSize + light tables + pointer tables
    void fillDesc(b200pt_scene_desc *d) const;
private:
    std::string modelsBaseDir, materialBaseDir, textureBaseDir;
    std::map<std::string, int> pathTextureId;
    void parseMitsubaSceneFile(const std::string &path);   // src/SceneLoader.cpp:329-348
    void parseJsonSceneFile(const std::string &path);      // src/SceneLoader.cpp:650-770
    int addTexture(const std::string &name);               // src/SceneLoader.cpp:184-236
    void calculateSceneSize();                             // src/SceneLoader.cpp:350-368
    void buildLightTables();                               // src/SceneLoader.cpp:847-944
    friend struct SceneBuilder;
};

// OBJ reader: vertex / normal / texcoord pools + triangulated faces (the subset of tinyobjloader the reference uses)
struct ObjIndex { int v, vt, vn; };
struct ObjMaterial {
    std::string name;
    float emission[3] = {0, 0, 0}, diffuse[3] = {0, 0, 0}, specular[3] = {0, 0, 0};      // InitMaterial of the reference's vendored tiny_obj_loader.h:1276-1310 (newer tinyobj versions default Kd to 0.6)
    float shininess = 1.0f, ior = 1.0f;
    int illum = 0;
    std::string diffuseTex, specularTex;
};
struct ObjData {
    std::vector<float> positions, normals, texcoords;
    std::vector<ObjIndex> faceIndices;      // 3 per triangle, shapes concatenated in file order
    std::vector<int> faceMaterial;          // per triangle, index into materials or -1
    std::vector<ObjMaterial> materials;
};
void readObj(const std::string &path, const std::string &mtlBaseDir, ObjData &out);

// src/SceneLoader.cpp:250-327 converteObjData
void convertObjData(const ObjData &obj, const std::vector<b200pt_material> &materials, int materialIndexOffset,
                    int materialIndexOverride, std::vector<b200pt_vertex> &outVertices,
                    std::vector<uint32_t> &outIndices, std::vector<int> &outEmissiveFaces);

// column-major 4x4 helpers (glm conventions)
void mat4Identity(float m[16]);
void mat4Mul(const float a[16], const float b[16], float out[16]);
bool mat4Inverse(const float m[16], float out[16]);
void mat4TransformPoint(const float m[16], const float p[3], float out[3]);

// EXR IO (scanline files; reader: NONE / RLE / ZIPS / ZIP / PIZ / PXR24 / B44 / B44A, HALF / FLOAT channels; writer: NONE, FLOAT)
void writeExrRGB(const std::string &path, const float *rgba, int width, int height);
void readExrRGBA(const std::string &path, std::vector<float> &rgba, int &width, int &height, bool viaHalf);
float halfToFloat(uint16_t h);
uint16_t floatToHalf(float f);

}  // namespace b200pt
