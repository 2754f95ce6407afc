// Headless scene front-end.  Follows the behaviour of the reference's src/SceneLoader.cpp (Mitsuba-0.6 XML subset
// and the JSON scene format), src/MitsubaXML.h, src/Model.cpp, src/WeightedSampler.cpp — re-written without
// Vulkan, tinyxml2, tinyobjloader, nlohmann::json or GLM.  Line references are to files under /root/reference.
#include "scene.h"
#include <cmath>
#include <cstring>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <random>
#include <unordered_map>
#include <memory>
#include <limits>
#include <algorithm>
#include <filesystem>

namespace b200pt {

// ------------------------------------------------------------------------------------------------ matrices
void mat4Identity(float m[16]) { for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.0f : 0.0f; }

void mat4Mul(const float a[16], const float b[16], float out[16]) {
    float r[16];
    for (int c = 0; c < 4; c++)
        for (int rw = 0; rw < 4; rw++) {
            float s = 0;
            for (int k = 0; k < 4; k++) s += a[k * 4 + rw] * b[c * 4 + k];
            r[c * 4 + rw] = s;
        }
    memcpy(out, r, sizeof(r));
}

// cofactor inverse (the algorithm family glm::inverse uses), column-major
bool mat4Inverse(const float m[16], float inv[16]) {
    float t[16];
    t[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    t[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    t[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    t[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    t[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    t[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    t[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    t[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    t[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    t[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    t[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    t[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    t[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    t[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    t[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    t[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    float det = m[0] * t[0] + m[1] * t[4] + m[2] * t[8] + m[3] * t[12];
    if (det == 0) return false;
    float id = 1.0f / det;
    for (int i = 0; i < 16; i++) inv[i] = t[i] * id;
    return true;
}

void mat4TransformPoint(const float m[16], const float p[3], float out[3]) {
    for (int r = 0; r < 3; r++) out[r] = ((m[0 + r] * p[0] + m[4 + r] * p[1]) + m[8 + r] * p[2]) + m[12 + r];
}

static void mat4Translate(float m[16], float x, float y, float z) {  // glm::translate(m, v)
    for (int r = 0; r < 4; r++) m[12 + r] = m[0 + r] * x + m[4 + r] * y + m[8 + r] * z + m[12 + r];
}
static void mat4Scale(float m[16], float x, float y, float z) {      // glm::scale(m, v)
    for (int r = 0; r < 4; r++) { m[0 + r] *= x; m[4 + r] *= y; m[8 + r] *= z; }
}
static void mat4Rotate(float m[16], float angle, float ax, float ay, float az) {  // glm::rotate(m, angle, axis)
    float c = cosf(angle), s = sinf(angle);
    float l = sqrtf(ax * ax + ay * ay + az * az);
    ax /= l; ay /= l; az /= l;
    float tx = (1 - c) * ax, ty = (1 - c) * ay, tz = (1 - c) * az;
    float R[16];
    mat4Identity(R);
    R[0] = c + tx * ax;      R[1] = tx * ay + s * az; R[2] = tx * az - s * ay;
    R[4] = ty * ax - s * az; R[5] = c + ty * ay;      R[6] = ty * az + s * ax;
    R[8] = tz * ax + s * ay; R[9] = tz * ay - s * ax; R[10] = c + tz * az;
    mat4Mul(m, R, m);
}

// ------------------------------------------------------------------------------------------------ Model
float Model::faceArea(int iFace, const float T[16]) const {      // src/Model.cpp:35-57
    float v[3][3];
    for (int k = 0; k < 3; k++) mat4TransformPoint(T, vertices[indices[3 * iFace + k]].pos, v[k]);
    float d01[3], d02[3];
    for (int a = 0; a < 3; a++) { d01[a] = v[1][a] - v[0][a]; d02[a] = v[2][a] - v[0][a]; }
    float l01 = sqrtf(d01[0] * d01[0] + d01[1] * d01[1] + d01[2] * d01[2]);
    float l02 = sqrtf(d02[0] * d02[0] + d02[1] * d02[1] + d02[2] * d02[2]);
    float cosAlpha = (d01[0] * d02[0] + d01[1] * d02[1] + d01[2] * d02[2]) / (l01 * l02);
    float sinAlpha = sqrtf(1 - cosAlpha * cosAlpha);
    return l01 * l02 * sinAlpha / 2.0f;
}

void Model::aabb(const float T[16], float mn[3], float mx[3]) const {   // src/Model.cpp:66-78
    const float inf = std::numeric_limits<float>::infinity();
    for (int a = 0; a < 3; a++) { mn[a] = inf; mx[a] = -inf; }
    for (const auto &v : vertices) {
        float p[3];
        mat4TransformPoint(T, v.pos, p);
        for (int a = 0; a < 3; a++) { mn[a] = fminf(mn[a], p[a]); mx[a] = fmaxf(mx[a], p[a]); }
    }
}

// ------------------------------------------------------------------------------------------------ WeightedSampler
struct WeightedSampler::Impl {
    std::mt19937 generator;                                   // default seed 5489, src/WeightedSampler.h:29
    std::uniform_real_distribution<float> distribution{0, 1};
};
WeightedSampler::WeightedSampler(const std::vector<float> &values) : values_(values), impl_(new Impl) {
    total_ = 0.0f;
    for (float v : values_) total_ += v;
}
WeightedSampler::~WeightedSampler() { delete impl_; }
int WeightedSampler::sample() {                               // src/WeightedSampler.cpp:8-25
    if (values_.empty()) return -1;
    float random = impl_->distribution(impl_->generator);
    random *= total_;
    float sum = values_[0];
    for (size_t i = 1; i < values_.size(); ++i) {
        if (sum > random) return int(i) - 1;
        sum += values_[i];
    }
    return int(values_.size()) - 1;
}
std::vector<float> WeightedSampler::probabilities() const {
    std::vector<float> p(values_.size());
    for (size_t i = 0; i < values_.size(); i++) p[i] = values_[i] / total_;
    return p;
}

// ------------------------------------------------------------------------------------------------ OBJ / MTL
static bool parseFloats(const char *s, float *out, int n) {
    char *end;
    for (int i = 0; i < n; i++) {
        out[i] = strtof(s, &end);
        if (end == s) return i > 0 && false;
        s = end;
    }
    return true;
}

// one line of an OBJ / MTL file: terminated by "\n", "\r\n" or a lone "\r" (old Mac exports — test-scene's sphere.obj is one), like
// tinyobjloader's safeGetline; std::getline would hand such a file over as a single line and every statement after the first be lost
static bool getLineAnyEol(std::istream &in, std::string &line) {
    line.clear();
    std::streambuf *sb = in.rdbuf();
    if (!in.good()) return false;
    for (;;) {
        const int c = sb->sbumpc();
        if (c == '\n') return true;
        if (c == '\r') { if (sb->sgetc() == '\n') sb->sbumpc(); return true; }
        if (c == std::streambuf::traits_type::eof()) { in.setstate(std::ios::eofbit); return !line.empty(); }
        line.push_back(char(c));
    }
}

static void readMtl(const std::string &path, std::vector<ObjMaterial> &mats, std::map<std::string, int> &byName) {
    std::ifstream in(path);
    if (!in) return;   // tinyobj only warns when a material file is missing
    std::string line;
    ObjMaterial cur;
    bool have = false;
    auto flush = [&]() { if (have) { byName[cur.name] = int(mats.size()); mats.push_back(cur); } };
    while (getLineAnyEol(in, line)) {
        size_t p = line.find_first_not_of(" \t\r");
        if (p == std::string::npos || line[p] == '#') continue;
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
        std::istringstream ss(line.substr(p));
        std::string key;
        ss >> key;
        if (key == "newmtl") { flush(); cur = ObjMaterial(); have = true; ss >> std::ws; std::getline(ss, cur.name); }
        else if (key == "Ke") ss >> cur.emission[0] >> cur.emission[1] >> cur.emission[2];
        else if (key == "Kd") ss >> cur.diffuse[0] >> cur.diffuse[1] >> cur.diffuse[2];
        else if (key == "Ks") ss >> cur.specular[0] >> cur.specular[1] >> cur.specular[2];
        else if (key == "Ns") ss >> cur.shininess;
        else if (key == "Ni") ss >> cur.ior;
        else if (key == "illum") ss >> cur.illum;
        else if (key == "map_Kd") { ss >> std::ws; std::getline(ss, cur.diffuseTex); }
        else if (key == "map_Ks") { ss >> std::ws; std::getline(ss, cur.specularTex); }
    }
    flush();
}

// Ear clipping in the polygon's dominant plane — the triangulation rule of the OBJ reader the reference uses
// (external/tiny_obj_loader.h:1371-1576).  Convex polygons come out as the fan (0,1,2),(0,2,3),...
static void triangulate(const std::vector<ObjIndex> &poly, const std::vector<float> &pos, std::vector<ObjIndex> &out) {
    size_t n = poly.size();
    if (n < 3) return;
    if (n == 3) { out.insert(out.end(), poly.begin(), poly.end()); return; }
    auto P = [&](const ObjIndex &i, int a) { return pos[size_t(i.v) * 3 + a]; };
    int ax0 = 1, ax1 = 2;
    for (size_t k = 0; k < n; ++k) {
        const ObjIndex &i0 = poly[k % n], &i1 = poly[(k + 1) % n], &i2 = poly[(k + 2) % n];
        float e0[3], e1[3];
        for (int a = 0; a < 3; a++) { e0[a] = P(i1, a) - P(i0, a); e1[a] = P(i2, a) - P(i1, a); }
        float cx = fabsf(e0[1] * e1[2] - e0[2] * e1[1]);
        float cy = fabsf(e0[2] * e1[0] - e0[0] * e1[2]);
        float cz = fabsf(e0[0] * e1[1] - e0[1] * e1[0]);
        const float eps = std::numeric_limits<float>::epsilon();
        if (cx > eps || cy > eps || cz > eps) {
            if (!(cx > cy && cx > cz)) { ax0 = 0; if (cz > cx && cz > cy) ax1 = 1; }
            break;
        }
    }
    float area = 0;
    for (size_t k = 0; k < n; ++k) {
        const ObjIndex &i0 = poly[k], &i1 = poly[(k + 1) % n];
        area += (P(i0, ax0) * P(i1, ax1) - P(i0, ax1) * P(i1, ax0)) * 0.5f;
    }
    std::vector<ObjIndex> rem = poly;
    size_t guess = 0, remainingIterations = n, previous = n;
    while (rem.size() > 3 && remainingIterations > 0) {
        size_t np = rem.size();
        if (guess >= np) guess -= np;
        if (previous != np) { previous = np; remainingIterations = np; } else remainingIterations--;
        ObjIndex ind[3];
        float vx[3], vy[3];
        for (int k = 0; k < 3; k++) { ind[k] = rem[(guess + k) % np]; vx[k] = P(ind[k], ax0); vy[k] = P(ind[k], ax1); }
        float cross = (vx[1] - vx[0]) * (vy[2] - vy[1]) - (vy[1] - vy[0]) * (vx[2] - vx[1]);
        if (cross * area < 0.0f) { guess++; continue; }
        bool overlap = false;
        for (size_t o = 3; o < np && !overlap; ++o) {
            const ObjIndex &oi = rem[(guess + o) % np];
            float tx = P(oi, ax0), ty = P(oi, ax1);
            bool c = false;   // point-in-triangle by crossing number
            for (int i = 0, j = 2; i < 3; j = i++)
                if (((vy[i] > ty) != (vy[j] > ty)) && (tx < (vx[j] - vx[i]) * (ty - vy[i]) / (vy[j] - vy[i]) + vx[i])) c = !c;
            overlap = c;
        }
        if (overlap) { guess++; continue; }
        out.push_back(ind[0]); out.push_back(ind[1]); out.push_back(ind[2]);
        rem.erase(rem.begin() + (guess + 1) % np);
    }
    if (rem.size() == 3) out.insert(out.end(), rem.begin(), rem.end());
}

void readObj(const std::string &path, const std::string &mtlBaseDir, ObjData &out) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("Cannot open OBJ file " + path);   // src/SceneLoader.cpp:133-136
    std::map<std::string, int> matByName;
    int curMat = -1;
    std::string line;
    std::vector<ObjIndex> poly, tris;
    while (getLineAnyEol(in, line)) {
        const char *s = line.c_str();
        while (*s == ' ' || *s == '\t') s++;
        if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
            float f[3] = {0, 0, 0};
            parseFloats(s + 2, f, 3);
            out.positions.insert(out.positions.end(), f, f + 3);
        } else if (s[0] == 'v' && s[1] == 'n' && (s[2] == ' ' || s[2] == '\t')) {
            float f[3] = {0, 0, 0};
            parseFloats(s + 3, f, 3);
            out.normals.insert(out.normals.end(), f, f + 3);
        } else if (s[0] == 'v' && s[1] == 't' && (s[2] == ' ' || s[2] == '\t')) {
            float f[2] = {0, 0};
            parseFloats(s + 3, f, 2);
            out.texcoords.insert(out.texcoords.end(), f, f + 2);
        } else if (s[0] == 'f' && (s[1] == ' ' || s[1] == '\t')) {
            poly.clear();
            const char *p = s + 2;
            int nv = int(out.positions.size() / 3), nt = int(out.texcoords.size() / 2), nn = int(out.normals.size() / 3);
            while (*p) {
                while (*p == ' ' || *p == '\t' || *p == '\r') p++;
                if (!*p) break;
                ObjIndex idx{-1, -1, -1};
                char *e;
                // 1-based, or negative = relative to what has been read so far.  The reference's parser resolves an index without
                // looking at it and the loader then reads out of bounds; a file is untrusted input here: anything outside the
                // arrays (0, a forward reference, an overflowing number) is an error of the loader, not a wild read.
                auto resolve = [&](long a, int count, const char *what) -> int {
                    const long r = a > 0 ? a - 1 : long(count) + a;
                    if (r < 0 || r >= long(count)) throw std::runtime_error(std::string("OBJ: ") + what + " index out of range in " + path);
                    return int(r);
                };
                long a = strtol(p, &e, 10);
                if (e == p) break;
                idx.v = resolve(a, nv, "vertex");
                p = e;
                if (*p == '/') {
                    p++;
                    if (*p != '/') { long b = strtol(p, &e, 10); if (e != p) idx.vt = resolve(b, nt, "texture coordinate"); p = e; }
                    if (*p == '/') { p++; long c = strtol(p, &e, 10); if (e != p) idx.vn = resolve(c, nn, "normal"); p = e; }
                }
                poly.push_back(idx);
            }
            tris.clear();
            triangulate(poly, out.positions, tris);
            for (size_t k = 0; k + 2 < tris.size(); k += 3) {
                out.faceIndices.push_back(tris[k]); out.faceIndices.push_back(tris[k + 1]); out.faceIndices.push_back(tris[k + 2]);
                out.faceMaterial.push_back(curMat);
            }
        } else if (!strncmp(s, "usemtl", 6)) {
            std::string name(s + 6);
            size_t a = name.find_first_not_of(" \t"), b = name.find_last_not_of(" \t\r");
            name = (a == std::string::npos) ? "" : name.substr(a, b - a + 1);
            auto it = matByName.find(name);
            curMat = it == matByName.end() ? -1 : it->second;
        } else if (!strncmp(s, "mtllib", 6)) {
            std::istringstream ss(s + 6);
            std::string f;
            while (ss >> f) readMtl((std::filesystem::path(mtlBaseDir) / f).string(), out.materials, matByName);
        }
    }
}

namespace {
struct VertexKey {
    b200pt_vertex v;
    bool operator==(const VertexKey &o) const {   // src/Model.h:30-32 (float compare: -0 == 0)
        for (int a = 0; a < 3; a++) if (!(v.pos[a] == o.v.pos[a]) || !(v.normal[a] == o.v.normal[a])) return false;
        return v.texCoord[0] == o.v.texCoord[0] && v.texCoord[1] == o.v.texCoord[1] && v.materialIndex == o.v.materialIndex;
    }
};
struct VertexHash {
    size_t operator()(const VertexKey &k) const {
        size_t h = 17;
        auto mix = [&](float f) { if (f == 0) f = 0; uint32_t u; memcpy(&u, &f, 4); h = h * 31 + std::hash<uint32_t>()(u); };
        for (int a = 0; a < 3; a++) { mix(k.v.pos[a]); mix(k.v.normal[a]); }
        mix(k.v.texCoord[0]); mix(k.v.texCoord[1]);
        return h * 31 + std::hash<int>()(k.v.materialIndex);
    }
};
}  // namespace

void convertObjData(const ObjData &obj, const std::vector<b200pt_material> &materials, int materialIndexOffset,
                    int materialIndexOverride, std::vector<b200pt_vertex> &outVertices,
                    std::vector<uint32_t> &outIndices, std::vector<int> &outEmissiveFaces) {
    bool hasNormals = !obj.normals.empty(), hasTex = !obj.texcoords.empty();
    std::unordered_map<VertexKey, uint32_t, VertexHash> unique;
    size_t numFaces = obj.faceIndices.size() / 3;
    for (size_t iFace = 0; iFace < numFaces; ++iFace) {
        int materialIndex = materialIndexOverride;
        if (materialIndexOverride < 0) materialIndex = materialIndexOffset + obj.faceMaterial[iFace];
        b200pt_vertex fv[3];
        for (int i = 0; i < 3; i++) {
            const ObjIndex &idx = obj.faceIndices[3 * iFace + i];
            b200pt_vertex v;
            memset(&v, 0, sizeof(v));
            for (int a = 0; a < 3; a++) v.pos[a] = obj.positions[3 * size_t(idx.v) + a];
            if (hasNormals && idx.vn >= 0) for (int a = 0; a < 3; a++) v.normal[a] = obj.normals[3 * size_t(idx.vn) + a];
            if (hasTex && idx.vt >= 0) {
                v.texCoord[0] = obj.texcoords[2 * size_t(idx.vt) + 0];
                v.texCoord[1] = 1.0f - obj.texcoords[2 * size_t(idx.vt) + 1];   // src/SceneLoader.cpp:288-291
            }
            v.materialIndex = materialIndex;
            fv[i] = v;
        }
        if (!hasNormals) {   // src/SceneLoader.cpp:299-308
            float ab[3], ac[3];
            for (int a = 0; a < 3; a++) { ab[a] = fv[1].pos[a] - fv[0].pos[a]; ac[a] = fv[2].pos[a] - fv[0].pos[a]; }
            float n[3] = {ab[1] * ac[2] - ac[1] * ab[2], ab[2] * ac[0] - ac[2] * ab[0], ab[0] * ac[1] - ac[0] * ab[1]};
            float il = 1.0f / sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            for (auto &v : fv) for (int a = 0; a < 3; a++) v.normal[a] = n[a] * il;
        }
        for (auto &v : fv) {
            VertexKey key{v};
            auto it = unique.find(key);
            if (it == unique.end()) { it = unique.emplace(key, uint32_t(outVertices.size())).first; outVertices.push_back(v); }
            outIndices.push_back(it->second);
        }
        if (materialIndex >= 0 && materialIndex < int(materials.size()) && materials[materialIndex].type == B200PT_MAT_LIGHT)
            outEmissiveFaces.push_back(int((outIndices.size() - 1) / 3));
    }
}

// ------------------------------------------------------------------------------------------------ mini XML
namespace {
struct XmlNode {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<XmlNode>> children;
    const char *attr(const char *k) const { for (auto &a : attrs) if (a.first == k) return a.second.c_str(); return nullptr; }
    const XmlNode *child(const char *n) const { for (auto &c : children) if (c->name == n) return c.get(); return nullptr; }
    std::vector<const XmlNode *> childrenNamed(const char *n) const {
        std::vector<const XmlNode *> r;
        for (auto &c : children) if (c->name == n) r.push_back(c.get());
        return r;
    }
};

struct XmlParser {
    const std::string &s; size_t i = 0;
    explicit XmlParser(const std::string &src) : s(src) {}
    void skipWs() { while (i < s.size() && isspace((unsigned char)s[i])) i++; }
    bool starts(const char *t) const { return s.compare(i, strlen(t), t) == 0; }
    void skipMisc() {
        for (;;) {
            skipWs();
            if (starts("<!--")) { size_t e = s.find("-->", i); if (e == std::string::npos) throw std::runtime_error("XML: unterminated comment"); i = e + 3; }
            else if (starts("<?")) { size_t e = s.find("?>", i); if (e == std::string::npos) throw std::runtime_error("XML: unterminated PI"); i = e + 2; }
            else if (starts("<!")) { size_t e = s.find('>', i); if (e == std::string::npos) throw std::runtime_error("XML: unterminated decl"); i = e + 1; }
            else return;
        }
    }
    static std::string unescape(const std::string &v) {
        std::string r;
        for (size_t k = 0; k < v.size(); k++) {
            if (v[k] == '&') {
                if (!v.compare(k, 4, "&lt;")) { r += '<'; k += 3; continue; }
                if (!v.compare(k, 4, "&gt;")) { r += '>'; k += 3; continue; }
                if (!v.compare(k, 5, "&amp;")) { r += '&'; k += 4; continue; }
                if (!v.compare(k, 6, "&quot;")) { r += '"'; k += 5; continue; }
                if (!v.compare(k, 6, "&apos;")) { r += '\''; k += 5; continue; }
            }
            r += v[k];
        }
        return r;
    }
    int depth = 0;
    std::unique_ptr<XmlNode> parseElement() {
        struct Depth { int &d; explicit Depth(int &x) : d(x) { if (++d > 256) throw std::runtime_error("XML: elements nested too deeply"); } ~Depth() { --d; } } guard(depth);
        skipMisc();
        if (i >= s.size() || s[i] != '<') throw std::runtime_error("XML: expected element");
        i++;
        auto node = std::make_unique<XmlNode>();
        size_t b = i;
        while (i < s.size() && !isspace((unsigned char)s[i]) && s[i] != '>' && s[i] != '/') i++;
        node->name = s.substr(b, i - b);
        for (;;) {
            skipWs();
            if (i >= s.size()) throw std::runtime_error("XML: unexpected end");
            if (s[i] == '/') { if (i + 1 < s.size() && s[i + 1] == '>') { i += 2; return node; } throw std::runtime_error("XML: bad tag"); }
            if (s[i] == '>') { i++; break; }
            size_t kb = i;
            while (i < s.size() && s[i] != '=' && !isspace((unsigned char)s[i])) i++;
            std::string key = s.substr(kb, i - kb);
            skipWs();
            if (i >= s.size() || s[i] != '=') throw std::runtime_error("XML: attribute without value");
            i++; skipWs();
            char q = s[i];
            if (q != '"' && q != '\'') throw std::runtime_error("XML: unquoted attribute");
            i++;
            size_t vb = i;
            while (i < s.size() && s[i] != q) i++;
            node->attrs.emplace_back(key, unescape(s.substr(vb, i - vb)));
            i++;
        }
        for (;;) {   // content
            size_t lt = s.find('<', i);
            if (lt == std::string::npos) throw std::runtime_error("XML: missing close tag for " + node->name);
            i = lt;
            if (starts("</")) {
                size_t e = s.find('>', i);
                if (e == std::string::npos) throw std::runtime_error("XML: unterminated close tag of " + node->name);     // (a truncated file; i = npos + 1 = 0 would parse it again, for ever)
                i = e + 1;
                return node;
            }
            if (starts("<!--") || starts("<?") || starts("<!")) { skipMisc(); continue; }
            node->children.push_back(parseElement());
        }
    }
};

// src/MitsubaXML.h helpers
const XmlNode *namedChild(const XmlNode *e, const std::string &name, const char *filter) {   // :15-28
    for (auto c : e->childrenNamed(filter)) { const char *n = c->attr("name"); if (n && name == n) return c; }
    return nullptr;
}
void parseVec3(const std::string &text, float out[3]) {   // parseCommaSpaceSeparatedVec3, :30-49
    std::stringstream ss(text);
    std::vector<float> values;
    for (float f; ss >> f;) { values.push_back(f); if (ss.peek() == ',' || ss.peek() == ' ') ss.ignore(); }
    if (values.size() == 1) { out[0] = out[1] = out[2] = values[0]; return; }
    if (values.size() != 3) throw std::runtime_error("Parsed text did not contain 3 values");
    out[0] = values[0]; out[1] = values[1]; out[2] = values[2];
}
float attrFloat(const XmlNode *n, const char *k, float def = 0.0f) { const char *v = n->attr(k); return v ? strtof(v, nullptr) : def; }
float childFloat(const XmlNode *e, const std::string &name) {
    const XmlNode *c = namedChild(e, name, "float");
    if (!c) throw std::runtime_error("Name not found: " + name);
    return attrFloat(c, "value");
}
float childSingleSpectrum(const XmlNode *e, const std::string &name) {
    const XmlNode *c = namedChild(e, name, "spectrum");
    if (!c) throw std::runtime_error("Name not found: " + name);
    return attrFloat(c, "value");
}
std::string childString(const XmlNode *e, const std::string &name) {
    const XmlNode *c = namedChild(e, name, "string");
    if (!c) throw std::runtime_error("Name not found: " + name);
    const char *v = c->attr("value");
    return v ? v : "";
}
void childRGB(const XmlNode *e, const std::string &name, float out[3]) {
    const XmlNode *c = namedChild(e, name, "rgb");
    if (!c) throw std::runtime_error("Name not found: " + name);
    parseVec3(c->attr("value") ? c->attr("value") : "", out);
}

b200pt_material blankMaterial() {
    b200pt_material m;
    memset(&m, 0, sizeof(m));
    m.textureIdDiffuse = -1;
    m.textureIdSpecular = -1;
    return m;
}

// src/SceneLoader.cpp:558-611 parseXmlBSDF.  The reference leaves unset fields uninitialised; we zero them.
b200pt_material parseXmlBSDF(const XmlNode *x, std::string &outId, std::map<std::string, int> &definedTextures) {
    if (const char *id = x->attr("id")) outId = id;
    std::string type = x->attr("type") ? x->attr("type") : "";
    b200pt_material mat = blankMaterial();
    if (type == "phong") {
        mat.type = B200PT_MAT_PHONG;
        childRGB(x, "specularReflectance", mat.specular);
        childRGB(x, "diffuseReflectance", mat.diffuse);
        mat.specularHighlight = childFloat(x, "exponent");
    } else if (type == "diffuse") {
        mat.type = B200PT_MAT_DIFFUSE;
        const XmlNode *ref = x->child("ref");
        if (ref && ref->attr("name") && std::string("reflectance") == ref->attr("name")) {
            mat.diffuse[0] = mat.diffuse[1] = mat.diffuse[2] = 1;
            mat.textureIdDiffuse = definedTextures[ref->attr("id") ? ref->attr("id") : ""];
        } else childRGB(x, "reflectance", mat.diffuse);
    } else if (type == "dielectric") {
        mat.type = B200PT_MAT_DIELECTRIC;
        mat.specular[0] = mat.specular[1] = mat.specular[2] = 1;
        mat.refractionIndex = childFloat(x, "intIOR") / childFloat(x, "extIOR");
        mat.refractionIndexInv = 1.0f / mat.refractionIndex;
    } else if (type == "conductor") {
        if (namedChild(x, "material", "string") && childString(x, "material") == "none") {
            mat.type = B200PT_MAT_SPECULAR;
            mat.specular[0] = mat.specular[1] = mat.specular[2] = 1;
        } else {
            mat.type = B200PT_MAT_CONDUCTOR;
            mat.eta = childSingleSpectrum(x, "eta");
            mat.k = childSingleSpectrum(x, "k");
        }
    } else if (type == "roughconductor") {
        mat.type = B200PT_MAT_ROUGH_CONDUCTOR;
        mat.roughness = childFloat(x, "alpha");
        mat.eta = childSingleSpectrum(x, "eta");
        mat.k = childSingleSpectrum(x, "k");
    } else {
        fprintf(stderr, "Encountered unknown material type: %s\n", type.c_str());
    }
    return mat;
}

// ---------------------------------------------------------------------------------------------- mini JSON
struct JsonValue {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    double num = 0; bool b = false; std::string str;
    std::vector<JsonValue> arr;
    std::vector<std::pair<std::string, JsonValue>> obj;   // insertion order kept
    const JsonValue *get(const std::string &k) const { for (auto &kv : obj) if (kv.first == k) return &kv.second; return nullptr; }
};
struct JsonParser {
    const std::string &s; size_t i = 0;
    explicit JsonParser(const std::string &src) : s(src) {}
    void ws() { while (i < s.size() && isspace((unsigned char)s[i])) i++; }
    JsonValue parse() {
        ws();
        JsonValue v;
        if (i >= s.size()) throw std::runtime_error("JSON: unexpected end");
        char c = s[i];
        if (c == '{') {
            v.kind = JsonValue::Obj; i++; ws();
            if (s[i] == '}') { i++; return v; }
            for (;;) {
                ws(); JsonValue k = parse(); ws();
                if (k.kind != JsonValue::Str || s[i] != ':') throw std::runtime_error("JSON: bad object");
                i++;
                v.obj.emplace_back(k.str, parse()); ws();
                if (s[i] == ',') { i++; continue; }
                if (s[i] == '}') { i++; return v; }
                throw std::runtime_error("JSON: bad object");
            }
        } else if (c == '[') {
            v.kind = JsonValue::Arr; i++; ws();
            if (s[i] == ']') { i++; return v; }
            for (;;) {
                v.arr.push_back(parse()); ws();
                if (s[i] == ',') { i++; continue; }
                if (s[i] == ']') { i++; return v; }
                throw std::runtime_error("JSON: bad array");
            }
        } else if (c == '"') {
            v.kind = JsonValue::Str; i++;
            while (i < s.size() && s[i] != '"') { if (s[i] == '\\' && i + 1 < s.size()) i++; v.str += s[i++]; }
            i++;
            return v;
        } else if (!s.compare(i, 4, "true")) { v.kind = JsonValue::Bool; v.b = true; i += 4; return v; }
        else if (!s.compare(i, 5, "false")) { v.kind = JsonValue::Bool; i += 5; return v; }
        else if (!s.compare(i, 4, "null")) { i += 4; return v; }
        char *e;
        v.kind = JsonValue::Num;
        v.num = strtod(s.c_str() + i, &e);
        if (e == s.c_str() + i) throw std::runtime_error("JSON: bad token");
        i = e - s.c_str();
        return v;
    }
};

std::string readTextFile(const std::string &path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("Cannot open " + path);
    std::stringstream ss;
    ss << in.rdbuf();
    return ss.str();
}
}  // namespace

// ------------------------------------------------------------------------------------------------ Scene
int Scene::addTexture(const std::string &name) {   // src/SceneLoader.cpp:184-236
    std::filesystem::path path = std::filesystem::path(textureBaseDir) / name;
    auto it = pathTextureId.find(path.string());
    if (it != pathTextureId.end()) return it->second;
    std::string ps = path.string();
    size_t bs = ps.find('\\');
    if (bs != std::string::npos) ps.replace(bs, 1, "/");
    path = std::filesystem::path(ps).lexically_normal();
    TextureImage t;
    t.format = B200PT_TEX_RGBA8_SRGB;       // vk::Format::eR8G8B8A8Srgb, 4 channels forced (stbi_load(..., 4))
    t.path = path.string();
    decodeImageFile(t.path, t.width, t.height, t.rgba8);
    textures.push_back(std::move(t));
    const int textureIndex = int(textures.size()) - 1;
    pathTextureId[path.string()] = textureIndex;      // sic: keyed by the normalised path, looked up by the raw one (:189,233)
    return textureIndex;
}

void Scene::parseMitsubaSceneFile(const std::string &filepath) {
    std::string src = readTextFile(filepath);
    XmlParser parser(src);
    std::unique_ptr<XmlNode> root = parser.parseElement();
    if (root->name != "scene") throw std::runtime_error("XML: no <scene> element");
    const XmlNode *xScene = root.get();

    // parseCameraSettings, src/SceneLoader.cpp:613-648
    if (const XmlNode *xSensor = xScene->child("sensor")) {
        if (const XmlNode *xTransform = xSensor->child("transform")) {
            const XmlNode *xLookAt = xTransform->child("lookAt");
            if (!xLookAt) xLookAt = xTransform->child("lookat");
            if (xLookAt) {
                if (!xLookAt->attr("origin") || !xLookAt->attr("target") || !xLookAt->attr("up"))
                    throw std::runtime_error("lookAt needs origin, target and up");
                parseVec3(xLookAt->attr("origin"), origin);
                parseVec3(xLookAt->attr("target"), target);
                parseVec3(xLookAt->attr("up"), upDir);
            }
            if (const XmlNode *xFov = namedChild(xSensor, "fov", "float")) vfov = attrFloat(xFov, "value");
        }
    }

    // parseXmlTextures, :516-534
    std::map<std::string, int> definedTextures;
    for (auto xTexture : xScene->childrenNamed("texture")) {
        std::string id = xTexture->attr("id") ? xTexture->attr("id") : "";
        if (definedTextures.count(id)) throw std::runtime_error("Duplicate texture id");
        definedTextures[id] = addTexture(childString(xTexture, "filename"));
    }

    // parseXmlBSDFs, :536-556
    std::map<std::string, int> definedMaterials;
    for (auto xBSDF : xScene->childrenNamed("bsdf")) {
        std::string id;
        b200pt_material mat = parseXmlBSDF(xBSDF, id, definedTextures);
        if (definedMaterials.count(id)) throw std::runtime_error("Duplicate BSDF id");
        definedMaterials[id] = int(materials.size());
        materials.push_back(mat);
    }

    // parseXmlShapes, :405-514
    for (auto xShape : xScene->childrenNamed("shape")) {
        std::string type = xShape->attr("type") ? xShape->attr("type") : "";
        int matIndex = -1;
        if (const XmlNode *xBSDF = xShape->child("bsdf")) {
            std::string tmp;
            b200pt_material mat = parseXmlBSDF(xBSDF, tmp, definedMaterials);   // sic: the reference passes definedMaterials
            matIndex = int(materials.size());
            materials.push_back(mat);
        }
        if (matIndex < 0) {
            if (const XmlNode *xRef = xShape->child("ref")) {
                std::string id = xRef->attr("id") ? xRef->attr("id") : "";
                matIndex = definedMaterials[id];
            }
        }
        if (matIndex < 0) throw std::runtime_error("Material index not set");
        if (const XmlNode *xEmitter = xShape->child("emitter")) {
            float radiance[3];
            childRGB(xEmitter, "radiance", radiance);
            b200pt_material mat = materials[matIndex];
            memcpy(mat.lightColor, radiance, sizeof(radiance));
            mat.type = B200PT_MAT_LIGHT;
            materials.push_back(mat);
            matIndex = int(materials.size()) - 1;
        }
        if (type == "obj") {
            std::string filename = (std::filesystem::path(modelsBaseDir) / childString(xShape, "filename")).string();
            ObjData obj;
            readObj(filename, materialBaseDir, obj);
            Model model;
            std::vector<int> emissiveFaces;
            convertObjData(obj, materials, -1, matIndex, model.vertices, model.indices, emissiveFaces);
            models.push_back(std::move(model));
            emissiveFacesPerModel.push_back(emissiveFaces);
            b200pt_instance inst;
            memset(&inst, 0, sizeof(inst));
            mat4Identity(inst.transform);
            mat4Identity(inst.normalTransform);
            inst.modelIndex = int(models.size()) - 1;
            inst.iLight = -1;
            if (!emissiveFaces.empty()) {
                b200pt_light light;
                memset(&light, 0, sizeof(light));
                light.type = B200PT_LIGHT_AREA;
                light.instanceIndex = uint32_t(instances.size());
                lights.push_back(light);
                inst.iLight = int(lights.size()) - 1;
            }
            instances.push_back(inst);
        } else if (type == "sphere") {
            b200pt_sphere sphere;
            sphere.radius = childFloat(xShape, "radius");
            const XmlNode *c = namedChild(xShape, "center", "point");
            if (!c) throw std::runtime_error("Name not found: center");
            sphere.center[0] = attrFloat(c, "x"); sphere.center[1] = attrFloat(c, "y"); sphere.center[2] = attrFloat(c, "z");
            if (!(std::isfinite(sphere.radius) && std::isfinite(sphere.center[0]) && std::isfinite(sphere.center[1]) && std::isfinite(sphere.center[2])))
                throw std::runtime_error("sphere with a non-finite radius or centre");      // (would make the scene box, hence every guiding region, infinite)
            sphere.materialIndex = matIndex;
            sphere.iLight = -1;
            if (materials[matIndex].type == B200PT_MAT_LIGHT) {
                b200pt_light light;
                memset(&light, 0, sizeof(light));
                light.type = B200PT_LIGHT_SPHERE;
                light.instanceIndex = uint32_t(spheres.size());
                lights.push_back(light);
                sphere.iLight = int(lights.size()) - 1;
            }
            spheres.push_back(sphere);
        }
    }

    // parseEnvMap, :370-403 — only the FIRST top-level <emitter> is looked at
    if (const XmlNode *xEnv = xScene->child("emitter")) {
        if (xEnv->attr("type") && std::string("envmap") == xEnv->attr("type")) {
            std::string envPath = (std::filesystem::path(textureBaseDir) / childString(xEnv, "filename")).string();
            TextureImage &t = textures[0];
            readExrRGBA(envPath, t.rgba32f, t.width, t.height, /*viaHalf=*/true);   // Imf::Rgba is half, CommonOps.cpp:40-65
            t.format = B200PT_TEX_RGBA32F;
            t.path = envPath;
            b200pt_light light;
            memset(&light, 0, sizeof(light));
            light.type = B200PT_LIGHT_ENV_MAP;
            lights.push_back(light);
        }
    }
}

void Scene::parseJsonSceneFile(const std::string &filepath) {   // src/SceneLoader.cpp:650-770
    std::string src = readTextFile(filepath);
    JsonParser jp(src);
    JsonValue j = jp.parse();
    std::map<std::string, int> nameIndex;
    const JsonValue *jModels = j.get("models");
    if (!jModels) throw std::runtime_error("JSON scene: no models");
    for (auto &jm : jModels->arr) {   // parseModels :760-770 → loadModel :114-126
        if (jm.obj.empty()) continue;
        const std::string &name = jm.obj[0].first;
        std::string path = (std::filesystem::path(modelsBaseDir) / jm.obj[0].second.str).string();
        ObjData obj;
        readObj(path, materialBaseDir, obj);
        int materialIndexOffset = int(materials.size());
        for (auto &tm : obj.materials) {   // addMaterials :139-182
            b200pt_material m = blankMaterial();
            memcpy(m.lightColor, tm.emission, 12); memcpy(m.diffuse, tm.diffuse, 12); memcpy(m.specular, tm.specular, 12);
            m.specularHighlight = tm.shininess;
            m.refractionIndex = tm.ior;
            m.refractionIndexInv = 1.0f / tm.ior;
            switch (tm.illum) {
                case 0: case 1: m.type = B200PT_MAT_DIFFUSE; break;
                case 2: m.type = B200PT_MAT_PHONG; break;
                case 3: m.type = B200PT_MAT_SPECULAR; break;
                case 4: case 7: m.type = B200PT_MAT_DIELECTRIC; break;
                case 11: m.type = B200PT_MAT_LIGHT; break;
                default: throw std::runtime_error("Unknown illum mode");
            }
            if (!tm.diffuseTex.empty()) m.textureIdDiffuse = addTexture(tm.diffuseTex);
            if (!tm.specularTex.empty()) m.textureIdDiffuse = addTexture(tm.specularTex);   // sic, :176-178
            materials.push_back(m);
        }
        Model model;
        std::vector<int> emissive;
        convertObjData(obj, materials, materialIndexOffset, -1, model.vertices, model.indices, emissive);
        models.push_back(std::move(model));
        emissiveFacesPerModel.push_back(emissive);
        nameIndex[name] = int(models.size()) - 1;
    }
    auto vec3Of = [](const JsonValue &v, float out[3]) { for (int a = 0; a < 3; a++) out[a] = float(v.arr.at(a).num); };
    if (const JsonValue *jInst = j.get("instances")) {   // parseInstances :704-758
        for (auto &ji : jInst->arr) {
            if (ji.obj.empty()) continue;
            const std::string &name = ji.obj[0].first;
            const JsonValue &props = ji.obj[0].second;
            b200pt_instance inst;
            memset(&inst, 0, sizeof(inst));
            inst.modelIndex = nameIndex[name];
            inst.iLight = -1;
            if (!emissiveFacesPerModel[inst.modelIndex].empty()) {
                b200pt_light light;
                memset(&light, 0, sizeof(light));
                light.type = B200PT_LIGHT_AREA;
                light.instanceIndex = uint32_t(instances.size());
                lights.push_back(light);
                inst.iLight = int(lights.size()) - 1;
            }
            float T[16], N[16];
            mat4Identity(T); mat4Identity(N);
            const float rad = 0.01745329251994329576923690768489f;
            if (const JsonValue *t = props.get("translate")) { float v[3]; vec3Of(*t, v); mat4Translate(T, v[0], v[1], v[2]); }
            if (const JsonValue *r = props.get("rotate")) {
                float v[3]; vec3Of(*r, v);
                mat4Rotate(T, v[0] * rad, 1, 0, 0); mat4Rotate(T, v[1] * rad, 0, 1, 0); mat4Rotate(T, v[2] * rad, 0, 0, 1);
                mat4Rotate(N, v[0] * rad, 1, 0, 0); mat4Rotate(N, v[1] * rad, 0, 1, 0); mat4Rotate(N, v[2] * rad, 0, 0, 1);
            }
            if (const JsonValue *s = props.get("scale")) {
                float v[3]; vec3Of(*s, v);
                mat4Scale(T, v[0], v[1], v[2]);
                mat4Scale(N, 1.0f / v[0], 1.0f / v[1], 1.0f / v[2]);
            }
            memcpy(inst.transform, T, sizeof(T));
            memcpy(inst.normalTransform, N, sizeof(N));
            instances.push_back(inst);
        }
    }
    if (const JsonValue *jl = j.get("lights")) {   // parsePointLights :686-702
        for (auto &l : jl->arr) {
            b200pt_light light;
            memset(&light, 0, sizeof(light));
            light.type = B200PT_LIGHT_POINT;
            if (const JsonValue *c = l.get("color")) vec3Of(*c, light.color);
            if (const JsonValue *p = l.get("position")) vec3Of(*p, light.pos);
            lights.push_back(light);
        }
    }
    if (const JsonValue *cam = j.get("camera")) {   // parseJsonCamera :668-680
        if (const JsonValue *v = cam->get("target")) vec3Of(*v, target);
        if (const JsonValue *v = cam->get("origin")) vec3Of(*v, origin);
        if (const JsonValue *v = cam->get("up")) vec3Of(*v, upDir);
    }
}

void Scene::calculateSceneSize() {   // src/SceneLoader.cpp:350-368
    const float inf = std::numeric_limits<float>::infinity();
    for (int a = 0; a < 3; a++) { sceneMin[a] = inf; sceneMax[a] = -inf; }
    for (const auto &inst : instances) {
        float mn[3], mx[3];
        models[inst.modelIndex].aabb(inst.transform, mn, mx);
        for (int a = 0; a < 3; a++) {   // Aabb::update is applied to min and to max, src/Shapes.h:55-62
            sceneMin[a] = fminf(fminf(sceneMin[a], mn[a]), mx[a]);
            sceneMax[a] = fmaxf(fmaxf(sceneMax[a], mn[a]), mx[a]);
        }
    }
    for (const auto &s : spheres)
        for (int a = 0; a < 3; a++) {
            sceneMin[a] = fminf(sceneMin[a], s.center[a] - s.radius);
            sceneMax[a] = fmaxf(sceneMax[a], s.center[a] + s.radius);
        }
}

void Scene::buildLightTables() {
    // getLightSamplingVector, src/SceneLoader.cpp:923-944 (every light has "power" 1.0, quirk 6)
    {
        std::vector<float> powers(lights.size(), 1.0f);
        WeightedSampler lightSampler(powers);
        std::vector<float> prob = lightSampler.probabilities();
        for (size_t i = 0; i < lights.size(); i++) lights[i].sampleProb = prob[i];
        randomLightIndex.resize(B200PT_SIZE_LIGHT_RANDOM);
        for (int i = 0; i < B200PT_SIZE_LIGHT_RANDOM; i++) randomLightIndex[i] = lightSampler.sample();
    }
    // getFaceSamplingVector, :861-921 — tables are appended for mesh lights only (quirk 5)
    randomTriIndex.clear();
    numFaceTables = 0;
    for (auto &light : lights) {
        if (light.type == B200PT_LIGHT_POINT || light.type == B200PT_LIGHT_ENV_MAP) continue;
        if (light.type == B200PT_LIGHT_SPHERE) {
            float radius = spheres[light.instanceIndex].radius;
            light.area = 4 * 3.14159265358979323846f * radius * radius;
            continue;
        }
        const b200pt_instance &inst = instances[light.instanceIndex];
        const Model &model = models[inst.modelIndex];
        const std::vector<int> &ef = emissiveFacesPerModel[inst.modelIndex];
        std::vector<float> areas(ef.size());
        for (size_t i = 0; i < ef.size(); i++) areas[i] = model.faceArea(ef[i], inst.transform);
        WeightedSampler faceSampler(areas);
        light.area = faceSampler.total();
        std::vector<float> prob = faceSampler.probabilities();
        for (int i = 0; i < B200PT_SIZE_TRI_RANDOM; i++) {
            int s = faceSampler.sample();
            b200pt_face_sample fs;
            fs.index = ef[s];
            fs.sampleProb = prob[s];
            fs.faceArea = areas[s];
            randomTriIndex.push_back(fs);
        }
        numFaceTables++;
    }
    if (randomTriIndex.empty()) {   // dummy table, :909-918
        b200pt_face_sample z;
        memset(&z, 0, sizeof(z));
        randomTriIndex.assign(B200PT_SIZE_TRI_RANDOM, z);
        numFaceTables = 1;
    }
}

void Scene::loadFile(const std::string &path_) {   // src/SceneLoader.cpp:29-65 (without SCENE_BASE_DIR: the CLI adds it)
    std::filesystem::path path(path_);
    std::filesystem::path directory = path.parent_path();
    std::string ext = path.extension().string();
    if (ext == ".xml") { modelsBaseDir = materialBaseDir = textureBaseDir = directory.string(); }
    else { modelsBaseDir = (directory / "models/").string(); materialBaseDir = (directory / "materials/").string(); textureBaseDir = (directory / "textures/").string(); }
    textures.clear();
    textures.emplace_back();   // slot 0 reserved for the env map, :54
    if (!std::filesystem::exists(path)) throw std::runtime_error("Scene file not found: " + path_);
    if (ext == ".xml") parseMitsubaSceneFile(path.string());
    else if (ext == ".json") parseJsonSceneFile(path.string());
    else throw std::runtime_error("Unknown scene file type: " + path_);
    finalize();
}

void Scene::finalize() {
    calculateSceneSize();
    buildLightTables();
    if (materials.empty()) materials.push_back(blankMaterial());   // :77-79
    if (textures.empty()) textures.emplace_back();
    if (textures[0].width == 0) {   // generateDefaultTexture, :89-112: 1x1 RGBA8 sRGB (0,0,0,0)
        textures[0].width = textures[0].height = 1;
        textures[0].format = B200PT_TEX_RGBA8_SRGB;
        textures[0].rgba8.assign(4, 0);
    }
    vertexPtrs.clear(); indexPtrs.clear(); numVertices.clear(); numIndices.clear(); textureDescs.clear();
    for (auto &m : models) {
        vertexPtrs.push_back(m.vertices.data()); indexPtrs.push_back(m.indices.data());
        numVertices.push_back(int(m.vertices.size())); numIndices.push_back(int(m.indices.size()));
    }
    for (auto &t : textures) {
        b200pt_texture d;
        d.width = t.width; d.height = t.height; d.format = t.format; d._pad = 0;
        d.pixels = t.format == B200PT_TEX_RGBA32F ? (const void *)t.rgba32f.data() : (const void *)t.rgba8.data();
        textureDescs.push_back(d);
    }
}

void Scene::fillDesc(b200pt_scene_desc *d) const {
    memset(d, 0, sizeof(*d));
    d->num_models = int(models.size());
    d->vertices = vertexPtrs.data(); d->num_vertices = numVertices.data();
    d->indices = indexPtrs.data(); d->num_indices = numIndices.data();
    d->num_materials = int(materials.size()); d->materials = materials.data();
    d->num_instances = int(instances.size()); d->instances = instances.data();
    d->num_lights = int(lights.size()); d->lights = lights.data();
    d->random_light_index = randomLightIndex.data();
    d->num_face_tables = numFaceTables; d->random_tri_index = randomTriIndex.data();
    d->num_spheres = int(spheres.size()); d->spheres = spheres.data();
    d->num_textures = int(textureDescs.size()); d->textures = textureDescs.data();
    memcpy(d->scene_min, sceneMin, 12); memcpy(d->scene_max, sceneMax, 12);
}

}  // namespace b200pt
