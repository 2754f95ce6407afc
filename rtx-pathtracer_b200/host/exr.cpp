// Minimal OpenEXR scanline IO for the headless host: replaces CommonOps::writeEXR / readEXR
// (src/CommonOps.cpp:12-65), which link OpenEXR 2.5.  Writer: three FLOAT channels, no compression.
// Reader: HALF/FLOAT channels, compression NONE / ZIPS / ZIP (what Mitsuba and the reference app write).
// PIZ (used by the scenes' envmap.exr files) is not implemented yet — reported as an error, never silently skipped.
#include "scene.h"
#include <cstring>
#include <cstdio>
#include <fstream>
#include <zlib.h>

namespace b200pt {

float halfToFloat(uint16_t h) {
    uint32_t s = (h >> 15) & 1, e = (h >> 10) & 0x1f, m = h & 0x3ff, out;
    if (e == 0) {
        if (m == 0) out = s << 31;
        else {
            e = 127 - 15 + 1;
            while (!(m & 0x400)) { m <<= 1; e--; }
            m &= 0x3ff;
            out = (s << 31) | (e << 23) | (m << 13);
        }
    } else if (e == 31) out = (s << 31) | 0x7f800000u | (m << 13);
    else out = (s << 31) | ((e + 127 - 15) << 23) | (m << 13);
    float f;
    memcpy(&f, &out, 4);
    return f;
}

uint16_t floatToHalf(float f) {   // round to nearest even
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t s = (x >> 16) & 0x8000;
    int32_t e = int32_t((x >> 23) & 0xff) - 127 + 15;
    uint32_t m = x & 0x7fffff;
    if (((x >> 23) & 0xff) == 0xff) return uint16_t(s | 0x7c00 | (m ? 0x200 : 0));
    if (e >= 31) return uint16_t(s | 0x7c00);
    if (e <= 0) {
        if (e < -10) return uint16_t(s);
        m |= 0x800000;
        uint32_t shift = uint32_t(14 - e);
        uint32_t r = m >> shift, rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (r & 1))) r++;
        return uint16_t(s | r);
    }
    uint32_t r = (uint32_t(e) << 10) | (m >> 13), rem = m & 0x1fff;
    if (rem > 0x1000 || (rem == 0x1000 && (r & 1))) r++;
    return uint16_t(s | r);
}

namespace {
void putStr(std::string &b, const char *s) { b.append(s); b.push_back('\0'); }
void putI32(std::string &b, int32_t v) { b.append(reinterpret_cast<const char *>(&v), 4); }
void putF32(std::string &b, float v) { b.append(reinterpret_cast<const char *>(&v), 4); }
void putAttr(std::string &b, const char *name, const char *type, const std::string &payload) {
    putStr(b, name); putStr(b, type); putI32(b, int32_t(payload.size())); b.append(payload);
}
}  // namespace

void writeExrRGB(const std::string &path, const float *rgba, int width, int height) {
    std::string h;
    putI32(h, 20000630);   // magic
    putI32(h, 2);          // version 2, scanline, single part
    {
        std::string ch;
        for (const char *c : {"B", "G", "R"}) {
            putStr(ch, c); putI32(ch, 2 /*FLOAT*/); ch.append(4, '\0'); putI32(ch, 1); putI32(ch, 1);
        }
        ch.push_back('\0');
        putAttr(h, "channels", "chlist", ch);
    }
    putAttr(h, "compression", "compression", std::string(1, '\0'));
    {
        std::string w;
        putI32(w, 0); putI32(w, 0); putI32(w, width - 1); putI32(w, height - 1);
        putAttr(h, "dataWindow", "box2i", w);
        putAttr(h, "displayWindow", "box2i", w);
    }
    putAttr(h, "lineOrder", "lineOrder", std::string(1, '\0'));
    { std::string v; putF32(v, 1.0f); putAttr(h, "pixelAspectRatio", "float", v); }
    { std::string v; putF32(v, 0.0f); putF32(v, 0.0f); putAttr(h, "screenWindowCenter", "v2f", v); }
    { std::string v; putF32(v, 1.0f); putAttr(h, "screenWindowWidth", "float", v); }
    h.push_back('\0');

    size_t rowBytes = size_t(width) * 4 * 3;
    uint64_t offset = h.size() + uint64_t(height) * 8;
    std::ofstream out(path, std::ios::binary);
    if (!out) throw std::runtime_error("Cannot write " + path);
    out.write(h.data(), std::streamsize(h.size()));
    for (int y = 0; y < height; y++) { uint64_t o = offset + uint64_t(y) * (8 + rowBytes); out.write(reinterpret_cast<const char *>(&o), 8); }
    std::vector<float> row(size_t(width) * 3);
    for (int y = 0; y < height; y++) {
        int32_t yy = y, sz = int32_t(rowBytes);
        out.write(reinterpret_cast<const char *>(&yy), 4);
        out.write(reinterpret_cast<const char *>(&sz), 4);
        const int order[3] = {2, 1, 0};   // B, G, R
        for (int c = 0; c < 3; c++)
            for (int x = 0; x < width; x++) row[size_t(c) * width + x] = rgba[(size_t(y) * width + x) * 4 + order[c]];
        out.write(reinterpret_cast<const char *>(row.data()), std::streamsize(rowBytes));
    }
    if (!out) throw std::runtime_error("Write failed: " + path);
}

void readExrRGBA(const std::string &path, std::vector<float> &rgba, int &width, int &height, bool viaHalf) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("Cannot open EXR file " + path);
    std::string data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (data.size() < 8) throw std::runtime_error("EXR: truncated " + path);
    int32_t magic, version;
    memcpy(&magic, data.data(), 4); memcpy(&version, data.data() + 4, 4);
    if (magic != 20000630) throw std::runtime_error("EXR: bad magic " + path);
    if (version & 0x1a00) throw std::runtime_error("EXR: tiled / multipart / deep files are not supported: " + path);
    size_t i = 8;
    struct Chan { std::string name; int type; };
    std::vector<Chan> chans;
    int comp = -1, dw[4] = {0, 0, -1, -1}, lineOrder = 0;
    for (;;) {
        if (i >= data.size()) throw std::runtime_error("EXR: truncated header");
        if (data[i] == 0) { i++; break; }
        std::string name(data.c_str() + i); i += name.size() + 1;
        std::string type(data.c_str() + i); i += type.size() + 1;
        int32_t sz; memcpy(&sz, data.data() + i, 4); i += 4;
        if (name == "channels") {
            size_t j = i;
            while (data[j] != 0) {
                Chan c; c.name = std::string(data.c_str() + j); j += c.name.size() + 1;
                int32_t t; memcpy(&t, data.data() + j, 4); c.type = t; j += 16;
                chans.push_back(c);
            }
        } else if (name == "compression") comp = (unsigned char)data[i];
        else if (name == "dataWindow") memcpy(dw, data.data() + i, 16);
        else if (name == "lineOrder") lineOrder = (unsigned char)data[i];
        i += size_t(sz);
    }
    (void)lineOrder;
    width = dw[2] - dw[0] + 1; height = dw[3] - dw[1] + 1;
    if (width <= 0 || height <= 0 || chans.empty()) throw std::runtime_error("EXR: bad header " + path);
    int linesPerBlock;
    if (comp == 0 || comp == 2) linesPerBlock = 1;
    else if (comp == 3) linesPerBlock = 16;
    else throw std::runtime_error("EXR: compression type " + std::to_string(comp) + " not supported (only NONE/ZIPS/ZIP): " + path);
    size_t bytesPerLine = 0;
    for (auto &c : chans) {
        if (c.type != 1 && c.type != 2) throw std::runtime_error("EXR: only HALF/FLOAT channels supported");
        bytesPerLine += size_t(width) * (c.type == 1 ? 2 : 4);
    }
    int numBlocks = (height + linesPerBlock - 1) / linesPerBlock;
    rgba.assign(size_t(width) * height * 4, 0.0f);
    bool hasA = false;
    for (auto &c : chans) if (c.name == "A") hasA = true;
    if (!hasA) for (size_t p = 0; p < size_t(width) * height; p++) rgba[p * 4 + 3] = 1.0f;
    std::vector<unsigned char> raw, tmp;
    for (int b = 0; b < numBlocks; b++) {
        uint64_t off; memcpy(&off, data.data() + i + size_t(b) * 8, 8);
        if (off + 8 > data.size()) throw std::runtime_error("EXR: bad chunk offset");
        int32_t y0, sz; memcpy(&y0, data.data() + off, 4); memcpy(&sz, data.data() + off + 4, 4);
        int lines = std::min(linesPerBlock, dw[3] - y0 + 1);
        size_t expect = bytesPerLine * size_t(lines);
        const unsigned char *src = reinterpret_cast<const unsigned char *>(data.data() + off + 8);
        raw.resize(expect);
        if (comp == 0 || size_t(sz) == expect) memcpy(raw.data(), src, expect);
        else {
            tmp.resize(expect);
            uLongf dl = uLongf(expect);
            if (uncompress(tmp.data(), &dl, src, uLong(sz)) != Z_OK || dl != expect) throw std::runtime_error("EXR: zlib error");
            for (size_t k = 1; k < expect; k++) tmp[k] = (unsigned char)(int(tmp[k - 1]) + int(tmp[k]) - 128);
            size_t half = (expect + 1) / 2;
            for (size_t k = 0; k < expect; k++) raw[k] = (k & 1) ? tmp[half + k / 2] : tmp[k / 2];
        }
        const unsigned char *p = raw.data();
        for (int l = 0; l < lines; l++) {
            int y = y0 - dw[1] + l;
            for (auto &c : chans) {
                int slot = c.name == "R" ? 0 : c.name == "G" ? 1 : c.name == "B" ? 2 : c.name == "A" ? 3 : -1;
                for (int x = 0; x < width; x++) {
                    float v;
                    if (c.type == 1) { uint16_t hv; memcpy(&hv, p, 2); p += 2; v = halfToFloat(hv); }
                    else { memcpy(&v, p, 4); p += 4; if (viaHalf) v = halfToFloat(floatToHalf(v)); }
                    if (slot >= 0) rgba[(size_t(y) * width + x) * 4 + slot] = v;
                }
            }
        }
    }
}

}  // namespace b200pt
