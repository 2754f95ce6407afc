// Minimal OpenEXR scanline IO for the headless host: replaces CommonOps::writeEXR / readEXR
// (src/CommonOps.cpp:12-65), which link OpenEXR 2.5.  Writer: three FLOAT channels, no compression.
// Reader: HALF / FLOAT channels, compression NONE / RLE / ZIPS / ZIP / PIZ / PXR24 / B44 / B44A — everything OpenEXR's
// RgbaInputFile reads from a scanline file except DWAA / DWAB, which (like tiled, multipart and deep files) are reported as
// an error, never silently skipped.  Checked against OpenEXR's own decode (tests/golden/exr_digests.json).
#include "scene.h"
#include <algorithm>
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

// ---- PIZ decompression (OpenEXR compression type 4) -------------------------------------------------------------------
// Independent implementation of the published format: a chunk holds, for all channels of up to 32 scanlines, 16-bit
// words that were (1) mapped to a dense range through a bitmap of the values that occur, (2) transformed by a 2-D Haar
// style wavelet (14-bit or 16-bit modular variant, chosen by the largest mapped value) and (3) Huffman coded with a
// canonical code whose length table is itself packed (6-bit lengths, zero runs) and one symbol reserved for run lengths.
namespace {
const int kHufEncBits = 16, kHufEncSize = (1 << kHufEncBits) + 1;

struct PizBits {            // MSB-first bit reader
    const unsigned char *p, *end;
    uint64_t c = 0;
    int lc = 0;
    uint32_t get(int n) {
        while (lc < n) { c = (c << 8) | (p < end ? *p : 0); p++; lc += 8; }
        lc -= n;
        return uint32_t((c >> lc) & ((uint64_t(1) << n) - 1));
    }
};

// canonical Huffman code from code lengths: codes of equal length are consecutive, shorter codes have larger prefixes
void pizCanonicalCodes(std::vector<uint64_t> &hcode) {
    uint64_t n[59] = {0};
    for (uint64_t l : hcode) n[l]++;
    uint64_t c = 0;
    for (int i = 58; i > 0; i--) { const uint64_t nc = (c + n[i]) >> 1; n[i] = c; c = nc; }
    for (uint64_t &h : hcode) { const uint64_t l = h; if (l > 0) h = l | (n[l]++ << 6); }
}

void pizHufDecode(const unsigned char *src, size_t srcLen, uint16_t *out, size_t outLen) {
    if (srcLen < 20) throw std::runtime_error("EXR/PIZ: truncated Huffman block");
    uint32_t im, iM, nBits;
    memcpy(&im, src, 4); memcpy(&iM, src + 4, 4); memcpy(&nBits, src + 12, 4);
    if (im >= uint32_t(kHufEncSize) || iM >= uint32_t(kHufEncSize) || im > iM) throw std::runtime_error("EXR/PIZ: bad Huffman header");
    // packed code-length table: 6 bits per symbol; 59..62 = run of 2..5 zeros, 63 = run of (next 8 bits) + 6 zeros
    std::vector<uint64_t> hcode(size_t(kHufEncSize), 0);
    PizBits tb{src + 20, src + srcLen};
    for (uint32_t i = im; i <= iM; i++) {
        const uint32_t l = tb.get(6);
        hcode[i] = l;
        if (l == 63) {
            uint32_t zerun = tb.get(8) + 6;
            if (i + zerun > iM + 1) throw std::runtime_error("EXR/PIZ: bad code table");
            while (zerun--) hcode[i++] = 0;
            i--;
        } else if (l >= 59) {
            uint32_t zerun = l - 59 + 2;
            if (i + zerun > iM + 1) throw std::runtime_error("EXR/PIZ: bad code table");
            while (zerun--) hcode[i++] = 0;
            i--;
        }
    }
    const unsigned char *dataStart = tb.p - (tb.lc / 8);       // the table ends on a byte boundary of what was consumed
    pizCanonicalCodes(hcode);
    // decoding tables by length: first code and first symbol index per length (symbols ordered by code)
    struct Sym { uint64_t code; int len; uint32_t sym; };
    std::vector<Sym> syms;
    for (uint32_t i = im; i <= iM; i++) if (hcode[i] & 63) syms.push_back({hcode[i] >> 6, int(hcode[i] & 63), i});
    std::vector<std::vector<Sym>> byLen(59);
    for (const Sym &sy : syms) byLen[size_t(sy.len)].push_back(sy);
    for (auto &v : byLen) std::sort(v.begin(), v.end(), [](const Sym &a, const Sym &b) { return a.code < b.code; });
    PizBits db{dataStart, src + srcLen};
    const uint32_t rlc = iM;                                   // run-length symbol
    size_t o = 0;
    uint64_t bitsLeft = nBits;
    while (bitsLeft > 0 && o <= outLen) {
        uint64_t code = 0;
        int len = 0;
        uint32_t sym = 0xffffffffu;
        while (len < 58 && bitsLeft > 0) {
            code = (code << 1) | db.get(1); len++; bitsLeft--;
            const auto &v = byLen[size_t(len)];
            if (!v.empty() && code >= v.front().code && code <= v.back().code) { sym = v[size_t(code - v.front().code)].sym; break; }
        }
        if (sym == 0xffffffffu) break;                         // trailing padding bits
        if (sym == rlc) {
            if (bitsLeft < 8 || o == 0) throw std::runtime_error("EXR/PIZ: bad run");
            uint32_t run = db.get(8); bitsLeft -= 8;
            if (o + run > outLen) throw std::runtime_error("EXR/PIZ: run past the end");
            const uint16_t prev = out[o - 1];
            while (run--) out[o++] = prev;
        } else {
            if (o >= outLen) throw std::runtime_error("EXR/PIZ: too much data");
            out[o++] = uint16_t(sym);
        }
    }
    if (o != outLen) throw std::runtime_error("EXR/PIZ: Huffman data ends early");
}

inline void wdec14(uint16_t l, uint16_t h, uint16_t &a, uint16_t &b) {
    const int16_t ls = int16_t(l), hs = int16_t(h);
    const int hi = hs, ai = ls + (hi & 1) + (hi >> 1);
    a = uint16_t(int16_t(ai)); b = uint16_t(int16_t(ai - hi));
}
inline void wdec16(uint16_t l, uint16_t h, uint16_t &a, uint16_t &b) {
    const int m = l, d = h;
    const int bb = (m - (d >> 1)) & 0xffff;
    const int aa = (d + bb - 0x8000) & 0xffff;
    b = uint16_t(bb); a = uint16_t(aa);
}
void pizWaveletDecode(uint16_t *in, int nx, int ox, int ny, int oy, uint16_t mx) {
    const bool w14 = mx < (1 << 14);
    const int n = std::min(nx, ny);
    int p = 1;
    while (p <= n) p <<= 1;
    p >>= 1;
    int p2 = p;
    p >>= 1;
    while (p >= 1) {
        uint16_t *py = in, *ey = in + ptrdiff_t(oy) * (ny - p2);
        const int oy1 = oy * p, oy2 = oy * p2, ox1 = ox * p, ox2 = ox * p2;
        uint16_t i00, i01, i10, i11;
        for (; py <= ey; py += oy2) {
            uint16_t *px = py, *ex = py + ptrdiff_t(ox) * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t *p01 = px + ox1, *p10 = px + oy1, *p11 = p10 + ox1;
                if (w14) { wdec14(*px, *p10, i00, i10); wdec14(*p01, *p11, i01, i11); wdec14(i00, i01, *px, *p01); wdec14(i10, i11, *p10, *p11); }
                else { wdec16(*px, *p10, i00, i10); wdec16(*p01, *p11, i01, i11); wdec16(i00, i01, *px, *p01); wdec16(i10, i11, *p10, *p11); }
            }
            if (nx & p) {
                uint16_t *p10 = px + oy1;
                if (w14) wdec14(*px, *p10, i00, *p10); else wdec16(*px, *p10, i00, *p10);
                *px = i00;
            }
        }
        if (ny & p) {
            uint16_t *px = py, *ex = py + ptrdiff_t(ox) * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t *p01 = px + ox1;
                if (w14) wdec14(*px, *p01, i00, *p01); else wdec16(*px, *p01, i00, *p01);
                *px = i00;
            }
        }
        p2 = p;
        p >>= 1;
    }
}

// one PIZ chunk -> raw scanline bytes (channels interleaved per scanline like the uncompressed layout)
void pizDecompress(const unsigned char *src, size_t srcLen, unsigned char *dst, size_t dstLen, int width, int lines, const std::vector<int> &wordsPerPixel) {
    if (srcLen < 4) throw std::runtime_error("EXR/PIZ: truncated chunk");
    uint16_t minNonZero, maxNonZero;
    memcpy(&minNonZero, src, 2); memcpy(&maxNonZero, src + 2, 2);
    size_t pos = 4;
    std::vector<unsigned char> bitmap(8192, 0);
    if (minNonZero <= maxNonZero) {
        const size_t nb = size_t(maxNonZero) - minNonZero + 1;
        if (maxNonZero >= 8192 || pos + nb > srcLen) throw std::runtime_error("EXR/PIZ: bad bitmap");
        memcpy(&bitmap[minNonZero], src + pos, nb);
        pos += nb;
    }
    std::vector<uint16_t> lut(65536, 0);                     // reverse LUT: dense index -> value (zero always occurs)
    uint32_t k = 0;
    for (uint32_t i = 0; i < 65536; i++) if (i == 0 || (bitmap[i >> 3] & (1 << (i & 7)))) lut[k++] = uint16_t(i);
    const uint16_t maxValue = uint16_t(k - 1);
    if (pos + 4 > srcLen) throw std::runtime_error("EXR/PIZ: truncated chunk");
    int32_t length;
    memcpy(&length, src + pos, 4); pos += 4;
    if (length < 0 || pos + size_t(length) > srcLen) throw std::runtime_error("EXR/PIZ: bad Huffman length");
    const size_t words = dstLen / 2;
    std::vector<uint16_t> tmp(words);
    pizHufDecode(src + pos, size_t(length), tmp.data(), words);
    // planes: channel after channel, each lines x width x wordsPerPixel; wavelet per 16-bit component
    size_t start = 0;
    std::vector<size_t> starts;
    for (int wp : wordsPerPixel) {
        starts.push_back(start);
        for (int j = 0; j < wp; j++) pizWaveletDecode(tmp.data() + start + j, width, wp, lines, width * wp, maxValue);
        start += size_t(width) * lines * wp;
    }
    for (uint16_t &v : tmp) v = lut[v];
    unsigned char *o = dst;
    for (int y = 0; y < lines; y++)
        for (size_t c = 0; c < wordsPerPixel.size(); c++) {
            const size_t n = size_t(width) * wordsPerPixel[c];
            memcpy(o, tmp.data() + starts[c] + size_t(y) * n, n * 2);
            o += n * 2;
        }
}

// ---- RLE (compression 1): runs of signed-count bytes, then the same predictor + byte de-interleave as ZIP
void rleDecompress(const unsigned char *src, size_t srcLen, unsigned char *dst, size_t dstLen) {
    size_t i = 0, o = 0;
    while (i < srcLen) {
        const int count = static_cast<signed char>(src[i++]);
        if (count < 0) {                                   // -count literal bytes
            const size_t n = size_t(-count);
            if (n > srcLen - i || n > dstLen - o) throw std::runtime_error("EXR/RLE: run past the end of the chunk");
            memcpy(dst + o, src + i, n);
            i += n; o += n;
        } else {                                           // the next byte count + 1 times
            const size_t n = size_t(count) + 1;
            if (i >= srcLen || n > dstLen - o) throw std::runtime_error("EXR/RLE: run past the end of the chunk");
            memset(dst + o, src[i++], n);
            o += n;
        }
    }
    if (o != dstLen) throw std::runtime_error("EXR/RLE: chunk of the wrong size");
}

// ---- PXR24 (compression 5): zlib over byte planes; per scanline and channel the bytes of the pixel-to-pixel differences are
// stored most significant plane first (HALF: 2 planes, FLOAT: 3 planes — the float's low byte is dropped)
void pxr24Decompress(const unsigned char *src, size_t srcLen, unsigned char *dst, int width, int lines, const std::vector<int> &chanTypes) {
    size_t planeBytes = 0;
    for (int t : chanTypes) planeBytes += size_t(width) * (t == 1 ? 2 : 3);
    std::vector<unsigned char> tmp(planeBytes * size_t(lines));
    uLongf dl = uLongf(tmp.size());
    if (uncompress(tmp.data(), &dl, src, uLong(srcLen)) != Z_OK || dl != tmp.size()) throw std::runtime_error("EXR/PXR24: zlib error");
    const unsigned char *p = tmp.data();
    unsigned char *o = dst;
    const size_t n = size_t(width);
    for (int y = 0; y < lines; y++)
        for (int t : chanTypes) {
            uint32_t pixel = 0;
            if (t == 1) {
                for (size_t x = 0; x < n; x++) {
                    pixel += (uint32_t(p[x]) << 8) | uint32_t(p[n + x]);
                    const uint16_t h = uint16_t(pixel);
                    memcpy(o, &h, 2); o += 2;
                }
                p += 2 * n;
            } else {
                for (size_t x = 0; x < n; x++) {
                    pixel += (uint32_t(p[x]) << 24) | (uint32_t(p[n + x]) << 16) | (uint32_t(p[2 * n + x]) << 8);
                    memcpy(o, &pixel, 4); o += 4;
                }
                p += 3 * n;
            }
        }
}

// ---- B44 / B44A (compression 6 / 7): HALF channels in 4x4 blocks of 14 bytes (B44A: 3 bytes for a block of one value), stored
// channel after channel; FLOAT channels uncompressed.  A block holds its first value and 6-bit running differences, on 16-bit
// codes in which the ordering of the half values is monotonic (sign bit flipped for positives, all bits for negatives).
inline uint16_t b44FromOrdered(uint16_t s) { return (s & 0x8000u) ? uint16_t(s & 0x7fffu) : uint16_t(~s); }
void b44Unpack14(const unsigned char b[14], uint16_t s[16]) {
    const uint32_t shift = b[2] >> 2, bias = 0x20u << shift;
    auto d = [&](uint32_t six) { return (six & 0x3fu) << shift; };
    s[0] = uint16_t((uint32_t(b[0]) << 8) | b[1]);
    s[4] = uint16_t(s[0] + d((uint32_t(b[2]) << 4) | (b[3] >> 4)) - bias);
    s[8] = uint16_t(s[4] + d((uint32_t(b[3]) << 2) | (b[4] >> 6)) - bias);
    s[12] = uint16_t(s[8] + d(b[4]) - bias);
    s[1] = uint16_t(s[0] + d(b[5] >> 2) - bias);
    s[5] = uint16_t(s[4] + d((uint32_t(b[5]) << 4) | (b[6] >> 4)) - bias);
    s[9] = uint16_t(s[8] + d((uint32_t(b[6]) << 2) | (b[7] >> 6)) - bias);
    s[13] = uint16_t(s[12] + d(b[7]) - bias);
    s[2] = uint16_t(s[1] + d(b[8] >> 2) - bias);
    s[6] = uint16_t(s[5] + d((uint32_t(b[8]) << 4) | (b[9] >> 4)) - bias);
    s[10] = uint16_t(s[9] + d((uint32_t(b[9]) << 2) | (b[10] >> 6)) - bias);
    s[14] = uint16_t(s[13] + d(b[10]) - bias);
    s[3] = uint16_t(s[2] + d(b[11] >> 2) - bias);
    s[7] = uint16_t(s[6] + d((uint32_t(b[11]) << 4) | (b[12] >> 4)) - bias);
    s[11] = uint16_t(s[10] + d((uint32_t(b[12]) << 2) | (b[13] >> 6)) - bias);
    s[15] = uint16_t(s[14] + d(b[13]) - bias);
    for (int i = 0; i < 16; i++) s[i] = b44FromOrdered(s[i]);
}
void b44Decompress(const unsigned char *src, size_t srcLen, unsigned char *dst, int width, int lines, const std::vector<int> &chanTypes) {
    // planes channel after channel, then re-interleaved per scanline like the uncompressed layout
    std::vector<std::vector<unsigned char>> planes(chanTypes.size());
    size_t pos = 0;
    for (size_t c = 0; c < chanTypes.size(); c++) {
        const size_t bytes = chanTypes[c] == 1 ? 2 : 4;
        planes[c].assign(size_t(width) * size_t(lines) * bytes, 0);
        if (chanTypes[c] != 1) {                           // not HALF: raw
            if (planes[c].size() > srcLen - pos) throw std::runtime_error("EXR/B44: truncated chunk");
            memcpy(planes[c].data(), src + pos, planes[c].size());
            pos += planes[c].size();
            continue;
        }
        uint16_t *plane = reinterpret_cast<uint16_t *>(planes[c].data());
        for (int y = 0; y < lines; y += 4)
            for (int x = 0; x < width; x += 4) {
                uint16_t s[16];
                if (srcLen - pos < 3) throw std::runtime_error("EXR/B44: truncated chunk");
                if (src[pos + 2] >= (13 << 2)) {           // one value for the whole block (3 bytes)
                    const uint16_t v = b44FromOrdered(uint16_t((uint32_t(src[pos]) << 8) | src[pos + 1]));
                    for (int k = 0; k < 16; k++) s[k] = v;
                    pos += 3;
                } else {
                    if (srcLen - pos < 14) throw std::runtime_error("EXR/B44: truncated chunk");
                    b44Unpack14(src + pos, s);
                    pos += 14;
                }
                for (int j = 0; j < 4 && y + j < lines; j++)
                    for (int k = 0; k < 4 && x + k < width; k++) plane[size_t(y + j) * size_t(width) + size_t(x + k)] = s[j * 4 + k];
            }
    }
    unsigned char *o = dst;
    for (int y = 0; y < lines; y++)
        for (size_t c = 0; c < chanTypes.size(); c++) {
            const size_t n = size_t(width) * (chanTypes[c] == 1 ? 2 : 4);
            memcpy(o, planes[c].data() + size_t(y) * n, n);
            o += n;
        }
}
}  // namespace

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
    const size_t size = data.size();
    // every read below is bounds-checked: a malformed or hostile file raises the loader's runtime_error instead of
    // reading or writing out of bounds
    auto need = [&](size_t at, size_t n) { if (at > size || n > size - at) throw std::runtime_error("EXR: truncated or corrupt header: " + path); };
    auto cstr = [&](size_t at) -> std::string {
        const void *z = at < size ? memchr(data.data() + at, 0, size - at) : nullptr;
        if (!z) throw std::runtime_error("EXR: unterminated string in header: " + path);
        return std::string(data.data() + at, static_cast<const char *>(z) - (data.data() + at));
    };
    for (;;) {
        need(i, 1);
        if (data[i] == 0) { i++; break; }
        std::string name = cstr(i); i += name.size() + 1;
        std::string type = cstr(i); i += type.size() + 1;
        need(i, 4);
        int32_t sz; memcpy(&sz, data.data() + i, 4); i += 4;
        if (sz < 0) throw std::runtime_error("EXR: negative attribute size: " + path);
        need(i, size_t(sz));
        if (name == "channels") {
            size_t j = i;
            const size_t end = i + size_t(sz);
            for (;;) {
                if (j >= end) throw std::runtime_error("EXR: unterminated channel list: " + path);
                if (data[j] == 0) break;
                Chan c; c.name = cstr(j); j += c.name.size() + 1;
                if (j + 16 > end) throw std::runtime_error("EXR: truncated channel list: " + path);
                int32_t t; memcpy(&t, data.data() + j, 4); c.type = t; j += 16;
                chans.push_back(c);
                if (chans.size() > 64) throw std::runtime_error("EXR: too many channels: " + path);
            }
        } else if (name == "compression") { if (sz < 1) throw std::runtime_error("EXR: bad compression attribute"); comp = (unsigned char)data[i]; }
        else if (name == "dataWindow") { if (sz < 16) throw std::runtime_error("EXR: bad dataWindow attribute"); memcpy(dw, data.data() + i, 16); }
        else if (name == "lineOrder") { if (sz < 1) throw std::runtime_error("EXR: bad lineOrder attribute"); lineOrder = (unsigned char)data[i]; }
        i += size_t(sz);
    }
    (void)lineOrder;
    const int64_t w64 = int64_t(dw[2]) - dw[0] + 1, h64 = int64_t(dw[3]) - dw[1] + 1;
    if (w64 <= 0 || h64 <= 0 || w64 > 65536 || h64 > 65536 || chans.empty()) throw std::runtime_error("EXR: bad header " + path);
    width = int(w64); height = int(h64);
    int linesPerBlock;
    if (comp == 0 || comp == 1 || comp == 2) linesPerBlock = 1;
    else if (comp == 3 || comp == 5) linesPerBlock = 16;
    else if (comp == 4 || comp == 6 || comp == 7) linesPerBlock = 32;
    else throw std::runtime_error("EXR: compression type " + std::to_string(comp) + " not supported (NONE / RLE / ZIPS / ZIP / PIZ / PXR24 / B44 / B44A are): " + path);
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
    need(i, size_t(numBlocks) * 8);                           // the chunk-offset table
    for (int b = 0; b < numBlocks; b++) {
        uint64_t off; memcpy(&off, data.data() + i + size_t(b) * 8, 8);
        if (off > size || size - off < 8) throw std::runtime_error("EXR: bad chunk offset");
        int32_t y0, sz; memcpy(&y0, data.data() + off, 4); memcpy(&sz, data.data() + off + 4, 4);
        if (y0 < dw[1] || y0 > dw[3]) throw std::runtime_error("EXR: chunk outside the data window");
        if (sz < 0 || size_t(sz) > size - off - 8) throw std::runtime_error("EXR: bad chunk size");
        const int lines = int(std::min<int64_t>(linesPerBlock, int64_t(dw[3]) - y0 + 1));
        size_t expect = bytesPerLine * size_t(lines);
        const unsigned char *src = reinterpret_cast<const unsigned char *>(data.data() + off + 8);
        raw.resize(expect);
        if (comp == 0 && size_t(sz) != expect) throw std::runtime_error("EXR: uncompressed chunk of the wrong size");
        if (comp == 0 || size_t(sz) == expect) memcpy(raw.data(), src, expect);
        else if (comp == 4) {
            std::vector<int> wordsPerPixel;
            for (auto &c : chans) wordsPerPixel.push_back(c.type == 1 ? 1 : 2);
            pizDecompress(src, size_t(sz), raw.data(), expect, width, lines, wordsPerPixel);
        } else if (comp == 5 || comp == 6 || comp == 7) {
            std::vector<int> chanTypes;
            for (auto &c : chans) chanTypes.push_back(c.type);
            if (comp == 5) pxr24Decompress(src, size_t(sz), raw.data(), width, lines, chanTypes);
            else b44Decompress(src, size_t(sz), raw.data(), width, lines, chanTypes);
        } else {
            tmp.resize(expect);
            if (comp == 1) rleDecompress(src, size_t(sz), tmp.data(), expect);
            else {
                uLongf dl = uLongf(expect);
                if (uncompress(tmp.data(), &dl, src, uLong(sz)) != Z_OK || dl != expect) throw std::runtime_error("EXR: zlib error");
            }
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
