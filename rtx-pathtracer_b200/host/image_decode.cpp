// Bitmap texture decoding for the scene loader: JPEG (baseline, extended sequential and progressive Huffman) and PNG to RGBA8.
// The reference decodes textures with the vendored stb_image (`stbi_load(path, &w, &h, &channels, 4)`,
// src/SceneLoader.cpp:198-207).  This file is an independent implementation of the two published formats
// (ITU-T T.81 sequential and progressive DCT + JFIF colour; PNG / RFC 2083 with zlib's inflate, Adam7 included).  So that texel bytes agree
// with what the reference uploads, the JPEG path uses the same well-known arithmetic stb_image documents:
//   - the 13-bit-constant "islow" integer inverse DCT of the IJG library (Loeffler–Ligtenberg–Moschytz), two extra
//     bits kept between the passes;
//   - triangle-filter ("fancy") chroma upsampling, weights 3/4 + 1/4 per axis, (9,3,3,1)/16 for 2x2;
//   - fixed-point YCbCr -> RGB with 12-bit coefficients shifted by 8, the Cb term of green truncated to 16 bits.
// tests/test_image_decode.py checks the output byte for byte against stb_image itself (compiled from the reference
// tree into oracle/_ref) and against committed digests.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <zlib.h>
#include "scene.h"

namespace b200pt {

namespace {

std::vector<uint8_t> readBinaryFile(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Could not load texture file " + path);
    std::vector<uint8_t> data;
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + n);
    fclose(f);
    return data;
}

// ================================================================================================ JPEG
const uint8_t kZigZag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                             35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {          // canonical code tables of T.81 Annex C / F.2.2.3
    bool defined = false;
    uint8_t values[256];
    int mincode[17], maxcode[18], valptr[17];
    void build(const uint8_t counts[16], const uint8_t *vals, int n) {
        memcpy(values, vals, size_t(n));
        int code = 0, k = 0;
        for (int len = 1; len <= 16; len++) {
            valptr[len] = k;
            mincode[len] = code;
            code += counts[len - 1];
            k += counts[len - 1];
            maxcode[len] = counts[len - 1] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        defined = true;
    }
};

struct BitReader {
    const uint8_t *p, *end;
    uint32_t acc = 0;
    int bits = 0;
    bool hitMarker = false;
    BitReader(const uint8_t *b, const uint8_t *e) : p(b), end(e) {}
    void fill() {
        while (bits <= 24) {
            uint32_t byte = 0;
            if (!hitMarker && p < end) {
                byte = *p;
                if (byte == 0xff) {
                    uint8_t next = p + 1 < end ? p[1] : 0xd9;
                    if (next == 0) p += 2;                 // stuffed zero
                    else { hitMarker = true; byte = 0; }   // a marker ends the entropy-coded segment: feed zeros
                } else p++;
            }
            acc |= byte << (24 - bits);
            bits += 8;
        }
    }
    int getBit() { if (bits < 1) fill(); int b = int(acc >> 31); acc <<= 1; bits--; return b; }
    int getBits(int n) {
        if (n == 0) return 0;
        if (bits < n) fill();
        int v = int(acc >> (32 - n));
        acc <<= n; bits -= n;
        return v;
    }
    void reset() { acc = 0; bits = 0; hitMarker = false; }
};

int decodeHuff(BitReader &br, const HuffTable &h) {
    int code = 0;
    for (int len = 1; len <= 16; len++) {
        code = (code << 1) | br.getBit();
        if (h.maxcode[len] >= 0 && code <= h.maxcode[len] && code >= h.mincode[len]) return h.values[h.valptr[len] + code - h.mincode[len]];
    }
    throw std::runtime_error("JPEG: bad Huffman code");
}

inline int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }   // T.81 F.2.2.1

inline uint8_t clampByte(int x) { return uint8_t(x < 0 ? 0 : (x > 255 ? 255 : x)); }

// 8x8 inverse DCT, integer "islow" variant with 12-bit constants (see the header comment)
inline int fix(double x) { return int(x * 4096 + 0.5); }
struct Idct1D { int x0, x1, x2, x3, t0, t1, t2, t3; };
inline Idct1D idct1d(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7) {
    static const int c0541 = fix(0.5411961f), c1847 = fix(-1.847759065f), c0765 = fix(0.765366865f), c1175 = fix(1.175875602f),
                     c0298 = fix(0.298631336f), c2053 = fix(2.053119869f), c3072 = fix(3.072711026f), c1501 = fix(1.501321110f),
                     c0899 = fix(-0.899976223f), c2562 = fix(-2.562915447f), c1961 = fix(-1.961570560f), c0390 = fix(-0.390180644f);
    Idct1D r;
    int p1 = (s2 + s6) * c0541;
    int e2 = p1 + s6 * c1847, e3 = p1 + s2 * c0765;
    int e0 = (s0 + s4) * 4096, e1 = (s0 - s4) * 4096;
    r.x0 = e0 + e3; r.x3 = e0 - e3; r.x1 = e1 + e2; r.x2 = e1 - e2;
    int o0 = s7, o1 = s5, o2 = s3, o3 = s1;
    int p3 = o0 + o2, p4 = o1 + o3;
    int q1 = o0 + o3, q2 = o1 + o2;
    int p5 = (p3 + p4) * c1175;
    o0 *= c0298; o1 *= c2053; o2 *= c3072; o3 *= c1501;
    q1 = p5 + q1 * c0899; q2 = p5 + q2 * c2562;
    p3 *= c1961; p4 *= c0390;
    r.t3 = o3 + q1 + p4; r.t2 = o2 + q2 + p3; r.t1 = o1 + q2 + p4; r.t0 = o0 + q1 + p3;
    return r;
}
void idctBlock(uint8_t *out, int stride, const short d[64]) {
    int val[64];
    for (int i = 0; i < 8; i++) {       // columns
        const short *c = d + i;
        int *v = val + i;
        if (c[8] == 0 && c[16] == 0 && c[24] == 0 && c[32] == 0 && c[40] == 0 && c[48] == 0 && c[56] == 0) {
            int dc = c[0] * 4;
            for (int k = 0; k < 8; k++) v[8 * k] = dc;
        } else {
            Idct1D r = idct1d(c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56]);
            r.x0 += 512; r.x1 += 512; r.x2 += 512; r.x3 += 512;
            v[0] = (r.x0 + r.t3) >> 10; v[56] = (r.x0 - r.t3) >> 10;
            v[8] = (r.x1 + r.t2) >> 10; v[48] = (r.x1 - r.t2) >> 10;
            v[16] = (r.x2 + r.t1) >> 10; v[40] = (r.x2 - r.t1) >> 10;
            v[24] = (r.x3 + r.t0) >> 10; v[32] = (r.x3 - r.t0) >> 10;
        }
    }
    for (int i = 0; i < 8; i++) {       // rows; 1<<17 to remove, rounding and the +128 level shift folded in
        const int *v = val + 8 * i;
        uint8_t *o = out + i * stride;
        Idct1D r = idct1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
        const int bias = 65536 + (128 << 17);
        r.x0 += bias; r.x1 += bias; r.x2 += bias; r.x3 += bias;
        o[0] = clampByte((r.x0 + r.t3) >> 17); o[7] = clampByte((r.x0 - r.t3) >> 17);
        o[1] = clampByte((r.x1 + r.t2) >> 17); o[6] = clampByte((r.x1 - r.t2) >> 17);
        o[2] = clampByte((r.x2 + r.t1) >> 17); o[5] = clampByte((r.x2 - r.t1) >> 17);
        o[3] = clampByte((r.x3 + r.t0) >> 17); o[4] = clampByte((r.x3 - r.t0) >> 17);
    }
}

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int dcPred = 0;
    int w2 = 0, h2 = 0;          // padded plane size
    int rows = 0, cols = 0;      // samples that carry image data: ceil(img * h / hmax)
    std::vector<uint8_t> plane;
};

// one output row of a component, upsampled to full width: near / far are the two source rows (triangle filter)
void upsampleRow(uint8_t *out, const uint8_t *nearRow, const uint8_t *farRow, int w, int hs, int vs) {
    if (hs == 1 && vs == 1) { memcpy(out, nearRow, size_t(w)); return; }
    if (hs == 1 && vs == 2) { for (int i = 0; i < w; i++) out[i] = uint8_t((3 * nearRow[i] + farRow[i] + 2) >> 2); return; }
    if (hs == 2 && vs == 1) {
        if (w == 1) { out[0] = out[1] = nearRow[0]; return; }
        out[0] = nearRow[0];
        out[1] = uint8_t((nearRow[0] * 3 + nearRow[1] + 2) >> 2);
        int i;
        for (i = 1; i < w - 1; i++) {
            int n = 3 * nearRow[i] + 2;
            out[2 * i] = uint8_t((n + nearRow[i - 1]) >> 2);
            out[2 * i + 1] = uint8_t((n + nearRow[i + 1]) >> 2);
        }
        out[2 * i] = uint8_t((nearRow[w - 2] * 3 + nearRow[w - 1] + 2) >> 2);
        out[2 * i + 1] = nearRow[w - 1];
        return;
    }
    if (hs == 2 && vs == 2) {
        if (w == 1) { out[0] = out[1] = uint8_t((3 * nearRow[0] + farRow[0] + 2) >> 2); return; }
        int t1 = 3 * nearRow[0] + farRow[0], t0;
        out[0] = uint8_t((t1 + 2) >> 2);
        for (int i = 1; i < w; i++) {
            t0 = t1;
            t1 = 3 * nearRow[i] + farRow[i];
            out[2 * i - 1] = uint8_t((3 * t0 + t1 + 8) >> 4);
            out[2 * i] = uint8_t((3 * t1 + t0 + 8) >> 4);
        }
        out[2 * w - 1] = uint8_t((t1 + 2) >> 2);
        return;
    }
    for (int i = 0; i < w; i++) for (int j = 0; j < hs; j++) out[i * hs + j] = nearRow[i];   // other ratios: nearest
}

inline int fixedColour(float x) { return int(x * 4096.0f + 0.5f) << 8; }

// One scan of a progressive frame works on the stored coefficients of the blocks it covers (T.81 Annex G): DC scans
// (first / refinement of the point transform Al) and AC band scans (first: run / size symbols with end-of-band runs;
// refinement: one correction bit per already-nonzero coefficient, new +-1 << Al coefficients placed after `r` zeros).
struct ScanParams { int ss = 0, se = 63, ah = 0, al = 0; };

void progressiveDc(BitReader &br, const HuffTable &h, Component &c, short *blk, const ScanParams &sp) {
    if (sp.ah == 0) {
        memset(blk, 0, 64 * sizeof(short));
        const int t = decodeHuff(br, h);
        if (t > 15) throw std::runtime_error("JPEG: bad DC size");
        c.dcPred += t ? extend(br.getBits(t), t) : 0;
        blk[0] = short(c.dcPred << sp.al);
    } else if (br.getBit()) blk[0] = short(blk[0] + short(1 << sp.al));
}

void progressiveAc(BitReader &br, const HuffTable &h, short *blk, const ScanParams &sp, int &eobRun) {
    if (sp.ah == 0) {
        if (eobRun) { eobRun--; return; }
        for (int k = sp.ss; k <= sp.se;) {
            const int rs = decodeHuff(br, h);
            const int r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (r < 15) {                       // end of band for this block and the next (1 << r) + bits - 1 blocks
                    eobRun = (1 << r) - 1;
                    if (r) eobRun += br.getBits(r);
                    break;
                }
                k += 16;
            } else {
                k += r;
                if (k > 63) throw std::runtime_error("JPEG: coefficient index out of range");
                blk[kZigZag[k++]] = short(extend(br.getBits(s), s) << sp.al);
            }
        }
        return;
    }
    const short bit = short(1 << sp.al);
    auto refine = [&](short &p) {                    // correction bit of a coefficient that is already nonzero
        if (br.getBit() && (p & bit) == 0) p = short(p > 0 ? p + bit : p - bit);
    };
    if (eobRun) {
        eobRun--;
        for (int k = sp.ss; k <= sp.se; k++) { short &p = blk[kZigZag[k]]; if (p != 0) refine(p); }
        return;
    }
    for (int k = sp.ss; k <= sp.se;) {
        const int rs = decodeHuff(br, h);
        int r = rs >> 4, s = rs & 15;
        if (s == 0) {
            if (r < 15) {
                eobRun = (1 << r) - 1;
                if (r) eobRun += br.getBits(r);
                r = 64;                              // the rest of the band only gets correction bits
            }                                        // r == 15: sixteen zeros (fifteen skipped + a zero "new coefficient")
        } else {
            if (s != 1) throw std::runtime_error("JPEG: bad refinement symbol");
            s = br.getBit() ? bit : -bit;
        }
        while (k <= sp.se) {
            short &p = blk[kZigZag[k++]];
            if (p != 0) refine(p);
            else {
                if (r == 0) { p = short(s); break; }
                r--;
            }
        }
    }
}

void decodeJpeg(const std::vector<uint8_t> &d, int &width, int &height, std::vector<uint8_t> &rgba) {
    if (d.size() < 4 || d[0] != 0xff || d[1] != 0xd8) throw std::runtime_error("JPEG: no SOI marker");
    uint16_t quant[4][64];
    bool quantDefined[4] = {false, false, false, false};
    HuffTable dc[4], ac[4];
    std::vector<Component> comps;
    std::vector<std::vector<short>> coeff;      // progressive frames: 64 coefficients per block, w2 / 8 blocks per row
    int hmax = 1, vmax = 1, restartInterval = 0, mcux = 0, mcuy = 0, scans = 0;
    bool haveFrame = false, progressive = false, adobeTransformKnown = false;
    int adobeTransform = -1;
    size_t pos = 2;
    width = height = 0;
    auto u16 = [&](size_t i) { if (i + 1 >= d.size()) throw std::runtime_error("JPEG: truncated"); return (int(d[i]) << 8) | d[i + 1]; };
    for (;;) {
        while (pos < d.size() && d[pos] != 0xff) pos++;
        while (pos < d.size() && d[pos] == 0xff) pos++;
        if (pos >= d.size()) { if (scans) break; throw std::runtime_error("JPEG: no image data"); }      // like stb_image, a missing EOI after the last scan is tolerated
        const int marker = d[pos++];
        if (marker == 0xd9) { if (scans) break; throw std::runtime_error("JPEG: no scan"); }
        if (marker == 0x00 || marker == 0x01 || (marker >= 0xd0 && marker <= 0xd7)) continue;          // stuffed byte of a scan's tail, TEM, RSTn
        const int len = u16(pos);
        const size_t seg = pos + 2, segEnd = pos + size_t(len);
        if (len < 2 || segEnd > d.size()) throw std::runtime_error("JPEG: bad segment length");
        // every index below stays inside [seg, segEnd): a truncated or corrupt segment raises instead of reading past it
        auto segNeed = [&](size_t at, size_t n) { if (at < seg || at > segEnd || n > segEnd - at) throw std::runtime_error("JPEG: truncated segment"); };
        if (marker == 0xdb) {                                    // DQT
            size_t i = seg;
            while (i < segEnd) {
                int pq = d[i] >> 4, tq = d[i] & 15;
                i++;
                if (tq > 3 || pq > 1) throw std::runtime_error("JPEG: bad DQT");
                segNeed(i, pq ? 128 : 64);
                for (int k = 0; k < 64; k++) { quant[tq][kZigZag[k]] = uint16_t(pq ? u16(i) : d[i]); i += pq ? 2 : 1; }
                quantDefined[tq] = true;
            }
        } else if (marker == 0xc4) {                             // DHT
            size_t i = seg;
            while (i < segEnd) {
                int tc = d[i] >> 4, th = d[i] & 15;
                i++;
                if (tc > 1 || th > 3 || i + 16 > segEnd) throw std::runtime_error("JPEG: bad DHT");
                uint8_t counts[16];
                int n = 0;
                for (int k = 0; k < 16; k++) { counts[k] = d[i + size_t(k)]; n += counts[k]; }
                i += 16;
                if (n > 256 || i + size_t(n) > segEnd) throw std::runtime_error("JPEG: bad DHT");
                (tc ? ac[th] : dc[th]).build(counts, &d[i], n);
                i += size_t(n);
            }
        } else if (marker == 0xc0 || marker == 0xc1 || marker == 0xc2) {      // SOF0 / SOF1 / SOF2: baseline / extended sequential / progressive, Huffman
            if (haveFrame) throw std::runtime_error("JPEG: more than one frame header");
            segNeed(seg, 6);
            if (d[seg] != 8) throw std::runtime_error("JPEG: only 8-bit samples are supported");
            height = u16(seg + 1); width = u16(seg + 3);
            const int n = d[seg + 5];
            if (width <= 0 || height <= 0 || (n != 1 && n != 3)) throw std::runtime_error("JPEG: unsupported frame (size or component count)");
            // a corrupt size must not turn into gigabytes of planes: every 8x8 block of every component costs the file at least one
            // bit (its DC code), so a frame that claims more blocks than the file has bits is refused before anything is allocated
            if (uint64_t(width) * uint64_t(height) > (1ull << 28) || (uint64_t(width + 7) / 8) * (uint64_t(height + 7) / 8) > 8ull * d.size() + 1024ull)
                throw std::runtime_error("JPEG: the frame size does not fit the file");
            segNeed(seg + 6, 3 * size_t(n));
            comps.assign(size_t(n), Component());
            for (int k = 0; k < n; k++) {
                Component &c = comps[size_t(k)];
                c.id = d[seg + 6 + 3 * size_t(k)];
                c.h = d[seg + 7 + 3 * size_t(k)] >> 4; c.v = d[seg + 7 + 3 * size_t(k)] & 15;
                c.tq = d[seg + 8 + 3 * size_t(k)];
                if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) throw std::runtime_error("JPEG: bad component");
                hmax = std::max(hmax, c.h); vmax = std::max(vmax, c.v);
            }
            haveFrame = true;
            progressive = marker == 0xc2;
            if (comps.size() == 1) { comps[0].h = comps[0].v = 1; hmax = vmax = 1; }      // a single-component scan is never interleaved
            const int mcuW = 8 * hmax, mcuH = 8 * vmax;
            mcux = (width + mcuW - 1) / mcuW; mcuy = (height + mcuH - 1) / mcuH;
            if (progressive) coeff.resize(comps.size());
            for (size_t k = 0; k < comps.size(); k++) {
                Component &c = comps[k];
                c.w2 = mcux * c.h * 8; c.h2 = mcuy * c.v * 8;
                c.cols = (width * c.h + hmax - 1) / hmax; c.rows = (height * c.v + vmax - 1) / vmax;
                c.plane.assign(size_t(c.w2) * size_t(c.h2), 0);
                if (progressive) coeff[k].assign(size_t(c.w2 / 8) * size_t(c.h2 / 8) * 64, 0);
            }
        } else if (marker >= 0xc3 && marker <= 0xcf && marker != 0xc4 && marker != 0xc8 && marker != 0xcc) {
            throw std::runtime_error("JPEG: lossless / hierarchical / arithmetic-coded files are not supported");
        } else if (marker == 0xdd) {
            segNeed(seg, 2);
            restartInterval = u16(seg);
        } else if (marker == 0xee && len >= 14 && memcmp(&d[seg], "Adobe", 5) == 0) {
            adobeTransformKnown = true; adobeTransform = d[seg + 11];
        } else if (marker == 0xda) {                             // SOS: one scan (a baseline file has one interleaved scan or one per component)
            if (!haveFrame) throw std::runtime_error("JPEG: scan before frame header");
            segNeed(seg, 1);
            const int ns = d[seg];
            if (ns < 1 || ns > int(comps.size())) throw std::runtime_error("JPEG: bad scan component count");
            segNeed(seg + 1, 2 * size_t(ns) + 3);
            std::vector<size_t> order;
            for (int k = 0; k < ns; k++) {
                const int cid = d[seg + 1 + 2 * size_t(k)], tbl = d[seg + 2 + 2 * size_t(k)];
                size_t which = comps.size();
                for (size_t q = 0; q < comps.size(); q++) if (comps[q].id == cid) { which = q; break; }
                if (which == comps.size()) throw std::runtime_error("JPEG: bad scan component");
                Component &c = comps[which];
                c.td = tbl >> 4; c.ta = tbl & 15;
                if (c.td > 3 || c.ta > 3) throw std::runtime_error("JPEG: bad table index");
                order.push_back(which);
            }
            ScanParams sp;
            sp.ss = d[seg + 1 + 2 * size_t(ns)]; sp.se = d[seg + 2 + 2 * size_t(ns)];
            sp.ah = d[seg + 3 + 2 * size_t(ns)] >> 4; sp.al = d[seg + 3 + 2 * size_t(ns)] & 15;
            if (progressive) {
                if (sp.ss > 63 || sp.se > 63 || sp.ss > sp.se || sp.ah > 13 || sp.al > 13) throw std::runtime_error("JPEG: bad scan parameters");
                if (sp.ss == 0 && sp.se != 0) throw std::runtime_error("JPEG: a progressive scan cannot mix DC and AC coefficients");
                if (sp.ss != 0 && ns != 1) throw std::runtime_error("JPEG: progressive AC scans hold one component");
            } else {
                if (sp.ss != 0 || sp.ah != 0 || sp.al != 0) throw std::runtime_error("JPEG: bad scan parameters");
                sp.se = 63;
            }
            for (size_t which : order) {
                const Component &c = comps[which];
                const bool needDc = !progressive || sp.ss == 0, needAc = !progressive || sp.ss != 0;
                if ((needDc && !(progressive && sp.ah) && !dc[c.td].defined) || (needAc && !ac[c.ta].defined) || !quantDefined[c.tq]) throw std::runtime_error("JPEG: missing table");
            }
            // ---- entropy-coded data of this scan
            BitReader br(d.data() + segEnd, d.data() + d.size());
            int todo = restartInterval ? restartInterval : 0x7fffffff, eobRun = 0;
            for (auto &c : comps) c.dcPred = 0;
            bool stop = false;
            auto mcuDone = [&]() {       // restart interval: skip to the RSTn marker, reset predictors; anything else there ends the scan
                if (--todo > 0) return;
                const uint8_t *p = br.p;
                while (p + 1 < br.end && !(p[0] == 0xff && p[1] >= 0xd0 && p[1] <= 0xd7)) {
                    if (p[0] == 0xff && p[1] != 0 && p[1] != 0xff) break;
                    p++;
                }
                if (p + 1 < br.end && p[0] == 0xff && p[1] >= 0xd0 && p[1] <= 0xd7) p += 2;
                else stop = true;
                br.p = p; br.reset();
                for (auto &c : comps) c.dcPred = 0;
                todo = restartInterval; eobRun = 0;
            };
            short block[64];
            auto codeBlock = [&](Component &c, size_t which, int bxAbs, int byAbs) {
                if (progressive) {
                    short *blk = &coeff[which][(size_t(byAbs) * size_t(c.w2 / 8) + size_t(bxAbs)) * 64];
                    if (sp.ss == 0) progressiveDc(br, dc[c.td], c, blk, sp);
                    else progressiveAc(br, ac[c.ta], blk, sp, eobRun);
                    return;
                }
                memset(block, 0, sizeof(block));
                const int t = decodeHuff(br, dc[c.td]);
                if (t > 15) throw std::runtime_error("JPEG: bad DC size");
                const int diff = t ? extend(br.getBits(t), t) : 0;
                c.dcPred += diff;
                block[0] = short(c.dcPred * quant[c.tq][0]);
                for (int k = 1; k < 64;) {
                    const int rs = decodeHuff(br, ac[c.ta]);
                    const int r = rs >> 4, s = rs & 15;
                    if (s == 0) {
                        if (rs != 0xf0) break;      // end of block
                        k += 16;
                    } else {
                        k += r;
                        if (k > 63) throw std::runtime_error("JPEG: coefficient index out of range");
                        const int zig = kZigZag[k];
                        block[zig] = short(extend(br.getBits(s), s) * quant[c.tq][zig]);
                        k++;
                    }
                }
                idctBlock(&c.plane[size_t(byAbs) * 8 * size_t(c.w2) + size_t(bxAbs) * 8], c.w2, block);
            };
            if (ns == 1) {               // not interleaved: the blocks that carry image samples, row by row
                Component &c = comps[order[0]];
                const int bw = (c.cols + 7) >> 3, bh = (c.rows + 7) >> 3;
                for (int by = 0; by < bh && !stop; by++)
                    for (int bx = 0; bx < bw && !stop; bx++) { codeBlock(c, order[0], bx, by); mcuDone(); }
            } else {
                for (int my = 0; my < mcuy && !stop; my++)
                    for (int mx = 0; mx < mcux && !stop; mx++) {
                        for (size_t which : order) {
                            Component &c = comps[which];
                            for (int by = 0; by < c.v; by++)
                                for (int bx = 0; bx < c.h; bx++) codeBlock(c, which, mx * c.h + bx, my * c.v + by);
                        }
                        mcuDone();
                    }
            }
            scans++;
            pos = size_t(br.p - d.data());       // at the marker that ended the entropy-coded data (or where the reader stopped)
            continue;
        }
        pos = segEnd;
    }
    if (progressive) {                           // dequantise (16-bit products, like the stored baseline coefficients) + inverse DCT
        for (size_t k = 0; k < comps.size(); k++) {
            Component &c = comps[k];
            if (!quantDefined[c.tq]) throw std::runtime_error("JPEG: missing table");
            const int bw = (c.cols + 7) >> 3, bh = (c.rows + 7) >> 3;
            for (int by = 0; by < bh; by++)
                for (int bx = 0; bx < bw; bx++) {
                    short *blk = &coeff[k][(size_t(by) * size_t(c.w2 / 8) + size_t(bx)) * 64];
                    for (int i = 0; i < 64; i++) blk[i] = short(blk[i] * quant[c.tq][i]);
                    idctBlock(&c.plane[size_t(by) * 8 * size_t(c.w2) + size_t(bx) * 8], c.w2, blk);
                }
        }
    }
    // ---- upsample + colour convert
    rgba.assign(size_t(width) * size_t(height) * 4, 255);
    std::vector<std::vector<uint8_t>> lines(comps.size(), std::vector<uint8_t>(size_t(width) + 8 * size_t(hmax) + 8));
    const bool ycc = comps.size() == 3 && !(adobeTransformKnown && adobeTransform == 0);
    for (int y = 0; y < height; y++) {
        for (size_t k = 0; k < comps.size(); k++) {
            const Component &c = comps[k];
            const int hs = hmax / c.h, vs = vmax / c.v;
            // source rows of the vertical triangle filter: the nearer row and its neighbour on the other side
            const int nearY = std::min(y / vs, c.rows - 1);
            int farY = nearY;
            if (vs == 2) farY = (y & 1) ? std::min(nearY + 1, c.rows - 1) : std::max(nearY - 1, 0);
            upsampleRow(lines[k].data(), &c.plane[size_t(nearY) * size_t(c.w2)], &c.plane[size_t(farY) * size_t(c.w2)], c.cols, hs, vs);
        }
        uint8_t *out = &rgba[size_t(y) * size_t(width) * 4];
        if (comps.size() == 1) {
            for (int x = 0; x < width; x++) { out[4 * x] = out[4 * x + 1] = out[4 * x + 2] = lines[0][size_t(x)]; out[4 * x + 3] = 255; }
        } else if (!ycc) {
            for (int x = 0; x < width; x++) { out[4 * x] = lines[0][size_t(x)]; out[4 * x + 1] = lines[1][size_t(x)]; out[4 * x + 2] = lines[2][size_t(x)]; out[4 * x + 3] = 255; }
        } else {
            static const int crR = fixedColour(1.40200f), crG = fixedColour(0.71414f), cbG = fixedColour(0.34414f), cbB = fixedColour(1.77200f);
            for (int x = 0; x < width; x++) {
                const int yf = (int(lines[0][size_t(x)]) << 20) + (1 << 19);
                const int cb = int(lines[1][size_t(x)]) - 128, cr = int(lines[2][size_t(x)]) - 128;
                int r = yf + cr * crR;
                int g = yf + cr * -crG + int(uint32_t(cb * -cbG) & 0xffff0000u);
                int b = yf + cb * cbB;
                out[4 * x] = clampByte(r >> 20); out[4 * x + 1] = clampByte(g >> 20); out[4 * x + 2] = clampByte(b >> 20); out[4 * x + 3] = 255;
            }
        }
    }
}

// ================================================================================================ PNG
inline uint32_t be32(const uint8_t *p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
inline int paeth(int a, int b, int c) {
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

void decodePng(const std::vector<uint8_t> &d, int &width, int &height, std::vector<uint8_t> &rgba) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (d.size() < 8 || memcmp(d.data(), sig, 8) != 0) throw std::runtime_error("PNG: bad signature");
    size_t pos = 8;
    int depth = 0, colour = 0, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    bool haveHeader = false;
    while (pos + 12 <= d.size()) {
        const uint32_t len = be32(&d[pos]);
        const char *type = reinterpret_cast<const char *>(&d[pos + 4]);
        const uint8_t *body = &d[pos + 8];
        if (pos + 12 + size_t(len) > d.size()) throw std::runtime_error("PNG: truncated chunk");
        if (!memcmp(type, "IHDR", 4)) {
            if (len < 13) throw std::runtime_error("PNG: bad IHDR");
            width = int(be32(body)); height = int(be32(body + 4)); depth = body[8]; colour = body[9]; interlace = body[12];
            haveHeader = true;
        } else if (!memcmp(type, "PLTE", 4)) palette.assign(body, body + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(body, body + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(type, "IEND", 4)) break;
        pos += 12 + size_t(len);
    }
    if (!haveHeader || width <= 0 || height <= 0) throw std::runtime_error("PNG: no header");
    if (interlace > 1) throw std::runtime_error("PNG: bad interlace method");
    if (!(depth == 8 || depth == 16 || (colour == 3 && (depth == 1 || depth == 2 || depth == 4)) || (colour == 0 && depth < 8)))
        throw std::runtime_error("PNG: unsupported bit depth");
    const int channels = colour == 0 ? 1 : colour == 2 ? 3 : colour == 3 ? 1 : colour == 4 ? 2 : colour == 6 ? 4 : 0;
    if (!channels) throw std::runtime_error("PNG: bad colour type");
    const size_t bpp = std::max<size_t>(1, size_t(channels) * size_t(depth) / 8);          // filter distance in bytes
    // the image is one pass, or the seven reduced images of Adam7 (RFC 2083 section 2.6): pixel (x, y) of a pass is pixel
    // (x0 + x * dx, y0 + y * dy) of the image; every pass is filtered on its own, empty passes have no bytes
    struct Pass { int x0, y0, dx, dy; };
    static const Pass adam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const Pass whole = {0, 0, 1, 1};
    const Pass *passes = interlace ? adam7 : &whole;
    const int numPasses = interlace ? 7 : 1;
    auto passWidth = [&](const Pass &ps) { return (width - ps.x0 + ps.dx - 1) / ps.dx; };
    auto passHeight = [&](const Pass &ps) { return (height - ps.y0 + ps.dy - 1) / ps.dy; };
    auto passStride = [&](int pw) { return (size_t(pw) * size_t(channels) * size_t(depth) + 7) / 8; };
    size_t rawSize = 0;
    for (int ip = 0; ip < numPasses; ip++) {
        const int pw = passWidth(passes[ip]), ph = passHeight(passes[ip]);
        if (pw > 0 && ph > 0) rawSize += (passStride(pw) + 1) * size_t(ph);
    }
    // deflate expands at most 1032 : 1: a header whose image cannot come out of the IDAT bytes present is refused before the
    // (possibly gigabytes of) scanlines and pixels are allocated
    if (uint64_t(width) * uint64_t(height) > (1ull << 28) || uint64_t(rawSize) > 1032ull * idat.size() + 1024ull)
        throw std::runtime_error("PNG: the image size does not fit the file");
    std::vector<uint8_t> raw(rawSize);
    uLongf rawLen = uLongf(raw.size());
    int zr = uncompress(raw.data(), &rawLen, idat.data(), uLong(idat.size()));
    if (zr != Z_OK || rawLen != raw.size()) throw std::runtime_error("PNG: inflate failed");
    rgba.assign(size_t(width) * size_t(height) * 4, 255);
    size_t rawPos = 0;
    for (int ip = 0; ip < numPasses; ip++) {
    const Pass ps = passes[ip];
    const int pw = passWidth(ps), ph = passHeight(ps);
    if (pw <= 0 || ph <= 0) continue;
    const size_t stride = passStride(pw);
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    for (int y = 0; y < ph; y++) {
        const uint8_t *line = &raw[rawPos + (stride + 1) * size_t(y)];
        const int filter = line[0];
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v = line[1 + i];
            switch (filter) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: throw std::runtime_error("PNG: bad filter");
            }
            cur[i] = uint8_t(v);
        }
        uint8_t *outRow = &rgba[size_t(ps.y0 + y * ps.dy) * size_t(width) * 4];
        for (int x = 0; x < pw; x++) {
            uint8_t *out = outRow + size_t(ps.x0 + x * ps.dx) * 4;
            int s[4] = {0, 0, 0, 255};
            for (int ch = 0; ch < channels; ch++) {
                if (depth == 8) s[ch] = cur[size_t(x) * size_t(channels) + size_t(ch)];
                else if (depth == 16) s[ch] = cur[(size_t(x) * size_t(channels) + size_t(ch)) * 2];      // high byte, like stbi_load's 8-bit API
                else {
                    const size_t bit = size_t(x) * size_t(depth);
                    s[ch] = (cur[bit >> 3] >> (8 - depth - int(bit & 7))) & ((1 << depth) - 1);
                }
            }
            if (colour == 0) {
                int g = s[0], a = 255;
                if (depth < 8) g = g * (255 / ((1 << depth) - 1));
                if (trns.size() >= 2 && depth <= 8 && cur.size() && s[0] == ((int(trns[0]) << 8 | trns[1]) & ((1 << depth) - 1))) a = 0;
                out[0] = out[1] = out[2] = uint8_t(g); out[3] = uint8_t(a);
            } else if (colour == 2) {
                int a = 255;
                if (trns.size() >= 6 && depth == 8 && s[0] == trns[1] && s[1] == trns[3] && s[2] == trns[5]) a = 0;
                out[0] = uint8_t(s[0]); out[1] = uint8_t(s[1]); out[2] = uint8_t(s[2]); out[3] = uint8_t(a);
            } else if (colour == 3) {
                const size_t idx = size_t(s[0]);
                if (idx * 3 + 2 >= palette.size()) throw std::runtime_error("PNG: palette index out of range");
                out[0] = palette[idx * 3]; out[1] = palette[idx * 3 + 1]; out[2] = palette[idx * 3 + 2];
                out[3] = idx < trns.size() ? trns[idx] : 255;
            } else if (colour == 4) {
                out[0] = out[1] = out[2] = uint8_t(s[0]); out[3] = uint8_t(s[1]);
            } else {
                out[0] = uint8_t(s[0]); out[1] = uint8_t(s[1]); out[2] = uint8_t(s[2]); out[3] = uint8_t(s[3]);
            }
        }
        prev.swap(cur);
    }
    rawPos += (stride + 1) * size_t(ph);
    }
}

}  // namespace

void decodeImageFile(const std::string &path, int &width, int &height, std::vector<uint8_t> &rgba) {
    std::vector<uint8_t> d = readBinaryFile(path);
    try {
        if (d.size() >= 2 && d[0] == 0xff && d[1] == 0xd8) decodeJpeg(d, width, height, rgba);
        else if (d.size() >= 4 && d[1] == 'P' && d[2] == 'N' && d[3] == 'G') decodePng(d, width, height, rgba);
        else throw std::runtime_error("unsupported image format (JPEG and PNG are)");
    } catch (const std::exception &e) {
        throw std::runtime_error("Could not load texture file " + path + ": " + e.what());
    }
}

}  // namespace b200pt
