// jpeg_go.hpp — baseline JPEG decoder that reproduces Go's image/jpeg + image/color texel for texel.
//
// The reference imports textures with image.Decode and reads them back through img.At(x, y).RGBA() >> 8
// (internal/imageloader/imageLoader.go:28-84).  Go's decoder differs from libjpeg's in three places that move
// texels by up to 3/255: its inverse DCT (the fixed-point Chen-Wang transform of the MPEG reference decoder,
// image/jpeg/idct.go), no chroma interpolation (YCbCr.At reads the co-sited chroma sample, image/ycbcr.go), and its
// own fixed-point YCbCr -> RGB (color.YCbCr.RGBA, image/color/ycbcr.go).  go.mod pins go 1.22.3; the algorithm below
// restates that version's published behaviour for what the reference's assets need: 8-bit baseline (SOF0), Huffman,
// 1 or 3 components, interleaved scans, restart intervals.  Anything else (progressive, 12-bit, CMYK) is refused.
//
// Pinned by the reference's own golden vector: internal/imageloader/imageLoader_test.go:33-62 holds the 25 RGB
// texels Go decodes from test.jpg (a 4:2:0 file) — tests/test_jpeg_go.py decodes the same bytes with this code.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <string>
#include <vector>

namespace grt {
namespace jpeg {

struct Image {
    int width = 0, height = 0;
    std::vector<uint8_t> rgb;   // width * height * 3
};

namespace detail {

static const int kUnzig[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                               41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                               30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// image/jpeg/idct.go: 2-D IDCT of one 8x8 block, in place.  Horizontal pass keeps 3 extra bits, vertical pass
// scales back by 2^14; w_k = 2048 * sqrt(2) * cos(k pi / 16).
inline void idct(int32_t* b) {
    const int32_t w1 = 2841, w2 = 2676, w3 = 2408, w5 = 1609, w6 = 1108, w7 = 565;
    const int32_t w1pw7 = w1 + w7, w1mw7 = w1 - w7, w2pw6 = w2 + w6, w2mw6 = w2 - w6, w3pw5 = w3 + w5, w3mw5 = w3 - w5;
    const int32_t r2 = 181;   // 256 / sqrt(2)
    for (int y = 0; y < 8; y++) {
        int32_t* s = b + 8 * y;
        if (s[1] == 0 && s[2] == 0 && s[3] == 0 && s[4] == 0 && s[5] == 0 && s[6] == 0 && s[7] == 0) {
            int32_t dc = (int32_t)((uint32_t)s[0] << 3);
            for (int k = 0; k < 8; k++) s[k] = dc;
            continue;
        }
        int32_t x0 = (int32_t)((uint32_t)s[0] << 11) + 128, x1 = (int32_t)((uint32_t)s[4] << 11);
        int32_t x2 = s[6], x3 = s[2], x4 = s[1], x5 = s[7], x6 = s[5], x7 = s[3];
        int32_t x8 = w7 * (x4 + x5);
        x4 = x8 + w1mw7 * x4;
        x5 = x8 - w1pw7 * x5;
        x8 = w3 * (x6 + x7);
        x6 = x8 - w3mw5 * x6;
        x7 = x8 - w3pw5 * x7;
        x8 = x0 + x1;
        x0 -= x1;
        x1 = w6 * (x3 + x2);
        x2 = x1 - w2pw6 * x2;
        x3 = x1 + w2mw6 * x3;
        x1 = x4 + x6;
        x4 -= x6;
        x6 = x5 + x7;
        x5 -= x7;
        x7 = x8 + x3;
        x8 -= x3;
        x3 = x0 + x2;
        x0 -= x2;
        x2 = (r2 * (x4 + x5) + 128) >> 8;
        x4 = (r2 * (x4 - x5) + 128) >> 8;
        s[0] = (x7 + x1) >> 8;
        s[1] = (x3 + x2) >> 8;
        s[2] = (x0 + x4) >> 8;
        s[3] = (x8 + x6) >> 8;
        s[4] = (x8 - x6) >> 8;
        s[5] = (x0 - x4) >> 8;
        s[6] = (x3 - x2) >> 8;
        s[7] = (x7 - x1) >> 8;
    }
    for (int x = 0; x < 8; x++) {
        int32_t* s = b + x;
        int32_t y0 = (int32_t)((uint32_t)s[8 * 0] << 8) + 8192, y1 = (int32_t)((uint32_t)s[8 * 4] << 8);
        int32_t y2 = s[8 * 6], y3 = s[8 * 2], y4 = s[8 * 1], y5 = s[8 * 7], y6 = s[8 * 5], y7 = s[8 * 3];
        int32_t y8 = w7 * (y4 + y5) + 4;
        y4 = (y8 + w1mw7 * y4) >> 3;
        y5 = (y8 - w1pw7 * y5) >> 3;
        y8 = w3 * (y6 + y7) + 4;
        y6 = (y8 - w3mw5 * y6) >> 3;
        y7 = (y8 - w3pw5 * y7) >> 3;
        y8 = y0 + y1;
        y0 -= y1;
        y1 = w6 * (y3 + y2) + 4;
        y2 = (y1 - w2pw6 * y2) >> 3;
        y3 = (y1 + w2mw6 * y3) >> 3;
        y1 = y4 + y6;
        y4 -= y6;
        y6 = y5 + y7;
        y5 -= y7;
        y7 = y8 + y3;
        y8 -= y3;
        y3 = y0 + y2;
        y0 -= y2;
        y2 = (r2 * (y4 + y5) + 128) >> 8;
        y4 = (r2 * (y4 - y5) + 128) >> 8;
        s[8 * 0] = (y7 + y1) >> 14;
        s[8 * 1] = (y3 + y2) >> 14;
        s[8 * 2] = (y0 + y4) >> 14;
        s[8 * 3] = (y8 + y6) >> 14;
        s[8 * 4] = (y8 - y6) >> 14;
        s[8 * 5] = (y0 - y4) >> 14;
        s[8 * 6] = (y3 - y2) >> 14;
        s[8 * 7] = (y7 - y1) >> 14;
    }
}

// color.YCbCr.RGBA (image/color/ycbcr.go), then imageLoader.go:66-68's uint8(v >> 8)
inline void ycbcr_to_rgb8(uint8_t Y, uint8_t Cb, uint8_t Cr, uint8_t* out) {
    const int32_t yy1 = (int32_t)Y * 0x10101, cb1 = (int32_t)Cb - 128, cr1 = (int32_t)Cr - 128;
    int32_t c[3] = {yy1 + 91881 * cr1, yy1 - 22554 * cb1 - 46802 * cr1, yy1 + 116130 * cb1};
    for (int k = 0; k < 3; k++) {
        int32_t v = c[k];
        if (((uint32_t)v & 0xff000000u) == 0) v >>= 8;
        else v = ~(v >> 31) & 0xffff;
        out[k] = (uint8_t)((uint32_t)v >> 8);
    }
}

struct Huff {
    bool present = false;
    uint8_t counts[16];
    uint8_t vals[256];
    int32_t mincode[16], maxcode[16], valptr[16];
    void build() {
        int32_t code = 0, k = 0;
        for (int i = 0; i < 16; i++) {
            valptr[i] = k;
            mincode[i] = code;
            code += counts[i];
            k += counts[i];
            maxcode[i] = counts[i] ? code - 1 : -1;
            code <<= 1;
        }
        present = true;
    }
};

struct Reader {
    const uint8_t* p;
    size_t n, pos = 0;
    uint32_t acc = 0;
    int nbits = 0;
    bool eof = false;
    int marker = 0;   // a marker met inside entropy-coded data (RSTn / EOI)
    int bit() {
        if (nbits == 0) {
            if (marker || pos >= n) { eof = true; return 0; }
            uint8_t c = p[pos++];
            if (c == 0xFF) {
                if (pos >= n) { eof = true; return 0; }
                uint8_t c2 = p[pos++];
                if (c2 != 0) { marker = c2; eof = true; return 0; }   // not byte stuffing: a marker ends the segment
            }
            acc = c;
            nbits = 8;
        }
        nbits--;
        return (acc >> nbits) & 1;
    }
    int32_t bits(int k) { int32_t v = 0; for (int i = 0; i < k; i++) v = (v << 1) | bit(); return v; }
    void reset() { acc = 0; nbits = 0; eof = false; marker = 0; }
};

inline int decode_huff(Reader& r, const Huff& h) {
    int32_t code = 0;
    for (int i = 0; i < 16; i++) {
        code = (code << 1) | r.bit();
        if (h.maxcode[i] >= 0 && code <= h.maxcode[i] && code >= h.mincode[i]) return h.vals[h.valptr[i] + (code - h.mincode[i])];
    }
    return -1;
}
// RECEIVE and EXTEND, ITU T.81 F.2.2.1 (image/jpeg/huffman.go receiveExtend)
inline int32_t receive_extend(Reader& r, int t) {
    if (t == 0) return 0;
    int32_t v = r.bits(t);
    if (v < (1 << (t - 1))) v += (int32_t)((uint32_t)(-1) << t) + 1;
    return v;
}

}  // namespace detail

// Returns "" on success.
inline std::string decode(const uint8_t* data, size_t n, Image& out) {
    using namespace detail;
    if (n < 4 || data[0] != 0xFF || data[1] != 0xD8) return "missing SOI marker";
    uint16_t quant[4][64];
    bool have_q[4] = {false, false, false, false};
    Huff huff[2][4];
    struct Comp { int id, h, v, tq, td = 0, ta = 0; };
    std::vector<Comp> comp;
    int width = 0, height = 0, ri = 0;
    std::vector<uint8_t> plane[3];
    int stride[3] = {0, 0, 0};
    size_t pos = 2;
    bool done = false, saw_scan = false;
    while (!done) {
        if (pos + 2 > n) return "truncated before EOI";
        if (data[pos] != 0xFF) return "expected a marker";
        while (pos < n && data[pos] == 0xFF) pos++;   // fill bytes
        if (pos >= n) return "truncated";
        const int m = data[pos++];
        if (m == 0xD9) { done = true; break; }
        if (m >= 0xD0 && m <= 0xD7) continue;
        if (pos + 2 > n) return "truncated segment";
        const size_t len = ((size_t)data[pos] << 8) | data[pos + 1];
        if (len < 2 || pos + len > n) return "bad segment length";
        const uint8_t* seg = data + pos + 2;
        const size_t sl = len - 2;
        pos += len;
        if (m == 0xDB) {   // DQT
            size_t i = 0;
            while (i < sl) {
                const int pq = seg[i] >> 4, tq = seg[i] & 15;
                i++;
                if (tq > 3) return "bad quantisation table id";
                if (pq == 0) { if (i + 64 > sl) return "short DQT"; for (int k = 0; k < 64; k++) quant[tq][k] = seg[i + k]; i += 64; }
                else if (pq == 1) { if (i + 128 > sl) return "short DQT"; for (int k = 0; k < 64; k++) quant[tq][k] = (uint16_t)((seg[i + 2 * k] << 8) | seg[i + 2 * k + 1]); i += 128; }
                else return "bad quantisation table precision";
                have_q[tq] = true;
            }
        } else if (m == 0xC0 || m == 0xC1) {   // SOF0 / SOF1: baseline / extended sequential, Huffman
            if (sl < 6) return "short SOF";
            if (seg[0] != 8) return "only 8-bit precision is supported";
            height = (seg[1] << 8) | seg[2];
            width = (seg[3] << 8) | seg[4];
            const int nc = seg[5];
            if (nc != 1 && nc != 3) return "only 1- or 3-component images are supported";
            if (sl < (size_t)(6 + 3 * nc) || width <= 0 || height <= 0) return "bad SOF";
            comp.resize(nc);
            for (int i = 0; i < nc; i++) {
                comp[i].id = seg[6 + 3 * i];
                comp[i].h = seg[7 + 3 * i] >> 4;
                comp[i].v = seg[7 + 3 * i] & 15;
                comp[i].tq = seg[8 + 3 * i];
                if (comp[i].h < 1 || comp[i].h > 4 || comp[i].v < 1 || comp[i].v > 4 || comp[i].tq > 3) return "bad component";
            }
            if (nc == 1) comp[0].h = comp[0].v = 1;   // image/jpeg/reader.go: a single component is never subsampled
            if (nc == 3 && (comp[1].h != comp[2].h || comp[1].v != comp[2].v || comp[0].h % comp[1].h || comp[0].v % comp[1].v)) return "unsupported chroma subsampling";
        } else if (m == 0xC2 || (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)) {
            return "only baseline (SOF0) JPEG is supported (the reference's assets are baseline)";
        } else if (m == 0xC4) {   // DHT
            size_t i = 0;
            while (i < sl) {
                if (i + 17 > sl) return "short DHT";
                const int tc = seg[i] >> 4, th = seg[i] & 15;
                if (tc > 1 || th > 3) return "bad Huffman table id";
                Huff& h = huff[tc][th];
                int total = 0;
                for (int k = 0; k < 16; k++) { h.counts[k] = seg[i + 1 + k]; total += h.counts[k]; }
                i += 17;
                if (total > 256 || i + total > sl) return "bad DHT";
                memcpy(h.vals, seg + i, total);
                i += total;
                h.build();
            }
        } else if (m == 0xDD) {   // DRI
            if (sl < 2) return "short DRI";
            ri = (seg[0] << 8) | seg[1];
        } else if (m == 0xDA) {   // SOS
            if (comp.empty()) return "SOS before SOF";
            const int ns = seg[0];
            if (ns != (int)comp.size()) return "only fully interleaved scans are supported";
            if (sl < (size_t)(4 + 2 * ns)) return "short SOS";
            for (int i = 0; i < ns; i++) {
                int cs = seg[1 + 2 * i], k = -1;
                for (size_t c = 0; c < comp.size(); c++) if (comp[c].id == cs) k = (int)c;
                if (k != i) return "scan components out of order";
                comp[i].td = seg[2 + 2 * i] >> 4;
                comp[i].ta = seg[2 + 2 * i] & 15;
                if (comp[i].td > 3 || comp[i].ta > 3 || !huff[0][comp[i].td].present || !huff[1][comp[i].ta].present) return "missing Huffman table";
                if (!have_q[comp[i].tq]) return "missing quantisation table";
            }
            const int h0 = comp[0].h, v0 = comp[0].v;
            const int mxx = (width + 8 * h0 - 1) / (8 * h0), myy = (height + 8 * v0 - 1) / (8 * v0);
            for (size_t c = 0; c < comp.size(); c++) {
                stride[c] = mxx * 8 * comp[c].h;
                plane[c].assign((size_t)stride[c] * myy * 8 * comp[c].v, 0);
            }
            Reader r{data, n};
            r.pos = pos;
            int32_t dc[3] = {0, 0, 0};
            int mcu = 0, expected_rst = 0;
            for (int my = 0; my < myy; my++)
                for (int mx = 0; mx < mxx; mx++) {
                    for (size_t c = 0; c < comp.size(); c++) {
                        const Comp& cc = comp[c];
                        for (int j = 0; j < cc.h * cc.v; j++) {
                            const int bx = cc.h * mx + j % cc.h, by = cc.v * my + j / cc.h;
                            int32_t b[64];
                            memset(b, 0, sizeof(b));
                            int t = decode_huff(r, huff[0][cc.td]);
                            if (t < 0 || t > 16 || r.eof) return "bad DC code";
                            dc[c] += receive_extend(r, t);
                            b[0] = dc[c] * (int32_t)quant[cc.tq][0];
                            for (int zig = 1; zig < 64; zig++) {
                                int value = decode_huff(r, huff[1][cc.ta]);
                                if (value < 0 || r.eof) return "bad AC code";
                                const int val0 = value >> 4, val1 = value & 15;
                                if (val1 != 0) {
                                    zig += val0;
                                    if (zig > 63) break;
                                    b[kUnzig[zig]] = receive_extend(r, val1) * (int32_t)quant[cc.tq][zig];
                                } else {
                                    if (val0 != 0x0f) break;   // EOB
                                    zig += 0x0f;
                                }
                            }
                            idct(b);
                            // level shift by +128, clip to [0, 255] (image/jpeg/scan.go reconstructBlock)
                            uint8_t* dst = plane[c].data() + (size_t)8 * ((size_t)by * stride[c] + bx);
                            for (int y = 0; y < 8; y++)
                                for (int x = 0; x < 8; x++) {
                                    int32_t v = b[8 * y + x];
                                    v = v < -128 ? 0 : (v > 127 ? 255 : v + 128);
                                    dst[(size_t)y * stride[c] + x] = (uint8_t)v;
                                }
                        }
                    }
                    mcu++;
                    if (ri > 0 && mcu % ri == 0 && mcu < mxx * myy) {
                        // a restart marker follows: byte-align, expect RSTn, reset the DC predictors
                        if (!r.marker) {
                            r.nbits = 0;
                            if (r.pos + 2 > n || data[r.pos] != 0xFF) return "missing restart marker";
                            r.marker = data[r.pos + 1];
                            r.pos += 2;
                        }
                        if (r.marker != 0xD0 + expected_rst) return "bad restart marker";
                        expected_rst = (expected_rst + 1) & 7;
                        r.reset();
                        dc[0] = dc[1] = dc[2] = 0;
                    }
                }
            pos = r.pos;
            if (r.marker) pos -= 2;   // hand the marker back to the segment loop
            saw_scan = true;
        }
        // every other segment (APPn, COM, ...) is skipped
    }
    if (!saw_scan) return "no scan data";
    out.width = width;
    out.height = height;
    out.rgb.assign((size_t)width * height * 3, 0);
    if (comp.size() == 1) {   // image.Gray: RGBA() replicates Y
        for (int y = 0; y < height; y++)
            for (int x = 0; x < width; x++) {
                uint8_t v = plane[0][(size_t)y * stride[0] + x];
                uint8_t* o = out.rgb.data() + ((size_t)y * width + x) * 3;
                o[0] = o[1] = o[2] = v;
            }
        return "";
    }
    // image.YCbCr.At: the chroma sample is the co-sited one of the subsampled plane, no interpolation (image/ycbcr.go COffset)
    const int hr = comp[0].h / comp[1].h, vr = comp[0].v / comp[1].v;
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            const uint8_t Y = plane[0][(size_t)y * stride[0] + x];
            const size_t co = (size_t)(y / vr) * stride[1] + (x / hr);
            detail::ycbcr_to_rgb8(Y, plane[1][co], plane[2][co], out.rgb.data() + ((size_t)y * width + x) * 3);
        }
    return "";
}

}  // namespace jpeg
}  // namespace grt
