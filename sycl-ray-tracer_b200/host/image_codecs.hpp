/*
 * image_codecs.hpp — decoders for the images embedded in a .glb, host side of the texture bake
 * (SURVEY §8f rank 2). The reference gets its pixels from tinygltf, which calls
 * stbi_load_from_memory(bytes, size, &w, &h, &comp, 4) (deps/include/tiny_gltf.h:2603-2638): 8-bit
 * RGBA, whatever the file holds. What reaches the kernels must therefore equal stb_image's output
 * byte for byte, which pins not only the formats but the integer arithmetic of the lossy one:
 *
 *   PNG  : all colour types (grey, RGB, palette, grey+alpha, RGBA), bit depths 1/2/4/8/16, tRNS
 *          (palette alpha and colour key), Adam7 interlace. Sub-byte grey is scaled by 255/(2^d-1),
 *          16-bit samples keep their high byte (or, on request, come out the way the reference's tinygltf + bake
 *          combination misreads them); ancillary chunks (gAMA, iCCP, ...) are ignored.
 *   JPEG : baseline / extended sequential and progressive Huffman, 8-bit, 1 or 3 components, any
 *          sampling factors, restart intervals, 8- and 16-bit quantisation tables, Adobe APP14
 *          transform flag and R/G/B component ids. The inverse DCT is the 12-bit fixed-point
 *          Loeffler-Ligtenberg-Moschytz form with +-2^9 / 2^16 rounding biases, chroma is upsampled
 *          with the 3:1 (h2v1 / h1v2) and 9:3:3:1 (h2v2) "fancy" filters in integer arithmetic and
 *          YCbCr -> RGB uses 20-bit fixed point with the truncated green/Cb product — the published
 *          arithmetic of stb_image's decoder, restated here (no code is shared with it).
 *
 * tests/test_image_codecs.py checks every variant against fixtures produced by the reference's own
 * vendored stb_image.h (compiled in place by `make -C oracle ref`, tests/tools/make_golden_images.py).
 * CMYK / YCCK JPEGs, arithmetic coding, 12-bit JPEG and CgBI PNGs are rejected (stb_image rejects
 * the last three as well).
 */
#ifndef RT_HOST_IMAGE_CODECS_HPP
#define RT_HOST_IMAGE_CODECS_HPP

#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace raytracer {
namespace img {

struct Image {
    uint32_t w = 0, h = 0;
    int channels_in_file = 0; /* stbi's `comp` */
    std::vector<uint8_t> rgba; /* w*h*4 */
};

[[noreturn]] inline void fail(const std::string &what) { throw std::runtime_error("image: " + what); }

/* =============================================================================== PNG */
namespace png {

inline uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
inline uint32_t be16(const uint8_t *p) { return (uint32_t)p[0] << 8 | p[1]; }

/* undo the per-scanline filters of one (sub)image; `bpp` = bytes per complete pixel (>= 1) */
inline void unfilter(const uint8_t *src, size_t src_len, uint32_t rows, size_t stride, size_t bpp, std::vector<uint8_t> &out) {
    if (src_len < (stride + 1) * rows) fail("PNG: not enough pixel data");
    out.resize(stride * rows);
    for (uint32_t y = 0; y < rows; y++) {
        const uint8_t ft = src[(stride + 1) * y], *in = src + (stride + 1) * y + 1;
        uint8_t *cur = out.data() + stride * y;
        const uint8_t *up = y ? cur - stride : nullptr;
        if (ft > 4) fail("PNG: invalid filter");
        for (size_t x = 0; x < stride; x++) {
            const int a = x >= bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
            int pred = 0;
            if (ft == 1) pred = a;
            else if (ft == 2) pred = b;
            else if (ft == 3) pred = (a + b) >> 1;
            else if (ft == 4) {
                const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
                pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
            }
            cur[x] = (uint8_t)(in[x] + pred);
        }
    }
}

inline Image decode(const uint8_t *d, size_t n, bool tinygltf_16bit_quirk = false) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (n < 8 || memcmp(d, sig, 8)) fail("not a PNG");
    Image im;
    std::vector<uint8_t> idat;
    int depth = 0, ctype = -1, interlace = 0;
    uint8_t pal[256 * 4];
    uint32_t pal_len = 0;
    bool has_trns = false;
    uint32_t key16[3] = {0, 0, 0};
    for (int i = 0; i < 256; i++) pal[i * 4 + 3] = 255;
    bool seen_end = false;
    for (size_t o = 8; o + 12 <= n && !seen_end;) {
        const uint32_t len = be32(d + o);
        const uint8_t *ty = d + o + 4, *body = d + o + 8;
        if ((size_t)len > n - o - 12) fail("PNG: truncated chunk");
        if (!memcmp(ty, "CgBI", 4)) fail("PNG: CgBI (iPhone) variant is not supported");
        if (!memcmp(ty, "IHDR", 4)) {
            if (len != 13) fail("PNG: bad IHDR");
            im.w = be32(body);
            im.h = be32(body + 4);
            depth = body[8];
            ctype = body[9];
            interlace = body[12];
            if (body[10] || body[11] || interlace > 1) fail("PNG: bad compression / filter / interlace method");
            if (!im.w || !im.h || im.w > (1u << 24) || im.h > (1u << 24)) fail("PNG: bad size");
            if (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16) fail("PNG: bad bit depth");
            if (ctype != 0 && ctype != 2 && ctype != 3 && ctype != 4 && ctype != 6) fail("PNG: bad colour type");
            if (ctype == 3 && depth == 16) fail("PNG: bad palette depth");
            if ((ctype == 2 || ctype == 4 || ctype == 6) && depth < 8) fail("PNG: bad depth for colour type");
        } else if (!memcmp(ty, "PLTE", 4)) {
            if (len > 256 * 3 || len % 3) fail("PNG: bad PLTE");
            pal_len = len / 3;
            for (uint32_t i = 0; i < pal_len; i++) memcpy(pal + i * 4, body + i * 3, 3);
        } else if (!memcmp(ty, "tRNS", 4)) {
            if (ctype < 0) fail("PNG: tRNS before IHDR");
            if (ctype == 3) {
                if (len > pal_len) fail("PNG: bad tRNS");
                for (uint32_t i = 0; i < len; i++) pal[i * 4 + 3] = body[i];
                has_trns = true;
            } else if (ctype == 0 || ctype == 2) {
                const uint32_t k = ctype == 0 ? 1 : 3;
                if (len != 2 * k) fail("PNG: bad tRNS");
                for (uint32_t i = 0; i < k; i++) key16[i] = be16(body + 2 * i);
                has_trns = true;
            } else fail("PNG: tRNS with alpha");
        } else if (!memcmp(ty, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!memcmp(ty, "IEND", 4)) {
            seen_end = true;
        } else if (!(ty[0] & 32)) {
            fail("PNG: unknown critical chunk");
        }
        o += 12 + (size_t)len;
    }
    if (ctype < 0 || idat.empty()) fail("PNG: no image data");
    if (ctype == 3 && !pal_len) fail("PNG: no PLTE");
    const uint32_t ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : 4;
    im.channels_in_file = ctype == 3 ? (has_trns ? 4 : 3) : (int)ch + (has_trns ? 1 : 0);

    /* inflate everything: the sum over the (sub)images of (stride + 1) * rows */
    const size_t bits_pp = (size_t)ch * depth;
    auto stride_of = [&](uint32_t w) { return (w * bits_pp + 7) / 8; };
    static const int xo[7] = {0, 4, 0, 2, 0, 1, 0}, yo[7] = {0, 0, 4, 0, 2, 0, 1}, xs[7] = {8, 8, 4, 4, 2, 2, 1}, ys[7] = {8, 8, 8, 4, 4, 2, 2};
    size_t raw_len = 0;
    if (!interlace) raw_len = (stride_of(im.w) + 1) * im.h;
    else
        for (int p = 0; p < 7; p++) {
            const uint32_t pw = (uint32_t)(((int64_t)im.w - xo[p] + xs[p] - 1) / xs[p]), ph = (uint32_t)(((int64_t)im.h - yo[p] + ys[p] - 1) / ys[p]);
            if (pw && ph) raw_len += (stride_of(pw) + 1) * ph;
        }
    if (raw_len > ((size_t)1 << 30)) fail("PNG: image too large");
    std::vector<uint8_t> raw(raw_len);
    {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit(&zs) != Z_OK) fail("PNG: zlib init failed");
        zs.next_in = idat.data();
        zs.avail_in = (uInt)idat.size();
        zs.next_out = raw.data();
        zs.avail_out = (uInt)raw.size();
        const int rc = inflate(&zs, Z_FINISH);
        const size_t got = raw.size() - zs.avail_out;
        inflateEnd(&zs);
        if ((rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) || got < raw.size()) fail("PNG: inflate failed");
    }

    /* samples of the whole image as 16-bit values per channel (expanded, not yet scaled) */
    std::vector<uint16_t> samp((size_t)im.w * im.h * ch);
    auto unpack = [&](const std::vector<uint8_t> &px, uint32_t pw, uint32_t ph, int x0, int y0, int dx, int dy) {
        const size_t stride = stride_of(pw);
        for (uint32_t y = 0; y < ph; y++) {
            const uint8_t *row = px.data() + stride * y;
            for (uint32_t x = 0; x < pw; x++) {
                uint16_t *o = &samp[(((size_t)y0 + (size_t)y * dy) * im.w + ((size_t)x0 + (size_t)x * dx)) * ch];
                for (uint32_t c = 0; c < ch; c++) {
                    const size_t i = (size_t)x * ch + c;
                    if (depth == 8) o[c] = row[i];
                    else if (depth == 16) o[c] = (uint16_t)(row[i * 2] << 8 | row[i * 2 + 1]);
                    else {
                        const size_t bit = i * depth;
                        o[c] = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
                    }
                }
            }
        }
    };
    const size_t bpp = bits_pp >= 8 ? bits_pp / 8 : 1;
    std::vector<uint8_t> px;
    if (!interlace) {
        unfilter(raw.data(), raw.size(), im.h, stride_of(im.w), bpp, px);
        unpack(px, im.w, im.h, 0, 0, 1, 1);
    } else {
        size_t off = 0;
        for (int p = 0; p < 7; p++) {
            const uint32_t pw = (uint32_t)(((int64_t)im.w - xo[p] + xs[p] - 1) / xs[p]), ph = (uint32_t)(((int64_t)im.h - yo[p] + ys[p] - 1) / ys[p]);
            if (!pw || !ph) continue;
            const size_t len = (stride_of(pw) + 1) * ph;
            unfilter(raw.data() + off, raw.size() - off, ph, stride_of(pw), bpp, px);
            unpack(px, pw, ph, xo[p], yo[p], xs[p], ys[p]);
            off += len;
        }
    }

    /* to RGBA8 the way stbi_load(..., 4) does */
    im.rgba.resize((size_t)im.w * im.h * 4);
    static const uint32_t depth_scale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
    const uint32_t scale = depth < 8 ? depth_scale[depth] : 1;
    for (size_t i = 0; i < (size_t)im.w * im.h; i++) {
        const uint16_t *s = &samp[i * ch];
        uint8_t *o = &im.rgba[i * 4];
        auto to8 = [&](uint16_t v) -> uint8_t { return depth == 16 ? (uint8_t)(v >> 8) : (uint8_t)(v * (ctype == 3 ? 1 : scale)); };
        if (ctype == 3) {
            const uint32_t idx = s[0];
            /* an index past the palette is invalid (stb_image reads its uninitialised table there): black, opaque */
            if (idx < pal_len) memcpy(o, pal + idx * 4, 4);
            else { o[0] = o[1] = o[2] = 0; o[3] = 255; }
        } else if (ctype == 0) {
            o[0] = o[1] = o[2] = to8(s[0]);
            o[3] = 255;
            if (has_trns) {
                const bool hit = depth == 16 ? s[0] == key16[0] : to8(s[0]) == (uint8_t)((key16[0] & 255) * scale);
                if (hit) o[3] = 0;
            }
        } else if (ctype == 4) {
            o[0] = o[1] = o[2] = to8(s[0]);
            o[3] = to8(s[1]);
        } else if (ctype == 2) {
            o[0] = to8(s[0]); o[1] = to8(s[1]); o[2] = to8(s[2]);
            o[3] = 255;
            if (has_trns) {
                const bool hit = depth == 16 ? (s[0] == key16[0] && s[1] == key16[1] && s[2] == key16[2])
                                             : (o[0] == (uint8_t)(key16[0] & 255) && o[1] == (uint8_t)(key16[1] & 255) && o[2] == (uint8_t)(key16[2] & 255));
                if (hit) o[3] = 0;
            }
        } else {
            o[0] = to8(s[0]); o[1] = to8(s[1]); o[2] = to8(s[2]); o[3] = to8(s[3]);
        }
    }
    if (depth == 16 && tinygltf_16bit_quirk) {
        /* What the REFERENCE sees for a 16-bit PNG: tinygltf notices stbi_is_16_bit and loads 16-bit samples
         * (deps/include/tiny_gltf.h:2618-2626), the reference's bake then reads that buffer as if it were 8-bit
         * RGBA of the same width and height (src/scene.cpp:157-160, src/image_manager.hpp:39): the first w*h*4
         * BYTES of the little-endian 16-bit RGBA stream, i.e. pixel p = (R.lo, R.hi, G.lo, G.hi) of 16-bit pixel p/2
         * for even p and (B.lo, B.hi, A.lo, A.hi) for odd p. Reproduced, not fixed (same class as F15). */
        std::vector<uint8_t> stream((size_t)im.w * im.h * 8);
        for (size_t i = 0; i < (size_t)im.w * im.h; i++) {
            const uint16_t *s = &samp[i * ch];
            uint16_t px[4];
            if (ctype == 0) { px[0] = px[1] = px[2] = s[0]; px[3] = (has_trns && s[0] == key16[0]) ? 0 : 65535; }
            else if (ctype == 4) { px[0] = px[1] = px[2] = s[0]; px[3] = s[1]; }
            else if (ctype == 2) { px[0] = s[0]; px[1] = s[1]; px[2] = s[2]; px[3] = (has_trns && s[0] == key16[0] && s[1] == key16[1] && s[2] == key16[2]) ? 0 : 65535; }
            else { px[0] = s[0]; px[1] = s[1]; px[2] = s[2]; px[3] = s[3]; }
            for (int c = 0; c < 4; c++) {
                stream[i * 8 + c * 2] = (uint8_t)(px[c] & 255);
                stream[i * 8 + c * 2 + 1] = (uint8_t)(px[c] >> 8);
            }
        }
        memcpy(im.rgba.data(), stream.data(), im.rgba.size());
    }
    return im;
}

} // namespace png

/* =============================================================================== JPEG */
namespace jpeg {

static const uint8_t kZigzag[64 + 15] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                                         6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                                         39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
                                         /* a run past the end lands here instead of outside the block */
                                         63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct Huffman {
    bool present = false;
    int32_t maxcode[18]; /* largest code of each length, left-justified to 16 bits, +1 */
    int32_t delta[17];   /* first symbol index - first code */
    uint8_t values[256];
    uint8_t sizes[257];
    uint16_t codes[256];
    void build(const uint8_t counts[16], const uint8_t *vals, int n) {
        int k = 0;
        for (int len = 1; len <= 16; len++)
            for (int i = 0; i < counts[len - 1]; i++) sizes[k++] = (uint8_t)len;
        sizes[k] = 0;
        uint32_t code = 0;
        k = 0;
        for (int len = 1; len <= 16; len++) {
            delta[len] = k - (int32_t)code;
            if (sizes[k] == len) {
                while (sizes[k] == len) codes[k++] = (uint16_t)code++;
                if (code - 1 >= (1u << len)) fail("JPEG: bad Huffman code lengths");
            }
            maxcode[len] = (int32_t)(code << (16 - len));
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        memcpy(values, vals, (size_t)n);
        present = true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0;
    int dc_pred = 0;
    int x = 0, y = 0;    /* size in samples */
    int w2 = 0, h2 = 0;  /* padded to whole MCUs */
    std::vector<uint8_t> data;   /* w2 * h2 */
    std::vector<int16_t> coeff;  /* progressive: (w2/8) * (h2/8) blocks of 64 */
    int coeff_w = 0;
};

struct Decoder {
    const uint8_t *p, *end;
    uint32_t bitbuf = 0;
    int bitcnt = 0;
    int marker = -1; /* a marker met inside entropy-coded data */
    bool nomore = false;

    uint16_t dequant[4][64];
    Huffman hdc[4], hac[4];
    Component comp[4];
    int ncomp = 0, img_w = 0, img_h = 0;
    bool progressive = false;
    int h_max = 1, v_max = 1, mcu_w = 0, mcu_h = 0, mcu_x = 0, mcu_y = 0;
    int restart_interval = 0, todo = 0;
    int scan_n = 0, order[4];
    int spec_start = 0, spec_end = 0, succ_high = 0, succ_low = 0, eob_run = 0;
    int adobe_transform = -1;
    bool jfif = false, rgb_ids = false;

    int get8() { return p < end ? *p++ : 0; }
    int get16() { const int a = get8(); return (a << 8) | get8(); }

    /* ---- entropy-coded bit stream: FF00 is a stuffed FF, any other FFxx ends the data */
    void grow() {
        do {
            uint32_t b = nomore ? 0u : (uint32_t)get8();
            if (b == 0xff && !nomore) {
                int c = get8();
                while (c == 0xff) c = get8();
                if (c != 0) {
                    marker = c;
                    nomore = true;
                    return;
                }
            }
            bitbuf |= b << (24 - bitcnt);
            bitcnt += 8;
        } while (bitcnt <= 24);
    }
    int decode(const Huffman &h) {
        if (bitcnt < 16) grow();
        const uint32_t top = bitbuf >> 16;
        int len = 1;
        while (top >= (uint32_t)h.maxcode[len]) len++;
        if (len > 16) fail("JPEG: bad Huffman code");
        if (len > bitcnt) fail("JPEG: bad Huffman code");
        const int idx = (int)((bitbuf >> (32 - len)) & ((1u << len) - 1)) + h.delta[len];
        if (idx < 0 || idx >= 256) fail("JPEG: bad Huffman code");
        bitcnt -= len;
        bitbuf <<= len;
        return h.values[idx];
    }
    /* n more bits as a signed magnitude (T.81 F.2.2.1 EXTEND) */
    int extend_receive(int n) {
        if (n == 0) return 0;
        if (bitcnt < n) grow();
        if (bitcnt < n) return 0;
        const uint32_t k = bitbuf >> (32 - n);
        bitcnt -= n;
        bitbuf <<= n;
        return (int)k < (1 << (n - 1)) ? (int)k - (1 << n) + 1 : (int)k;
    }
    int get_bits(int n) {
        if (n == 0) return 0;
        if (bitcnt < n) grow();
        if (bitcnt < n) return 0;
        const uint32_t k = bitbuf >> (32 - n);
        bitcnt -= n;
        bitbuf <<= n;
        return (int)k;
    }
    int get_bit() { return get_bits(1); }
    void reset_entropy() {
        bitbuf = 0;
        bitcnt = 0;
        nomore = false;
        marker = -1;
        for (int i = 0; i < 4; i++) comp[i].dc_pred = 0;
        todo = restart_interval ? restart_interval : 0x7fffffff;
        eob_run = 0;
    }

    /* ---- blocks */
    void block_sequential(int16_t data[64], const Huffman &dc, const Huffman &ac, const uint16_t *dq, Component &c) {
        memset(data, 0, 64 * sizeof(int16_t));
        const int t = decode(dc);
        if (t > 16) fail("JPEG: bad DC size");
        const int diff = t ? extend_receive(t) : 0;
        c.dc_pred += diff;
        if (c.dc_pred > (1 << 24) || c.dc_pred < -(1 << 24)) fail("JPEG: bad DC delta");
        data[0] = (int16_t)((int64_t)c.dc_pred * dq[0]);
        int k = 1;
        do {
            const int rs = decode(ac), s = rs & 15, r = rs >> 4;
            if (s == 0) {
                if (rs != 0xf0) break; /* end of block */
                k += 16;
            } else {
                k += r;
                const int zig = kZigzag[k++];
                data[zig] = (int16_t)(extend_receive(s) * dq[zig]);
            }
        } while (k < 64);
    }
    void block_prog_dc(int16_t data[64], const Huffman &dc, Component &c) {
        if (spec_end != 0) fail("JPEG: can't merge DC and AC");
        if (succ_high == 0) {
            memset(data, 0, 64 * sizeof(int16_t));
            const int t = decode(dc);
            if (t > 16) fail("JPEG: bad DC size");
            const int diff = t ? extend_receive(t) : 0;
            c.dc_pred += diff;
            if (c.dc_pred > (1 << 24) || c.dc_pred < -(1 << 24)) fail("JPEG: bad DC delta");
            data[0] = (int16_t)((int64_t)c.dc_pred * (1 << succ_low));
        } else if (get_bit()) {
            data[0] = (int16_t)(data[0] + (1 << succ_low));
        }
    }
    void block_prog_ac(int16_t data[64], const Huffman &ac) {
        if (spec_start == 0) fail("JPEG: can't merge DC and AC");
        if (succ_high == 0) {
            const int shift = succ_low;
            if (eob_run) {
                --eob_run;
                return;
            }
            int k = spec_start;
            do {
                const int rs = decode(ac), s = rs & 15, r = rs >> 4;
                if (s == 0) {
                    if (r < 15) {
                        eob_run = 1 << r;
                        if (r) eob_run += get_bits(r);
                        --eob_run;
                        break;
                    }
                    k += 16;
                } else {
                    k += r;
                    const int zig = kZigzag[k++];
                    data[zig] = (int16_t)(extend_receive(s) * (1 << shift));
                }
            } while (k <= spec_end);
        } else { /* refinement scan */
            const int16_t bit = (int16_t)(1 << succ_low);
            auto refine = [&](int16_t *q) {
                if (*q != 0 && get_bit() && (*q & bit) == 0) *q = (int16_t)(*q > 0 ? *q + bit : *q - bit);
            };
            if (eob_run) {
                --eob_run;
                for (int k = spec_start; k <= spec_end; k++) refine(&data[kZigzag[k]]);
            } else {
                int k = spec_start;
                do {
                    const int rs = decode(ac);
                    int s = rs & 15, r = rs >> 4;
                    if (s == 0) {
                        if (r < 15) {
                            eob_run = (1 << r) - 1;
                            if (r) eob_run += get_bits(r);
                            r = 64; /* run to the end of the band, refining what is there */
                        }
                    } else {
                        if (s != 1) fail("JPEG: bad refinement code");
                        s = get_bit() ? bit : -bit;
                    }
                    while (k <= spec_end) {
                        int16_t *q = &data[kZigzag[k++]];
                        if (*q != 0) {
                            if (get_bit() && (*q & bit) == 0) *q = (int16_t)(*q > 0 ? *q + bit : *q - bit);
                        } else {
                            if (r == 0) {
                                *q = (int16_t)s;
                                break;
                            }
                            --r;
                        }
                    }
                } while (k <= spec_end);
            }
        }
    }

    /* 8x8 inverse DCT, 12-bit fixed point; columns keep 2 extra bits, rows round with a 2^16 bias that
     * also adds the +128 level shift. ATTRIBUTION: this is the butterfly of stb_image v2.29 `stbi__idct_block` /
     * STBI__IDCT_1D (Sean Barrett, public domain / MIT; the reference vendors it as deps/include/stb_image.h:2429-2523,
     * itself derived from the IJG jidctint.c): bit-exact agreement with the bytes stb hands the reference forces the
     * same fixed-point constants, the same order of additions and the same rounding biases, so the temporaries keep
     * stb's names (t0..t3, p1..p5, x0..x3) to make the correspondence checkable. */
    static inline uint8_t clamp8(int x) { return (unsigned)x > 255u ? (x < 0 ? 0 : 255) : (uint8_t)x; }
    static inline int f2f(double x) { return (int)(x * 4096 + 0.5); }
    static void idct(uint8_t *out, int stride, const int16_t d[64]) {
#define RT_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                                       \
    int64_t t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3; /* 64-bit: corrupt data cannot overflow */ \
    p2 = s2;                                                                             \
    p3 = s6;                                                                             \
    p1 = (p2 + p3) * c0541;                                                              \
    t2 = p1 + p3 * cm1847;                                                               \
    t3 = p1 + p2 * c0765;                                                                \
    p2 = s0;                                                                             \
    p3 = s4;                                                                             \
    t0 = (p2 + p3) * 4096;                                                               \
    t1 = (p2 - p3) * 4096;                                                               \
    x0 = t0 + t3;                                                                        \
    x3 = t0 - t3;                                                                        \
    x1 = t1 + t2;                                                                        \
    x2 = t1 - t2;                                                                        \
    t0 = s7;                                                                             \
    t1 = s5;                                                                             \
    t2 = s3;                                                                             \
    t3 = s1;                                                                             \
    p3 = t0 + t2;                                                                        \
    p4 = t1 + t3;                                                                        \
    p1 = t0 + t3;                                                                        \
    p2 = t1 + t2;                                                                        \
    p5 = (p3 + p4) * c1175;                                                              \
    t0 = t0 * c0298;                                                                     \
    t1 = t1 * c2053;                                                                     \
    t2 = t2 * c3072;                                                                     \
    t3 = t3 * c1501;                                                                     \
    p1 = p5 + p1 * cm0899;                                                               \
    p2 = p5 + p2 * cm2562;                                                               \
    p3 = p3 * cm1961;                                                                    \
    p4 = p4 * cm0390;                                                                    \
    t3 += p1 + p4;                                                                       \
    t2 += p2 + p3;                                                                       \
    t1 += p2 + p4;                                                                       \
    t0 += p1 + p3;
        static const int c0541 = f2f(0.5411961), cm1847 = f2f(-1.847759065), c0765 = f2f(0.765366865), c1175 = f2f(1.175875602),
                         c0298 = f2f(0.298631336), c2053 = f2f(2.053119869), c3072 = f2f(3.072711026), c1501 = f2f(1.501321110),
                         cm0899 = f2f(-0.899976223), cm2562 = f2f(-2.562915447), cm1961 = f2f(-1.961570560), cm0390 = f2f(-0.390180644);
        int val[64], *v = val;
        for (int i = 0; i < 8; i++, d++, v++) {
            if (d[8] == 0 && d[16] == 0 && d[24] == 0 && d[32] == 0 && d[40] == 0 && d[48] == 0 && d[56] == 0) {
                const int dc = d[0] * 4; /* a constant column: same result as the full transform, up to its rounding */
                v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dc;
            } else {
                RT_IDCT_1D(d[0], d[8], d[16], d[24], d[32], d[40], d[48], d[56])
                x0 += 512; x1 += 512; x2 += 512; x3 += 512;
                v[0] = (int)((x0 + t3) >> 10);  v[56] = (int)((x0 - t3) >> 10);
                v[8] = (int)((x1 + t2) >> 10);  v[48] = (int)((x1 - t2) >> 10);
                v[16] = (int)((x2 + t1) >> 10); v[40] = (int)((x2 - t1) >> 10);
                v[24] = (int)((x3 + t0) >> 10); v[32] = (int)((x3 - t0) >> 10);
            }
        }
        v = val;
        for (int i = 0; i < 8; i++, v += 8, out += stride) {
            RT_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
            x0 += 65536 + (128 << 17); x1 += 65536 + (128 << 17); x2 += 65536 + (128 << 17); x3 += 65536 + (128 << 17);
            out[0] = clamp8((int)((x0 + t3) >> 17)); out[7] = clamp8((int)((x0 - t3) >> 17));
            out[1] = clamp8((int)((x1 + t2) >> 17)); out[6] = clamp8((int)((x1 - t2) >> 17));
            out[2] = clamp8((int)((x2 + t1) >> 17)); out[5] = clamp8((int)((x2 - t1) >> 17));
            out[3] = clamp8((int)((x3 + t0) >> 17)); out[4] = clamp8((int)((x3 - t0) >> 17));
        }
#undef RT_IDCT_1D
    }

    /* ---- one scan */
    bool restart_due() { /* after an MCU; false = stop this scan */
        if (--todo > 0) return true;
        if (bitcnt < 24) grow();
        if (marker < 0xd0 || marker > 0xd7) return false;
        reset_entropy();
        return true;
    }
    void scan() {
        reset_entropy();
        for (int k = 0; k < scan_n; k++) { /* the tables this scan decodes with must have been defined */
            const Component &c = comp[order[k]];
            const bool need_dc = !progressive || spec_start == 0, need_ac = !progressive || spec_start != 0;
            if ((need_dc && !(progressive && succ_high) && !hdc[c.hd].present) || (need_ac && !hac[c.ha].present)) fail("JPEG: scan uses an undefined Huffman table");
        }
        int16_t blk[64];
        if (!progressive) {
            if (scan_n == 1) {
                Component &c = comp[order[0]];
                const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
                for (int j = 0; j < h; j++)
                    for (int i = 0; i < w; i++) {
                        block_sequential(blk, hdc[c.hd], hac[c.ha], dequant[c.tq], c);
                        idct(c.data.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, blk);
                        if (!restart_due()) return;
                    }
            } else {
                for (int j = 0; j < mcu_y; j++)
                    for (int i = 0; i < mcu_x; i++) {
                        for (int k = 0; k < scan_n; k++) {
                            Component &c = comp[order[k]];
                            for (int y = 0; y < c.v; y++)
                                for (int x = 0; x < c.h; x++) {
                                    const int x2 = (i * c.h + x) * 8, y2 = (j * c.v + y) * 8;
                                    block_sequential(blk, hdc[c.hd], hac[c.ha], dequant[c.tq], c);
                                    idct(c.data.data() + (size_t)c.w2 * y2 + x2, c.w2, blk);
                                }
                        }
                        if (!restart_due()) return;
                    }
            }
        } else {
            if (scan_n == 1) {
                Component &c = comp[order[0]];
                const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
                for (int j = 0; j < h; j++)
                    for (int i = 0; i < w; i++) {
                        int16_t *data = c.coeff.data() + 64 * ((size_t)i + (size_t)j * c.coeff_w);
                        if (spec_start == 0) block_prog_dc(data, hdc[c.hd], c);
                        else block_prog_ac(data, hac[c.ha]);
                        if (!restart_due()) return;
                    }
            } else {
                for (int j = 0; j < mcu_y; j++)
                    for (int i = 0; i < mcu_x; i++) {
                        for (int k = 0; k < scan_n; k++) {
                            Component &c = comp[order[k]];
                            for (int y = 0; y < c.v; y++)
                                for (int x = 0; x < c.h; x++) {
                                    const int x2 = i * c.h + x, y2 = j * c.v + y;
                                    block_prog_dc(c.coeff.data() + 64 * ((size_t)x2 + (size_t)y2 * c.coeff_w), hdc[c.hd], c);
                                }
                        }
                        if (!restart_due()) return;
                    }
            }
        }
    }
    void finish_progressive() {
        for (int n = 0; n < ncomp; n++) {
            Component &c = comp[n];
            const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
            for (int j = 0; j < h; j++)
                for (int i = 0; i < w; i++) {
                    int16_t *data = c.coeff.data() + 64 * ((size_t)i + (size_t)j * c.coeff_w);
                    for (int k = 0; k < 64; k++) data[k] = (int16_t)(data[k] * dequant[c.tq][k]);
                    idct(c.data.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, data);
                }
        }
    }

    /* ---- marker segments */
    void read_dqt(int len) {
        while (len > 0) {
            const int q = get8(), prec = q >> 4, t = q & 15;
            if ((prec != 0 && prec != 1) || t > 3) fail("JPEG: bad DQT");
            for (int i = 0; i < 64; i++) dequant[t][kZigzag[i]] = (uint16_t)(prec ? get16() : get8());
            len -= prec ? 129 : 65;
        }
        if (len != 0) fail("JPEG: bad DQT length");
    }
    void read_dht(int len) {
        while (len > 0) {
            const int q = get8(), tc = q >> 4, th = q & 15;
            if (tc > 1 || th > 3) fail("JPEG: bad DHT");
            uint8_t counts[16], vals[256];
            int n = 0;
            for (int i = 0; i < 16; i++) n += (counts[i] = (uint8_t)get8());
            if (n > 256) fail("JPEG: bad DHT");
            for (int i = 0; i < n; i++) vals[i] = (uint8_t)get8();
            (tc ? hac[th] : hdc[th]).build(counts, vals, n);
            len -= 17 + n;
        }
        if (len != 0) fail("JPEG: bad DHT length");
    }
    void read_sof(int m) {
        progressive = m == 0xc2;
        const int len = get16(), prec = get8();
        if (prec != 8) fail("JPEG: only 8-bit precision");
        img_h = get16();
        img_w = get16();
        if (!img_h || !img_w) fail("JPEG: bad size");
        if ((size_t)img_h * (size_t)img_w > ((size_t)1 << 28)) fail("JPEG: image too large");
        ncomp = get8();
        if (ncomp != 1 && ncomp != 3) fail(ncomp == 4 ? "JPEG: CMYK / YCCK is not supported" : "JPEG: bad component count");
        if (len != 8 + 3 * ncomp) fail("JPEG: bad SOF length");
        static const char rgb[3] = {'R', 'G', 'B'};
        rgb_ids = ncomp == 3;
        h_max = v_max = 1;
        for (int i = 0; i < ncomp; i++) {
            Component &c = comp[i];
            c.id = get8();
            if (ncomp == 3 && c.id != rgb[i]) rgb_ids = false;
            const int q = get8();
            c.h = q >> 4;
            c.v = q & 15;
            c.tq = get8();
            if (!c.h || c.h > 4 || !c.v || c.v > 4 || c.tq > 3) fail("JPEG: bad component");
            if (c.h > h_max) h_max = c.h;
            if (c.v > v_max) v_max = c.v;
        }
        for (int i = 0; i < ncomp; i++)
            if (h_max % comp[i].h || v_max % comp[i].v) fail("JPEG: fractional sampling ratios are not supported");
        mcu_w = h_max * 8;
        mcu_h = v_max * 8;
        mcu_x = (img_w + mcu_w - 1) / mcu_w;
        mcu_y = (img_h + mcu_h - 1) / mcu_h;
        for (int i = 0; i < ncomp; i++) {
            Component &c = comp[i];
            c.x = (img_w * c.h + h_max - 1) / h_max;
            c.y = (img_h * c.v + v_max - 1) / v_max;
            c.w2 = mcu_x * c.h * 8;
            c.h2 = mcu_y * c.v * 8;
            c.data.assign((size_t)c.w2 * c.h2, 0);
            if (progressive) {
                c.coeff_w = c.w2 / 8;
                c.coeff.assign((size_t)c.w2 * c.h2, 0);
            }
        }
    }
    void read_sos() {
        const int len = get16();
        scan_n = get8();
        if (scan_n < 1 || scan_n > 4 || scan_n > ncomp) fail("JPEG: bad SOS component count");
        if (len != 6 + 2 * scan_n) fail("JPEG: bad SOS length");
        for (int i = 0; i < scan_n; i++) {
            const int id = get8(), q = get8();
            int which = 0;
            for (; which < ncomp; which++)
                if (comp[which].id == id) break;
            if (which == ncomp) fail("JPEG: SOS names an unknown component");
            comp[which].hd = q >> 4;
            comp[which].ha = q & 15;
            if (comp[which].hd > 3 || comp[which].ha > 3) fail("JPEG: bad Huffman table index");
            order[i] = which;
        }
        spec_start = get8();
        spec_end = get8();
        const int a = get8();
        succ_high = a >> 4;
        succ_low = a & 15;
        if (progressive) {
            if (spec_start > 63 || spec_end > 63 || spec_start > spec_end || succ_high > 13 || succ_low > 13) fail("JPEG: bad SOS");
        } else {
            if (spec_start != 0 || succ_high != 0 || succ_low != 0) fail("JPEG: bad SOS");
            spec_end = 63;
        }
    }

    /* ---- chroma upsampling (the "fancy" triangle filters, integer) */
    static const uint8_t *up_1(uint8_t *, const uint8_t *near, const uint8_t *, int, int) { return near; }
    static const uint8_t *up_v2(uint8_t *out, const uint8_t *near, const uint8_t *far, int w, int) {
        for (int i = 0; i < w; i++) out[i] = (uint8_t)((3 * near[i] + far[i] + 2) >> 2);
        return out;
    }
    static const uint8_t *up_h2(uint8_t *out, const uint8_t *in, const uint8_t *, int w, int) {
        if (w == 1) {
            out[0] = out[1] = in[0];
            return out;
        }
        out[0] = in[0];
        out[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
        int i;
        for (i = 1; i < w - 1; i++) {
            const int n = 3 * in[i] + 2;
            out[i * 2] = (uint8_t)((n + in[i - 1]) >> 2);
            out[i * 2 + 1] = (uint8_t)((n + in[i + 1]) >> 2);
        }
        out[i * 2] = (uint8_t)((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
        out[i * 2 + 1] = in[w - 1];
        return out;
    }
    static const uint8_t *up_hv2(uint8_t *out, const uint8_t *near, const uint8_t *far, int w, int) {
        if (w == 1) {
            out[0] = out[1] = (uint8_t)((3 * near[0] + far[0] + 2) >> 2);
            return out;
        }
        int t1 = 3 * near[0] + far[0];
        out[0] = (uint8_t)((t1 + 2) >> 2);
        for (int i = 1; i < w; i++) {
            const int t0 = t1;
            t1 = 3 * near[i] + far[i];
            out[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
            out[i * 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
        }
        out[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
        return out;
    }
    static const uint8_t *up_generic(uint8_t *out, const uint8_t *near, const uint8_t *, int w, int hs) {
        for (int i = 0; i < w; i++)
            for (int j = 0; j < hs; j++) out[i * hs + j] = near[i];
        return out;
    }

    Image run() {
        if (get8() != 0xff || get8() != 0xd8) fail("not a JPEG");
        bool have_sof = false, done = false;
        int pending = -1;
        while (!done) {
            int m;
            if (pending >= 0) {
                m = pending;
                pending = -1;
            } else {
                int c = get8();
                if (p >= end && c == 0) fail("JPEG: no EOI");
                if (c != 0xff) continue; /* junk between segments */
                while ((m = get8()) == 0xff) {}
                if (m == 0) continue;
            }
            if (m == 0xd9) {
                done = true;
            } else if (m == 0xc0 || m == 0xc1 || m == 0xc2) {
                if (have_sof) fail("JPEG: more than one frame");
                read_sof(m);
                have_sof = true;
            } else if ((m >= 0xc3 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc)) {
                fail("JPEG: unsupported coding process (lossless / hierarchical / arithmetic)");
            } else if (m == 0xc4) {
                read_dht(get16() - 2);
            } else if (m == 0xdb) {
                read_dqt(get16() - 2);
            } else if (m == 0xdd) {
                if (get16() != 4) fail("JPEG: bad DRI");
                restart_interval = get16();
            } else if (m == 0xda) {
                if (!have_sof) fail("JPEG: SOS before SOF");
                read_sos();
                scan();
                /* next marker: the one the bit reader ran into, or scan forward for it */
                if (marker >= 0) {
                    pending = marker;
                } else {
                    while (p < end) {
                        if (get8() != 0xff) continue;
                        int c = get8();
                        while (c == 0xff && p < end) c = get8();
                        if (c != 0) {
                            pending = c;
                            break;
                        }
                    }
                    if (pending < 0) fail("JPEG: no EOI");
                }
                if (pending >= 0xd0 && pending <= 0xd7) pending = -1; /* a stray restart marker */
            } else if (m == 0xe0 || m == 0xee) {
                int len = get16() - 2;
                if (len < 0) fail("JPEG: bad APP length");
                if (m == 0xe0 && len >= 5) {
                    static const char tag[5] = {'J', 'F', 'I', 'F', 0};
                    bool ok = true;
                    for (int i = 0; i < 5; i++) ok &= get8() == tag[i];
                    len -= 5;
                    if (ok) jfif = true;
                } else if (m == 0xee && len >= 12) {
                    static const char tag[6] = {'A', 'd', 'o', 'b', 'e', 0};
                    bool ok = true;
                    for (int i = 0; i < 6; i++) ok &= get8() == tag[i];
                    len -= 6;
                    if (ok) {
                        get8();  /* version */
                        get16(); /* flags0 */
                        get16(); /* flags1 */
                        adobe_transform = get8();
                        len -= 6;
                    }
                }
                p += (len < end - p) ? len : end - p;
            } else if ((m >= 0xe1 && m <= 0xef) || m == 0xfe) {
                const int len = get16() - 2;
                if (len < 0) fail("JPEG: bad segment length");
                p += (len < end - p) ? len : end - p;
            } else if (m >= 0xd0 && m <= 0xd7) {
                /* restart marker outside a scan: ignore */
            } else {
                fail("JPEG: unknown marker");
            }
        }
        if (!have_sof) fail("JPEG: no frame");
        if (progressive) finish_progressive();

        /* upsample + colour conversion, one output row at a time */
        Image im;
        im.w = (uint32_t)img_w;
        im.h = (uint32_t)img_h;
        im.channels_in_file = ncomp >= 3 ? 3 : 1;
        im.rgba.resize((size_t)img_w * img_h * 4);
        typedef const uint8_t *(*Resample)(uint8_t *, const uint8_t *, const uint8_t *, int, int);
        struct Up {
            Resample fn;
            const uint8_t *line0, *line1;
            int hs, vs, w_lores, ystep, ypos;
            std::vector<uint8_t> buf;
        } up[3];
        for (int k = 0; k < ncomp; k++) {
            Up &r = up[k];
            r.hs = h_max / comp[k].h;
            r.vs = v_max / comp[k].v;
            r.ystep = r.vs >> 1;
            r.w_lores = (img_w + r.hs - 1) / r.hs;
            r.ypos = 0;
            r.line0 = r.line1 = comp[k].data.data();
            r.buf.resize((size_t)img_w + 3 + 16);
            r.fn = (r.hs == 1 && r.vs == 1) ? up_1 : (r.hs == 1 && r.vs == 2) ? up_v2 : (r.hs == 2 && r.vs == 1) ? up_h2 : (r.hs == 2 && r.vs == 2) ? up_hv2 : up_generic;
        }
        const bool to_rgb = ncomp == 3 && !rgb_ids && !(adobe_transform == 0 && !jfif);
        for (int j = 0; j < img_h; j++) {
            const uint8_t *row[3] = {nullptr, nullptr, nullptr};
            for (int k = 0; k < ncomp; k++) {
                Up &r = up[k];
                const bool y_bot = r.ystep >= (r.vs >> 1);
                row[k] = r.fn(r.buf.data(), y_bot ? r.line1 : r.line0, y_bot ? r.line0 : r.line1, r.w_lores, r.hs);
                if (++r.ystep >= r.vs) {
                    r.ystep = 0;
                    r.line0 = r.line1;
                    if (++r.ypos < comp[k].y) r.line1 += comp[k].w2;
                }
            }
            uint8_t *out = &im.rgba[(size_t)j * img_w * 4];
            if (ncomp == 1) {
                for (int i = 0; i < img_w; i++, out += 4) {
                    out[0] = out[1] = out[2] = row[0][i];
                    out[3] = 255;
                }
            } else if (!to_rgb) {
                for (int i = 0; i < img_w; i++, out += 4) {
                    out[0] = row[0][i];
                    out[1] = row[1][i];
                    out[2] = row[2][i];
                    out[3] = 255;
                }
            } else {
                /* 20-bit fixed point; the Cb term of green is truncated to its upper 16 bits first */
                auto fx = [](float x) { return ((int)(x * 4096.0f + 0.5f)) << 8; };
                static const int k_r = fx(1.40200f), k_g1 = fx(0.71414f), k_g2 = fx(0.34414f), k_b = fx(1.77200f);
                for (int i = 0; i < img_w; i++, out += 4) {
                    const int yf = (row[0][i] << 20) + (1 << 19), cr = row[2][i] - 128, cb = row[1][i] - 128;
                    int r = yf + cr * k_r;
                    int g = yf + cr * -k_g1 + (int)(((unsigned)(cb * -k_g2)) & 0xffff0000u);
                    int b = yf + cb * k_b;
                    r >>= 20;
                    g >>= 20;
                    b >>= 20;
                    out[0] = clamp8(r);
                    out[1] = clamp8(g);
                    out[2] = clamp8(b);
                    out[3] = 255;
                }
            }
        }
        return im;
    }
};

inline Image decode(const uint8_t *d, size_t n) {
    Decoder dec;
    dec.p = d;
    dec.end = d + n;
    memset(dec.dequant, 0, sizeof(dec.dequant));
    return dec.run();
}

} // namespace jpeg

/* what stbi_load_from_memory(bytes, size, &w, &h, &comp, 4) returns for the formats glTF allows */
inline Image decode_rgba8(const uint8_t *d, size_t n, bool tinygltf_16bit_quirk = false) {
    if (n >= 8 && d[0] == 0x89 && d[1] == 'P' && d[2] == 'N' && d[3] == 'G') return png::decode(d, n, tinygltf_16bit_quirk);
    if (n >= 3 && d[0] == 0xff && d[1] == 0xd8) return jpeg::decode(d, n);
    fail("unsupported image format (glTF allows image/png and image/jpeg)");
}

} // namespace img
} // namespace raytracer

#endif /* RT_HOST_IMAGE_CODECS_HPP */
