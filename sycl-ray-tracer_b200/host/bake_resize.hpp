/*
 * bake_resize.hpp — the texture bake's resize to one 512x512 RGBA8 layer, bit for bit.
 *
 * The reference bakes every glTF image with
 *     stbir_resize_uint8_srgb(pixels, w, h, 0, out, 512, 512, 0, STBIR_RGBA)      (src/image_manager.hpp:52-62)
 * from its vendored stb_image_resize2 v2.04 (deps/include/stb_image_resize2.h, public domain / MIT, Jeff Roberts
 * and Jorge Rodriguez). Texels feed the renderer as bytes, so "close" is not parity: this file restates the
 * ARITHMETIC of that one call — same single-precision operations in the same order — as plain scalar C++ written for
 * this path only (one pixel layout, one data type, clamp edges, default filters, no sub-rectangles, no callbacks, no
 * threading). stb's SIMD and scalar paths are bit-identical by design ("not on by default to maintain bit identical
 * simd to non-simd", :1311), which is what makes a scalar restatement possible; FMA contraction must stay off
 * (host/Makefile passes -ffp-contract=off). Checked against the reference's own library compiled in place
 * (oracle/_ref/libstbref.so) on enlarging, reducing, mixed, tiny, huge and transparent inputs
 * (tests/test_image_codecs.py) and through committed golden texels (tests/golden/images.npz).
 *
 * What decides the bits (file:line = deps/include/stb_image_resize2.h):
 *   - per axis: scale = (float)(out / in), inv_scale = (float)(1 / scale) (:7369-7372); scale == 1 -> point sampling
 *     (a copy), scale > 1 -> Catmull-Rom over input pixels, scale < 1 -> Mitchell stretched by 1 / scale (:6312-6330);
 *   - filter taps in single precision per output pixel (enlarging, :3213-3268) or per input pixel (reducing,
 *     :3317-3393), taps below 2^-120 dropped, sum normalised by ONE reciprocal (:3395-3446), only the first
 *     `numerator` phases computed when the scale is a small rational and the rest copied (:3448-3464), taps that fall
 *     off the image folded into the edge pixel nearest-first on the left, in order on the right (:3496-3532);
 *   - horizontal taps are stored `widest` per pixel, and windows that would read past the row end are moved back with
 *     zero taps in front (:3699-3773) — this shifts which taps land on even and odd positions;
 *   - a pixel is 7 floats: linear r, g, b (256-entry table, :1104), alpha / 255, and the three alpha-weighted colours
 *     (:3980-4069);
 *   - horizontal sum: up to 3 taps in sequence; from 4 on, even and odd positions are summed separately and added at the
 *     end (:5686-5823); vertical sum: in sequence (:9784-9925; the scatter form used for reductions beyond 8x adds the
 *     same products in the same order);
 *   - which axis goes first is a cost model with tuned weights (:6587-6720);
 *   - back to bytes: colour = weighted / alpha with ONE reciprocal, or the plainly filtered colour when alpha < 2^-120
 *     (:4140-4186); sRGB through the 104-entry piecewise-linear table of Fabian Giesen's float->sRGB8 routine (:1139-1178);
 *     alpha = trunc(clamp(a * 255 + 0.5)).
 */
#ifndef RT_BAKE_RESIZE_HPP
#define RT_BAKE_RESIZE_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace raytracer {
namespace glb {
namespace bake {

constexpr int kChannels = 7; /* r g b a, r*a g*a b*a */

inline float tiny() { return std::ldexp(1.0f, -120); } /* stbir__small_float */

enum Filter { kPoint, kCatmullRom, kMitchell };

inline float filter_support(Filter f) { return f == kPoint ? 0.5f : 2.0f; }
inline float filter_kernel(Filter f, float x) {
    if (f == kPoint) return 1.0f;
    if (x < 0.0f) x = -x;
    if (f == kCatmullRom) {
        if (x < 1.0f) return 1.0f - x * x * (2.5f - 1.5f * x);
        if (x < 2.0f) return 2.0f - x * (4.0f + x * (0.5f * x - 2.5f));
        return 0.0f;
    }
    if (x < 1.0f) return (16.0f + x * x * (21.0f * x - 36.0f)) / 18.0f;
    if (x < 2.0f) return (32.0f + x * (-60.0f + x * (36.0f - 7.0f * x))) / 18.0f;
    return 0.0f;
}

/* continued-fraction search for scale = numer / denom with the bounded term <= limit (:7268-7346): decides whether only
 * `numer` filter phases are computed and the others copied */
inline bool scale_as_rational(double f, uint32_t limit, bool bound_denominator, uint32_t &numer, uint32_t &denom) {
    uint64_t top = (uint64_t)(f * (double)(1 << 25)), bot = 1 << 25;
    uint64_t n_prev = 0, d_prev = 1, n = 1, d = 0;
    const double one_ulp = 1.0 / (double)(1 << 24);
    for (;;) {
        if ((bound_denominator ? d : n) >= limit) break;
        if (d) {
            double err = (double)n / (double)d - f;
            if (err < 0.0) err = -err;
            if (err < one_ulp) {
                numer = (uint32_t)n;
                denom = (uint32_t)d;
                return true;
            }
        }
        if (bot == 0) break;
        const uint64_t q = top / bot, r = top % bot;
        top = bot;
        bot = r;
        uint64_t t = q * d + d_prev;
        d_prev = d;
        d = t;
        t = q * n + n_prev;
        n_prev = n;
        n = t;
    }
    if (bound_denominator) {
        n = (uint64_t)(f * (double)limit + 0.5);
        d = limit;
    } else {
        n = limit;
        d = (uint64_t)((double)limit / f + 0.5);
    }
    numer = (uint32_t)n;
    denom = (uint32_t)d;
    double err = d ? ((double)(uint32_t)n / (double)(uint32_t)d - f) : 1.0;
    if (err < 0.0) err = -err;
    return err < one_ulp;
}

/* one axis of the resize: for every output index the window [n0, n1] of input indices and its taps */
struct Axis {
    int in_n = 0, out_n = 0;
    float scale = 1.0f, inv_scale = 1.0f;
    bool rational = false;
    uint32_t numer = 0, denom = 0;
    Filter filter = kPoint;
    int pixel_width = 1; /* input pixels under one output pixel's filter */
    int margin = 0;
    int mode = 1;        /* 1 = enlarging, 2 = reducing; the vertical axis beyond 32 rows per window "scatters" (gathers == false) */
    bool gathers = true;
    int stride = 0;      /* floats per output index in taps */
    std::vector<int> n0, n1;
    std::vector<float> taps;
    int widest = -1;
};

inline int floor_to_int(float x) { return (int)std::floor(x); }
inline int ceil_to_int(float x) { return (int)std::ceil(x); }

inline void add_tap_at(Axis &a, int o, int pixel, float value) { /* fold a tap that fell off the image onto `pixel` (:3270-3298) */
    float *c = &a.taps[(size_t)o * a.stride];
    if (pixel <= a.n1[o]) {
        if (pixel < a.n0[o]) throw std::runtime_error("bake resize: filter window lies outside the image");
        c[pixel - a.n0[o]] += value;
    } else {
        const int e = pixel - a.n0[o];
        for (int j = a.n1[o] - a.n0[o] + 1; j < e; j++) c[j] = 0.0f;
        c[e] = value;
        a.n1[o] = pixel;
    }
}

inline void taps_enlarging(Axis &a, int count) { /* :3213-3268 with pixel_shift 0 */
    const float radius = filter_support(a.filter) * a.scale;
    for (int o = 0; o < count; o++) {
        const float out_center = (float)o + 0.5f;
        const float in_center = (out_center + 0.0f) * a.inv_scale;
        const float lo = out_center - radius, hi = out_center + radius;
        int first = floor_to_int((lo + 0.0f) * a.inv_scale + 0.5f);
        int last = floor_to_int((hi + 0.0f) * a.inv_scale - 0.5f);
        float *c = &a.taps[(size_t)o * a.stride];
        int last_nonzero = -1;
        for (int i = 0; i <= last - first; i++) {
            const float in_pixel_center = (float)(i + first) + 0.5f;
            float v = filter_kernel(a.filter, in_center - in_pixel_center);
            if (v < tiny() && v > -tiny()) {
                if (i == 0) { /* a vanishing first tap is dropped, the window starts one pixel later */
                    ++first;
                    i--;
                    continue;
                }
                v = 0.0f;
            } else {
                last_nonzero = i;
            }
            c[i] = v;
        }
        a.n0[o] = first;
        a.n1[o] = last_nonzero + first;
    }
}

inline void taps_reducing(Axis &a) { /* :3317-3393: input pixels hand their weight to the output pixels they reach */
    const float radius = filter_support(a.filter) * a.inv_scale;
    const bool polyphase = a.rational && (int)a.numer < a.out_n;
    int first_seen = -1;
    for (int in_pixel = -a.margin; in_pixel < a.in_n + a.margin; in_pixel++) {
        const float in_center = (float)in_pixel + 0.5f;
        const float out_center_of_in = in_center * a.scale - 0.0f;
        const float lo = in_center - radius, hi = in_center + radius;
        int first = floor_to_int((lo * a.scale - 0.0f) + 0.5f);
        int last = floor_to_int((hi * a.scale - 0.0f) - 0.5f);
        if (first < 0) first = 0;
        if (last >= a.out_n) last = a.out_n - 1;
        if (first > last) continue;
        if (polyphase) {
            if (first == (int)a.numer) break;
            if (last >= (int)a.numer) last = (int)a.numer - 1;
        }
        for (int i = 0; i <= last - first; i++) {
            const float out_pixel_center = (float)(i + first) + 0.5f;
            const float x = out_pixel_center - out_center_of_in;
            float v = filter_kernel(a.filter, x) * a.scale;
            if (v < tiny() && v > -tiny()) v = 0.0f;
            const int o = i + first;
            float *c = &a.taps[(size_t)o * a.stride];
            if (o > first_seen) {
                first_seen = o;
                a.n0[o] = in_pixel;
                a.n1[o] = in_pixel;
                c[0] = v;
            } else {
                if (c[0] == 0.0f) a.n0[o] = in_pixel; /* the window's first tap vanished: start here instead */
                a.n1[o] = in_pixel;
                if (in_pixel - a.n0[o] >= a.stride) throw std::runtime_error("bake resize: filter window wider than expected");
                c[in_pixel - a.n0[o]] = v;
            }
        }
    }
}

inline void normalise_and_clamp(Axis &a) { /* :3395-3560, clamp edges */
    const int n = a.out_n;
    const bool polyphase = a.rational && (int)a.numer < n;
    const int computed = polyphase ? (int)a.numer : n;
    for (int o = 0; o < computed; o++) {
        float *c = &a.taps[(size_t)o * a.stride];
        const int e = a.n1[o] - a.n0[o];
        float total = 0.0f;
        for (int i = 0; i <= e; i++) total += c[i];
        if (total < tiny() && total > -tiny()) {
            a.n1[o] = a.n0[o];
            c[0] = 0.0f;
        } else if (total < (1.0f - tiny()) || total > (1.0f + tiny())) {
            const float s = 1.0f / total;
            for (int i = 0; i <= e; i++) c[i] *= s;
        }
    }
    if (polyphase) { /* the remaining phases repeat the first `numer` ones, `denom` input pixels further on */
        for (int o = (int)a.numer; o < n; o++) {
            a.n0[o] = a.n0[o - (int)a.numer] + (int)a.denom;
            a.n1[o] = a.n1[o - (int)a.numer] + (int)a.denom;
        }
        const size_t period = (size_t)a.numer * a.stride, total = (size_t)n * a.stride;
        for (size_t i = period; i < total; i++) a.taps[i] = a.taps[i - period];
    }
    a.widest = -1;
    const int last_in = a.in_n - 1;
    for (int o = 0; o < n; o++) {
        float *c = &a.taps[(size_t)o * a.stride];
        if (a.n1[o] > last_in) { /* right edge: fold in increasing order */
            const int start = a.n0[o], end = a.n1[o];
            a.n1[o] = last_in;
            for (int i = a.in_n; i <= end; i++) add_tap_at(a, o, last_in, c[i - start]);
        }
        if (a.n0[o] < 0) { /* left edge: fold nearest first, the window's own first tap last */
            const int old_n0 = a.n0[o];
            for (int i = -1; i > old_n0; i--) add_tap_at(a, o, 0, c[i - old_n0]);
            const float first_tap = c[0];
            a.n0[o] = 0;
            for (int i = 0; i <= a.n1[o]; i++) c[i] = c[i - old_n0];
            add_tap_at(a, o, 0, first_tap);
        }
        if (a.n0[o] <= a.n1[o]) {
            int width = a.n1[o] - a.n0[o] + 1;
            while (width && c[width - 1] == 0.0f) --width;
            a.n1[o] = a.n0[o] + width - 1;
            if (a.n0[o] <= a.n1[o] && width > a.widest) a.widest = width;
            for (int i = width; i < a.stride; i++) c[i] = 0.0f;
        }
    }
}

inline Axis make_axis(int in_n, int out_n, bool horizontal) { /* :7349-7397, :6312-6376 */
    Axis a;
    a.in_n = in_n;
    a.out_n = out_n;
    const double scale = ((double)out_n / (double)in_n) * (1.0 / 1.0);
    a.scale = (float)scale;
    a.inv_scale = (float)(1.0 / scale);
    a.rational = scale_as_rational(scale, scale <= 1.0 ? (uint32_t)out_n : (uint32_t)in_n, scale >= 1.0, a.numer, a.denom);
    const bool enlarging = a.scale >= (1.0f - tiny());
    a.filter = enlarging ? (a.scale <= (1.0f + tiny()) ? kPoint : kCatmullRom) : kMitchell;
    const float support = filter_support(a.filter);
    a.pixel_width = enlarging ? ceil_to_int(support * 2.0f) : ceil_to_int(support * 2.0f / a.scale);
    a.margin = a.pixel_width / 2;
    a.mode = enlarging ? 1 : 2;
    a.gathers = enlarging || horizontal || a.pixel_width <= 32;
    /* taps per window while they are computed: a scattering axis computes them as a reducing gather over pixel_width slots */
    a.stride = enlarging ? ceil_to_int(support * 2.0f) : (a.gathers ? ceil_to_int(support * 2.0f / a.scale) : a.pixel_width);
    a.n0.assign((size_t)out_n, 0);
    a.n1.assign((size_t)out_n, -1);
    a.taps.assign((size_t)out_n * a.stride + 1, 0.0f);
    if (enlarging) {
        const bool polyphase = a.rational && (int)a.numer < out_n;
        taps_enlarging(a, polyphase ? (int)a.numer : out_n);
    } else {
        taps_reducing(a);
    }
    normalise_and_clamp(a);
    return a;
}

/* the input span one row is decoded over: only its end matters here — windows are kept inside [0, row_end] (:6378-6468) */
inline int decoded_row_end(const Axis &a) {
    auto in_range = [&](float center, float radius, int &first, int &last) {
        const float lo = center - radius, hi = center + radius;
        first = floor_to_int((lo + 0.0f) * a.inv_scale + 0.5f);
        last = floor_to_int((hi + 0.0f) * a.inv_scale - 0.5f);
    };
    auto out_range = [&](float in_center, float radius, int &first, int &last) {
        const float lo = in_center - radius, hi = in_center + radius;
        first = floor_to_int((lo * a.scale - 0.0f) + 0.5f);
        last = floor_to_int((hi * a.scale - 0.0f) - 0.5f);
        if (first < 0) first = 0;
        if (last >= a.out_n) last = a.out_n - 1;
    };
    int first, last, end;
    if (a.mode == 1) {
        in_range((float)(a.out_n - 1) + 0.5f, filter_support(a.filter) * a.scale, first, last);
        end = last;
    } else {
        const float radius = filter_support(a.filter) * a.inv_scale;
        in_range((float)a.out_n, 0.0f, first, last);
        end = last;
        int n = end - 1;
        const int stop = n + 1 + a.margin;
        while (n <= stop) {
            int f, l;
            out_range((float)n + 0.5f, radius, f, l);
            if (f > l) break;
            if (f < a.out_n || l >= 0) end = n;
            ++n;
        }
    }
    if (end >= a.in_n) end = a.in_n - 1;
    return end;
}

/* horizontal taps as the row filter reads them: `widest` per pixel, windows moved back from the row end (:3563-3778) */
inline void pack_for_rows(Axis &a, int row_width) {
    const int w = a.widest;
    if (w < 1) throw std::runtime_error("bake resize: empty filter");
    std::vector<float> packed((size_t)a.out_n * w + 1, 0.0f);
    for (int o = 0; o < a.out_n; o++)
        for (int i = 0; i < w && i < a.stride; i++) packed[(size_t)o * w + i] = a.taps[(size_t)o * a.stride + i];
    a.taps.swap(packed);
    a.stride = w;
    auto reach = [&](int o) { /* how far the row filter reads from n0: generic loops run in steps of four past 12 taps */
        if (w <= 12) return w;
        const int mod = w & 3;
        int r = (((a.n1[o] - a.n0[o] + 1) - mod + 3) & ~3) + mod;
        if (r < 8 + mod) r = 8 + mod;
        return r;
    };
    for (int o = a.out_n - 1; o >= 0 && a.n0[o] + w * 2 >= row_width; o--) {
        if (a.n0[o] + w > row_width && a.n0[o] + reach(o) > row_width) {
            const int new_n0 = row_width - reach(o), count = a.n1[o] - a.n0[o] + 1, back = a.n0[o] - new_n0;
            if (new_n0 < 0 || back + count > w) throw std::runtime_error("bake resize: image narrower than the filter");
            float *c = &a.taps[(size_t)o * w];
            for (int i = count - 1; i >= 0; i--) c[i + back] = c[i];
            for (int i = 0; i < back; i++) c[i] = 0.0f;
            a.n0[o] = new_n0;
        }
    }
}

/* one row, horizontally: in = in_n pixels of 7 floats, out = out_n pixels */
inline void filter_row(const Axis &a, const float *in, float *out) {
    if (a.filter == kPoint && a.scale == 1.0f) {
        std::memcpy(out, in, (size_t)a.out_n * kChannels * sizeof(float));
        return;
    }
    const int w = a.stride;
    for (int o = 0; o < a.out_n; o++) {
        const float *c = &a.taps[(size_t)o * w];
        const float *p = in + (size_t)a.n0[o] * kChannels;
        const int count = std::min(w, a.n1[o] - a.n0[o] + 1); /* taps beyond are zero: adding 0 * pixel changes nothing */
        float *q = out + (size_t)o * kChannels;
        if (w <= 3) {
            for (int ch = 0; ch < kChannels; ch++) {
                float t = p[ch] * c[0];
                for (int i = 1; i < count; i++) t += p[(size_t)i * kChannels + ch] * c[i];
                q[ch] = t;
            }
        } else { /* even and odd positions summed apart, then added */
            for (int ch = 0; ch < kChannels; ch++) {
                float even = p[ch] * c[0];
                float odd = count > 1 ? p[kChannels + ch] * c[1] : 0.0f;
                for (int i = 2; i < count; i += 2) even += p[(size_t)i * kChannels + ch] * c[i];
                for (int i = 3; i < count; i += 2) odd += p[(size_t)i * kChannels + ch] * c[i];
                q[ch] = even + odd;
            }
        }
    }
}

/* tuned cost weights of the 7-float pixel (:6622-6631), rows = class of the vertical resize */
inline bool vertical_pass_first(const Axis &h, const Axis &v) { /* :6659-6720 */
    static const float weights[8][4] = {
        {0.00000f, 0.59375f, 0.00000f, 0.96875f}, {0.06250f, 0.81250f, 0.06250f, 0.59375f},
        {0.75000f, 0.43750f, 0.12500f, 0.96875f}, {0.87500f, 0.06250f, 0.18750f, 0.43750f},
        {1.00000f, 1.00000f, 1.00000f, 1.00000f}, {0.15625f, 0.12500f, 1.00000f, 1.00000f},
        {0.06250f, 0.12500f, 0.00000f, 1.00000f}, {0.00000f, 1.00000f, 0.03125f, 0.34375f}};
    int cls;
    if (v.out_n <= 4 || h.out_n <= 4) cls = v.out_n < h.out_n ? 6 : 7;
    else if (v.scale <= 1.0f) cls = v.gathers ? 1 : 0;
    else if (v.scale <= 2.0f) cls = 2;
    else if (v.scale <= 3.0f) cls = 3;
    else if (v.scale <= 4.0f) cls = 5;
    else cls = 6;
    const float *w = weights[cls];
    const double h_cost = (float)h.pixel_width * w[0] + h.scale * (float)v.pixel_width * w[1];
    const double v_cost = (float)v.pixel_width * w[2] + v.scale * (float)h.pixel_width * w[3];
    return v_cost <= h_cost;
}

struct Tables {
    float to_linear[256];
    Tables() {
        /* stb's table holds the sRGB decode evaluated in single precision and printed with six decimals (:1104-1129);
         * generated the same way here, compared entry by entry with the reference's in tests/test_image_codecs.py */
        for (int v = 0; v < 256; v++) {
            const float c = (float)v / 255.0f;
            const float lin = c <= 0.04045f ? c / 12.92f : std::pow((c + 0.055f) / 1.055f, 2.4f);
            to_linear[v] = (float)(std::round((double)lin * 1e6) / 1e6);
        }
    }
};
inline const Tables &tables() {
    static const Tables t;
    return t;
}

/* float -> sRGB8 after Fabian Giesen (https://gist.github.com/rygorous/2203834, public domain), the table stb uses (:1139-1178):
 * per binade eighth a bias and a slope, interpolated with the next eight mantissa bits */
inline uint8_t linear_to_srgb8(float v) {
    static const uint32_t tab[104] = {
        0x0073000d, 0x007a000d, 0x0080000d, 0x0087000d, 0x008d000d, 0x0094000d, 0x009a000d, 0x00a1000d, 0x00a7001a, 0x00b4001a, 0x00c1001a,
        0x00ce001a, 0x00da001a, 0x00e7001a, 0x00f4001a, 0x0101001a, 0x010e0033, 0x01280033, 0x01410033, 0x015b0033, 0x01750033, 0x018f0033,
        0x01a80033, 0x01c20033, 0x01dc0067, 0x020f0067, 0x02430067, 0x02760067, 0x02aa0067, 0x02dd0067, 0x03110067, 0x03440067, 0x037800ce,
        0x03df00ce, 0x044600ce, 0x04ad00ce, 0x051400ce, 0x057b00c5, 0x05dd00bc, 0x063b00b5, 0x06970158, 0x07420142, 0x07e30130, 0x087b0120,
        0x090b0112, 0x09940106, 0x0a1700fc, 0x0a9500f2, 0x0b0f01cb, 0x0bf401ae, 0x0ccb0195, 0x0d950180, 0x0e56016e, 0x0f0d015e, 0x0fbc0150,
        0x10630143, 0x11070264, 0x1238023e, 0x1357021d, 0x14660201, 0x156601e9, 0x165a01d3, 0x174401c0, 0x182401af, 0x18fe0331, 0x1a9602fe,
        0x1c1502d2, 0x1d7e02ad, 0x1ed4028d, 0x201a0270, 0x21520256, 0x227d0240, 0x239f0443, 0x25c003fe, 0x27bf03c4, 0x29a10392, 0x2b6a0367,
        0x2d1d0341, 0x2ebe031f, 0x304d0300, 0x31d105b0, 0x34a80555, 0x37520507, 0x39d504c5, 0x3c37048b, 0x3e7c0458, 0x40a8042a, 0x42bd0401,
        0x44c20798, 0x488e071e, 0x4c1c06b6, 0x4f76065d, 0x52a50610, 0x55ac05cc, 0x5892058f, 0x5b590559, 0x5e0c0a23, 0x631c0980, 0x67db08f6,
        0x6c55087f, 0x70940818, 0x74a007bd, 0x787d076c, 0x7c330723};
    const uint32_t lowest = (127u - 13u) << 23, almost_one = 0x3f7fffffu;
    float lo, hi;
    std::memcpy(&lo, &lowest, 4);
    std::memcpy(&hi, &almost_one, 4);
    if (!(v > lo)) return 0; /* also NaN */
    if (v > hi) return 255;
    uint32_t bits;
    std::memcpy(&bits, &v, 4);
    const uint32_t e = tab[(bits - lowest) >> 20];
    const uint32_t bias = (e >> 16) << 9, slope = e & 0xffffu, t = (bits >> 12) & 0xffu;
    return (uint8_t)((bias + slope * t) >> 16);
}

inline void decode_row(const uint8_t *src, int n, float *out) { /* :8611-8625, :3980-4069 */
    const float *lin = tables().to_linear;
    for (int i = 0; i < n; i++) {
        const float r = lin[src[i * 4 + 0]], g = lin[src[i * 4 + 1]], b = lin[src[i * 4 + 2]];
        const float a = (float)src[i * 4 + 3] * (1.0f / 255.0f);
        float *p = out + (size_t)i * kChannels;
        p[0] = r; p[1] = g; p[2] = b; p[3] = a;
        p[4] = r * a; p[5] = g * a; p[6] = b * a;
    }
}

inline void encode_row(const float *in, int n, uint8_t *dst) { /* :4140-4186, :8627-8687 */
    for (int i = 0; i < n; i++) {
        const float *p = in + (size_t)i * kChannels;
        const float a = p[3];
        float r, g, b;
        if (a < tiny()) {
            r = p[0]; g = p[1]; b = p[2];
        } else {
            const float ia = 1.0f / a;
            r = p[4] * ia; g = p[5] * ia; b = p[6] * ia;
        }
        dst[i * 4 + 0] = linear_to_srgb8(r);
        dst[i * 4 + 1] = linear_to_srgb8(g);
        dst[i * 4 + 2] = linear_to_srgb8(b);
        float f = a * 255.0f + 0.5f;
        if (f < 0.0f) f = 0.0f;
        if (f > 255.0f) f = 255.0f;
        dst[i * 4 + 3] = (uint8_t)f;
    }
}

/* rows[k] (k = n0 .. n1) weighted into out, in sequence */
inline void blend_rows(const Axis &v, int o, const std::vector<const float *> &rows, size_t floats, float *out) {
    const float *c = &v.taps[(size_t)o * v.stride];
    const int count = v.n1[o] - v.n0[o] + 1;
    const float *r0 = rows[0];
    for (size_t i = 0; i < floats; i++) out[i] = r0[i] * c[0];
    for (int k = 1; k < count; k++) {
        const float *r = rows[(size_t)k];
        const float ck = c[k];
        for (size_t i = 0; i < floats; i++) out[i] += r[i] * ck;
    }
}

/* src: w x h RGBA8 (non-premultiplied, sRGB colour, linear alpha) -> out: ow x oh RGBA8 */
inline void resize_srgb_rgba(const uint8_t *src, int w, int h, uint8_t *out, int ow, int oh) {
    if (w < 1 || h < 1 || ow < 1 || oh < 1) throw std::runtime_error("bake resize: empty image");
    Axis ax = make_axis(w, ow, true), ay = make_axis(h, oh, false);
    const bool v_first = vertical_pass_first(ax, ay);
    pack_for_rows(ax, decoded_row_end(ax) + 1);
    /* rows the vertical pass reads: decoded input rows (vertical first) or horizontally filtered ones; kept in a ring as wide
     * as the tallest window */
    const size_t row_floats = (size_t)(v_first ? w : ow) * kChannels;
    int tallest = 1;
    for (int o = 0; o < oh; o++) tallest = std::max(tallest, ay.n1[o] - ay.n0[o] + 1);
    std::vector<float> ring((size_t)tallest * row_floats), decoded((size_t)w * kChannels), blended(row_floats), filtered((size_t)ow * kChannels);
    std::vector<int> ring_row((size_t)tallest, -1);
    std::vector<const float *> rows;
    auto fetch = [&](int y) -> const float * {
        const int slot = y % tallest;
        float *r = &ring[(size_t)slot * row_floats];
        if (ring_row[(size_t)slot] != y) {
            if (v_first) {
                decode_row(src + (size_t)y * w * 4, w, r);
            } else {
                decode_row(src + (size_t)y * w * 4, w, decoded.data());
                filter_row(ax, decoded.data(), r);
            }
            ring_row[(size_t)slot] = y;
        }
        return r;
    };
    for (int o = 0; o < oh; o++) {
        if (ay.n1[o] < ay.n0[o] || ay.n0[o] < 0 || ay.n1[o] >= h) throw std::runtime_error("bake resize: bad vertical window");
        rows.clear();
        for (int y = ay.n0[o]; y <= ay.n1[o]; y++) rows.push_back(fetch(y));
        blend_rows(ay, o, rows, row_floats, blended.data());
        const float *line = blended.data();
        if (v_first) {
            filter_row(ax, blended.data(), filtered.data());
            line = filtered.data();
        }
        encode_row(line, ow, out + (size_t)o * ow * 4);
    }
}

} /* namespace bake */
} /* namespace glb */
} /* namespace raytracer */
#endif /* RT_BAKE_RESIZE_HPP */
