/*
 * glb_loader.hpp — binary glTF (.glb) -> rt_scene_desc, following the reference loader's rules
 * (src/scene.cpp:54-129 Scene::Scene, :148-162 load_images, :164-442 load_primitives,
 * :444-510 load_node). The reference uses tinygltf + stb; this is a from-scratch reader (own JSON
 * parser, own PNG / JPEG decoders) that reproduces what reaches the kernels:
 *
 *   - one instance per glTF node x mesh primitive, in node-index order then primitive order
 *     (= Embree attach order = instID, F11); POSITION / NORMAL / TEXCOORD_0 and indices required
 *     (:256-276), indices u8/u16/u32 widened to u32 (:359-401), byteStride honoured (:289-292);
 *   - node transform local = T * R * S * matrix (:18-21; tinygltf drops T/R/S when "matrix" is given), global =
 *     parents... * (local * scale(gs))
 *     (:137-146); a node without rotation keeps glm::quat{} = (0,0,0,0) whose mat4 cast is the
 *     identity (src/scene.hpp:48);
 *   - material classification (:188-254): KHR_materials_ior && KHR_materials_transmission ->
 *     dielectric(ior); else metallicFactor > 0.01 -> metallic(albedo tex|colour, roughness, emissive);
 *     else diffuse(albedo tex|colour, emissive); emissive = emissiveFactor * emissiveStrength, and the
 *     strength is 0 when KHR_materials_emissive_strength is absent (:203-211);
 *   - scene extras sky_color (3 numbers) and sky_strength (:80-94), default sky (0.5,0.7,1.0);
 *   - camera node: position = global[3], direction = normalize(rotation * (0,0,-1)),
 *     focal = 1 / tan(yfov / 2) (:109-128);
 *   - images: every image is resized to 512x512 RGBA8 and baked into the layer array
 *     (src/image_manager.hpp:39-100). Embedded PNG and JPEG are decoded by image_codecs.hpp to exactly
 *     the bytes stb_image hands the reference, and bake_resize.hpp restates the arithmetic of the reference's
 *     stbir_resize_uint8_srgb(..., 512, 512, 0, STBIR_RGBA) call operation for operation, so the baked layers are
 *     the reference's bytes (tests/test_image_codecs.py, against the reference's own library).
 *
 * Explicit fallbacks where the reference relies on undefined behaviour (F15): a primitive without a
 * material -> diffuse 0.8 grey; no camera node -> position (0,0,0), direction (0,0,-1), focal 1. (The reference
 * tests `if (camera_node_index)` with -1 meaning "none", :109, so it also skips a camera that is node 0 and
 * leaves the camera fields uninitialised; here node 0 is a camera like any other.)
 */
#pragma once

#include <zlib.h>

#include "bake_resize.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rt_api.h"
#include "image_codecs.hpp"

namespace raytracer {
namespace glb {

/* ------------------------------------------------------------------ minimal JSON */
struct Json {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    double num = 0;
    bool b = false;
    std::string str;
    std::vector<Json> arr;
    std::map<std::string, Json> obj;
    bool has(const std::string &k) const { return type == Obj && obj.count(k); }
    const Json &operator[](const std::string &k) const {
        static const Json null_json;
        auto it = obj.find(k);
        return it == obj.end() ? null_json : it->second;
    }
    const Json &operator[](size_t i) const { /* out of range (also a negative index cast to size_t): null, like a missing key */
        static const Json null_json;
        return (type == Arr && i < arr.size()) ? arr[i] : null_json;
    }
    size_t size() const { return type == Arr ? arr.size() : 0; }
    double number(double dflt) const { return type == Num ? num : dflt; }
    int integer(int dflt) const {
        if (type != Num || !(num == num)) return dflt;
        return num >= 2147483647.0 ? 2147483647 : (num <= -2147483648.0 ? (-2147483647 - 1) : (int)num);
    }
};

class JsonParser {
    const char *p, *end;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    }
    [[noreturn]] void fail(const char *m) { throw std::runtime_error(std::string("Failed to load .glTF : JSON ") + m); }
    Json value() {
        ws();
        if (p >= end) fail("truncated");
        Json j;
        if (*p == '{') {
            j.type = Json::Obj;
            p++;
            ws();
            if (p < end && *p == '}') { p++; return j; }
            for (;;) {
                ws();
                Json k = string();
                ws();
                if (p >= end || *p != ':') fail("expected ':'");
                p++;
                j.obj[k.str] = value();
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; break; }
                fail("expected ',' or '}'");
            }
        } else if (*p == '[') {
            j.type = Json::Arr;
            p++;
            ws();
            if (p < end && *p == ']') { p++; return j; }
            for (;;) {
                j.arr.push_back(value());
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; break; }
                fail("expected ',' or ']'");
            }
        } else if (*p == '"') {
            j = string();
        } else if (!strncmp(p, "true", 4)) {
            j.type = Json::Bool; j.b = true; p += 4;
        } else if (!strncmp(p, "false", 5)) {
            j.type = Json::Bool; p += 5;
        } else if (!strncmp(p, "null", 4)) {
            p += 4;
        } else {
            char *e = nullptr;
            j.type = Json::Num;
            j.num = strtod(p, &e);
            if (e == p) fail("bad token");
            p = e;
        }
        return j;
    }
    Json string() {
        if (p >= end || *p != '"') fail("expected string");
        p++;
        Json j;
        j.type = Json::Str;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                p++;
                switch (*p) {
                case 'n': j.str += '\n'; break;
                case 't': j.str += '\t'; break;
                case 'u': j.str += '?'; p += 4; break;
                default: j.str += *p;
                }
                p++;
            } else j.str += *p++;
        }
        if (p >= end) fail("unterminated string");
        p++;
        return j;
    }

  public:
    static Json parse(const char *s, size_t n) {
        JsonParser q;
        q.p = s;
        q.end = s + n;
        return q.value();
    }
};

/* ------------------------------------------------------------------ small column-major mat4 */
struct Mat4 {
    float m[16]; /* m[col * 4 + row], like glm */
    static Mat4 identity() {
        Mat4 r{};
        r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
        return r;
    }
};
inline Mat4 mul(const Mat4 &a, const Mat4 &b) { /* glm order: (a * b)[c] = sum_k a[k] * b[c][k] */
    Mat4 r{};
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++)
            r.m[c * 4 + row] = a.m[0 * 4 + row] * b.m[c * 4 + 0] + a.m[1 * 4 + row] * b.m[c * 4 + 1] +
                               a.m[2 * 4 + row] * b.m[c * 4 + 2] + a.m[3 * 4 + row] * b.m[c * 4 + 3];
    return r;
}
inline Mat4 translate(const float t[3]) {
    Mat4 r = Mat4::identity();
    r.m[12] = t[0]; r.m[13] = t[1]; r.m[14] = t[2];
    return r;
}
inline Mat4 scale(const float s[3]) {
    Mat4 r = Mat4::identity();
    r.m[0] = s[0]; r.m[5] = s[1]; r.m[10] = s[2];
    return r;
}
/* glm::mat4_cast(quat(w,x,y,z)); the all-zero default quaternion gives the identity */
inline Mat4 from_quat(const float q[4] /* x y z w */) {
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    Mat4 r = Mat4::identity();
    r.m[0] = 1 - 2 * (y * y + z * z); r.m[1] = 2 * (x * y + w * z);     r.m[2] = 2 * (x * z - w * y);
    r.m[4] = 2 * (x * y - w * z);     r.m[5] = 1 - 2 * (x * x + z * z); r.m[6] = 2 * (y * z + w * x);
    r.m[8] = 2 * (x * z + w * y);     r.m[9] = 2 * (y * z - w * x);     r.m[10] = 1 - 2 * (x * x + y * y);
    return r;
}

/* ------------------------------------------------------------------ images */
/* embedded PNG / JPEG -> RGBA8 exactly as stbi_load_from_memory(..., 4) delivers it to the reference
 * (image_codecs.hpp) */
inline std::vector<uint8_t> png_decode(const uint8_t *d, size_t n, uint32_t &w, uint32_t &h) {
    img::Image im = img::decode_rgba8(d, n, /* tinygltf_16bit_quirk = */ true);
    w = im.w;
    h = im.h;
    return std::move(im.rgba);
}

/* RGBA8 PNG writer (filter 0, one zlib stream) — what stbi_write_png does for out.png (src/util.hpp:27) */
inline bool png_write(const std::string &path, const uint8_t *rgba, uint32_t w, uint32_t h) {
    std::vector<uint8_t> raw((size_t)(w * 4 + 1) * h);
    for (uint32_t y = 0; y < h; y++) {
        raw[(size_t)(w * 4 + 1) * y] = 0;
        memcpy(&raw[(size_t)(w * 4 + 1) * y + 1], rgba + (size_t)y * w * 4, (size_t)w * 4);
    }
    uLongf clen = compressBound(raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), raw.size(), 6) != Z_OK) return false;
    std::ofstream f(path, std::ios::binary);
    if (!f) return false;
    auto chunk = [&](const char *type, const uint8_t *data, uint32_t len) {
        uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len,
                          (uint8_t)type[0], (uint8_t)type[1], (uint8_t)type[2], (uint8_t)type[3]};
        f.write((const char *)hdr, 8);
        if (len) f.write((const char *)data, len);
        uLong crc = crc32(0L, hdr + 4, 4);
        if (len) crc = crc32(crc, data, len);
        const uint8_t c[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
        f.write((const char *)c, 4);
    };
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    f.write((const char *)sig, 8);
    const uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w, (uint8_t)(h >> 24),
                              (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 8, 6, 0, 0, 0};
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", comp.data(), (uint32_t)clen);
    chunk("IEND", nullptr, 0);
    return (bool)f;
}

/* Resize to one 512x512 layer exactly as the reference's bake does (stbir_resize_uint8_srgb(..., STBIR_RGBA),
 * src/image_manager.hpp:52-62): bake_resize.hpp restates that call's single-precision arithmetic in its order, so the
 * texels are the reference's bytes (tests/test_image_codecs.py compares them with the reference's own library). The
 * reference resizes EVERY image, 512x512 ones included (both axes then point-sample, which is the identity), so they are
 * passed through untouched. */
inline std::vector<uint8_t> resize_to_layer(const std::vector<uint8_t> &src, uint32_t w, uint32_t h) {
    const uint32_t N = RT_TEX_SIZE;
    if (w == N && h == N) return src;
    std::vector<uint8_t> out((size_t)N * N * 4);
    bake::resize_srgb_rgba(src.data(), (int)w, (int)h, out.data(), (int)N, (int)N);
    return out;
}

/* ------------------------------------------------------------------ the loaded scene */
struct LoadedScene {
    struct Inst {
        std::vector<float> positions, normals, uvs;
        std::vector<uint32_t> indices;
        Mat4 transform;
        rt_material material;
        int node, mesh, primitive;
    };
    std::vector<Inst> instances;
    std::vector<uint8_t> texture_layers; /* n * 512 * 512 * 4 */
    uint32_t texture_layer_count = 0;
    float sky_color[3] = {0.5f, 0.7f, 1.0f};
    float camera_position[3] = {0, 0, 0}, camera_direction[3] = {0, 0, -1};
    float camera_focal_length = 1.0f;
    bool has_camera = false;
    std::vector<rt_instance> rt_instances;

    rt_scene_desc desc() {
        rt_instances.clear();
        for (auto &i : instances) {
            rt_instance r{};
            r.positions = i.positions.data();
            r.normals = i.normals.data();
            r.uvs = i.uvs.data();
            r.indices = i.indices.data();
            r.vertex_count = (uint32_t)(i.positions.size() / 3);
            r.index_count = (uint32_t)i.indices.size();
            memcpy(r.transform, i.transform.m, sizeof(r.transform));
            r.material = i.material;
            rt_instances.push_back(r);
        }
        rt_scene_desc d{};
        d.instances = rt_instances.data();
        d.instance_count = (uint32_t)rt_instances.size();
        d.texture_layers = texture_layer_count ? texture_layers.data() : nullptr;
        d.texture_layer_count = texture_layer_count;
        memcpy(d.sky_color, sky_color, sizeof(sky_color));
        return d;
    }
};

namespace detail {
struct Reader {
    Json j;
    std::vector<uint8_t> bin;
    [[noreturn]] static void bad(const char *what) { throw std::runtime_error(std::string("Failed to load .glTF : ") + what); }
    /* start of a buffer view (+ extra bytes) and how many bytes remain in it; every read is checked against that */
    const uint8_t *view_ptr(int view, size_t extra, size_t &stride_out, size_t *avail_out = nullptr) const {
        if (view < 0 || (size_t)view >= j["bufferViews"].size()) bad("bufferView index out of range");
        const Json &v = j["bufferViews"][(size_t)view];
        if (v["buffer"].integer(0) != 0) bad("only the GLB-embedded buffer 0 is supported");
        const int off_i = v["byteOffset"].integer(0), len_i = v["byteLength"].integer(0), stride_i = v["byteStride"].integer(0);
        if (off_i < 0 || len_i < 0 || stride_i < 0 || stride_i > 4096) bad("bad bufferView");
        stride_out = (size_t)stride_i;
        const size_t off = (size_t)off_i, len = (size_t)len_i;
        if (off > bin.size() || len > bin.size() - off || extra > len) bad("buffer view out of range");
        if (avail_out) *avail_out = len - extra;
        return bin.data() + off + extra;
    }
    const Json &accessor_at(int accessor) const {
        if (accessor < 0 || (size_t)accessor >= j["accessors"].size()) bad("accessor index out of range");
        const Json &a = j["accessors"][(size_t)accessor];
        if (a["byteOffset"].integer(0) < 0 || a["count"].integer(0) < 0) bad("bad accessor");
        if (a.has("sparse")) bad("sparse accessors are not supported");
        return a;
    }
    std::vector<float> floats(int accessor, int comps) const {
        const Json &a = accessor_at(accessor);
        if (a["componentType"].integer(0) != 5126) bad("float accessor expected");
        size_t stride = 0, avail = 0;
        const uint8_t *p = view_ptr(a["bufferView"].integer(-1), (size_t)a["byteOffset"].integer(0), stride, &avail);
        if (!stride) stride = (size_t)comps * 4;
        const size_t n = (size_t)a["count"].integer(0);
        if (n && ((n - 1) > (avail / stride) || (n - 1) * stride + (size_t)comps * 4 > avail)) bad("accessor reads past its buffer view");
        std::vector<float> out(n * comps);
        for (size_t i = 0; i < n; i++) memcpy(&out[i * comps], p + i * stride, (size_t)comps * 4);
        return out;
    }
    std::vector<uint32_t> indices(int accessor) const {
        const Json &a = accessor_at(accessor);
        size_t stride = 0, avail = 0;
        const uint8_t *p = view_ptr(a["bufferView"].integer(-1), (size_t)a["byteOffset"].integer(0), stride, &avail);
        const size_t n = (size_t)a["count"].integer(0);
        const int ct = a["componentType"].integer(0);
        const size_t es = ct == 5125 ? 4 : ct == 5123 ? 2 : ct == 5121 ? 1 : 0;
        if (!es) throw std::runtime_error("Index component type not supported!"); /* src/scene.cpp:394-400 */
        if (n > avail / es) bad("accessor reads past its buffer view");
        std::vector<uint32_t> out(n);
        for (size_t i = 0; i < n; i++) {
            if (ct == 5125) { uint32_t v; memcpy(&v, p + i * 4, 4); out[i] = v; }
            else if (ct == 5123) { uint16_t v; memcpy(&v, p + i * 2, 2); out[i] = v; }
            else out[i] = p[i];
        }
        return out;
    }
};
} // namespace detail

/* bytes behind a glTF "uri": RFC 2397 data URI (base64) or a file relative to the directory of the .glb */
inline std::vector<uint8_t> read_uri(const std::string &uri, const std::string &glb_path) {
    if (uri.compare(0, 5, "data:") == 0) {
        const size_t comma = uri.find(',');
        if (comma == std::string::npos || comma < 7 || uri.compare(comma - 7, 7, ";base64") != 0)
            throw std::runtime_error("Failed to load .glTF : unsupported data URI");
        std::vector<uint8_t> out;
        uint32_t acc = 0;
        int bits = 0;
        for (size_t i = comma + 1; i < uri.size(); i++) {
            const char c = uri[i];
            int v;
            if (c >= 'A' && c <= 'Z') v = c - 'A';
            else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
            else if (c >= '0' && c <= '9') v = c - '0' + 52;
            else if (c == '+' || c == '-') v = 62;
            else if (c == '/' || c == '_') v = 63;
            else continue; /* '=' padding, whitespace */
            acc = (acc << 6) | (uint32_t)v;
            bits += 6;
            if (bits >= 8) {
                bits -= 8;
                out.push_back((uint8_t)(acc >> bits));
            }
        }
        return out;
    }
    std::string decoded; /* percent-decoding, as tinygltf does before opening the file */
    for (size_t i = 0; i < uri.size(); i++) {
        if (uri[i] == '%' && i + 2 < uri.size() && isxdigit((unsigned char)uri[i + 1]) && isxdigit((unsigned char)uri[i + 2])) {
            decoded.push_back((char)std::stoi(uri.substr(i + 1, 2), nullptr, 16));
            i += 2;
        } else decoded.push_back(uri[i]);
    }
    const size_t slash = glb_path.find_last_of("/\\");
    const std::string full = (slash == std::string::npos || decoded[0] == '/') ? decoded : glb_path.substr(0, slash + 1) + decoded;
    std::ifstream f(full, std::ios::binary);
    if (!f) throw std::runtime_error("Failed to load .glTF : cannot open " + full);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

inline LoadedScene load(const std::string &path, const float global_scale[3] = nullptr) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("Failed to load .glTF : cannot open " + path); /* src/scene.cpp:68-70 */
    std::vector<uint8_t> file((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (file.size() < 20 || memcmp(file.data(), "glTF", 4)) throw std::runtime_error("Failed to load .glTF : not a binary glTF");
    auto le32 = [&](size_t o) { uint32_t v; memcpy(&v, &file[o], 4); return v; };
    detail::Reader rd;
    for (size_t o = 12; o + 8 <= file.size();) {
        const uint32_t len = le32(o), type = le32(o + 4);
        if (o + 8 + len > file.size()) throw std::runtime_error("Failed to load .glTF : truncated chunk");
        if (type == 0x4E4F534A) rd.j = JsonParser::parse((const char *)&file[o + 8], len);
        else if (type == 0x004E4942) rd.bin.assign(file.begin() + o + 8, file.begin() + o + 8 + len);
        o += 8 + ((len + 3) & ~3u);
    }
    const Json &j = rd.j;
    LoadedScene out;
    const float gs_default[3] = {1, 1, 1};
    const float *gs = global_scale ? global_scale : gs_default;

    /* images -> 512x512 layers (src/scene.cpp:148-162) */
    if (j.type != Json::Obj) throw std::runtime_error("Failed to load .glTF : no JSON chunk");
    const size_t n_images = j["images"].size();
    if (n_images > RT_MAX_IMAGES) throw std::runtime_error("Too many images uploaded"); /* src/image_manager.hpp:41-44 */
    for (size_t i = 0; i < n_images; i++) {
        const Json &im = j["images"][i];
        std::vector<uint8_t> ext; /* an image given by "uri" (data: URI or a file next to the .glb, as tinygltf resolves it) */
        const uint8_t *p = nullptr;
        size_t n_bytes = 0;
        if (im.has("bufferView")) {
            size_t stride;
            p = rd.view_ptr(im["bufferView"].integer(-1), 0, stride, &n_bytes);
        } else if (im.has("uri")) {
            ext = read_uri(im["uri"].str, path);
            p = ext.data();
            n_bytes = ext.size();
        } else {
            throw std::runtime_error("Failed to load .glTF : image without bufferView or uri");
        }
        uint32_t w = 0, h = 0;
        std::vector<uint8_t> px = png_decode(p, n_bytes, w, h);
        std::vector<uint8_t> layer = resize_to_layer(px, w, h);
        out.texture_layers.insert(out.texture_layers.end(), layer.begin(), layer.end());
        out.texture_layer_count++;
    }

    /* node hierarchy: parent links by traversal from the default scene (src/scene.cpp:96-99,444-480) */
    const size_t n_nodes = j["nodes"].size();
    std::vector<int> parent(n_nodes, -1);
    std::vector<char> reached(n_nodes, 0);
    std::vector<Mat4> local(n_nodes);
    for (size_t n = 0; n < n_nodes; n++) {
        const Json &nd = j["nodes"][n];
        float t[3] = {0, 0, 0}, s[3] = {1, 1, 1}, q[4] = {0, 0, 0, 0};
        Mat4 m = Mat4::identity();
        /* tinygltf treats "matrix" and T/R/S as exclusive: once a node has a numeric "matrix" array it does not even
         * parse translation / rotation / scale (deps/include/tiny_gltf.h:5113-5118) */
        bool has_matrix = nd["matrix"].type == Json::Arr;
        for (size_t k = 0; has_matrix && k < nd["matrix"].size(); k++) has_matrix = nd["matrix"][k].type == Json::Num;
        if (!has_matrix) {
            if (nd["translation"].size() == 3) for (int k = 0; k < 3; k++) t[k] = (float)nd["translation"][k].num;
            if (nd["rotation"].size() == 4) for (int k = 0; k < 4; k++) q[k] = (float)nd["rotation"][k].num;
            if (nd["scale"].size() == 3) for (int k = 0; k < 3; k++) s[k] = (float)nd["scale"][k].num;
        }
        if (has_matrix && nd["matrix"].size() == 16) for (int k = 0; k < 16; k++) m.m[k] = (float)nd["matrix"][k].num;
        local[n] = mul(mul(mul(translate(t), from_quat(q)), scale(s)), m); /* T * R * S * matrix */
    }
    const Json &scene = j["scenes"][(size_t)std::max(0, j["scene"].integer(0))];
    int camera_node = -1;
    /* Scene::load_node (src/scene.cpp:444-480) recurses in PRE-ORDER: scene roots in order, a node before its children,
     * children in order. The last node WITH a camera that this order visits wins (:455-457), and a node reached a
     * second time (not a valid glTF, but nothing stops it) gets the later visitor as its parent (:452). The explicit
     * stack below reproduces that order (children pushed in reverse); a node is never followed into itself (the
     * reference would recurse until its stack overflows), and a pathological DAG is cut off after 64 visits per node. */
    struct Visit { int node, parent; bool leave; };
    std::vector<Visit> stack;
    std::vector<char> on_path(n_nodes, 0);
    size_t visits = 0;
    for (size_t k = scene["nodes"].size(); k-- > 0;) stack.push_back({scene["nodes"][k].integer(0), -1, false});
    while (!stack.empty()) {
        const Visit v = stack.back();
        stack.pop_back();
        const int n = v.node;
        if (v.leave) {
            on_path[(size_t)n] = 0;
            continue;
        }
        if (n < 0 || (size_t)n >= n_nodes) throw std::runtime_error("Failed to load .glTF : node index out of range");
        if (on_path[(size_t)n]) continue; /* a cycle: never follow a node into itself */
        if (++visits > 64 * n_nodes) throw std::runtime_error("Failed to load .glTF : node graph is not a forest");
        reached[(size_t)n] = 1;
        parent[(size_t)n] = v.parent;
        const Json &nd = j["nodes"][(size_t)n];
        if (nd.has("camera")) camera_node = n;
        on_path[(size_t)n] = 1;
        stack.push_back({n, v.parent, true});
        for (size_t k = nd["children"].size(); k-- > 0;) {
            const int c = nd["children"][k].integer(-1);
            if (c < 0 || (size_t)c >= n_nodes) throw std::runtime_error("Failed to load .glTF : node index out of range");
            stack.push_back({c, n, false});
        }
    }
    auto global_matrix = [&](int n) { /* src/scene.cpp:137-146 */
        Mat4 m = mul(local[(size_t)n], scale(gs));
        for (int p = parent[(size_t)n]; p >= 0; p = parent[(size_t)p]) m = mul(local[(size_t)p], m);
        return m;
    };

    /* scene extras (src/scene.cpp:80-94) */
    const Json &extras = scene["extras"];
    if (extras["sky_color"].size() == 3) for (int k = 0; k < 3; k++) out.sky_color[k] = (float)extras["sky_color"][k].num;
    if (extras["sky_strength"].type == Json::Num) for (int k = 0; k < 3; k++) out.sky_color[k] *= (float)extras["sky_strength"].num;

    /* instances in node-index order (src/scene.cpp:101-106), primitives in order */
    for (size_t n = 0; n < n_nodes; n++) {
        const Json &nd = j["nodes"][n];
        if (!reached[n] || !nd.has("mesh")) continue;
        const int mesh = nd["mesh"].integer(0);
        const Json &prims = j["meshes"][(size_t)mesh]["primitives"];
        for (size_t pi = 0; pi < prims.size(); pi++) {
            const Json &pr = prims[pi];
            const Json &at = pr["attributes"];
            if (!pr.has("indices") || !at.has("POSITION") || !at.has("NORMAL") || !at.has("TEXCOORD_0"))
                throw std::runtime_error("Failed to load .glTF : primitives need indices, POSITION, NORMAL and TEXCOORD_0"); /* :256-276 */
            LoadedScene::Inst in;
            in.node = (int)n; in.mesh = mesh; in.primitive = (int)pi;
            in.positions = rd.floats(at["POSITION"].integer(0), 3);
            in.normals = rd.floats(at["NORMAL"].integer(0), 3);
            in.uvs = rd.floats(at["TEXCOORD_0"].integer(0), 2);
            in.indices = rd.indices(pr["indices"].integer(0));
            in.transform = global_matrix((int)n);
            rt_material m{};
            m.albedo_image = -1;
            m.ior = 1.5f;
            if (!pr.has("material")) { /* F15 fallback: the reference indexes materials[-1] */
                m.type = RT_MAT_DIFFUSE;
                m.albedo_color[0] = m.albedo_color[1] = m.albedo_color[2] = 0.8f;
            } else {
                const Json &mat = j["materials"][(size_t)pr["material"].integer(0)];
                const Json &pbr = mat["pbrMetallicRoughness"];
                for (int k = 0; k < 3; k++) m.albedo_color[k] = pbr["baseColorFactor"].size() >= 3 ? (float)pbr["baseColorFactor"][k].num : 1.0f;
                const Json &ext = mat["extensions"];
                float strength = 0.0f; /* 0 when the extension is absent (src/scene.cpp:203-211) */
                if (ext.has("KHR_materials_emissive_strength")) strength = (float)ext["KHR_materials_emissive_strength"]["emissiveStrength"].number(0.0); /* a missing key reads as 0 in tinygltf */
                for (int k = 0; k < 3; k++) m.emissive[k] = (mat["emissiveFactor"].size() == 3 ? (float)mat["emissiveFactor"][k].num : 0.0f) * strength;
                const int tex = pbr["baseColorTexture"]["index"].integer(-1);
                int image = tex >= 0 ? j["textures"][(size_t)tex]["source"].integer(-1) : -1;
                if (image < -1 || image >= (int)n_images) throw std::runtime_error("Failed to load .glTF : texture source out of range");
                if (ext.has("KHR_materials_ior") && ext.has("KHR_materials_transmission")) {
                    m.type = RT_MAT_DIELECTRIC;
                    m.ior = (float)ext["KHR_materials_ior"]["ior"].number(0.0); /* tinygltf: Value::Get of a missing key -> 0 */
                } else if ((float)pbr["metallicFactor"].number(1.0) > 0.01f) { /* glTF default metallicFactor = 1 */
                    m.type = RT_MAT_METALLIC;
                    m.roughness = (float)pbr["roughnessFactor"].number(1.0);
                    m.albedo_image = image;
                } else {
                    m.type = RT_MAT_DIFFUSE;
                    m.albedo_image = image;
                }
            }
            in.material = m;
            out.instances.push_back(std::move(in));
        }
    }

    /* camera (src/scene.cpp:109-128) */
    if (camera_node >= 0) {
        const Mat4 m = global_matrix(camera_node);
        for (int k = 0; k < 3; k++) out.camera_position[k] = m.m[12 + k];
        /* glm::quat_cast of the global matrix's upper 3x3 — scale and all, as the reference does (:116) —
         * then normalize(q * (0,0,-1)) (:120-121). With a pure rotation (and uniform scale) this is minus the
         * third column, normalised; a non-uniform scale above the camera skews it exactly like the reference. */
        const float m00 = m.m[0], m01 = m.m[1], m02 = m.m[2], m10 = m.m[4], m11 = m.m[5], m12 = m.m[6], m20 = m.m[8], m21 = m.m[9], m22 = m.m[10];
        const float fx = m00 - m11 - m22, fy = m11 - m00 - m22, fz = m22 - m00 - m11, fw = m00 + m11 + m22;
        int big = 0;
        float fb = fw;
        if (fx > fb) { fb = fx; big = 1; }
        if (fy > fb) { fb = fy; big = 2; }
        if (fz > fb) { fb = fz; big = 3; }
        const float bv = std::sqrt(fb + 1.0f) * 0.5f, mult = 0.25f / bv;
        float qw, qx, qy, qz;
        if (big == 0) { qw = bv; qx = (m12 - m21) * mult; qy = (m20 - m02) * mult; qz = (m01 - m10) * mult; }
        else if (big == 1) { qw = (m12 - m21) * mult; qx = bv; qy = (m01 + m10) * mult; qz = (m20 + m02) * mult; }
        else if (big == 2) { qw = (m20 - m02) * mult; qx = (m01 + m10) * mult; qy = bv; qz = (m12 + m21) * mult; }
        else { qw = (m01 - m10) * mult; qx = (m20 + m02) * mult; qy = (m12 + m21) * mult; qz = bv; }
        /* q * v = v + 2 * (w * (qv x v) + qv x (qv x v)), v = (0, 0, -1) */
        const float v[3] = {0.0f, 0.0f, -1.0f}, qv[3] = {qx, qy, qz};
        const float uv[3] = {qv[1] * v[2] - v[1] * qv[2], qv[2] * v[0] - v[2] * qv[0], qv[0] * v[1] - v[0] * qv[1]};
        const float uuv[3] = {qv[1] * uv[2] - uv[1] * qv[2], qv[2] * uv[0] - uv[2] * qv[0], qv[0] * uv[1] - uv[0] * qv[1]};
        float d[3];
        for (int k = 0; k < 3; k++) d[k] = v[k] + ((uv[k] * qw) + uuv[k]) * 2.0f;
        const float inv = 1.0f / std::sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
        for (int k = 0; k < 3; k++) out.camera_direction[k] = d[k] * inv;
        const Json &cam = j["cameras"][(size_t)j["nodes"][(size_t)camera_node]["camera"].integer(0)];
        const float yfov = (float)cam["perspective"]["yfov"].number(1.0);
        out.camera_focal_length = 1.0f / std::tan(yfov / 2.0f);
        out.has_camera = true;
    }
    return out;
}

} // namespace glb
} // namespace raytracer
