/*
 * rt_shade.h — per-ray shading of the path: camera rays, hit attribute interpolation, the three
 * material scatters, sky miss shading, fp16 ray-state quantisation, output quantisation.
 *
 * New code that follows the behaviour of (not copied from):
 *   src/camera.hpp:12-131       RayData / Camera::get_ray / pixel_sample_square
 *   src/trace_ray.hpp:11-82     one path segment
 *   src/material.hpp:45-53,68-238   Texture::sample, Material*::scatter / emitted
 *   src/util.hpp:82-125         linear_to_gamma, near_zero, length_squared, reflect, refract
 *   src/util.hpp:16-22          image read-back quantisation
 * Operation order is pinned by the arithmetic contract in DESIGN.md; the oracle
 * (oracle/rt_oracle.cpp, an independent restatement) uses the same order, so the two agree
 * bit for bit wherever the same hit is found.
 */
#ifndef RT_SHADE_H
#define RT_SHADE_H

#include "rt_traverse.h"

#define RT_TEX_DIM 512 /* src/image_manager.hpp:14 */

/* Instance record, 80 bytes = 5 x float4 (GeometryData minus the buffers, src/scene.hpp:17-24) */
struct RtInstance {
    float nmat[9];        /* column-major transpose(inverse(mat3(T))) (src/scene.cpp:502) */
    int32_t type;         /* rt_material_type */
    int32_t albedo_image; /* >= 0: layer; < 0: albedo colour */
    float albedo[3];
    float roughness;
    float ior;
    float emissive[3];
    uint32_t first_tri;   /* global id of primitive 0 */
};

/* Per-triangle shading record in leaf order, 64 bytes = 4 x float4:
 *   s0 = {n0.x, n0.y, n0.z, n1.x}  s1 = {n1.y, n1.z, n2.x, n2.y}
 *   s2 = {n2.z, uv0.x, uv0.y, uv1.x}  s3 = {uv1.y, uv2.x, uv2.y, inst_id_bits}
 * (the three index / normal / uv gathers of src/trace_ray.hpp:29-41 pre-resolved at commit) */

struct RtCamera { /* src/camera.hpp:65-72 */
    f3 center, pixel00, du, dv;
    int32_t w, h;
};

struct RtScene {
    RtBvh bvh;
    const rt_float4 *shade;    /* 4 per triangle, leaf order */
    const RtInstance *inst;
    const uint8_t *tex_raw;    /* layer-major RGBA8, used by the host emulation only */
#if defined(__CUDACC__)
    cudaTextureObject_t tex;   /* layered 512x512xN uchar4, point, element read */
#else
    unsigned long long tex;
#endif
    uint32_t n_layers;
    f3 sky;
};

/* fp32 origin + fp16-rounded direction / attenuation / radiance (RayData, F6) */
struct RtRayState {
    f3 org, dir, att, rad;
};

/* src/render_megakernel.cpp:144-146 (x * H_pad + y) / src/render_wavefront.cpp:69-73 (x + y*W);
 * std::hash<size_t> is the identity in libstdc++ (F2, F3) */
RT_HD uint32_t rt_pixel_seed(int wavefront, int x, int y, int w, int h) {
    if (!wavefront) {
        uint64_t h_pad = (uint64_t)((h + 7) / 8) * 8;
        return (uint32_t)((uint64_t)x * h_pad + (uint64_t)y);
    }
    return (uint32_t)((uint64_t)x + (uint64_t)y * (uint64_t)w);
}

/* src/camera.hpp:109-131 */
RT_HD RtRayState rt_camera_ray(const RtCamera &c, int x, int y, XorShift32 &rng) {
    f3 pixel_center = (c.pixel00 + ((float)x * c.du)) + ((float)y * c.dv);
    float px = -0.5f + rng.next();
    float py = -0.5f + rng.next();
    f3 pixel_sample = pixel_center + ((px * c.du) + (py * c.dv));
    RtRayState r;
    r.org = c.center;
    r.dir = round_half3(pixel_sample - c.center);
    r.att = mk3(1.0f, 1.0f, 1.0f);
    r.rad = mk3(0.0f, 0.0f, 0.0f);
    return r;
}

/* nearest / repeat / normalised addressing (sampler of src/render_megakernel.cpp:99-103):
 * i = floor((s - floor(s)) * 512), wrapped (OpenCL 3.0 section 8.2) */
RT_HD int rt_wrap_texel(float s) {
    float u = (s - floorf(s)) * (float)RT_TEX_DIM;
    int i = (int)floorf(u);
    if (i > RT_TEX_DIM - 1) i -= RT_TEX_DIM;
    if (i < 0) i = 0;
    return i;
}

RT_HD f3 rt_texture_sample(const RtScene &s, int layer, float su, float sv) {
    if (layer < 0 || (uint32_t)layer >= s.n_layers) return mk3(0.0f, 0.0f, 0.0f);
    int ix = rt_wrap_texel(su), iy = rt_wrap_texel(sv);
#if RT_DEVICE_CODE
    uchar4 t = tex2DLayered<uchar4>(s.tex, (float)ix, (float)iy, layer);
    return mk3(rt_u8_to_unit((float)t.x), rt_u8_to_unit((float)t.y), rt_u8_to_unit((float)t.z));
#else
    const uint8_t *p = s.tex_raw + (((size_t)layer * RT_TEX_DIM + (size_t)iy) * RT_TEX_DIM + (size_t)ix) * 4;
    return mk3(rt_u8_to_unit((float)p[0]), rt_u8_to_unit((float)p[1]), rt_u8_to_unit((float)p[2]));
#endif
}

/* src/util.hpp:103-107 */
RT_HD bool rt_near_zero(f3 e) {
    const float s = 1e-8f;
    return (fabsf(e.x) < s) && (fabsf(e.y) < s) && (fabsf(e.z) < s);
}
/* src/util.hpp:114-116 */
RT_HD f3 rt_reflect(f3 v, f3 n) { return v - (2.0f * dot3(v, n)) * n; }
/* src/util.hpp:118-125; |perp|^2 evaluated as length() squared (:109-112) */
RT_HD f3 rt_refract(f3 uv, f3 n, float eta) {
    float cos_theta = rt_min(dot3(-uv, n), 1.0f);
    f3 perp = eta * (uv + cos_theta * n);
    float l = length3(perp);
    f3 par = (-rt_sqrt(fabsf(1.0f - l * l))) * n;
    return perp + par;
}
/* src/material.hpp:124-129, pow(x,5) = ((x*x)*(x*x))*x */
RT_HD float rt_reflectance(float cosine, float ref_idx) {
    float r0 = rt_div(1.0f - ref_idx, 1.0f + ref_idx);
    r0 = r0 * r0;
    float x = 1.0f - cosine;
    float x2 = x * x;
    return r0 + (1.0f - r0) * ((x2 * x2) * x);
}

/* One path segment after the closest hit is known (src/trace_ray.hpp:24-81).
 * Returns true when the path ended; `result` is then the sample colour. Otherwise the ray state
 * has been advanced (origin fp32, the rest still fp32: the caller re-quantises, F6). */
RT_HD bool rt_shade_segment(const RtScene &s, const RtHit &h, XorShift32 &rng, f3 &org, f3 &dir,
                            f3 &att, f3 &rad, f3 &result) {
    if (h.tri == RT_MISS) {
        result = att * (s.sky + rad); /* :25-27 */
        return true;
    }
    const rt_float4 *sp = s.shade + (size_t)h.tri * 4;
    rt_float4 s0, s1, s2, s3;
#if RT_USE_LDG256
    rt_ldg2(sp, s0, s1);
    rt_ldg2(sp + 2, s2, s3);
#else
    s0 = rt_ldg_hint<RT_SHADE_LD_HINT>(sp);
    s1 = rt_ldg_hint<RT_SHADE_LD_HINT>(sp + 1);
    s2 = rt_ldg_hint<RT_SHADE_LD_HINT>(sp + 2);
    s3 = rt_ldg_hint<RT_SHADE_LD_HINT>(sp + 3);
#endif
    const RtInstance &g = s.inst[rt_f2u(s3.w)];
    const float bx = h.u, by = h.v;
    const float bw = (1.0f - bx) - by;
    const float su = (bw * s2.y + bx * s2.w) + by * s3.y; /* :43-44 */
    const float sv = (bw * s2.z + bx * s3.x) + by * s3.z;
    const f3 n0 = mk3(s0.x, s0.y, s0.z), n1 = mk3(s0.w, s1.x, s1.y), n2 = mk3(s1.z, s1.w, s2.x);
    const f3 vn = normalize3(((bw * n0) + (bx * n1)) + (by * n2)); /* :47-50 */
    const float *m = g.nmat;
    const f3 gn = mk3((m[0] * vn.x + m[3] * vn.y) + m[6] * vn.z,
                      (m[1] * vn.x + m[4] * vn.y) + m[7] * vn.z,
                      (m[2] * vn.x + m[5] * vn.y) + m[8] * vn.z); /* :52 */
    const f3 normal = normalize3(gn);                             /* :53-54 */
    const f3 ndir = normalize3(dir);                              /* :56-57 */
    const int type = g.type;
    if (type == 1 || type == 2) rad = rad + mk3(g.emissive[0], g.emissive[1], g.emissive[2]); /* :59 */

    f3 sdir, satt;
    bool scattered;
    if (type == 1) { /* MaterialDiffuse::scatter, src/material.hpp:72-86 */
        sdir = normal + rng.random_unit_vector();
        if (rt_near_zero(ndir)) sdir = normal; /* F8 */
        satt = g.albedo_image >= 0 ? rt_texture_sample(s, g.albedo_image, su, sv)
                                   : mk3(g.albedo[0], g.albedo[1], g.albedo[2]);
        scattered = true;
    } else if (type == 2) { /* MaterialMetallic::scatter, :99-110 */
        f3 reflected = rt_reflect(ndir, normal);
        sdir = reflected + g.roughness * rng.random_unit_vector();
        satt = g.albedo_image >= 0 ? rt_texture_sample(s, g.albedo_image, su, sv)
                                   : mk3(g.albedo[0], g.albedo[1], g.albedo[2]);
        scattered = dot3(sdir, normal) > 0.0f;
    } else if (type == 3) { /* MaterialDielectric::scatter, :131-160 */
        satt = mk3(1.0f, 1.0f, 1.0f);
        bool front_face = dot3(ndir, normal) < 0.0f;
        f3 n = front_face ? normal : -normal;
        float ratio = front_face ? rt_div(1.0f, g.ior) : g.ior;
        f3 unit_direction = normalize3(ndir);
        float cos_theta = rt_min(dot3(-unit_direction, n), 1.0f);
        float sin_theta = rt_sqrt(1.0f - cos_theta * cos_theta);
        bool cannot_refract = ratio * sin_theta > 1.0f;
        bool refl = cannot_refract;
        if (!refl) refl = rt_reflectance(cos_theta, ratio) > rng.next(0.0f, 1.0f); /* F5 */
        sdir = refl ? rt_reflect(unit_direction, n) : rt_refract(unit_direction, n, ratio);
        scattered = true;
    } else {
        scattered = false;
        sdir = mk3(0.0f, 0.0f, 0.0f);
        satt = sdir;
    }
    if (scattered) {
        org = org + dir * h.t; /* :62-64 */
        dir = sdir;
        att = att * satt;
        return false;
    }
    result = att * rad; /* :73 */
    return true;
}

/* F10: unorm8 image store (saturate, round to nearest even) followed by the reference's
 * read-back `* 255.0f` truncating cast (src/util.hpp:16-22) */
RT_HD uint8_t rt_output_byte(float g) {
    float c = g * 255.0f;
    float q;
    if (!(c > 0.0f)) q = 0.0f;
    else if (c >= 255.0f) q = 255.0f;
    else q = rintf(c);
    float back = rt_div(q, 255.0f) * 255.0f;
    return (uint8_t)back;
}

/* mean over samples, sqrt gamma (src/render_megakernel.cpp:154-156, src/util.hpp:82-84) */
RT_HD uint32_t rt_resolve_pixel(float sr, float sg, float sb, float spp) {
    uint32_t r = rt_output_byte(rt_sqrt(rt_div(sr, spp)));
    uint32_t g = rt_output_byte(rt_sqrt(rt_div(sg, spp)));
    uint32_t b = rt_output_byte(rt_sqrt(rt_div(sb, spp)));
    return r | (g << 8) | (b << 16) | (255u << 24);
}

RT_HD float rt_clamp01(float v) { return rt_min(rt_max(v, 0.0f), 1.0f); }

/* RenderContext by value (src/render_context.hpp:12-24) + the renderer's scalars */
struct RtFrameParams {
    RtCamera cam;
    uint32_t max_depth, spp;
    uint32_t seed_salt;
    uint32_t rank, world, tile_size;
    int32_t wavefront_seed; /* F3: 0 = x*H_pad + y, 1 = x + y*W */
    int32_t clamp_samples;  /* F9: wavefront clamps every sample to [0,1] */
    int32_t resume;         /* continue the streams / accumulation left by the previous frame */
    int32_t keep_foreign;   /* tile shards: leave other ranks' RGBA8 pixels alone (this image is a gather destination) */
    /* scheduling knobs of the persistent kernels (no effect on results) */
    int32_t tune_refill;    /* leave the traversal loop once this many lanes have finished */
    int32_t tune_ctx;       /* megakernel: ray contexts per lane (0 = the round-1 one-pixel-in-registers kernel) */
    int32_t tune_shade;     /* megakernel contexts: shade once this many lanes have a hit waiting ... */
    int32_t tune_idle;      /* ... or this many lanes can neither traverse nor switch */
    int32_t tune_inflight;  /* wavefront (queue-driven warps): pixels a warp keeps in flight, a multiple of 32 */
    /* options that are NOT the reference's sampling order (off by default; PLAN.md:23-27 lists both as open) */
    int32_t roulette;       /* RT_RENDER_ROULETTE: Russian roulette from the third bounce on */
    uint32_t chains;        /* rt_render_params.sample_chains: independent sample chains per pixel (<= 1: one stream per pixel, F4) */
    /* SAMPLE PARTS (scheduling only, results unchanged): a pixel's samples are handed out as n_parts work items, part k =
     * samples [part_end[k-1], part_end[k]) of the SAME stream, continued through the frame buffers exactly like
     * RT_RENDER_RESUME; all first parts are handed out before any second part. The frame's tail is then one LAST-part
     * chain (a few samples) instead of one whole-pixel chain, and the tail of the long first parts is filled with the
     * later parts of other pixels (profiles/README.md, "The tail of a frame"). */
    uint32_t order_region;  /* cost-ordered hand-out: side of the regions whose probed cost classes order the blocks (pixels, a power of two >= 8) */
    uint32_t order_probes;  /* ... and throw-away probe paths per 8x4 block */
    uint32_t n_parts;       /* 1 .. 3 */
    uint32_t part_end[3];   /* part_end[n_parts - 1] == spp */
};

#define RT_ROULETTE_MIN_DEPTH 3u
#define RT_CHAIN_SALT 0x9E3779B9u /* chain c of a pixel runs on the stream seed ^ c * RT_CHAIN_SALT (chain 0 = the reference stream) */

/* What follows rt_shade_segment for one path segment (src/render_megakernel.cpp:41-62, src/render_wavefront.cpp:272-280):
 * count the segment, cut the path at max_depth (survivors are black, F7), optionally play Russian roulette, and clamp a
 * finished sample in the wavefront formulation (F9). Returns true when the path ended; `res` is then the sample.
 * Russian roulette (optional, not in the reference): from the third bounce on the path survives with probability
 * q = clamp(max(att), 0.05, 1) — ONE extra draw — and its attenuation is divided by q; a killed path is black. The
 * caller quantises `att` afterwards (F6). */
RT_HD bool rt_after_segment(const RtFrameParams &p, bool done, uint32_t &depth, f3 &att, XorShift32 &rng, f3 &res) {
    depth++;
    if (!done && depth == p.max_depth) { /* :62, survivors are black (F7) */
        done = true;
        res = mk3(0.0f, 0.0f, 0.0f);
    }
    if (!done && p.roulette && depth >= RT_ROULETTE_MIN_DEPTH) {
        const float q = rt_min(rt_max(rt_max(rt_max(att.x, att.y), att.z), 0.05f), 1.0f);
        if (rng.next() > q) {
            done = true;
            res = mk3(0.0f, 0.0f, 0.0f);
        } else {
            att = att * rt_div(1.0f, q);
        }
    }
    if (done && p.clamp_samples) res = mk3(rt_clamp01(res.x), rt_clamp01(res.y), rt_clamp01(res.z)); /* :277 */
    return done;
}

/* samples chain `c` of `chains` contributes to a pixel of `spp` samples (the first spp % chains chains take one more) */
RT_HD uint32_t rt_chain_spp(uint32_t spp, uint32_t chains, uint32_t c) {
    return chains <= 1u ? spp : spp / chains + (c < spp % chains ? 1u : 0u);
}

/* image-tile sharding: tile t (row major, tile_size^2 pixels) belongs to rank t % world */
RT_HD bool rt_owns_pixel(const RtFrameParams &p, int x, int y) {
    if (p.world <= 1 || p.tile_size == 0) return true;
    const uint32_t tiles_x = ((uint32_t)p.cam.w + p.tile_size - 1) / p.tile_size;
    const uint32_t t = ((uint32_t)y / p.tile_size) * tiles_x + (uint32_t)x / p.tile_size;
    return (t % p.world) == p.rank;
}

/* All samples of one pixel, back to back on the pixel's own stream (F4):
 * src/render_megakernel.cpp:144-153 (seed, sample loop) around :20-63 (render_pixel). A finished
 * path regenerates the next sample inside the same loop, so a lane is never idle while its
 * pixel has samples left. Returns the linear sum; `rays` counts rtcIntersect1-equivalents. */
RT_HD f3 rt_megakernel_pixel(const RtScene &scene, const RtFrameParams &p, int x, int y, XorShift32 &rng,
                             unsigned long long &rays, f3 sum = mk3(0.0f, 0.0f, 0.0f), bool resume = false) {
    if (!resume) rng.a = rt_pixel_seed(p.wavefront_seed, x, y, p.cam.w, p.cam.h) ^ p.seed_salt;
    uint32_t s = 0, depth = 0;
    RtRayState r;
    r.org = r.dir = r.att = r.rad = mk3(0.0f, 0.0f, 0.0f);
    bool need_ray = true;
    for (;;) {
        if (need_ray) {
            if (s == p.spp) break;
            r = rt_camera_ray(p.cam, x, y, rng); /* 2 draws (F5) */
            depth = 0;
            need_ray = false;
            if (p.max_depth == 0) { /* the bounce loop never runs: black sample */
                s++;
                need_ray = true;
                continue;
            }
        }
        rays++;
        const RtHit h = rt_traverse(scene.bvh, r.org, r.dir, 0.0001f, INFINITY);
        f3 org = r.org, dir = r.dir, att = r.att, rad = r.rad, res = mk3(0.0f, 0.0f, 0.0f);
        bool done = rt_shade_segment(scene, h, rng, org, dir, att, rad, res);
        done = rt_after_segment(p, done, depth, att, rng, res);
        r.org = org; /* src/render_megakernel.cpp:41-55: re-quantise the ray state (F6) */
        r.dir = round_half3(dir);
        r.att = round_half3(att);
        r.rad = round_half3(rad);
        if (done) {
            sum = sum + res;
            s++;
            need_ray = true;
        }
    }
    return sum;
}

#endif /* RT_SHADE_H */
