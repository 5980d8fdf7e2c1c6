/*
 * rt_wavefront.h — per-item logic of the streaming wavefront renderer (K2-K5,
 * src/render_wavefront.cpp:62-355). The kernels in render.cu add the launch shape, the
 * device-side queue counters and the ballot / shared-memory compaction around these.
 */
#ifndef RT_WAVEFRONT_H
#define RT_WAVEFRONT_H

#include "rt_shade.h"

struct RtFrameOut {
    rt_float4 *accum; /* W*H: linear sum rgb, sample count in w */
    uint32_t *rgba8;  /* W*H */
    uint32_t *rng;    /* W*H: final xorshift state */
    uint32_t *gather; /* optional second destination of owned pixels' RGBA8 (peer memory, tile shards) */
    uint32_t *part_done; /* W*H, sample parts (RtFrameParams.n_parts > 1): parts of the pixel's chain finished in this frame */
};

/* per-pixel ray state (Buffers, src/render_wavefront.hpp:10-37): fp32 origin padded to 16 B
 * like sycl::float3, direction / attenuation / radiance as 3 x fp16 in 8 B like sycl::half3 */
struct RtWavefrontState {
    rt_float4 *org;
    rt_uint2 *dir, *att, *rad;
    rt_float4 *hit;     /* t, u, v, triangle slot bits */
    rt_uint2 *prog;     /* x = sample index, y = depth */
    uint32_t *rng;      /* XorShift32State per pixel (src/render_wavefront.hpp:52) */
    uint32_t *queue[2]; /* live pixel ids, double buffered */
    uint32_t *count[2];
    uint32_t *head;     /* extend kernel's fetch cursor into queue[cur] */
};

RT_HD rt_uint2 rt_pack_half3(f3 v) {
    return rt_mk_uint2((uint32_t)rt_float_to_half_bits(v.x) | ((uint32_t)rt_float_to_half_bits(v.y) << 16),
                       (uint32_t)rt_float_to_half_bits(v.z));
}
RT_HD f3 rt_unpack_half3(rt_uint2 u) {
    return mk3(rt_half_bits_to_float((uint16_t)(u.x & 0xffffu)), rt_half_bits_to_float((uint16_t)(u.x >> 16)),
               rt_half_bits_to_float((uint16_t)(u.y & 0xffffu)));
}

RT_HD void rt_wf_store_ray(const RtWavefrontState &w, uint32_t pix, f3 org, f3 dir, f3 att, f3 rad) {
    w.org[pix] = rt_mk_float4(org.x, org.y, org.z, 0.0f);
    w.dir[pix] = rt_pack_half3(dir); /* the fp16 store IS the reference's quantisation (F6) */
    w.att[pix] = rt_pack_half3(att);
    w.rad[pix] = rt_pack_half3(rad);
}

/* K2 + first K3 (src/render_wavefront.cpp:62-74,106-124): seed, zero, first camera ray.
 * Returns true when the pixel enters the queue. */
/* With sample chains (RtFrameParams.chains > 1, not the reference's order) the unit of work is a VIRTUAL pixel
 * vp = chain * n_pix + pix: every per-pixel array (ray state, rng, accumulation) has `chains` planes, chain c runs its
 * share of the samples on the stream seed ^ c * RT_CHAIN_SALT, and k_combine_chains sums the planes in chain order. */
struct RtVirtualPixel {
    uint32_t pix, chain, spp; /* image pixel, chain index, samples this chain contributes */
};
RT_HD RtVirtualPixel rt_virtual_pixel(const RtFrameParams &p, uint32_t vp) {
    RtVirtualPixel v;
    if (p.chains > 1u) {
        const uint32_t n_pix = (uint32_t)p.cam.w * (uint32_t)p.cam.h;
        v.chain = vp / n_pix;
        v.pix = vp - v.chain * n_pix;
    } else {
        v.chain = 0u;
        v.pix = vp;
    }
    v.spp = rt_chain_spp(p.spp, p.chains, v.chain);
    return v;
}

RT_HD bool rt_wf_generate_pixel(const RtFrameParams &p, const RtWavefrontState &w, const RtFrameOut &out,
                                uint32_t vp) {
    const RtVirtualPixel v = rt_virtual_pixel(p, vp);
    const int x = (int)(v.pix % (uint32_t)p.cam.w), y = (int)(v.pix / (uint32_t)p.cam.w);
    if (!p.resume) out.accum[vp] = rt_mk_float4(0.0f, 0.0f, 0.0f, 0.0f); /* the reference forgets combined_image (:56-57) */
    if (!rt_owns_pixel(p, x, y)) {
        w.rng[vp] = 0u;
        return false;
    }
    XorShift32 rng;
    rng.a = p.resume ? w.rng[vp] : (rt_pixel_seed(p.wavefront_seed, x, y, p.cam.w, p.cam.h) ^ p.seed_salt ^ (v.chain * RT_CHAIN_SALT));
    bool live = false;
    if (v.spp > 0 && p.max_depth > 0) {
        const RtRayState r = rt_camera_ray(p.cam, x, y, rng);
        rt_wf_store_ray(w, vp, r.org, r.dir, r.att, r.rad);
        w.prog[vp] = rt_mk_uint2(0u, 0u);
        live = true;
    } else {
        for (uint32_t s = 0; s < v.spp; s++) { /* depth 0: two draws per black sample */
            rng.next();
            rng.next();
        }
        rt_float4 a = p.resume ? out.accum[vp] : rt_mk_float4(0.0f, 0.0f, 0.0f, 0.0f);
        a.w += (float)v.spp;
        out.accum[vp] = a;
    }
    w.rng[vp] = rng.a;
    return live;
}

/* extend: traversal only (the rtcIntersect1 half of K4, src/render_wavefront.cpp:256-271) */
RT_HD void rt_wf_extend_pixel(const RtScene &scene, const RtWavefrontState &w, uint32_t pix) {
    const rt_float4 o = w.org[pix];
    const f3 d = rt_unpack_half3(w.dir[pix]);
    const RtHit h = rt_traverse(scene.bvh, mk3(o.x, o.y, o.z), d, 0.0001f, INFINITY);
    w.hit[pix] = rt_mk_float4(h.t, h.u, h.v, rt_u2f(h.tri));
}

/* shade + connect + regenerate (the rest of K4, K5 and the next sample's K3,
 * src/render_wavefront.cpp:244-296,340-355). Returns true when the pixel stays queued. */
RT_HD bool rt_wf_shade_pixel(const RtScene &scene, const RtFrameParams &p, const RtWavefrontState &w,
                             const RtFrameOut &out, uint32_t pix /* virtual pixel */) {
    const rt_float4 o4 = w.org[pix];
    const rt_float4 h4 = w.hit[pix];
    RtHit h;
    h.t = h4.x;
    h.u = h4.y;
    h.v = h4.z;
    h.tri = rt_f2u(h4.w);
    h.gid = 0;
    f3 org = mk3(o4.x, o4.y, o4.z), dir = rt_unpack_half3(w.dir[pix]);
    f3 att = rt_unpack_half3(w.att[pix]), rad = rt_unpack_half3(w.rad[pix]);
    rt_uint2 prog = w.prog[pix];
    XorShift32 rng;
    rng.a = w.rng[pix]; /* ScopedRng load, src/render_wavefront.cpp:15-32 */
    f3 res = mk3(0.0f, 0.0f, 0.0f);
    bool keep;
    bool done = rt_shade_segment(scene, h, rng, org, dir, att, rad, res);
    done = rt_after_segment(p, done, prog.y, att, rng, res); /* :277-280 */
    if (done) {
        rt_float4 a = out.accum[pix]; /* merge_samples, :340-355 */
        a.x += res.x;
        a.y += res.y;
        a.z += res.z;
        a.w += 1.0f;
        out.accum[pix] = a;
        prog.x++;
        prog.y = 0;
        const RtVirtualPixel v = rt_virtual_pixel(p, pix);
        keep = prog.x < v.spp;
        if (keep) { /* regenerate: this pixel's next sample */
            const int x = (int)(v.pix % (uint32_t)p.cam.w), y = (int)(v.pix / (uint32_t)p.cam.w);
            const RtRayState r = rt_camera_ray(p.cam, x, y, rng);
            rt_wf_store_ray(w, pix, r.org, r.dir, r.att, r.rad);
        }
    } else {
        rt_wf_store_ray(w, pix, org, dir, att, rad);
        keep = true;
    }
    w.prog[pix] = prog;
    w.rng[pix] = rng.a; /* ScopedRng store */
    return keep;
}

#endif /* RT_WAVEFRONT_H */
