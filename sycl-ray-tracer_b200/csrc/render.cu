/*
 * render.cu — the render kernels of the path for sm_100a.
 *
 *  k_intersect        batch closest hit (rtcIntersect1, src/trace_ray.hpp:18-22)
 *  k_megakernel       persistent-thread megakernel (K1, src/render_megakernel.cpp:116-168):
 *                     warps fetch 8x4 pixel tiles through a global atomic; every lane owns one
 *                     pixel and its xorshift stream, runs that pixel's samples back to back
 *                     (path regeneration inside the loop keeps lanes busy), accumulates in
 *                     registers and resolves the pixel.
 *  wavefront          streaming formulation of K2-K6 (src/render_wavefront.cpp:62-391):
 *    k_wf_generate    seed rng, zero accumulation, first camera ray, initial id queue
 *    k_wf_extend      traversal only, over the queue of live pixels -> hit records
 *    k_wf_shade       material + rng + terminate/accumulate ("connect") + regeneration of the
 *                     pixel's next sample, survivors compacted into the next id queue with
 *                     warp ballots, a shared-memory block scan and one global atomic per block
 *  k_resolve          mean, sqrt gamma, F10 byte rule (K6/K7)
 *
 *  Per-pixel ray state stays in SoA arrays indexed by pixel id (fp32 origin, fp16 direction /
 *  attenuation / radiance exactly like the reference's Buffers, src/render_wavefront.hpp:10-37);
 *  only 4-byte pixel ids move through the queues. Because a pixel has at most one ray in flight
 *  and consumes its stream sequentially (F4), both formulations produce the reference's
 *  per-pixel results; they differ only in seed mapping (F3) and per-sample clamp (F9).
 */
#ifdef RT_GPU_COUNTERS /* development build only (tools/): node visits / triangle tests of the render kernels */
#include <cstdint>
__device__ unsigned long long g_rt_counters[2];
#define RT_COUNTERS 1
__host__ __device__ inline void rt_count(int i) {
#ifdef __CUDA_ARCH__
    atomicAdd(&g_rt_counters[i], 1ull);
#else
    (void)i;
#endif
}
#define RT_COUNT_NODE() rt_count(0)
#define RT_COUNT_TRI() rt_count(1)
#endif
#include "rt_render.h"
#include "rt_blocks.h"

#include <cub/device/device_radix_sort.cuh>

namespace {

constexpr int kMegaBlock = 128;
#ifndef RT_WF_BLOCK
#define RT_WF_BLOCK 256
#endif
constexpr int kWfBlock = RT_WF_BLOCK;

/* ------------------------------------------------------------------------------ intersect */
__global__ void k_intersect(RtScene scene, const RtInstance *inst, uint64_t n, const float *org,
                            const float *dir, float tnear, float tfar, int32_t *o_inst, int32_t *o_prim,
                            float *o_u, float *o_v, float *o_t) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const f3 o = mk3(org[i * 3], org[i * 3 + 1], org[i * 3 + 2]);
    const f3 d = mk3(dir[i * 3], dir[i * 3 + 1], dir[i * 3 + 2]);
    const RtHit h = rt_traverse(scene.bvh, o, d, tnear, tfar);
    if (h.tri == RT_MISS) {
        o_inst[i] = -1;
        o_prim[i] = -1;
        o_u[i] = 0.0f;
        o_v[i] = 0.0f;
        o_t[i] = tfar;
    } else {
        const uint32_t ii = rt_f2u(rt_ldg(scene.shade + (size_t)h.tri * 4 + 3).w);
        o_inst[i] = (int32_t)ii;
        o_prim[i] = (int32_t)(h.gid - inst[ii].first_tri);
        o_u[i] = h.u;
        o_v[i] = h.v;
        o_t[i] = h.t;
    }
}

/* ------------------------------------------------------------------------------ block order */
/* Cost probe for the block order. A frame ends when its slowest lane finishes its last pixel, and a pixel
 * is a strictly sequential chain (spp samples on one xorshift stream, F4), so expensive pixels handed out
 * last stretch the tail of the kernel by up to a whole pixel time. The probe traces one throw-away path
 * per block (own seed: the pixels' streams are untouched, nothing is accumulated or counted) and adds the
 * traversal steps it took to the block's 64x64-pixel region; blocks are then handed out by decreasing cost
 * class of their REGION (128 probes each: smooth, and neighbouring blocks stay together), image order
 * within a class. Scheduling only: results are bit-identical for any order. */
constexpr uint32_t kRegionMin = 8; /* the cost table is sized for the smallest region */
__device__ __forceinline__ uint32_t region_of(const RtFrameParams &p, uint32_t x0, uint32_t y0) {
    const uint32_t rg = p.order_region;
    return (y0 / rg) * (((uint32_t)p.cam.w + rg - 1u) / rg) + x0 / rg;
}
__global__ void __launch_bounds__(128) k_block_cost(RtScene scene, RtFrameParams p, RtBlockGeom g, uint32_t *region_cost, uint32_t *vals) {
    const uint32_t blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= g.n_blocks) return;
    uint32_t x0, y0;
    rt_block_origin(p, g, blk, x0, y0);
    uint32_t cost = 0;
    if (x0 < (uint32_t)p.cam.w && y0 < (uint32_t)p.cam.h) {
        const int x = min((int)x0 + 3, p.cam.w - 1), y = min((int)y0 + 1, p.cam.h - 1);
        XorShift32 rng;
        rng.a = (((uint32_t)x * 0x9E3779B1u) ^ ((uint32_t)y * 0x85EBCA77u)) | 1u;
        for (uint32_t probe = 0; probe < p.order_probes; probe++) {
        const RtRayState r = rt_camera_ray(p.cam, x, y, rng);
        f3 org = r.org, dir = r.dir, att = r.att, rad = r.rad, res;
        for (uint32_t depth = 0; depth < p.max_depth; depth++) {
            RtTravState tv;
            RtTravStacks ks;
            rt_trav_init(tv, org, dir, 0.0001f, INFINITY);
            cost += 2; /* ray set-up + shading, in units of one traversal step */
            const RtRayTri rtri = rt_trav_ray_tri(tv);
            while (rt_trav_has_node(tv)) {
                rt_trav_node_step(scene.bvh, tv, ks);
                cost++;
                while (rt_trav_has_tri(tv)) {
                    rt_trav_tri_step(scene.bvh, tv, ks, rtri);
                    cost++;
                }
            }
            if (rt_shade_segment(scene, rt_trav_hit_noid(tv), rng, org, dir, att, rad, res)) break;
        }
        }
        atomicAdd(region_cost + region_of(p, x0, y0), cost);
    }
    vals[blk] = (y0 << 16) | x0;
}

__global__ void __launch_bounds__(128) k_block_key(RtFrameParams p, RtBlockGeom g, const uint32_t *region_cost, const uint32_t *vals,
                                                   uint32_t *keys) {
    const uint32_t blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= g.n_blocks) return;
    const uint32_t x0 = vals[blk] & 0xffffu, y0 = vals[blk] >> 16;
    /* cost class: log2 with two mantissa bits (8 bits, one radix pass); 0 = outside the image */
    uint32_t key = 0;
    if (x0 < (uint32_t)p.cam.w && y0 < (uint32_t)p.cam.h) {
        const uint32_t cost = region_cost[region_of(p, x0, y0)] | 1u;
        const int e = 31 - __clz(cost);
        key = ((uint32_t)(e << 2) | ((e >= 2 ? cost >> (e - 2) : cost << (2 - e)) & 3u)) + 1u;
    }
    keys[blk] = key;
}

/* ------------------------------------------------------------------------------ megakernel */
/* Persistent lanes. Every lane owns one pixel at a time and walks that pixel's samples back to
 * back on the pixel's own xorshift stream (F4). The warp alternates between two phases:
 *   regenerate : lanes whose traversal finished shade their hit (rt_shade_segment), start the next
 *                bounce / the pixel's next sample, or write the finished pixel and fetch a new one
 *                (one warp-aggregated atomicAdd on the global pixel counter);
 *   traverse   : all lanes with a live ray advance one wide node per iteration; the warp leaves the
 *                loop as soon as kRefill lanes have finished, so finished lanes never idle for long
 *                (terminated-ray replacement, after Aila & Laine, HPG 2009).
 * The per-pixel arithmetic is the same as rt_megakernel_pixel (rt_shade.h), which the host
 * emulation and the oracle comparison exercise; only the scheduling differs. */
enum { kNeedPixel = 0, kNeedRay = 1, kStart = 2, kTraversing = 3, kHitPending = 4, kExhausted = 5, kWaitPart = 6 };

/* Traverse phase shared by the megakernel and the wavefront extend kernel (warp-uniform control
 * flow): node steps for every lane with node work until `refill` lanes have run out of nodes, or
 * a lane's triangle stack is full; then all pending
 * triangles are tested together. Lanes left with neither nodes nor triangles become kHitPending. */
/* Traversal stacks of the persistent kernels. The triangle-group stack and the first RT_SMEM_NODE entries
 * of the node-group stack can live in shared memory ([entry][thread]: conflict-free when the lanes of a
 * warp are at the same depth); deeper node entries stay in local memory. RT_SMEM_TRI / RT_SMEM_NODE = 0
 * select plain local-memory stacks — the default: measured on B200, every shared-memory split (tri only,
 * tri + 4, tri + 8, 8 node entries only) is 3-4 % SLOWER on C2/C3/C4 in both renderers, because the shared
 * memory comes out of the L1 that caches the nodes and the (small, hot) local stacks. */
#ifndef RT_SMEM_TRI
#define RT_SMEM_TRI 0
#endif
#ifndef RT_SMEM_NODE
#define RT_SMEM_NODE 0
#endif
#define RT_SMEM_ENTRIES ((RT_SMEM_TRI ? RT_TSTACK_SIZE : 0) + RT_SMEM_NODE)
/* octant permutation table of the child-hit flags (rt_trav_node_step): s_perm_tbl[x * 8 + k] =
 * rt_xor_perm8(x, k). One LDS at a compile-time shared address instead of a dozen ALU instructions per
 * node visit; 2 KB per CTA of every kernel that traverses with SmemStacks. Indexed [x][k]: most lanes
 * carry a small x (0-2 inner children hit) and differ in k, which the [k][x] layout put on ONE bank
 * (ncu: 7.2 wavefronts per LDS); here lanes with equal x share a word or sit on adjacent banks.
 * RT_PERM_MODE 2 = no table, the three conditional swaps in registers. */
#ifndef RT_PERM_MODE
#define RT_PERM_MODE 1
#endif
__shared__ __align__(16) uint8_t s_perm_tbl[8 * 256];
__device__ __forceinline__ void fill_perm_table() { /* the caller synchronises the CTA afterwards */
#if RT_PERM_MODE == 0
    for (uint32_t i = threadIdx.x; i < 8u * 256u; i += blockDim.x) s_perm_tbl[i] = (uint8_t)rt_xor_perm8(i & 255u, i >> 8);
#elif RT_PERM_MODE == 1
    for (uint32_t i = threadIdx.x; i < 8u * 256u; i += blockDim.x) s_perm_tbl[i] = (uint8_t)rt_xor_perm8(i >> 3, i & 7u);
#endif
}
__device__ __forceinline__ uint32_t perm_lookup(uint32_t k, uint32_t x) {
#if RT_PERM_MODE == 0
    return s_perm_tbl[k * 256u + x];
#elif RT_PERM_MODE == 1
    return s_perm_tbl[x * 8u + k];
#else
    return rt_xor_perm8(x, k);
#endif
}

template <int BLOCK>
struct SmemStacks {
    uint64_t node[RT_STACK_SIZE - RT_SMEM_NODE];
#if !RT_SMEM_TRI
    uint64_t tri[RT_TSTACK_SIZE];
#endif
    uint64_t *sm; /* this thread's column of the CTA's shared array */
    __device__ __forceinline__ uint32_t perm(uint32_t k, uint32_t x) const { return perm_lookup(k, x); }
    __device__ __forceinline__ uint64_t node_get(int i) const {
#if RT_SMEM_NODE
        if (i < RT_SMEM_NODE) return sm[((RT_SMEM_TRI ? RT_TSTACK_SIZE : 0) + i) * BLOCK];
        return node[i - RT_SMEM_NODE];
#else
        return node[i];
#endif
    }
    __device__ __forceinline__ void node_put(int i, uint64_t v) {
#if RT_SMEM_NODE
        if (i < RT_SMEM_NODE) sm[((RT_SMEM_TRI ? RT_TSTACK_SIZE : 0) + i) * BLOCK] = v;
        else node[i - RT_SMEM_NODE] = v;
#else
        node[i] = v;
#endif
    }
#if RT_SMEM_TRI
    __device__ __forceinline__ uint64_t tri_get(int i) const { return sm[i * BLOCK]; }
    __device__ __forceinline__ void tri_put(int i, uint64_t v) { sm[i * BLOCK] = v; }
#else
    __device__ __forceinline__ uint64_t tri_get(int i) const { return tri[i]; }
    __device__ __forceinline__ void tri_put(int i, uint64_t v) { tri[i] = v; }
#endif
};

/* the megakernel extracts ALL plane bytes with IDP.4A (measured against x, y only: C3 3405 -> 3507, C2 6830 -> 6909,
 * C4 4042 -> 4168 Mrays/s; the queue-driven wavefront kernel is 1 % faster with x, y only — its shading and queue code
 * keeps the FMA pipe busier) */
#ifndef RT_MEGA_IDP_MASK
#define RT_MEGA_IDP_MASK 7
#endif
/* UNROLL = node steps per exit vote. Two votes, a population count and the branches are 17 of the ~285 instructions of a
 * node-loop round; with several steps per round a lane that runs dry in one sits out the rest. The steps run as a ROLLED loop
 * over one copy of the node test (RT_NODE_STEPS_ROLLED): the megakernel is at the edge of the instruction cache, three
 * unrolled copies (3248 instead of 2752 instructions) are 11 % slower than one step, the rolled loop is not. Measured at 72
 * registers, C3 / C2 / C4 Mrays/s: 1 step 3680 / 7501 / 4255, 2 unrolled 3739 / 7431 / 4345, 3 rolled 3763 / 7746 / 4381,
 * 4 rolled 3770 / 7721 (profiles/r02_ab_regs_unroll.log, r02_ab_rolled_steps.log); neutral for the queue-driven wavefront
 * kernel (r02_ab_wf_rolled.log), which keeps one step. */
#ifndef RT_MEGA_NODE_UNROLL
#define RT_MEGA_NODE_UNROLL 3
#endif
#ifndef RT_MEGA_NODE_NEAR /* within this many dry lanes of the refill threshold the vote is taken after every step again (0 = never) */
#define RT_MEGA_NODE_NEAR 0
#endif
#ifndef RT_WF_NODE_UNROLL
#define RT_WF_NODE_UNROLL 1
#endif
#ifndef RT_NODE_STEPS_ROLLED
#define RT_NODE_STEPS_ROLLED 1
#endif
template <int MASK = RT_BYTE_IDP_MASK, int UNROLL = 1, int NEAR = 0, class Stacks>
__device__ __forceinline__ void traverse_phase(const RtBvh &bvh, RtTravState &tv, Stacks &ks, int &mode, int refill_carry /* tune_refill | tune_carry << 8 */) {
    const unsigned full = 0xffffffffu;
    const int refill = refill_carry & 0xff, carry_cfg = refill_carry >> 8;
    const bool trav = mode == kTraversing;
    const unsigned m_trav = __ballot_sync(full, trav); /* does not change inside the node loop */
    /* lanes that carried triangles over from the previous drain (tune_carry) have no nodes to begin with: they neither count as
     * lanes that ran dry nor as lanes that could */
    const unsigned m_run = __ballot_sync(full, trav && rt_trav_has_node(tv));
    /* the refill threshold scales with the lanes that still have work: in the drain of a frame (most lanes
     * exhausted) a finished lane must not wait for every other ray of its warp */
    const int thr = max(1, (__popc(m_run) * refill + 31) >> 5);
    int steps = UNROLL;
    for (;;) {
#if RT_NODE_STEPS_ROLLED /* one copy of the node test, run `steps` times */
#pragma unroll 1
        for (int u = 0; u < steps; u++)
            if (trav && rt_trav_has_node(tv) && !rt_trav_tri_full(tv)) rt_trav_node_step<MASK>(bvh, tv, ks);
#else
        if (trav && rt_trav_has_node(tv)) rt_trav_node_step<MASK>(bvh, tv, ks);
#pragma unroll
        for (int u = 1; u < UNROLL; u++)
            if (trav && rt_trav_has_node(tv) && !rt_trav_tri_full(tv)) rt_trav_node_step<MASK>(bvh, tv, ks);
#endif
        const bool node = trav && rt_trav_has_node(tv);
        const unsigned m_node = __ballot_sync(full, node);
        /* leave when no lane has nodes left, `refill` lanes have run dry, or a lane's triangle stack
         * is full (draining on a pending-triangle count instead was measured and never paid off) */
        const bool out = trav && rt_trav_tri_full(tv);
        const int dry = __popc(m_run & ~m_node);
        if (!m_node || dry >= thr || __any_sync(full, out)) break;
        if (UNROLL > 1 && NEAR > 0) steps = thr - dry <= NEAR ? 1 : UNROLL; /* close to the refill threshold: vote after every step */
    }
    /* drain: lanes that are out of nodes (or nearly out of triangle-stack room) must finish their
     * triangles; every other lane with triangles pending joins in, and carries what is left to the
     * next drain. The nine-float shear form lives in registers only here (rt_traverse.h, RtTravState).
     * tune_carry > 0: the drain's length is set by the lane with the MOST triangles pending while the others idle, so it stops
     * once no more than `carry` lanes (scaled like the refill threshold) still have to finish; those lanes stay in the traversing
     * state without nodes, sit out the next node loop and finish in the next drain, where other lanes' new triangles fill the
     * warp. Every drain runs at least one round, so they always make progress; lanes short of stack room are never carried. */
    {
        const int carry = (__popc(m_trav) * carry_cfg + 31) >> 5;
        bool tri = mode == kTraversing && rt_trav_has_tri(tv);
        bool must_room = tri && tv.tsp >= RT_TSTACK_SIZE - 2, must_dry = tri && !rt_trav_has_node(tv);
        if (__any_sync(full, must_room || must_dry)) {
            const RtRayTri rtri = rt_trav_ray_tri(tv);
            bool again;
            do {
                if (tri) rt_trav_tri_step(bvh, tv, ks, rtri);
                tri = mode == kTraversing && rt_trav_has_tri(tv);
                must_room = tri && tv.tsp >= RT_TSTACK_SIZE - 2;
                must_dry = tri && !rt_trav_has_node(tv);
                again = __any_sync(full, must_room) || __popc(__ballot_sync(full, must_dry)) > carry; /* carry 0: while any lane must */
            } while (again);
        }
    }
    if (mode == kTraversing && !rt_trav_has_node(tv) && !rt_trav_has_tri(tv)) mode = kHitPending;
}

/* resident CTAs per SM the megakernel is compiled for: 7 x 128 threads = 72 registers, no spills; 8 (64 registers) spills
 * ~20 words of which two are re-loaded in every node visit, 6 (80 registers) has too few warps. Measured C3 / C2 / C4:
 * 8: 3605 / 7326 / 4167, 7: 3685 / 7509 / 4258, 6: 3503 / 7292 / 4059 Mrays/s (profiles/r02_ab_stacktop_regs.log,
 * r02_ab_regs_unroll.log). The queue-driven wavefront kernel stays at 4 x 256 threads / 64 registers (8 bytes of spill);
 * 128 x 7, 128 x 6 and 256 x 3 are 5-8 % slower there. */
#ifndef RT_MEGA_MIN_BLOCKS
#define RT_MEGA_MIN_BLOCKS 7
#endif
/* Nine fp16 values in five registers. Between two bounces direction, attenuation and radiance ARE fp16 values
 * (RayData, F6), so holding them packed across the traversal loses nothing: cvt.rn.f16x2.f32 rounds each half
 * exactly like the scalar conversion of round_half3. */
struct HalfRay {
    uint32_t dxy, dz_ax, ayz, rxy, rz;
};
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ float h_lo(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v & 0xffffu))); }
__device__ __forceinline__ float h_hi(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v >> 16))); }
__device__ __forceinline__ HalfRay pack_ray(f3 dir, f3 att, f3 rad) {
    HalfRay h;
    h.dxy = pack_h2(dir.x, dir.y);
    h.dz_ax = pack_h2(dir.z, att.x);
    h.ayz = pack_h2(att.y, att.z);
    h.rxy = pack_h2(rad.x, rad.y);
    h.rz = pack_h2(rad.z, 0.0f);
    return h;
}
__device__ __forceinline__ f3 ray_dir(const HalfRay &h) { return mk3(h_lo(h.dxy), h_hi(h.dxy), h_lo(h.dz_ax)); }
__device__ __forceinline__ f3 ray_att(const HalfRay &h) { return mk3(h_hi(h.dz_ax), h_lo(h.ayz), h_hi(h.ayz)); }
__device__ __forceinline__ f3 ray_rad(const HalfRay &h) { return mk3(h_lo(h.rxy), h_hi(h.rxy), h_lo(h.rz)); }

/* What a lane carries across the traversal loop is kept to 13 registers next to the traversal state (18): pixel
 * coordinates in one word, sample and bounce counters, the stream, the running sum, the fp16 ray state packed
 * (HalfRay), the ray counter in 32 bits; the ray origin lives in the traversal state only. Sample chains and the
 * resume count are rare options: chains are a template parameter (no registers in the default kernel) and the count
 * of earlier frames is re-read from the accumulation buffer when the pixel finishes. */
template <bool CHAINS>
__global__ void __launch_bounds__(kMegaBlock, RT_MEGA_MIN_BLOCKS) k_megakernel(RtScene scene, RtFrameParams p, RtFrameOut out,
                                                           uint32_t *work_counter, unsigned long long *ray_counter,
                                                           const uint32_t *__restrict__ order, const RtBlockGeom g /* = rt_block_geom(p), from the host: lives in the constant bank, not in registers */) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    /* uniform quantities of the hand-out are macros over kernel parameters on purpose: named locals stayed in registers across
     * the node loop and brought its spill re-loads back */
#define N_CHAINS (CHAINS ? p.chains : 1u)
#define N_PARTS (CHAINS ? 1u : p.n_parts)                    /* sample parts (RtFrameParams): all first parts, then all second parts, ... */
#define N_SLOTS (g.n_blocks * N_CHAINS * 32u)                /* pixel slots of one part, 32 per (8x4 block, sample chain) */
#define N_PIX ((size_t)p.cam.w * (size_t)p.cam.h)
    uint32_t rays = 0; /* per lane: < 2^32 for any frame that finishes */
    int mode = kNeedPixel;
    uint32_t xy = 0;
    uint32_t s = 0, depth = 0, chain = 0, spp_c = p.spp; /* sample chain of the lane's pixel, samples it contributes (CHAINS only) */
    /* sample parts cost no register: the part index rides in the top two bits of the bounce counter, and a lane that waits
     * for the previous part of its pixel (mode kWaitPart) keeps the work index in xy */
    XorShift32 rng;
    rng.a = 0;
    f3 sum = mk3(0.0f, 0.0f, 0.0f);
    HalfRay hr = pack_ray(sum, sum, sum);
    RtTravState tv;
    SmemStacks<kMegaBlock> ks;
#if RT_SMEM_TRI || RT_SMEM_NODE
    __shared__ uint64_t s_stacks[RT_SMEM_ENTRIES * kMegaBlock];
    ks.sm = s_stacks + threadIdx.x;
#endif
    fill_perm_table();
    __syncthreads();
    tv.sp = 0;
    tv.tsp = 0;
    tv.ng_y = 0;
    tv.org = sum;

    for (;;) {
        /* ---------------- regenerate ---------------- */
        if (mode == kHitPending) {
            f3 org = tv.org, dir = ray_dir(hr), att = ray_att(hr), rad = ray_rad(hr), res = mk3(0.0f, 0.0f, 0.0f);
            bool done = rt_shade_segment(scene, rt_trav_hit_noid(tv), rng, org, dir, att, rad, res);
            {
                uint32_t d = depth & 0x3fffffffu;
                done = rt_after_segment(p, done, d, att, rng, res);
                depth = (depth & 0xc0000000u) | d;
            }
            tv.org = org;                 /* src/render_megakernel.cpp:41-55: re-quantise the ray state (F6) */
            hr = pack_ray(dir, att, rad); /* fp32 -> fp16, round to nearest even */
            if (done) {
                sum = sum + res;
                s++;
                mode = kNeedRay;
            } else {
                mode = kStart;
            }
        }
        bool may_retry = mode == kWaitPart; /* ONE look per round: the previous part may be running in a lane of this very warp */
        for (;;) { /* warp-uniform: runs until no lane is waiting for a pixel */
            if (mode == kNeedRay) {
                const int x = (int)(xy & 0xffffu), y = (int)(xy >> 16);
                const uint32_t part = CHAINS ? 0u : depth >> 30;
                const uint32_t spp_l = CHAINS ? spp_c : p.part_end[part];
                while (p.max_depth == 0 && s < spp_l) { /* the bounce loop never runs: black sample */
                    rng.next();
                    rng.next();
                    s++;
                }
                if (s == spp_l) { /* pixel finished: :154-158 mean, gamma, image write */
                    const size_t pix = (size_t)y * (size_t)p.cam.w + (size_t)x, vp = (CHAINS ? (size_t)chain * N_PIX : 0) + pix;
                    const uint32_t first = part ? p.part_end[part - 1u] : 0u; /* this part added the samples [first, spp_l) */
                    const float base_count = (p.resume || part) ? __ldcg(&out.accum[vp]).w : 0.0f; /* samples accumulated by earlier frames / parts */
                    const float count = base_count + (float)(spp_l - first);
                    out.accum[vp] = make_float4(sum.x, sum.y, sum.z, count);
                    out.rng[vp] = rng.a;
                    if (!CHAINS && part + 1u < N_PARTS) { /* hand the pixel over to its next part: state first, then the flag */
                        __threadfence();
                        *(volatile uint32_t *)&out.part_done[pix] = part + 1u;
                    } else if (!CHAINS) { /* with sample chains k_combine_chains sums the planes and writes the image */
                        const uint32_t px = rt_resolve_pixel(sum.x, sum.y, sum.z, count);
                        out.rgba8[pix] = px;
                        if (out.gather) out.gather[pix] = px; /* tile shards: straight into the destination rank's image */
                    }
                    mode = kNeedPixel;
                } else {
                    const RtRayState r = rt_camera_ray(p.cam, x, y, rng); /* 2 draws (F5) */
                    tv.org = r.org;
                    hr = pack_ray(r.dir, r.att, r.rad);
                    depth &= 0xc0000000u;
                    mode = kStart;
                }
            }
            /* a lane whose later part found the previous part unfinished looks again (it does not take a new slot) */
            const bool retry = mode == kWaitPart && may_retry;
            may_retry = false;
            const unsigned need = __ballot_sync(full, mode == kNeedPixel);
            if (!need && !__any_sync(full, retry)) break;
            uint32_t base = 0;
            if (need) {
                const int leader = __ffs(need) - 1;
                if (lane == leader) base = atomicAdd(work_counter, (uint32_t)__popc(need));
                base = __shfl_sync(full, base, leader);
            }
            if (mode == kNeedPixel || retry) {
                const uint32_t idx = retry ? xy : base + (uint32_t)__popc(need & ((1u << lane) - 1u));
                mode = kNeedPixel;
                if (idx >= N_SLOTS * N_PARTS) {
                    mode = kExhausted;
                } else {
                    const uint32_t part = (CHAINS || N_PARTS == 1u) ? 0u : idx / N_SLOTS, slot = idx - part * N_SLOTS;
                    const uint32_t in = slot & 31u, item = slot >> 5, blk = CHAINS ? item / N_CHAINS : item; /* (block, chain) pairs */
                    uint32_t x0, y0;
                    if (order) { /* blocks sorted by decreasing cost class (k_block_cost) */
                        const uint32_t e = __ldg(order + blk);
                        x0 = e & 0xffffu;
                        y0 = e >> 16;
                    } else {
                        rt_block_origin(p, g, blk, x0, y0);
                    }
                    const int x = (int)(x0 + (in & 7u)), y = (int)(y0 + (in >> 3));
                    if (x < p.cam.w && y < p.cam.h && rt_owns_pixel(p, x, y)) {
                        const size_t pix = (size_t)y * (size_t)p.cam.w + (size_t)x;
                        if (part && *(volatile const uint32_t *)&out.part_done[pix] != part) {
                            xy = idx; /* the previous part of this pixel is still running (in another lane): come back after the next traversal phase */
                            mode = kWaitPart;
                        } else {
                            xy = (uint32_t)x | ((uint32_t)y << 16);
                            if (CHAINS) {
                                chain = item - blk * N_CHAINS;
                                spp_c = rt_chain_spp(p.spp, p.chains, chain);
                            }
                            depth = part << 30;
                            if (p.resume || part) { /* carry on where the previous frame / part stopped */
                                if (part) __threadfence(); /* the flag was read above: order the state loads after it */
                                const size_t vp = (CHAINS ? (size_t)chain * N_PIX : 0) + pix;
                                const float4 a = __ldcg(&out.accum[vp]);
                                rng.a = __ldcg(&out.rng[vp]);
                                sum = mk3(a.x, a.y, a.z);
                            } else {
                                rng.a = rt_pixel_seed(p.wavefront_seed, x, y, p.cam.w, p.cam.h) ^ p.seed_salt ^ (CHAINS ? chain * RT_CHAIN_SALT : 0u);
                                sum = mk3(0.0f, 0.0f, 0.0f);
                            }
                            s = part ? p.part_end[part - 1u] : 0u;
                            mode = kNeedRay;
                        }
                    } /* else: padding / another rank's pixel, fetch again */
                }
            }
        }
        if (mode == kStart) {
            rt_trav_init(tv, tv.org, ray_dir(hr), 0.0001f, INFINITY);
            rays++;
            mode = kTraversing;
        }
        /* ---------------- traverse ---------------- */
        const unsigned act0 = __ballot_sync(full, mode == kTraversing);
        if (!act0) {
            if (!__any_sync(full, mode == kWaitPart)) break; /* every lane is exhausted */
            __nanosleep(500); /* only lanes waiting for a part that another warp is finishing: look again */
            continue;
        }
        traverse_phase<RT_MEGA_IDP_MASK, RT_MEGA_NODE_UNROLL, RT_MEGA_NODE_NEAR>(scene.bvh, tv, ks, mode, p.tune_refill);
    }
    unsigned long long total = rays;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(full, total, o);
    if (lane == 0 && total) atomicAdd(ray_counter, total);
#undef N_CHAINS
#undef N_PARTS
#undef N_SLOTS
#undef N_PIX
}

/* ------------------------------------------------------------------------------ megakernel, K contexts per lane */
/* Round-2 scheduling of K1. A lane owns K pixels ("contexts"); everything a pixel needs between two
 * rays is PARKED in local memory (7 x 16 bytes per context) instead of being held in registers across
 * the traversal loop, so the node loop runs on the traversal state alone (the round-1 kernel re-loaded
 * spilled ray constants on every node visit). One context at a time is under traversal; the others
 * are READY (a ray waiting to be traced, its direction-only set-up already computed) or HIT (a
 * closest hit waiting to be shaded). The warp alternates between
 *   shade   : every lane with a HIT (or brand-new) context shades ONE of them, regenerates the next
 *             ray / sample / pixel exactly like k_megakernel, runs the six divisions of the ray set-up
 *             and parks the context as READY. Runs when `tune_shade` lanes have something to shade, or
 *             `tune_idle` lanes can neither traverse nor switch — so shading runs with most lanes active
 *             instead of the 8-13 of the one-context scheme;
 *   switch  : lanes whose traversal finished park the hit (16 bytes) and load a READY context
 *             (3 x 16 bytes + a few selects): cheap, so the refill threshold of the node loop is low;
 *   traverse: traverse_phase(), as before.
 * Per-pixel arithmetic and the order of a pixel's draws are those of rt_megakernel_pixel (a pixel still
 * has one ray in flight and its own stream, F4); only the interleaving of pixels changes, so results are
 * bit-identical to k_megakernel and to the oracle.
 *   context: q0 = {org, rng}  q1 = {dir, s}  q2 = {att, depth}  q3 = {rad, base_count}  q4 = {sum, x}
 *            q5 = READY: {nSx, nSy, Sz, code}, HIT: {t, u, v, tri}   q6 = {rcp, y} */
constexpr int kCtxQ = 7;
enum { cNew = 0, cReady = 1, cTrav = 2, cHit = 3, cDone = 4 };

__device__ __forceinline__ int opaque_index(int i) { /* keeps the context arrays in local memory for K = 1 too */
    asm volatile("" : "+r"(i));
    return i;
}

template <int K>
__global__ void __launch_bounds__(kMegaBlock, RT_MEGA_MIN_BLOCKS) k_megakernel_ctx(RtScene scene, RtFrameParams p, RtFrameOut out,
                                                                                  uint32_t *work_counter, unsigned long long *ray_counter,
                                                                                  const uint32_t *__restrict__ order) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const RtBlockGeom g = rt_block_geom(p);
    const uint32_t n_work = g.n_blocks * 32u; /* pixel slots, 32 per 8x4 block */
    float4 cx[K * kCtxQ];
    uint32_t st = 0; /* 3 bits per context, all cNew */
    int cur = -1;    /* context under traversal */
    unsigned long long rays = 0;
    RtTravState tv;
    SmemStacks<kMegaBlock> ks;
#if RT_SMEM_TRI || RT_SMEM_NODE
    __shared__ uint64_t s_stacks[RT_SMEM_ENTRIES * kMegaBlock];
    ks.sm = s_stacks + threadIdx.x;
#endif
    fill_perm_table();
    __syncthreads();
    tv.sp = 0;
    tv.tsp = 0;
    tv.ng_y = 0;

    for (;;) {
        /* ---------------- shade passes ---------------- */
        for (;;) {
            int c = -1;
            bool ready = false, alive = false;
#pragma unroll
            for (int i = K - 1; i >= 0; i--) {
                const uint32_t s_i = (st >> (3 * i)) & 7u;
                if (s_i == cNew || s_i == cHit) c = i;
                ready |= s_i == cReady;
                alive |= s_i != cDone;
            }
            const unsigned m_pend = __ballot_sync(full, c >= 0);
            if (!m_pend) break;
            const unsigned m_go = __ballot_sync(full, cur >= 0 || ready);
            const unsigned m_alive = __ballot_sync(full, alive);
            /* thresholds scale with the lanes that still have work (the drain of a frame) */
            const int n_alive = __popc(m_alive);
            const int t_shade = max(1, (n_alive * p.tune_shade + 31) >> 5), t_idle = max(1, (n_alive * p.tune_idle + 31) >> 5);
            if (m_go && __popc(m_pend) < t_shade && __popc(m_pend & ~m_go) < t_idle) break;

            /* ---- one pass: lanes with c >= 0 shade / regenerate context c ---- */
            int mode = kExhausted; /* not taking part */
            int x = 0, y = 0;
            uint32_t s = 0, depth = 0;
            XorShift32 rng;
            rng.a = 0;
            f3 sum = mk3(0.0f, 0.0f, 0.0f);
            float base_count = 0.0f;
            RtRayState r;
            r.org = r.dir = r.att = r.rad = sum;
            float4 *q = cx + opaque_index((c < 0 ? 0 : c) * kCtxQ);
            if (c >= 0) {
                if (((st >> (3 * c)) & 7u) == cNew) {
                    mode = kNeedPixel;
                } else {
                    const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4], q5 = q[5];
                    rng.a = __float_as_uint(q0.w);
                    s = __float_as_uint(q1.w);
                    depth = __float_as_uint(q2.w);
                    base_count = q3.w;
                    sum = mk3(q4.x, q4.y, q4.z);
                    x = (int)__float_as_uint(q4.w);
                    y = (int)__float_as_uint(q[6].w);
                    RtHit h;
                    h.t = q5.x;
                    h.u = q5.y;
                    h.v = q5.z;
                    h.tri = __float_as_uint(q5.w);
                    h.gid = 0;
                    f3 org = mk3(q0.x, q0.y, q0.z), dir = mk3(q1.x, q1.y, q1.z), att = mk3(q2.x, q2.y, q2.z),
                       rad = mk3(q3.x, q3.y, q3.z), res = mk3(0.0f, 0.0f, 0.0f);
                    bool done = rt_shade_segment(scene, h, rng, org, dir, att, rad, res);
                    done = rt_after_segment(p, done, depth, att, rng, res);
                    r.org = org; /* src/render_megakernel.cpp:41-55: re-quantise the ray state (F6) */
                    r.dir = round_half3(dir);
                    r.att = round_half3(att);
                    r.rad = round_half3(rad);
                    if (done) {
                        sum = sum + res;
                        s++;
                        mode = kNeedRay;
                    } else {
                        mode = kStart;
                    }
                }
            }
            for (;;) { /* warp-uniform: runs until no lane is waiting for a pixel */
                if (mode == kNeedRay) {
                    while (p.max_depth == 0 && s < p.spp) { /* the bounce loop never runs: black sample */
                        rng.next();
                        rng.next();
                        s++;
                    }
                    if (s == p.spp) { /* pixel finished: :154-158 mean, gamma, image write */
                        const size_t pix = (size_t)y * (size_t)p.cam.w + (size_t)x;
                        const float count = base_count + (float)p.spp;
                        out.accum[pix] = make_float4(sum.x, sum.y, sum.z, count);
                        const uint32_t px = rt_resolve_pixel(sum.x, sum.y, sum.z, count);
                        out.rgba8[pix] = px;
                        if (out.gather) out.gather[pix] = px; /* tile shards: straight into the destination rank's image */
                        out.rng[pix] = rng.a;
                        mode = kNeedPixel;
                    } else {
                        r = rt_camera_ray(p.cam, x, y, rng); /* 2 draws (F5) */
                        depth = 0;
                        mode = kStart;
                    }
                }
                const unsigned need = __ballot_sync(full, mode == kNeedPixel);
                if (!need) break;
                uint32_t base = 0;
                const int leader = __ffs(need) - 1;
                if (lane == leader) base = atomicAdd(work_counter, (uint32_t)__popc(need));
                base = __shfl_sync(full, base, leader);
                if (mode == kNeedPixel) {
                    const uint32_t idx = base + (uint32_t)__popc(need & ((1u << lane) - 1u));
                    if (idx >= n_work) {
                        mode = kExhausted;
                        st = (st & ~(7u << (3 * c))) | ((uint32_t)cDone << (3 * c));
                    } else {
                        const uint32_t in = idx & 31u;
                        uint32_t x0, y0;
                        if (order) { /* blocks sorted by decreasing cost class (k_block_cost) */
                            const uint32_t e = __ldg(order + (idx >> 5));
                            x0 = e & 0xffffu;
                            y0 = e >> 16;
                        } else {
                            rt_block_origin(p, g, idx >> 5, x0, y0);
                        }
                        x = (int)(x0 + (in & 7u));
                        y = (int)(y0 + (in >> 3));
                        if (x < p.cam.w && y < p.cam.h && rt_owns_pixel(p, x, y)) {
                            if (p.resume) { /* carry on where the previous frame stopped */
                                const size_t pix = (size_t)y * (size_t)p.cam.w + (size_t)x;
                                const float4 a = out.accum[pix];
                                rng.a = out.rng[pix];
                                sum = mk3(a.x, a.y, a.z);
                                base_count = a.w;
                            } else {
                                rng.a = rt_pixel_seed(p.wavefront_seed, x, y, p.cam.w, p.cam.h) ^ p.seed_salt;
                                sum = mk3(0.0f, 0.0f, 0.0f);
                                base_count = 0.0f;
                            }
                            s = 0;
                            mode = kNeedRay;
                        } /* else: padding / another rank's pixel, fetch again */
                    }
                }
            }
            if (mode == kStart) { /* park the context with its ray ready to be traced */
                const RtRayPre pre = rt_ray_pre(r.dir);
                q[0] = make_float4(r.org.x, r.org.y, r.org.z, __uint_as_float(rng.a));
                q[1] = make_float4(r.dir.x, r.dir.y, r.dir.z, __uint_as_float(s));
                q[2] = make_float4(r.att.x, r.att.y, r.att.z, __uint_as_float(depth));
                q[3] = make_float4(r.rad.x, r.rad.y, r.rad.z, base_count);
                q[4] = make_float4(sum.x, sum.y, sum.z, __uint_as_float((uint32_t)x));
                q[5] = make_float4(pre.nSx, pre.nSy, pre.Sz, __uint_as_float(pre.code));
                q[6] = make_float4(pre.rcp.x, pre.rcp.y, pre.rcp.z, __uint_as_float((uint32_t)y));
                st = (st & ~(7u << (3 * c))) | ((uint32_t)cReady << (3 * c));
                rays++;
            }
        }
        /* ---------------- switch ---------------- */
        if (cur < 0) {
            int c = -1;
#pragma unroll
            for (int i = K - 1; i >= 0; i--)
                if (((st >> (3 * i)) & 7u) == cReady) c = i;
            if (c >= 0) {
                const float4 *q = cx + opaque_index(c * kCtxQ);
                const float4 q0 = q[0], q5 = q[5], q6 = q[6];
                RtRayPre pre;
                pre.rcp = mk3(q6.x, q6.y, q6.z);
                pre.nSx = q5.x;
                pre.nSy = q5.y;
                pre.Sz = q5.z;
                pre.code = __float_as_uint(q5.w);
                rt_trav_init_pre(tv, mk3(q0.x, q0.y, q0.z), pre, 0.0001f, INFINITY);
                st = (st & ~(7u << (3 * c))) | ((uint32_t)cTrav << (3 * c));
                cur = c;
            }
        }
        /* ---------------- traverse ---------------- */
        int mode = cur >= 0 ? kTraversing : kExhausted;
        const unsigned act0 = __ballot_sync(full, mode == kTraversing);
        if (!act0) {
            bool alive = false;
#pragma unroll
            for (int i = 0; i < K; i++) alive |= ((st >> (3 * i)) & 7u) != cDone;
            if (!__any_sync(full, alive)) break; /* every context of every lane is exhausted */
            continue;                            /* only shading work is left: the next pass runs it */
        }
        traverse_phase(scene.bvh, tv, ks, mode, p.tune_refill);
        if (mode == kHitPending) { /* park the hit; the lane switches at the top of the next round */
            cx[opaque_index(cur * kCtxQ + 5)] = make_float4(tv.t, tv.u, tv.v, __uint_as_float(tv.tri));
            st = (st & ~(7u << (3 * cur))) | ((uint32_t)cHit << (3 * cur));
            cur = -1;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rays += __shfl_xor_sync(full, rays, o);
    if (lane == 0 && rays) atomicAdd(ray_counter, rays);
}

/* ------------------------------------------------------------------------------ wavefront */
/* block-wide order-preserving compaction of `keep` flags: returns the global slot or ~0u */
__device__ __forceinline__ uint32_t block_compact(bool keep, uint32_t *queue_count, uint32_t *s_warp, uint32_t *s_base) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
    const uint32_t rank = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
        uint32_t c = lane < n_warps ? s_warp[lane] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane < n_warps) s_warp[lane] = incl - c; /* exclusive */
        if (lane == 31) {
            const uint32_t total = incl;
            *s_base = total ? atomicAdd(queue_count, total) : 0u;
        }
    }
    __syncthreads();
    const uint32_t slot = keep ? (*s_base + s_warp[warp] + rank) : 0xffffffffu;
    __syncthreads(); /* s_warp / s_base are reused by the next loop iteration */
    return slot;
}

__global__ void __launch_bounds__(kWfBlock) k_wf_generate(RtFrameParams p, RtWavefrontState w, RtFrameOut out) {
    __shared__ uint32_t s_warp[kWfBlock / 32];
    __shared__ uint32_t s_base;
    const uint32_t n_pix = (uint32_t)p.cam.w * (uint32_t)p.cam.h;
    const uint32_t rounds = (n_pix + blockDim.x * gridDim.x - 1) / (blockDim.x * gridDim.x);
    for (uint32_t it = 0; it < rounds; it++) {
        const uint32_t pix = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
        const bool live = pix < n_pix ? rt_wf_generate_pixel(p, w, out, pix) : false;
        const uint32_t slot = block_compact(live, w.count[0], s_warp, &s_base);
        if (live) w.queue[0][slot] = pix;
    }
}

/* extend: traversal only. Persistent warps pull rays from the id queue through a device-side head
 * counter and replace finished rays inside the traversal loop (same scheme as the megakernel). */
#ifndef RT_EXT_MIN_BLOCKS
#define RT_EXT_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(kWfBlock, RT_EXT_MIN_BLOCKS) k_wf_extend(RtScene scene, RtWavefrontState w, int cur,
                                                         unsigned long long *ray_counter, int tune_refill) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t count = *w.count[cur];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *w.count[cur ^ 1] = 0u; /* the shade kernel of this bounce appends here */
        atomicAdd(ray_counter, (unsigned long long)count); /* src/render_wavefront.cpp:407 */
    }
    const uint32_t *queue = w.queue[cur];
    int mode = kNeedPixel; /* kNeedPixel = idle, kTraversing, kHitPending, kExhausted */
    uint32_t pix = 0;
    RtTravState tv;
    SmemStacks<kWfBlock> ks;
#if RT_SMEM_TRI || RT_SMEM_NODE
    __shared__ uint64_t s_stacks[RT_SMEM_ENTRIES * kWfBlock];
    ks.sm = s_stacks + threadIdx.x;
#endif
    fill_perm_table();
    __syncthreads();
    tv.sp = 0;
    tv.tsp = 0;
    tv.ng_y = 0;
    for (;;) {
        if (mode == kHitPending) {
            w.hit[pix] = make_float4(tv.t, tv.u, tv.v, __uint_as_float(tv.tri));
            mode = kNeedPixel;
        }
        const unsigned need = __ballot_sync(full, mode == kNeedPixel);
        if (need) {
            uint32_t base = 0;
            const int leader = __ffs(need) - 1;
            if (lane == leader) base = atomicAdd(w.head, (uint32_t)__popc(need));
            base = __shfl_sync(full, base, leader);
            if (mode == kNeedPixel) {
                const uint32_t idx = base + (uint32_t)__popc(need & ((1u << lane) - 1u));
                if (idx >= count) {
                    mode = kExhausted;
                } else {
                    pix = queue[idx];
                    const float4 o = w.org[pix];
                    rt_trav_init(tv, mk3(o.x, o.y, o.z), rt_unpack_half3(w.dir[pix]), 0.0001f, INFINITY);
                    mode = kTraversing;
                }
            }
        }
        const unsigned act0 = __ballot_sync(full, mode == kTraversing);
        if (!act0) break;
        traverse_phase(scene.bvh, tv, ks, mode, tune_refill);
    }
}

__global__ void __launch_bounds__(kWfBlock) k_wf_shade(RtScene scene, RtFrameParams p, RtWavefrontState w,
                                                        RtFrameOut out, int cur) {
    __shared__ uint32_t s_warp[kWfBlock / 32];
    __shared__ uint32_t s_base;
    const uint32_t count = *w.count[cur];
    if (blockIdx.x == 0 && threadIdx.x == 0) *w.head = 0u; /* next bounce's extend starts at slot 0 */
    const uint32_t *queue = w.queue[cur];
    uint32_t *next_queue = w.queue[cur ^ 1];
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (count + stride - 1) / stride;
    for (uint32_t it = 0; it < rounds; it++) {
        const uint32_t i = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        bool keep = false;
        uint32_t pix = 0;
        if (i < count) {
            pix = queue[i];
            keep = rt_wf_shade_pixel(scene, p, w, out, pix);
        }
        const uint32_t slot = block_compact(keep, w.count[cur ^ 1], s_warp, &s_base);
        if (keep) next_queue[slot] = pix;
    }
}

/* ------------------------------------------------------------------------------ wavefront, persistent */
/* The whole wavefront frame in ONE launch (K2-K6 of src/render_wavefront.cpp:62-417). The streaming kernels above pay
 * a grid-wide ramp and tail for every one of the spp * max_depth bounce iterations (thousands of launches per frame).
 * Here every CTA owns a fixed share of the image — the 8x4 pixel blocks b with b % gridDim.x == blockIdx.x, a
 * regular lattice over the frame, so the shares cost the same to within a few per cent — and runs generate /
 * {extend, shade}* / resolve for its own pixels with CTA-local queues: the iteration boundary is a __syncthreads() of
 * one CTA, not a grid-wide barrier, the CTAs of an SM drift apart and fill each other's ramps and tails, and the L1
 * keeps the upper BVH levels across iterations. A pixel's state, its queue entries and its accumulation are only
 * ever touched by its own CTA (same SM, same L1), so no device-scope fence is needed. Per-pixel arithmetic is
 * rt_wf_generate_pixel / traverse / rt_wf_shade_pixel, unchanged: results are bit-identical to the streaming form. */
__device__ __forceinline__ bool wf_block_pixel(const RtFrameParams &p, const RtBlockGeom &g, uint32_t blk, int lane, uint32_t &pix,
                                               const uint32_t *order = nullptr) {
    if (blk >= g.n_blocks) return false;
    uint32_t x0, y0;
    if (order) { /* blocks sorted by decreasing cost class (k_block_cost), like the megakernel's hand-out */
        const uint32_t e = __ldg(order + blk);
        x0 = e & 0xffffu;
        y0 = e >> 16;
    } else {
        rt_block_origin(p, g, blk, x0, y0);
    }
    const int x = (int)(x0 + ((uint32_t)lane & 7u)), y = (int)(y0 + ((uint32_t)lane >> 3));
    if (x >= p.cam.w || y >= p.cam.h || !rt_owns_pixel(p, x, y)) return false;
    pix = (uint32_t)y * (uint32_t)p.cam.w + (uint32_t)x;
    return true;
}

__global__ void __launch_bounds__(kWfBlock, RT_EXT_MIN_BLOCKS) k_wf_persistent(RtScene scene, RtFrameParams p, RtWavefrontState w, RtFrameOut out,
                                                                              unsigned long long *ray_counter, uint32_t cap) {
    __shared__ uint32_t s_warp[kWfBlock / 32];
    __shared__ uint32_t s_base;
    __shared__ uint32_t s_count[2];
    __shared__ uint32_t s_head;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr uint32_t kWarps = kWfBlock / 32;
    const RtBlockGeom g = rt_block_geom(p);
    uint32_t *queue[2] = {w.queue[0] + (size_t)blockIdx.x * cap, w.queue[1] + (size_t)blockIdx.x * cap};
    const uint32_t my_blocks = g.n_blocks > blockIdx.x ? (g.n_blocks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    const uint32_t rounds = (my_blocks + kWarps - 1) / kWarps;
    fill_perm_table();
    if (threadIdx.x == 0) {
        s_count[0] = 0u;
        s_count[1] = 0u;
        s_head = 0u;
    }
    __syncthreads();
    /* ---- generate (K2 + first K3): seed, zero, first camera ray of every owned pixel ---- */
    for (uint32_t it = 0; it < rounds; it++) {
        const uint32_t k = it * kWarps + (uint32_t)warp;
        uint32_t pix = 0;
        bool live = k < my_blocks && wf_block_pixel(p, g, blockIdx.x + k * gridDim.x, lane, pix);
        if (live) live = rt_wf_generate_pixel(p, w, out, pix);
        const uint32_t slot = block_compact(live, &s_count[0], s_warp, &s_base);
        if (live) queue[0][slot] = pix;
    }
    unsigned long long rays = 0;
    RtTravState tv;
    SmemStacks<kWfBlock> ks;
#if RT_SMEM_TRI || RT_SMEM_NODE
    __shared__ uint64_t s_stacks[RT_SMEM_ENTRIES * kWfBlock];
    ks.sm = s_stacks + threadIdx.x;
#endif
    int cur = 0;
    for (;;) {
        __syncthreads(); /* queue[cur] and its length are complete */
        const uint32_t count = s_count[cur];
        if (count == 0u) break;
        __syncthreads(); /* every thread has read the length before it is reused */
        if (threadIdx.x == 0) {
            s_count[cur ^ 1] = 0u;
            s_head = 0u;
            rays += count; /* src/render_wavefront.cpp:407 */
        }
        __syncthreads();
        /* ---- extend: traversal only, persistent lanes with replacement from the CTA's queue ---- */
        {
            const uint32_t *q = queue[cur];
            int mode = kNeedPixel;
            uint32_t pix = 0;
            tv.sp = 0;
            tv.tsp = 0;
            tv.ng_y = 0;
            for (;;) {
                if (mode == kHitPending) {
                    w.hit[pix] = make_float4(tv.t, tv.u, tv.v, __uint_as_float(tv.tri));
                    mode = kNeedPixel;
                }
                const unsigned need = __ballot_sync(full, mode == kNeedPixel);
                if (need) {
                    uint32_t base = 0;
                    const int leader = __ffs(need) - 1;
                    if (lane == leader) base = atomicAdd(&s_head, (uint32_t)__popc(need));
                    base = __shfl_sync(full, base, leader);
                    if (mode == kNeedPixel) {
                        const uint32_t idx = base + (uint32_t)__popc(need & ((1u << lane) - 1u));
                        if (idx >= count) {
                            mode = kExhausted;
                        } else {
                            pix = q[idx];
                            const float4 o = w.org[pix];
                            rt_trav_init(tv, mk3(o.x, o.y, o.z), rt_unpack_half3(w.dir[pix]), 0.0001f, INFINITY);
                            mode = kTraversing;
                        }
                    }
                }
                const unsigned act0 = __ballot_sync(full, mode == kTraversing);
                if (!act0) break;
                traverse_phase(scene.bvh, tv, ks, mode, p.tune_refill);
            }
        }
        __syncthreads(); /* every hit record of this iteration is written */
        /* ---- shade + connect + regenerate, survivors compacted into the other queue ---- */
        {
            const uint32_t *q = queue[cur];
            uint32_t *nq = queue[cur ^ 1];
            const uint32_t srounds = (count + kWfBlock - 1) / kWfBlock;
            for (uint32_t it = 0; it < srounds; it++) {
                const uint32_t i = it * kWfBlock + threadIdx.x;
                bool keep = false;
                uint32_t pix = 0;
                if (i < count) {
                    pix = q[i];
                    keep = rt_wf_shade_pixel(scene, p, w, out, pix);
                }
                const uint32_t slot = block_compact(keep, &s_count[cur ^ 1], s_warp, &s_base);
                if (keep) nq[slot] = pix;
            }
        }
        cur ^= 1;
    }
    /* ---- resolve (K6/K7) of the owned pixels ---- */
    for (uint32_t it = 0; it < rounds; it++) {
        const uint32_t k = it * kWarps + (uint32_t)warp;
        uint32_t pix = 0;
        if (k < my_blocks && wf_block_pixel(p, g, blockIdx.x + k * gridDim.x, lane, pix)) {
            const float4 a = out.accum[pix];
            const uint32_t px = rt_resolve_pixel(a.x, a.y, a.z, a.w); /* a.w = samples accumulated (= spp, or more after resumes) */
            out.rgba8[pix] = px;
            if (out.gather) out.gather[pix] = px;
            out.rng[pix] = w.rng[pix];
        }
    }
    if (threadIdx.x == 0 && rays) atomicAdd(ray_counter, rays);
}

/* ------------------------------------------------------------------------------ wavefront, queue-driven warps */
/* k_wf_persistent still ends every bounce iteration with a drain (the CTA's last rays run in half-empty warps): making
 * its CTAs smaller, down to one warp, changes nothing (profiles/README.md), the loss is the iteration itself. Here the
 * unit is the WARP and there are no iterations. A warp keeps `tune_inflight` pixels in flight, claimed 8x4 block by
 * block from a global counter (like the megakernel's hand-out, so the frame ends within one pixel time on every warp),
 * and two private ring queues in global memory — pixels whose ray is ready to be traced, pixels whose hit is ready to
 * be shaded — through which its pixels flow continuously:
 *   claim  : while fewer than tune_inflight pixels are in flight, fetch the next block and generate its camera rays;
 *   hits   : lanes whose traversal finished store the hit record and append the pixel to the hit ring;
 *   shade  : as soon as 32 hits are queued (or nothing else can be done) the warp shades 32 of them with all lanes
 *            (rt_wf_shade_pixel: material, rng, accumulate, regenerate), appends the survivors to the ray ring and
 *            resolves the pixels that finished their samples (K6/K7);
 *   refill : idle lanes take the next rays from the ray ring and start traversing;
 *   traverse_phase(): node steps / triangle drain with the warp converged.
 * No barrier, no atomics on the queues, no fences: a warp's rings and the state of the pixels it claimed are private to
 * it, and warp-level program order (__syncwarp) is all the ordering they need. Per-pixel arithmetic is unchanged: a
 * pixel has one ray in flight and its own stream, so results are bit-identical to the other two forms. */
__device__ __forceinline__ void wf_resolve_pixel(const RtWavefrontState &w, const RtFrameOut &out, uint32_t pix) {
    const float4 a = out.accum[pix];
    const uint32_t px = rt_resolve_pixel(a.x, a.y, a.z, a.w); /* a.w = samples accumulated (= spp, or more after resumes) */
    out.rgba8[pix] = px;
    if (out.gather) out.gather[pix] = px;
    out.rng[pix] = w.rng[pix];
}

__global__ void __launch_bounds__(kWfBlock, RT_EXT_MIN_BLOCKS) k_wf_flow(RtScene scene, RtFrameParams p, RtWavefrontState w, RtFrameOut out,
                                                                        uint32_t *work_counter, unsigned long long *ray_counter,
                                                                        uint32_t cap /* slots per ring, a power of two >= tune_inflight */,
                                                                        const uint32_t *__restrict__ order) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t wid = blockIdx.x * (uint32_t)(kWfBlock / 32) + (threadIdx.x >> 5);
    const RtBlockGeom g = rt_block_geom(p);
    uint32_t *rayq = w.queue[0] + (size_t)wid * cap, *hitq = w.queue[1] + (size_t)wid * cap;
    const uint32_t mask = cap - 1u;
    const uint32_t n_chains = p.chains > 1u ? p.chains : 1u, n_pix = (uint32_t)p.cam.w * (uint32_t)p.cam.h;
    fill_perm_table();
    __syncthreads();
    uint32_t ray_head = 0, ray_tail = 0, hit_head = 0, hit_tail = 0; /* warp-uniform ring cursors */
    uint32_t inflight = 0;                                             /* pixels claimed and not finished */
    bool more = true;                                                  /* blocks are left on the global counter */
    unsigned long long rays = 0;
    RtTravState tv;
    SmemStacks<kWfBlock> ks;
#if RT_SMEM_TRI || RT_SMEM_NODE
    __shared__ uint64_t s_stacks[RT_SMEM_ENTRIES * kWfBlock];
    ks.sm = s_stacks + threadIdx.x;
#endif
    tv.sp = 0;
    tv.tsp = 0;
    tv.ng_y = 0;
    int mode = kNeedPixel; /* kNeedPixel = idle */
    uint32_t pix = 0;
    for (;;) {
        /* ---- claim + generate (K2 + first K3) ---- */
        while (more && inflight + 32u <= (uint32_t)p.tune_inflight) {
            uint32_t item = 0; /* (block, chain) pairs, the chains of a block next to each other */
            if (lane == 0) item = atomicAdd(work_counter, 1u);
            item = __shfl_sync(full, item, 0);
            if (item >= g.n_blocks * n_chains) {
                more = false;
                break;
            }
            uint32_t gp = 0;
            bool live = wf_block_pixel(p, g, item / n_chains, lane, gp, order);
            if (live) {
                gp += (item % n_chains) * n_pix; /* virtual pixel */
                live = rt_wf_generate_pixel(p, w, out, gp);
                if (!live && n_chains == 1u) wf_resolve_pixel(w, out, gp); /* no samples to trace */
            }
            const unsigned m = __ballot_sync(full, live);
            if (live) rayq[(ray_tail + (uint32_t)__popc(m & lt)) & mask] = gp;
            ray_tail += (uint32_t)__popc(m);
            inflight += (uint32_t)__popc(m);
            __syncwarp();
        }
        /* ---- finished traversals -> hit ring ---- */
        {
            const bool fin = mode == kHitPending;
            const unsigned m = __ballot_sync(full, fin);
            if (m) {
                if (fin) {
                    w.hit[pix] = make_float4(tv.t, tv.u, tv.v, __uint_as_float(tv.tri));
                    hitq[(hit_tail + (uint32_t)__popc(m & lt)) & mask] = pix;
                    mode = kNeedPixel;
                }
                hit_tail += (uint32_t)__popc(m);
                __syncwarp();
            }
        }
        const unsigned m_trav = __ballot_sync(full, mode == kTraversing);
        /* ---- shade: 32 wide whenever possible; narrower only when the warp would otherwise starve ---- */
        for (;;) {
            const uint32_t n_hits = hit_tail - hit_head, n_rays = ray_tail - ray_head;
            const int n_idle = 32 - __popc(m_trav);
            if (!(n_hits >= 32u || (n_hits > 0u && n_rays == 0u && (n_idle >= (p.tune_refill & 0xff) || !m_trav)))) break;
            const uint32_t n = n_hits < 32u ? n_hits : 32u;
            bool keep = false;
            uint32_t spix = 0;
            if ((uint32_t)lane < n) {
                spix = hitq[(hit_head + (uint32_t)lane) & mask];
                keep = rt_wf_shade_pixel(scene, p, w, out, spix);
                if (!keep && n_chains == 1u) wf_resolve_pixel(w, out, spix); /* the pixel's last sample: K6/K7 (chains: k_combine_chains) */
            }
            const unsigned mk = __ballot_sync(full, keep);
            if (keep) rayq[(ray_tail + (uint32_t)__popc(mk & lt)) & mask] = spix;
            hit_head += n;
            ray_tail += (uint32_t)__popc(mk);
            inflight -= n - (uint32_t)__popc(mk);
            __syncwarp();
        }
        /* ---- refill idle lanes from the ray ring ---- */
        {
            const bool idle = mode == kNeedPixel;
            const unsigned m = __ballot_sync(full, idle);
            const uint32_t n_rays = ray_tail - ray_head;
            if (m && n_rays) {
                const uint32_t r = (uint32_t)__popc(m & lt);
                if (idle && r < n_rays) {
                    pix = rayq[(ray_head + r) & mask];
                    const float4 o = w.org[pix];
                    rt_trav_init(tv, mk3(o.x, o.y, o.z), rt_unpack_half3(w.dir[pix]), 0.0001f, INFINITY);
                    mode = kTraversing;
                }
                const uint32_t take = min((uint32_t)__popc(m), n_rays);
                ray_head += take;
                rays += take; /* one rtcIntersect1 per segment, src/render_wavefront.cpp:407 */
            }
        }
        if (!__ballot_sync(full, mode == kTraversing)) {
            if (inflight == 0u && !more) break; /* every pixel this warp claimed has finished its samples */
            continue;
        }
        traverse_phase<RT_BYTE_IDP_MASK, RT_WF_NODE_UNROLL>(scene.bvh, tv, ks, mode, p.tune_refill);
    }
    if (lane == 0 && rays) atomicAdd(ray_counter, rays);
}

/* ------------------------------------------------------------------------------ resolve */
__global__ void k_resolve(const float4 *accum, uint32_t *rgba8, uint32_t n_pix, float spp) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const float4 a = accum[i];
    rgba8[i] = rt_resolve_pixel(a.x, a.y, a.z, spp);
}

__global__ void k_resolve_owned(RtFrameParams p, const float4 *accum, const uint32_t *rng_state, RtFrameOut out) {
    const uint32_t n_pix = (uint32_t)p.cam.w * (uint32_t)p.cam.h;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const int x = (int)(i % (uint32_t)p.cam.w), y = (int)(i / (uint32_t)p.cam.w);
    if (rt_owns_pixel(p, x, y)) {
        const float4 a = accum[i];
        const uint32_t px = rt_resolve_pixel(a.x, a.y, a.z, a.w); /* a.w = samples accumulated (= spp, or more after resumes) */
        out.rgba8[i] = px;
        if (out.gather) out.gather[i] = px;
    } else if (!p.keep_foreign) {
        out.rgba8[i] = 0u;
    }
    out.rng[i] = rng_state[i];
}

/* sample chains: sum the chain planes of every owned pixel in chain order, resolve, store (also to the gather target) */
__global__ void k_combine_chains(RtFrameParams p, const float4 *chain_accum, const uint32_t *chain_rng, RtFrameOut out) {
    const uint32_t n_pix = (uint32_t)p.cam.w * (uint32_t)p.cam.h;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const int x = (int)(i % (uint32_t)p.cam.w), y = (int)(i / (uint32_t)p.cam.w);
    if (!rt_owns_pixel(p, x, y)) return;
    float4 s = chain_accum[i];
    for (uint32_t c = 1; c < p.chains; c++) {
        const float4 a = chain_accum[(size_t)c * n_pix + i];
        s.x += a.x;
        s.y += a.y;
        s.z += a.z;
        s.w += a.w;
    }
    out.accum[i] = s;
    const uint32_t px = rt_resolve_pixel(s.x, s.y, s.z, s.w);
    out.rgba8[i] = px;
    if (out.gather) out.gather[i] = px;
    out.rng[i] = chain_rng[i]; /* chain 0: the reference stream's state */
}

/* contraction self-test: (a*b)+c must be two roundings (DESIGN.md arithmetic contract) */
__global__ void k_selftest(float a, float b, float c, float *o) {
    o[0] = a * b + c;
    o[1] = __fmaf_rn(a, b, c);
}

} // namespace

/* =================================================================== host-side launchers */
cudaError_t rt_launch_intersect(cudaStream_t st, const RtScene &scene, const RtInstance *inst, uint64_t n,
                                const float *org, const float *dir, float tnear, float tfar, int32_t *o_inst,
                                int32_t *o_prim, float *o_u, float *o_v, float *o_t) {
    if (n == 0) return cudaSuccess;
    const int block = 128;
    k_intersect<<<(unsigned)((n + block - 1) / block), block, 0, st>>>(scene, inst, n, org, dir, tnear, tfar, o_inst,
                                                                       o_prim, o_u, o_v, o_t);
    return cudaGetLastError();
}

cudaError_t rt_megakernel_grid(int sm_count, int tune_ctx, int *grid) {
    int per_sm = 0;
    cudaError_t e = cudaSuccess;
    switch (tune_ctx) {
    case 1: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_megakernel_ctx<1>, kMegaBlock, 0); break;
    case 2: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_megakernel_ctx<2>, kMegaBlock, 0); break;
    case 3: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_megakernel_ctx<3>, kMegaBlock, 0); break;
    case 4: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_megakernel_ctx<4>, kMegaBlock, 0); break;
    default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_megakernel<false>, kMegaBlock, 0); break;
    }
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    *grid = sm_count * per_sm;
    return cudaSuccess;
}

cudaError_t rt_launch_megakernel(cudaStream_t st, int grid, const RtScene &scene, const RtFrameParams &p,
                                 const RtFrameOut &out, uint32_t *work_counter, unsigned long long *ray_counter,
                                 const uint32_t *order) {
    switch (p.tune_ctx) {
    case 1: k_megakernel_ctx<1><<<grid, kMegaBlock, 0, st>>>(scene, p, out, work_counter, ray_counter, order); break;
    case 2: k_megakernel_ctx<2><<<grid, kMegaBlock, 0, st>>>(scene, p, out, work_counter, ray_counter, order); break;
    case 3: k_megakernel_ctx<3><<<grid, kMegaBlock, 0, st>>>(scene, p, out, work_counter, ray_counter, order); break;
    case 4: k_megakernel_ctx<4><<<grid, kMegaBlock, 0, st>>>(scene, p, out, work_counter, ray_counter, order); break;
    default:
        if (p.chains > 1u) k_megakernel<true><<<grid, kMegaBlock, 0, st>>>(scene, p, out, work_counter, ray_counter, order, rt_block_geom(p));
        else k_megakernel<false><<<grid, kMegaBlock, 0, st>>>(scene, p, out, work_counter, ray_counter, order, rt_block_geom(p));
        break;
    }
    return cudaGetLastError();
}

void rt_counters_read(unsigned long long out[2], bool reset) {
    out[0] = out[1] = 0;
#ifdef RT_GPU_COUNTERS
    cudaMemcpyFromSymbol(out, g_rt_counters, sizeof(unsigned long long) * 2);
    if (reset) {
        const unsigned long long z[2] = {0, 0};
        cudaMemcpyToSymbol(g_rt_counters, z, sizeof(z));
    }
#else
    (void)reset;
#endif
}

uint32_t rt_block_count(const RtFrameParams &p) { return rt_block_geom(p).n_blocks; }

cudaError_t rt_block_order_temp_bytes(uint32_t n_blocks, size_t *bytes) {
    *bytes = 0;
    return cub::DeviceRadixSort::SortPairsDescending(nullptr, *bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                                     (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n_blocks, 0, 8);
}

uint32_t rt_region_count(int w, int h) { return (((uint32_t)w + kRegionMin - 1u) / kRegionMin) * (((uint32_t)h + kRegionMin - 1u) / kRegionMin); }

cudaError_t rt_launch_block_order(cudaStream_t st, const RtScene &scene, const RtFrameParams &p, uint32_t *region_cost, uint32_t *keys_in,
                                  uint32_t *keys_out, uint32_t *vals_in, uint32_t *vals_out, void *temp, size_t temp_bytes) {
    const RtBlockGeom g = rt_block_geom(p);
    if (g.n_blocks == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(region_cost, 0, rt_region_count(p.cam.w, p.cam.h) * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    k_block_cost<<<(g.n_blocks + 127) / 128, 128, 0, st>>>(scene, p, g, region_cost, vals_in);
    k_block_key<<<(g.n_blocks + 127) / 128, 128, 0, st>>>(p, g, region_cost, vals_in, keys_in);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    /* stable, descending: expensive classes first, enumeration (image) order inside a class */
    return cub::DeviceRadixSort::SortPairsDescending(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, (int)g.n_blocks, 0, 8, st);
}

cudaError_t rt_wavefront_grid(int sm_count, int *grid_extend, int *grid_shade) {
    int a = 0, b = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_wf_extend, kWfBlock, 0);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_wf_shade, kWfBlock, 0);
    if (e != cudaSuccess) return e;
    *grid_extend = sm_count * (a < 1 ? 1 : a);
    *grid_shade = sm_count * (b < 1 ? 1 : b);
    return cudaSuccess;
}

cudaError_t rt_launch_wf_persistent(cudaStream_t st, int grid, uint32_t cap, const RtScene &scene, const RtFrameParams &p,
                                    const RtWavefrontState &w, const RtFrameOut &out, unsigned long long *ray_counter) {
    k_wf_persistent<<<grid, kWfBlock, 0, st>>>(scene, p, w, out, ray_counter, cap);
    return cudaGetLastError();
}

cudaError_t rt_launch_wf_flow(cudaStream_t st, int grid, uint32_t cap, const RtScene &scene, const RtFrameParams &p,
                              const RtWavefrontState &w, const RtFrameOut &out, uint32_t *work_counter, unsigned long long *ray_counter,
                              const uint32_t *order) {
    k_wf_flow<<<grid, kWfBlock, 0, st>>>(scene, p, w, out, work_counter, ray_counter, cap, order);
    return cudaGetLastError();
}

cudaError_t rt_wf_flow_grid(int sm_count, int *grid, int *warps_per_block) {
    int a = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_wf_flow, kWfBlock, 0);
    if (e != cudaSuccess) return e;
    *grid = sm_count * (a < 1 ? 1 : a);
    *warps_per_block = kWfBlock / 32;
    return cudaSuccess;
}

cudaError_t rt_wf_persistent_grid(int sm_count, int *grid) {
    int a = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_wf_persistent, kWfBlock, 0);
    if (e != cudaSuccess) return e;
    *grid = sm_count * (a < 1 ? 1 : a);
    return cudaSuccess;
}

cudaError_t rt_launch_wf_generate(cudaStream_t st, int grid, const RtFrameParams &p, const RtWavefrontState &w,
                                  const RtFrameOut &out) {
    k_wf_generate<<<grid, kWfBlock, 0, st>>>(p, w, out);
    return cudaGetLastError();
}

cudaError_t rt_launch_wf_extend(cudaStream_t st, int grid, const RtScene &scene, const RtWavefrontState &w, int cur,
                                unsigned long long *ray_counter, const RtFrameParams &p) {
    k_wf_extend<<<grid, kWfBlock, 0, st>>>(scene, w, cur, ray_counter, p.tune_refill);
    return cudaGetLastError();
}

cudaError_t rt_launch_wf_shade(cudaStream_t st, int grid, const RtScene &scene, const RtFrameParams &p,
                               const RtWavefrontState &w, const RtFrameOut &out, int cur) {
    k_wf_shade<<<grid, kWfBlock, 0, st>>>(scene, p, w, out, cur);
    return cudaGetLastError();
}

cudaError_t rt_launch_resolve(cudaStream_t st, const float *accum, uint32_t *rgba8, uint32_t n_pix, float spp) {
    if (n_pix == 0) return cudaSuccess;
    k_resolve<<<(n_pix + 255) / 256, 256, 0, st>>>((const float4 *)accum, rgba8, n_pix, spp);
    return cudaGetLastError();
}

cudaError_t rt_launch_resolve_owned(cudaStream_t st, const RtFrameParams &p, const float *accum,
                                    const uint32_t *rng_state, const RtFrameOut &out) {
    const uint32_t n_pix = (uint32_t)p.cam.w * (uint32_t)p.cam.h;
    if (n_pix == 0) return cudaSuccess;
    k_resolve_owned<<<(n_pix + 255) / 256, 256, 0, st>>>(p, (const float4 *)accum, rng_state, out);
    return cudaGetLastError();
}

cudaError_t rt_launch_combine_chains(cudaStream_t st, const RtFrameParams &p, const float *chain_accum, const uint32_t *chain_rng,
                                     const RtFrameOut &out) {
    const uint32_t n_pix = (uint32_t)p.cam.w * (uint32_t)p.cam.h;
    if (n_pix == 0) return cudaSuccess;
    k_combine_chains<<<(n_pix + 255) / 256, 256, 0, st>>>(p, (const float4 *)chain_accum, chain_rng, out);
    return cudaGetLastError();
}

cudaError_t rt_launch_selftest(cudaStream_t st, float a, float b, float c, float *o) {
    k_selftest<<<1, 1, 0, st>>>(a, b, c, o);
    return cudaGetLastError();
}
