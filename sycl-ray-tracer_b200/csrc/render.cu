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
#include "rt_render.h"

namespace {

constexpr int kMegaBlock = 128;
constexpr int kWfBlock = 256;

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

/* ------------------------------------------------------------------------------ megakernel */
__global__ void __launch_bounds__(kMegaBlock) k_megakernel(RtScene scene, RtFrameParams p, RtFrameOut out,
                                                           uint32_t *work_counter, unsigned long long *ray_counter) {
    const int lane = threadIdx.x & 31;
    const uint32_t tiles_x = ((uint32_t)p.cam.w + 7u) / 8u, tiles_y = ((uint32_t)p.cam.h + 3u) / 4u;
    const uint32_t n_work = tiles_x * tiles_y;
    unsigned long long rays = 0;
    for (;;) { /* persistent warps: fetch the next 8x4 pixel tile */
        uint32_t work = 0;
        if (lane == 0) work = atomicAdd(work_counter, 1u);
        work = __shfl_sync(0xffffffffu, work, 0);
        if (work >= n_work) break;
        const int x = (int)((work % tiles_x) * 8u) + (lane & 7);
        const int y = (int)((work / tiles_x) * 4u) + (lane >> 3);
        if (x < p.cam.w && y < p.cam.h && rt_owns_pixel(p, x, y)) {
            XorShift32 rng;
            const f3 sum = rt_megakernel_pixel(scene, p, x, y, rng, rays);
            const size_t pix = (size_t)y * (size_t)p.cam.w + (size_t)x;
            out.accum[pix] = make_float4(sum.x, sum.y, sum.z, (float)p.spp);
            out.rgba8[pix] = rt_resolve_pixel(sum.x, sum.y, sum.z, (float)p.spp);
            out.rng[pix] = rng.a;
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rays += __shfl_xor_sync(0xffffffffu, rays, o);
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

__global__ void __launch_bounds__(kWfBlock) k_wf_extend(RtScene scene, RtWavefrontState w, int cur,
                                                         unsigned long long *ray_counter) {
    const uint32_t count = *w.count[cur];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *w.count[cur ^ 1] = 0u; /* the shade kernel of this bounce appends here */
        atomicAdd(ray_counter, (unsigned long long)count); /* src/render_wavefront.cpp:407 */
    }
    const uint32_t *queue = w.queue[cur];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        rt_wf_extend_pixel(scene, w, queue[i]);
    }
}

__global__ void __launch_bounds__(kWfBlock) k_wf_shade(RtScene scene, RtFrameParams p, RtWavefrontState w,
                                                        RtFrameOut out, int cur) {
    __shared__ uint32_t s_warp[kWfBlock / 32];
    __shared__ uint32_t s_base;
    const uint32_t count = *w.count[cur];
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
        out.rgba8[i] = rt_resolve_pixel(a.x, a.y, a.z, (float)p.spp);
    } else {
        out.rgba8[i] = 0u;
    }
    out.rng[i] = rng_state[i];
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

cudaError_t rt_megakernel_grid(int sm_count, int *grid) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_megakernel, kMegaBlock, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    *grid = sm_count * per_sm;
    return cudaSuccess;
}

cudaError_t rt_launch_megakernel(cudaStream_t st, int grid, const RtScene &scene, const RtFrameParams &p,
                                 const RtFrameOut &out, uint32_t *work_counter, unsigned long long *ray_counter) {
    k_megakernel<<<grid, kMegaBlock, 0, st>>>(scene, p, out, work_counter, ray_counter);
    return cudaGetLastError();
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

cudaError_t rt_launch_wf_generate(cudaStream_t st, int grid, const RtFrameParams &p, const RtWavefrontState &w,
                                  const RtFrameOut &out) {
    k_wf_generate<<<grid, kWfBlock, 0, st>>>(p, w, out);
    return cudaGetLastError();
}

cudaError_t rt_launch_wf_extend(cudaStream_t st, int grid, const RtScene &scene, const RtWavefrontState &w, int cur,
                                unsigned long long *ray_counter) {
    k_wf_extend<<<grid, kWfBlock, 0, st>>>(scene, w, cur, ray_counter);
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

cudaError_t rt_launch_selftest(cudaStream_t st, float a, float b, float c, float *o) {
    k_selftest<<<1, 1, 0, st>>>(a, b, c, o);
    return cudaGetLastError();
}
