/*
 * rt_render.h — kernel argument blocks and launchers shared by render.cu and rt_api.cu.
 */
#ifndef RT_RENDER_H
#define RT_RENDER_H

#include <cuda_runtime.h>

#include "rt_wavefront.h"

cudaError_t rt_launch_intersect(cudaStream_t st, const RtScene &scene, const RtInstance *inst, uint64_t n,
                                const float *org, const float *dir, float tnear, float tfar, int32_t *o_inst,
                                int32_t *o_prim, float *o_u, float *o_v, float *o_t);
cudaError_t rt_megakernel_grid(int sm_count, int tune_ctx, int *grid);
cudaError_t rt_launch_megakernel(cudaStream_t st, int grid, const RtScene &scene, const RtFrameParams &p,
                                 const RtFrameOut &out, uint32_t *work_counter, unsigned long long *ray_counter,
                                 const uint32_t *order /* NULL: enumeration order */);
/* block order of the megakernel: cost probe + one-pass stable radix sort (k_block_cost in render.cu) */
uint32_t rt_block_count(const RtFrameParams &p);
cudaError_t rt_block_order_temp_bytes(uint32_t n_blocks, size_t *bytes);
uint32_t rt_region_count(int w, int h);
cudaError_t rt_launch_block_order(cudaStream_t st, const RtScene &scene, const RtFrameParams &p, uint32_t *region_cost, uint32_t *keys_in,
                                  uint32_t *keys_out, uint32_t *vals_in, uint32_t *vals_out, void *temp, size_t temp_bytes);
cudaError_t rt_wavefront_grid(int sm_count, int *grid_extend, int *grid_shade);
/* the whole wavefront frame in one launch: CTA-local queues of `cap` slots each (render.cu) */
cudaError_t rt_wf_persistent_grid(int sm_count, int *grid);
/* queue-driven warps: one ray ring and one hit ring of `cap` (power of two >= tune_inflight) slots per warp */
cudaError_t rt_wf_flow_grid(int sm_count, int *grid, int *warps_per_block);
cudaError_t rt_launch_wf_flow(cudaStream_t st, int grid, uint32_t cap, const RtScene &scene, const RtFrameParams &p,
                              const RtWavefrontState &w, const RtFrameOut &out, uint32_t *work_counter, unsigned long long *ray_counter,
                              const uint32_t *order /* NULL: enumeration order */);
cudaError_t rt_launch_wf_persistent(cudaStream_t st, int grid, uint32_t cap, const RtScene &scene, const RtFrameParams &p,
                                    const RtWavefrontState &w, const RtFrameOut &out, unsigned long long *ray_counter);
cudaError_t rt_launch_wf_generate(cudaStream_t st, int grid, const RtFrameParams &p, const RtWavefrontState &w,
                                  const RtFrameOut &out);
cudaError_t rt_launch_wf_extend(cudaStream_t st, int grid, const RtScene &scene, const RtWavefrontState &w, int cur,
                                unsigned long long *ray_counter, const RtFrameParams &p);
cudaError_t rt_launch_wf_shade(cudaStream_t st, int grid, const RtScene &scene, const RtFrameParams &p,
                               const RtWavefrontState &w, const RtFrameOut &out, int cur);
cudaError_t rt_launch_resolve(cudaStream_t st, const float *accum, uint32_t *rgba8, uint32_t n_pix, float spp);
cudaError_t rt_launch_resolve_owned(cudaStream_t st, const RtFrameParams &p, const float *accum,
                                    const uint32_t *rng_state, const RtFrameOut &out);
/* RT_GPU_COUNTERS builds: traversal steps since the last reset (zeros otherwise) */
void rt_counters_read(unsigned long long out[2], bool reset);
/* sample chains (rt_render_params.sample_chains > 1): sum of the chain planes -> accum / rgba8 / rng of the frame */
cudaError_t rt_launch_combine_chains(cudaStream_t st, const RtFrameParams &p, const float *chain_accum, const uint32_t *chain_rng,
                                     const RtFrameOut &out);
/* reduce-scatter + resolve + gather over peer memory (rt_group.cu): pixels [first, first + count) summed over `world`
 * accumulation buffers in rank order; sum_out may be NULL */
cudaError_t rt_launch_reduce_resolve_peer(cudaStream_t st, const float4 *const *accum, uint32_t world, uint32_t first, uint32_t count,
                                          float4 *sum_out, uint32_t *rgba8_out);
cudaError_t rt_launch_selftest(cudaStream_t st, float a, float b, float c, float *o);

#endif
