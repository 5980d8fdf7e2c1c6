/*
 * rt_blocks.h — how the megakernel's pixel slots map to pixels (shared by render.cu and the host
 * emulation, which checks the index arithmetic on the CPU).
 *
 * Pixels are handed to the persistent lanes in 8x4 blocks (32 consecutive slots of the global
 * counter = one coherent block; the reference's work-group is 8x8, src/render_megakernel.cpp:85-93).
 * Unsharded: blocks row-major over the image. Image-tile shards (rt_shard.tile_size != 0): only this
 * rank's tiles are enumerated — the k-th owned tile is tile k * world + rank (rt_owns_pixel) — and
 * blocks are row-major inside each tile.
 */
#ifndef RT_BLOCKS_H
#define RT_BLOCKS_H

#include "rt_shade.h"

struct RtBlockGeom {
    uint32_t tiled, ts, blocks_x, tiles_px, per_tile, n_blocks;
};

RT_HD RtBlockGeom rt_block_geom(const RtFrameParams &p) {
    RtBlockGeom g;
    g.tiled = (p.world > 1 && p.tile_size != 0) ? 1u : 0u;
    g.ts = g.tiled ? p.tile_size : 0u;
    g.blocks_x = ((uint32_t)p.cam.w + 7u) / 8u;
    const uint32_t blocks_y = ((uint32_t)p.cam.h + 3u) / 4u;
    g.tiles_px = g.tiled ? ((uint32_t)p.cam.w + g.ts - 1u) / g.ts : 0u;
    const uint32_t tiles_py = g.tiled ? ((uint32_t)p.cam.h + g.ts - 1u) / g.ts : 0u;
    const uint32_t n_tiles = g.tiles_px * tiles_py;
    const uint32_t owned = g.tiled ? (n_tiles > p.rank ? (n_tiles - p.rank + p.world - 1u) / p.world : 0u) : 0u;
    g.per_tile = (g.ts >> 3) * (g.ts >> 2); /* tile_size is a multiple of 8 */
    g.n_blocks = g.tiled ? owned * g.per_tile : g.blocks_x * blocks_y;
    return g;
}

/* top-left pixel of block `blk` of this rank's enumeration (may lie outside the image in a partial
 * edge tile; the caller checks x < w && y < h per pixel) */
RT_HD void rt_block_origin(const RtFrameParams &p, const RtBlockGeom &g, uint32_t blk, uint32_t &x0, uint32_t &y0) {
    uint32_t bx0 = 0, by0 = 0, bw = g.blocks_x;
    if (g.tiled) {
        const uint32_t t = (blk / g.per_tile) * p.world + p.rank;
        blk %= g.per_tile;
        bw = g.ts >> 3;
        bx0 = (t % g.tiles_px) * g.ts;
        by0 = (t / g.tiles_px) * g.ts;
    }
    x0 = bx0 + (blk % bw) * 8u;
    y0 = by0 + (blk / bw) * 4u;
}

#endif /* RT_BLOCKS_H */
