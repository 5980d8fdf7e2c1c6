#!/bin/bash
# round 2, call Y (1 GPU): the tail of 8-way tile shards of C4, rank by rank on one GPU (tools/tile_probe.py)
mkdir -p gpurun_out; : > gpurun_out/y_tile.log
for kind in megakernel wavefront; do
  for cfg in "X=0" "RT_BLOCK_ORDER_MIN_SPP=1" "RT_TUNE_INFLIGHT=32"; do
    echo "== $kind $cfg" >> gpurun_out/y_tile.log
    env $cfg timeout 600 python tools/tile_probe.py c4_heightfield_10m 8 64 $kind 2>&1 | grep -v "^$" >> gpurun_out/y_tile.log
  done
done
echo "== tile 32" >> gpurun_out/y_tile.log
timeout 600 python tools/tile_probe.py c4_heightfield_10m 8 32 megakernel 2>&1 >> gpurun_out/y_tile.log
cat gpurun_out/y_tile.log
