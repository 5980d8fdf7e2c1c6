#!/bin/bash
mkdir -p gpurun_out
for k in megakernel wavefront; do timeout 600 python tools/tail_fit.py c4_heightfield_10m $k 2>&1 | grep -v "^$"; done | tee gpurun_out/z_tail_fit.log
