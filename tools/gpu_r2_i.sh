#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/i_wf.log
for wl in c3_sponza_scale c2_cornell c4_heightfield_10m; do
  s=64; [ $wl = c4_heightfield_10m ] && s=16
  timeout 600 python tools/tune.py --workload $wl --renderer wavefront --spp $s --frames 3 --configs "RT_TUNE_INFLIGHT=32;RT_TUNE_INFLIGHT=64;RT_TUNE_INFLIGHT=96;RT_TUNE_INFLIGHT=64,RT_TUNE_REFILL=6;RT_TUNE_INFLIGHT=64,RT_TUNE_REFILL=8;RT_TUNE_INFLIGHT=64,RT_TUNE_REFILL=10;RT_TUNE_INFLIGHT=64,RT_TUNE_REFILL=14;RT_TUNE_INFLIGHT=64,RT_TUNE_REFILL=16;RT_TUNE_INFLIGHT=96,RT_TUNE_REFILL=16" 2>&1 | grep -E "Mrays|rror" >> gpurun_out/i_wf.log
done
cat gpurun_out/i_wf.log
