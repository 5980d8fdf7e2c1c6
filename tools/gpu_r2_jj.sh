#!/bin/bash
# drain carry-over (RT_TUNE_CARRY): sweep on the three scenes, both renderers; the previous build (variants/librt_h_base.so) beside it
mkdir -p gpurun_out; : > gpurun_out/jj_carry.log
C="RT_TUNE_CARRY=0;RT_TUNE_CARRY=1;RT_TUNE_CARRY=2;RT_TUNE_CARRY=3;RT_TUNE_CARRY=4;RT_TUNE_CARRY=6;RT_TUNE_CARRY=8;RT_TUNE_CARRY=12"
for wl in c3_sponza_scale c2_cornell; do
  for r in megakernel wavefront; do
    RT_LIB_PATH=$PWD/variants/librt_h_base.so timeout 300 python tools/tune.py --workload $wl --renderer $r --spp 64 --frames 5 2>&1 | grep -E "Mrays|rror" | sed 's/^/h_base /; s/1920x1080 //; s/depth=10 //' >> gpurun_out/jj_carry.log
    timeout 600 python tools/tune.py --workload $wl --renderer $r --spp 64 --frames 5 --configs "$C" 2>&1 | grep -E "Mrays|rror" | sed 's/1920x1080 //; s/depth=10 //' >> gpurun_out/jj_carry.log
  done
done
timeout 600 python tools/tune.py --workload c4_heightfield_10m --renderer megakernel --spp 16 --frames 5 --configs "RT_TUNE_CARRY=0;RT_TUNE_CARRY=2;RT_TUNE_CARRY=4;RT_TUNE_CARRY=8" 2>&1 | grep -E "Mrays|rror" >> gpurun_out/jj_carry.log
cat gpurun_out/jj_carry.log
