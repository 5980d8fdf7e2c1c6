#!/bin/bash
# refill threshold x drain carry-over, jointly
mkdir -p gpurun_out; : > gpurun_out/kk_refill_carry.log
C="RT_TUNE_CARRY=0;RT_TUNE_REFILL=12,RT_TUNE_CARRY=2;RT_TUNE_REFILL=14,RT_TUNE_CARRY=2;RT_TUNE_REFILL=16,RT_TUNE_CARRY=2;RT_TUNE_REFILL=18,RT_TUNE_CARRY=2;RT_TUNE_REFILL=12,RT_TUNE_CARRY=3;RT_TUNE_REFILL=16,RT_TUNE_CARRY=3;RT_TUNE_REFILL=18,RT_TUNE_CARRY=3;RT_TUNE_REFILL=20,RT_TUNE_CARRY=3"
for wl in c3_sponza_scale c2_cornell stadium; do
  for r in megakernel wavefront; do
    timeout 600 python tools/tune.py --workload $wl --renderer $r --spp 64 --frames 5 --configs "$C" 2>&1 | grep -E "Mrays|rror" | sed 's/1920x1080 //; s/depth=10 //' >> gpurun_out/kk_refill_carry.log
  done
done
cat gpurun_out/kk_refill_carry.log
