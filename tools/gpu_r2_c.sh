#!/bin/bash
# round 2, call C: compile-flag variants (perm table layout, registers / occupancy, SAH leaf cost), megakernel + wavefront
mkdir -p gpurun_out; : > gpurun_out/c_variants.log
C="RT_MEGA_CTX=0;RT_MEGA_CTX=1,RT_TUNE_REFILL=12;RT_MEGA_CTX=2,RT_TUNE_REFILL=8"
for v in default perm2 mb6 mb7 prim10 prim15 prim25; do
  if [ $v = default ]; then unset RT_LIB_PATH; else export RT_LIB_PATH=$PWD/variants/librt_$v.so; fi
  for wl in c3_sponza_scale c2_cornell; do
    echo "== $v $wl" >> gpurun_out/c_variants.log
    timeout 300 python tools/tune.py --workload $wl --spp 64 --frames 3 --configs "$C" 2>&1 | grep -E "Mrays|Error|error" >> gpurun_out/c_variants.log
    timeout 300 python tools/tune.py --workload $wl --renderer wavefront --spp 64 --frames 3 2>&1 | grep -E "Mrays|Error|error" >> gpurun_out/c_variants.log
  done
done
cat gpurun_out/c_variants.log | sed 's/1920x1080 spp=64 depth=10 //'
