#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/f_wfblock.log
for v in wfb128 wfb64 wfb32; do
  export RT_LIB_PATH=$PWD/variants/librt_$v.so
  for wl in c3_sponza_scale c2_cornell; do
    echo "== $v" >> gpurun_out/f_wfblock.log
    timeout 300 python tools/tune.py --workload $wl --renderer wavefront --spp 64 --frames 3 --configs "RT_WF_PERSIST=1" 2>&1 | grep -E "Mrays|rror" >> gpurun_out/f_wfblock.log
  done
done
cat gpurun_out/f_wfblock.log
