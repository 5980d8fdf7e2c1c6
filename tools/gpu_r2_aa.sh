#!/bin/bash
mkdir -p gpurun_out; log=gpurun_out/aa_mega_mask.log; : > $log
for wl in c3_sponza_scale c2_cornell c4_heightfield_10m; do
  s=64; [ $wl = c4_heightfield_10m ] && s=16
  for v in default mega7 mega5 mega6 default mega7; do
    if [ $v = default ]; then unset RT_LIB_PATH; else export RT_LIB_PATH=$PWD/variants/librt_$v.so; fi
    echo -n "$v: " >> $log
    timeout 300 python tools/tune.py --workload $wl --renderer megakernel --spp $s --frames 5 2>&1 | grep -E "Mrays|rror" | sed 's/1920x1080 //; s/3840x2160 //; s/depth=10 \[defaults\]//' >> $log
  done
done
cat $log
