#!/bin/bash
# A/B of compile-time variants (variants/librt_<name>.so): tools/gpu_ab_variants.sh "<names>" ["<workloads>"] [spp]
mkdir -p gpurun_out; log=gpurun_out/ab_$(echo $1 | tr ' ' '_' | cut -c1-40).log; : > $log
wls=${2:-"c3_sponza_scale c2_cornell c4_heightfield_10m"}
for wl in $wls; do
  s=${3:-64}; [ $wl = c4_heightfield_10m ] && s=16
  for v in $1; do
    if [ $v = default ]; then unset RT_LIB_PATH; else export RT_LIB_PATH=$PWD/variants/librt_$v.so; fi
    for r in megakernel wavefront; do
      echo -n "$v: " >> $log
      timeout 300 python tools/tune.py --workload $wl --renderer $r --spp $s --frames 3 2>&1 | grep -E "Mrays|rror" | sed 's/1920x1080 //; s/3840x2160 //; s/depth=10 \[defaults\]//' >> $log
    done
  done
done
cat $log
