#!/bin/bash
# node steps per vote as a rolled loop (one copy of the node test): 2, 3, 4 steps against the unrolled default
mkdir -p gpurun_out; : > gpurun_out/qq_rolled.log
run() { if [ $1 = default ]; then unset RT_LIB_PATH; else export RT_LIB_PATH=$PWD/variants/librt_$1.so; fi
  timeout 300 python tools/tune.py --workload $2 --renderer $3 --spp $4 --frames 5 2>&1 | grep -E "Mrays|rror" | sed "s/^/$1 /; s/1920x1080 //; s/depth=10 //" >> gpurun_out/qq_rolled.log; unset RT_LIB_PATH; }
for wl in c3_sponza_scale c2_cornell; do for v in default l2 l3 l4; do run $v $wl megakernel 64; done; done
for v in default l2 l3; do run $v c4_heightfield_10m megakernel 16; done
cat gpurun_out/qq_rolled.log
