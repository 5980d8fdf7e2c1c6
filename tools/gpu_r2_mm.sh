#!/bin/bash
# node-stack top in registers (RT_STACK_TOP) and 72 registers / 7 CTAs for the megakernel (RT_MEGA_MIN_BLOCKS=7)
mkdir -p gpurun_out; : > gpurun_out/mm_stacktop.log
run() { RT_LIB_PATH=$PWD/variants/librt_$1.so timeout 300 python tools/tune.py --workload $2 --renderer $3 --spp $4 --frames 5 2>&1 | grep -E "Mrays|rror" | sed "s/^/$1 /; s/1920x1080 //; s/depth=10 //" >> gpurun_out/mm_stacktop.log; }
for v in t_base t_top t_mb7 t_top_mb7; do run $v c3_sponza_scale megakernel 64; done
for v in t_base t_top; do run $v c3_sponza_scale wavefront 64; done
for v in t_base t_top t_mb7 t_top_mb7; do run $v c2_cornell megakernel 64; done
for v in t_base t_top t_mb7 t_top_mb7; do run $v c4_heightfield_10m megakernel 16; done
for v in t_base t_top; do run $v c4_heightfield_10m wavefront 16; done
cat gpurun_out/mm_stacktop.log
