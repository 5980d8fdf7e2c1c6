#!/bin/bash
# registers / resident CTAs of the two render kernels, two node steps per vote
mkdir -p gpurun_out; : > gpurun_out/nn_regs.log
run() { RT_LIB_PATH=$PWD/variants/librt_$1.so timeout 300 python tools/tune.py --workload $2 --renderer $3 --spp $4 --frames 5 2>&1 | grep -E "Mrays|rror" | sed "s/^/$1 /; s/1920x1080 //; s/depth=10 //" >> gpurun_out/nn_regs.log; }
for v in r_mb7 r_mb6 r_unroll2; do run $v c3_sponza_scale megakernel 64; done
for v in r_mb7 r_w128x7 r_w128x6 r_w256x3 r_unroll2; do run $v c3_sponza_scale wavefront 64; done
for v in r_mb7 r_mb6 r_unroll2; do run $v c2_cornell megakernel 64; done
for v in r_mb7 r_w128x7 r_w128x6 r_w256x3; do run $v c2_cornell wavefront 64; done
for v in r_mb7 r_mb6; do run $v c4_heightfield_10m megakernel 16; done
for v in r_mb7 r_w128x7 r_w128x6 r_w256x3; do run $v c4_heightfield_10m wavefront 16; done
cat gpurun_out/nn_regs.log
