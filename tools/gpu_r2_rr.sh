#!/bin/bash
# queue-driven wavefront: 1 (default) / 2 / 3 rolled node steps per vote
mkdir -p gpurun_out; : > gpurun_out/rr_wf_rolled.log
run() { if [ $1 = default ]; then unset RT_LIB_PATH; else export RT_LIB_PATH=$PWD/variants/librt_$1.so; fi
  timeout 300 python tools/tune.py --workload $2 --renderer $3 --spp $4 --frames 5 2>&1 | grep -E "Mrays|rror" | sed "s/^/$1 /; s/1920x1080 //; s/depth=10 //" >> gpurun_out/rr_wf_rolled.log; unset RT_LIB_PATH; }
for wl in c3_sponza_scale c2_cornell; do for v in default w2 w3; do run $v $wl wavefront 64; done; done
for v in default w2 w3; do run $v c4_heightfield_10m wavefront 16; done
run default c3_sponza_scale megakernel 64
cat gpurun_out/rr_wf_rolled.log
