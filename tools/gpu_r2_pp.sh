#!/bin/bash
# after 72 registers + two node steps per vote: refill x carry again, three steps per vote
mkdir -p gpurun_out; : > gpurun_out/pp_retune.log
C=";RT_TUNE_REFILL=14,RT_TUNE_CARRY=2;RT_TUNE_REFILL=18,RT_TUNE_CARRY=2;RT_TUNE_REFILL=20,RT_TUNE_CARRY=2;RT_TUNE_REFILL=16,RT_TUNE_CARRY=1;RT_TUNE_REFILL=16,RT_TUNE_CARRY=3;RT_TUNE_REFILL=18,RT_TUNE_CARRY=3;RT_TUNE_REFILL=20,RT_TUNE_CARRY=4"
for wl in c3_sponza_scale c2_cornell; do
  timeout 600 python tools/tune.py --workload $wl --renderer megakernel --spp 64 --frames 5 --configs "$C" 2>&1 | grep -E "Mrays|rror" | sed 's/1920x1080 //; s/depth=10 //' >> gpurun_out/pp_retune.log
done
timeout 600 python tools/tune.py --workload c4_heightfield_10m --renderer megakernel --spp 16 --frames 5 --configs "$C" 2>&1 | grep -E "Mrays|rror" >> gpurun_out/pp_retune.log
for wl in c3_sponza_scale c2_cornell; do
  RT_LIB_PATH=$PWD/variants/librt_u3.so timeout 300 python tools/tune.py --workload $wl --renderer megakernel --spp 64 --frames 5 2>&1 | grep -E "Mrays|rror" | sed "s/^/u3 /; s/1920x1080 //; s/depth=10 //" >> gpurun_out/pp_retune.log
done
RT_LIB_PATH=$PWD/variants/librt_u3.so timeout 300 python tools/tune.py --workload c4_heightfield_10m --renderer megakernel --spp 16 --frames 5 2>&1 | grep -E "Mrays|rror" | sed "s/^/u3 /" >> gpurun_out/pp_retune.log
cat gpurun_out/pp_retune.log
