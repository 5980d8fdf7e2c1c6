#!/bin/bash
# round 2, call B: ncu --set full of three megakernel variants on C3 (8 spp)
mkdir -p gpurun_out
run() { # tag env...
  tag=$1; shift
  env "$@" python tools/profile_run.py --spp 8 --frames 2 > gpurun_out/b_plain_$tag.log 2>&1 &&
  env "$@" ncu --set full --clock-control none --import-source on -k regex:k_megakernel -s 1 -c 1 -o gpurun_out/b_$tag -f python tools/profile_run.py --spp 8 --frames 2 > gpurun_out/b_ncu_$tag.log 2>&1
  tail -2 gpurun_out/b_plain_$tag.log
}
run ctx0 RT_MEGA_CTX=0
run ctx1 RT_MEGA_CTX=1 RT_TUNE_REFILL=12
run ctx2 RT_MEGA_CTX=2 RT_TUNE_REFILL=8
ls -la gpurun_out/b_*.ncu-rep
