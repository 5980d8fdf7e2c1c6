#!/bin/bash
# round 2, call E: ncu of the persistent wavefront kernel on C2 and C3 (8 spp)
mkdir -p gpurun_out
for wl in c2_cornell c3_sponza_scale; do
  python tools/profile_run.py --workload $wl --renderer wavefront --spp 8 --frames 2 > gpurun_out/e_plain_$wl.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_wf_persistent -s 1 -c 1 -o gpurun_out/e_wfp_$wl -f python tools/profile_run.py --workload $wl --renderer wavefront --spp 8 --frames 2 > gpurun_out/e_ncu_$wl.log 2>&1
  tail -2 gpurun_out/e_plain_$wl.log
done
