#!/bin/bash
# round 2, call T: variants on top of the slim state (C3, C2), then ncu --set full of both render kernels on C3
mkdir -p gpurun_out
bash tools/gpu_ab_variants.sh "default s_ldg256 s_idp7 s_idp0 s_idp5 s_smtri s_mb6 s_f2" "c3_sponza_scale c2_cornell"
unset RT_LIB_PATH
prof() { # workload renderer spp kernel-regex tag
  python tools/profile_run.py --workload $1 --renderer $2 --spp $3 --frames 2 > gpurun_out/t_plain_$5.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$4 -s 1 -c 1 -o gpurun_out/r02b_$5 -f python tools/profile_run.py --workload $1 --renderer $2 --spp $3 --frames 2 > gpurun_out/t_ncu_$5.log 2>&1
  tail -2 gpurun_out/t_plain_$5.log | head -1
}
prof c3_sponza_scale megakernel 32 k_megakernel mega_c3
prof c3_sponza_scale wavefront 32 k_wf_flow flow_c3
