#!/bin/bash
# round 2, call O: full GPU test suite, ncu --set full of both render kernels on C2/C3/C4, the reference's benchmark protocol
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest.log; tail -3 gpurun_out/o_pytest.log
prof() { # workload renderer spp kernel-regex tag
  python tools/profile_run.py --workload $1 --renderer $2 --spp $3 --frames 2 > gpurun_out/o_plain_$5.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$4 -s 1 -c 1 -o gpurun_out/r02_$5 -f python tools/profile_run.py --workload $1 --renderer $2 --spp $3 --frames 2 > gpurun_out/o_ncu_$5.log 2>&1
  tail -2 gpurun_out/o_plain_$5.log | head -1
}
prof c3_sponza_scale megakernel 32 k_megakernel mega_c3
prof c3_sponza_scale wavefront 32 k_wf_flow flow_c3
prof c2_cornell megakernel 64 k_megakernel mega_c2
prof c2_cornell wavefront 64 k_wf_flow flow_c2
prof c4_heightfield_10m megakernel 16 k_megakernel mega_c4
prof c4_heightfield_10m wavefront 16 k_wf_flow flow_c4
timeout 1500 python tools/benchmark_protocol.py gpurun_out/r02_benchmark_protocol.csv 3 > gpurun_out/o_protocol.log 2>&1; tail -40 gpurun_out/o_protocol.log
