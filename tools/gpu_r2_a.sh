#!/bin/bash
# round 2, call A: parity + first A/B of the K-context megakernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -3 gpurun_out/a_pytest.log
C="RT_MEGA_CTX=0;RT_MEGA_CTX=1,RT_TUNE_REFILL=12;RT_MEGA_CTX=1,RT_TUNE_REFILL=8;RT_MEGA_CTX=2,RT_TUNE_REFILL=2;RT_MEGA_CTX=2,RT_TUNE_REFILL=4;RT_MEGA_CTX=2,RT_TUNE_REFILL=8;RT_MEGA_CTX=2,RT_TUNE_REFILL=4,RT_TUNE_SHADE=16;RT_MEGA_CTX=2,RT_TUNE_REFILL=4,RT_TUNE_SHADE=28,RT_TUNE_IDLE=8;RT_MEGA_CTX=3,RT_TUNE_REFILL=2;RT_MEGA_CTX=3,RT_TUNE_REFILL=4;RT_MEGA_CTX=3,RT_TUNE_REFILL=8;RT_MEGA_CTX=4,RT_TUNE_REFILL=4"
timeout 900 python tools/tune.py --workload c3_sponza_scale --spp 64 --frames 3 --configs "$C" > gpurun_out/a_tune_c3.log 2>&1
RT_LIB_PATH=$PWD/variants/librt_r01.so timeout 300 python tools/tune.py --workload c3_sponza_scale --spp 64 --frames 3 >> gpurun_out/a_tune_c3.log 2>&1
cat gpurun_out/a_tune_c3.log
C2="RT_MEGA_CTX=0;RT_MEGA_CTX=1,RT_TUNE_REFILL=8;RT_MEGA_CTX=2,RT_TUNE_REFILL=4;RT_MEGA_CTX=3,RT_TUNE_REFILL=4"
timeout 600 python tools/tune.py --workload c2_cornell --frames 3 --configs "$C2" > gpurun_out/a_tune_c2.log 2>&1
RT_LIB_PATH=$PWD/variants/librt_r01.so timeout 300 python tools/tune.py --workload c2_cornell --frames 3 >> gpurun_out/a_tune_c2.log 2>&1
cat gpurun_out/a_tune_c2.log
timeout 900 python tools/tune.py --workload c4_heightfield_10m --frames 3 --configs "$C2" > gpurun_out/a_tune_c4.log 2>&1
RT_LIB_PATH=$PWD/variants/librt_r01.so timeout 300 python tools/tune.py --workload c4_heightfield_10m --frames 3 >> gpurun_out/a_tune_c4.log 2>&1
cat gpurun_out/a_tune_c4.log
# wavefront with the new node format vs round 1
for wl in c3_sponza_scale c2_cornell; do
  timeout 300 python tools/tune.py --workload $wl --renderer wavefront --spp 64 --frames 3 > gpurun_out/a_wf_$wl.log 2>&1
  RT_LIB_PATH=$PWD/variants/librt_r01.so timeout 300 python tools/tune.py --workload $wl --renderer wavefront --spp 64 --frames 3 >> gpurun_out/a_wf_$wl.log 2>&1
  cat gpurun_out/a_wf_$wl.log
done
