#!/bin/bash
# round 2, call U: early-outs + approximate box reciprocal: parity, A/B; re-tune of the scheduling knobs in the issue-bound regime
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/u_pytest.log; tail -4 gpurun_out/u_pytest.log
bash tools/gpu_ab_variants.sh "slimidp3 default"
unset RT_LIB_PATH
: > gpurun_out/u_tune.log
C="RT_TUNE_REFILL=8;RT_TUNE_REFILL=10;RT_TUNE_REFILL=12;RT_TUNE_REFILL=14;RT_TUNE_REFILL=16;RT_MEGA_CTX=1,RT_TUNE_REFILL=8;RT_MEGA_CTX=1,RT_TUNE_REFILL=12;RT_MEGA_CTX=2,RT_TUNE_REFILL=4;RT_MEGA_CTX=2,RT_TUNE_REFILL=8;RT_MEGA_CTX=3,RT_TUNE_REFILL=4;RT_MEGA_CTX=3,RT_TUNE_REFILL=8;RT_MEGA_CTX=2,RT_TUNE_REFILL=8,RT_TUNE_SHADE=16"
timeout 600 python tools/tune.py --workload c3_sponza_scale --spp 64 --frames 3 --configs "$C" 2>&1 | grep -E "Mrays|rror" | sed 's/1920x1080 spp=64 depth=10 //' >> gpurun_out/u_tune.log
W="RT_TUNE_INFLIGHT=32;RT_TUNE_INFLIGHT=64;RT_TUNE_INFLIGHT=96;RT_TUNE_INFLIGHT=128;RT_TUNE_INFLIGHT=64,RT_TUNE_REFILL=10;RT_TUNE_INFLIGHT=64,RT_TUNE_REFILL=12;RT_TUNE_INFLIGHT=64,RT_TUNE_REFILL=16;RT_TUNE_INFLIGHT=64,RT_TUNE_REFILL=18;RT_TUNE_INFLIGHT=96,RT_TUNE_REFILL=16"
timeout 600 python tools/tune.py --workload c3_sponza_scale --renderer wavefront --spp 64 --frames 3 --configs "$W" 2>&1 | grep -E "Mrays|rror" | sed 's/1920x1080 spp=64 depth=10 //' >> gpurun_out/u_tune.log
cat gpurun_out/u_tune.log
