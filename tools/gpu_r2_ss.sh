#!/bin/bash
# last check of round 2: the knob-invariance test at the final build, and the vote-after-every-step-near-the-threshold variants
mkdir -p gpurun_out; : > gpurun_out/ss_near.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "knobs or handout or determinism" > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log; tail -3 gpurun_out/s_pytest.log
run() { if [ $1 = default ]; then unset RT_LIB_PATH; else export RT_LIB_PATH=$PWD/variants/librt_$1.so; fi
  timeout 300 python tools/tune.py --workload $2 --renderer $3 --spp $4 --frames 5 2>&1 | grep -E "Mrays|rror" | sed "s/^/$1 /; s/1920x1080 //; s/depth=10 //" >> gpurun_out/ss_near.log; unset RT_LIB_PATH; }
for wl in c3_sponza_scale c2_cornell; do for v in default near3 near6; do run $v $wl megakernel 64; done; done
cat gpurun_out/ss_near.log
