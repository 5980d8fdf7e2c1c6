#!/bin/bash
# round 2, call W: full GPU suite with the widened parity tests (1024 spp on C2/C3, true 10 M C4, C5 4K progressive, stadium),
# size-class key bit A/B, stadium throughput with / without the size class and the optional split
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/w_pytest.log; tail -14 gpurun_out/w_pytest.log
bash tools/gpu_ab_variants.sh "nocls default" "stadium c2_cornell c3_sponza_scale"
unset RT_LIB_PATH
echo "== stadium with RT_SPLIT=1" >> gpurun_out/ab_nocls_default.log
for r in megakernel wavefront; do RT_SPLIT=1 timeout 300 python tools/tune.py --workload stadium --renderer $r --frames 3 2>&1 | grep -E "Mrays|rror" >> gpurun_out/ab_nocls_default.log; done
tail -3 gpurun_out/ab_nocls_default.log
