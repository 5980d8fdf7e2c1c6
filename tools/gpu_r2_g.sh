#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
tail -4 gpurun_out/g_pytest.log
: > gpurun_out/g_wf.log
for wl in c3_sponza_scale c2_cornell c4_heightfield_10m; do
  timeout 600 python tools/tune.py --workload $wl --renderer wavefront --frames 3 --configs "RT_WF_PERSIST=1;RT_WF_PERSIST=2;RT_WF_PERSIST=2,RT_TUNE_REFILL=8;RT_WF_PERSIST=2,RT_TUNE_REFILL=16" 2>&1 | grep -E "Mrays|rror" >> gpurun_out/g_wf.log
done
cat gpurun_out/g_wf.log
