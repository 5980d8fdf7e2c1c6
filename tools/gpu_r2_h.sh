#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -4 gpurun_out/h_pytest.log
: > gpurun_out/h_wf.log
for wl in c3_sponza_scale c2_cornell c4_heightfield_10m; do
  timeout 600 python tools/tune.py --workload $wl --renderer wavefront --frames 3 --configs "RT_WF_PERSIST=1;RT_WF_PERSIST=2,RT_TUNE_INFLIGHT=64;RT_WF_PERSIST=2,RT_TUNE_INFLIGHT=128;RT_WF_PERSIST=2,RT_TUNE_INFLIGHT=256;RT_WF_PERSIST=2,RT_TUNE_INFLIGHT=512;RT_WF_PERSIST=2,RT_TUNE_INFLIGHT=128,RT_TUNE_REFILL=16;RT_WF_PERSIST=2,RT_TUNE_INFLIGHT=256,RT_TUNE_REFILL=16" 2>&1 | grep -E "Mrays|rror" >> gpurun_out/h_wf.log
done
cat gpurun_out/h_wf.log
