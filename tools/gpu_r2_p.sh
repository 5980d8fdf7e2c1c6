#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p_pytest.log; tail -5 gpurun_out/p_pytest.log
timeout 600 python tools/tune.py --workload c3_sponza_scale --renderer wavefront --spp 128 --depth 50 --frames 3 --configs "RT_BLOCK_ORDER=1;RT_BLOCK_ORDER=0" 2>&1 | grep -E "Mrays|rror"
timeout 600 python tools/tune.py --workload c3_sponza_scale --renderer wavefront --spp 128 --frames 3 --configs "RT_BLOCK_ORDER=1;RT_BLOCK_ORDER=0" 2>&1 | grep -E "Mrays|rror"
timeout 600 python tools/tune.py --workload c2_cornell --renderer wavefront --frames 3 --configs "RT_BLOCK_ORDER=1;RT_BLOCK_ORDER=0" 2>&1 | grep -E "Mrays|rror"
