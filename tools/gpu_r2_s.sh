#!/bin/bash
# round 2, call S: slim traversal state / compact pixel state: parity, then A/B against the previous build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log; tail -4 gpurun_out/s_pytest.log
bash tools/gpu_ab_variants.sh "base slim slimidp3"
