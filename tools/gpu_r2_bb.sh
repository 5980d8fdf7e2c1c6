#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/probe_h2d.py 2>&1 | grep -v "^$" | tee gpurun_out/bb_probe_h2d.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/bb_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/bb_pytest.log; tail -3 gpurun_out/bb_pytest.log
