#!/bin/bash
# round 2, call Q: pipe microbenchmark, GPU test suite at the restored HEAD, default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/q_smi.log 2>&1
timeout 300 tools/microbench/pipe_bench > gpurun_out/q_pipe_bench.log 2>&1; cat gpurun_out/q_pipe_bench.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log; tail -4 gpurun_out/q_pytest.log
timeout 900 python bench.py > gpurun_out/q_bench_n1.json 2> gpurun_out/q_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/q_bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/q_bench_n1.json"))
print({k:d.get(k) for k in ("value","ms_per_step","msamples_per_s","gpu_launches")})
print(d.get("renderers")); print(d.get("roofline")); print(d.get("cpu_baseline")); print(d.get("e2e")); print(d.get("clocks"))
PY
