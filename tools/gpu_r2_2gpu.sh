#!/bin/bash
# two GPUs at the final build: the multi-GPU tests and the driver's strong-scaling line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_parity.py -m gpu -q -k "group or peer or shard or tile or spp" > gpurun_out/y_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/y_pytest_2gpu.log; tail -3 gpurun_out/y_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/y_bench_n2.json 2> gpurun_out/y_bench_n2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/y_bench_n2.json"))
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "scaling", "gpu_launches")}, "e2e", d.get("e2e") and round(d["e2e"]["value"], 1), d.get("config", {}).get("sharding"), d.get("verified"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/y_bench_n2.err").read()[-2000:])
PY
