#!/bin/bash
# round 2, call X (8 GPUs): rt_group tests on every GPU, bench.py under torchrun: C3 strong (the driver's line) and C4 tiles
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/x_smi.log 2>&1
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -x -q -rs > gpurun_out/x_pytest_group_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x_pytest_group_8gpu.log; tail -4 gpurun_out/x_pytest_group_8gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/x_n8_strong.json 2> gpurun_out/x_n8_strong.err; echo "strong rc=$?"; tail -2 gpurun_out/x_n8_strong.err
timeout 900 $TR bench.py --gpus 8 --steps 5 --warmup 3 --workload c4_heightfield_10m --no-e2e > gpurun_out/x_n8_c4.json 2> gpurun_out/x_n8_c4.err; echo "c4 rc=$?"; tail -2 gpurun_out/x_n8_c4.err
python - <<'PY'
import json
for f in ("x_n8_strong","x_n8_c4"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, {k:d[k] for k in ("value","ms_per_step","scaling","gpu_launches")}, d["config"]["sharding"][:60])
        print("   ", {k:(v["mrays_per_s"], v["ms_per_step"], v.get("verified")) for k,v in d["renderers"].items()}, d.get("also_weak"), d["e2e"] and (d["e2e"]["value"], d["e2e"]["scene_upload_and_build_ms_per_step"], d["e2e"]["render_call_ms_per_step"], d["e2e"]["render_device_ms_per_step"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
