#!/bin/bash
# round 2, call V (2 GPUs): GPU test suite incl. the rt_group tests on real peers, bench.py under torchrun (strong spp; C4 tiles)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/v_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -rs > gpurun_out/v_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest_2gpu.log; tail -6 gpurun_out/v_pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/v_n2_strong.json 2> gpurun_out/v_n2_strong.err; echo "strong rc=$?"; tail -2 gpurun_out/v_n2_strong.err
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --workload c4_heightfield_10m --no-e2e > gpurun_out/v_n2_c4.json 2> gpurun_out/v_n2_c4.err; echo "c4 rc=$?"; tail -2 gpurun_out/v_n2_c4.err
python - <<'PY'
import json
for f in ("v_n2_strong","v_n2_c4"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, {k:d[k] for k in ("value","ms_per_step","scaling","gpu_launches")}, d["config"]["sharding"][:60])
        print("   ", {k:(v["mrays_per_s"], v.get("verified")) for k,v in d["renderers"].items()}, d.get("also_weak"), d["e2e"] and (d["e2e"]["value"], d["e2e"]["scene_upload_and_build_ms_per_step"], d["e2e"]["render_call_ms_per_step"], d["e2e"]["render_device_ms_per_step"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
