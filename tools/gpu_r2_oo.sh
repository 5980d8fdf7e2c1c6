#!/bin/bash
# 72-register megakernel with two node steps per vote (default) against one step (variants/librt_u1.so); then tests and the bench lines
mkdir -p gpurun_out; : > gpurun_out/oo_unroll.log
run() { if [ $1 = default ]; then unset RT_LIB_PATH; else export RT_LIB_PATH=$PWD/variants/librt_$1.so; fi
  timeout 300 python tools/tune.py --workload $2 --renderer $3 --spp $4 --frames 5 2>&1 | grep -E "Mrays|rror" | sed "s/^/$1 /; s/1920x1080 //; s/depth=10 //" >> gpurun_out/oo_unroll.log; unset RT_LIB_PATH; }
for wl in c3_sponza_scale c2_cornell stadium; do for v in default u1; do run $v $wl megakernel 64; done; done
for v in default u1; do run $v c4_heightfield_10m megakernel 16; done
cat gpurun_out/oo_unroll.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest.log; tail -3 gpurun_out/o_pytest.log
timeout 900 python bench.py > gpurun_out/o_bench_n1.json 2> gpurun_out/o_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("o_bench_n1",):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, {k: round(v["mrays_per_s"], 1) for k, v in d.get("renderers", {}).items()},
              "e2e", d.get("e2e") and round(d["e2e"]["value"], 1), "frac", d.get("roofline") and round(d["roofline"]["frac"], 3), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"], 2))
    except Exception as e:
        print(f, "FAILED", e)
PY
