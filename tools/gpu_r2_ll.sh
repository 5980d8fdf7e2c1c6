#!/bin/bash
# new defaults (refill 16 / carry 2 megakernel, 12 / 3 wavefront; texel bytes through rt_u8_to_unit; shared dominant-axis reciprocal): tests, defaults vs the old knobs, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log; tail -4 gpurun_out/l_pytest.log
: > gpurun_out/ll_defaults.log
C=";RT_TUNE_REFILL=14,RT_TUNE_CARRY=0"
for wl in c3_sponza_scale c2_cornell stadium; do
  for r in megakernel wavefront; do
    timeout 600 python tools/tune.py --workload $wl --renderer $r --spp 64 --frames 5 --configs "$C" 2>&1 | grep -E "Mrays|rror" | sed 's/1920x1080 //; s/depth=10 //' >> gpurun_out/ll_defaults.log
  done
done
for r in megakernel wavefront; do
  timeout 600 python tools/tune.py --workload c4_heightfield_10m --renderer $r --spp 16 --frames 5 --configs "$C" 2>&1 | grep -E "Mrays|rror" >> gpurun_out/ll_defaults.log
done
cat gpurun_out/ll_defaults.log
timeout 900 python bench.py > gpurun_out/l_bench_n1.json 2> gpurun_out/l_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("l_bench_n1",):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, {k: round(v["mrays_per_s"], 1) for k, v in d.get("renderers", {}).items()},
              "e2e", d.get("e2e") and round(d["e2e"]["value"], 1), "frac", d.get("roofline") and round(d["roofline"]["frac"], 3), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"], 2))
    except Exception as e:
        print(f, "FAILED", e)
PY
