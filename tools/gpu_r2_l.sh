#!/bin/bash
# round 2, call L: bench.py end to end (quick, then the default run), on 1 GPU
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --spp 8 --cpu-seconds 1 > gpurun_out/l_quick.json 2> gpurun_out/l_quick.err; echo "quick rc=$?"; tail -3 gpurun_out/l_quick.err; head -c 600 gpurun_out/l_quick.json; echo
timeout 1200 python bench.py > gpurun_out/l_bench_n1.json 2> gpurun_out/l_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/l_bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/l_bench_n1.json"))
print({k:d[k] for k in ("value","ms_per_step","msamples_per_s","gpu_launches")})
print(d["renderers"]); print(d["also"]); print(d["roofline"]); print(d["cpu_baseline"]); print(d["e2e"]); print(d["clocks"])
PY
