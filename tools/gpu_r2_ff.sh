#!/bin/bash
mkdir -p gpurun_out
RT_TRACE=1 timeout 900 python bench.py --workload c4_heightfield_10m --steps 5 --no-cpu-baseline > gpurun_out/ff_c4.json 2> gpurun_out/ff_c4.err; echo "c4 rc=$?"
grep "rt trace" gpurun_out/ff_c4.err | tail -16
python -c "
import json; d=json.load(open('gpurun_out/ff_c4.json')); print(d['e2e'])"
free -g | head -2; nproc; cat /sys/kernel/mm/transparent_hugepage/enabled 2>/dev/null; numactl -H 2>/dev/null | head -5
