#!/bin/bash
mkdir -p gpurun_out
RT_TRACE=1 timeout 600 python bench.py --workload c4_heightfield_10m --steps 3 --no-cpu-baseline --no-also --renderer wavefront > gpurun_out/cc_c4.json 2> gpurun_out/cc_c4.err
grep "rt trace" gpurun_out/cc_c4.err | tail -8
python -c "
import json; d=json.load(open('gpurun_out/cc_c4.json')); print(d['e2e'])"
RT_TRACE=1 timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-also --renderer megakernel > gpurun_out/cc_c3.json 2> gpurun_out/cc_c3.err
grep "rt trace" gpurun_out/cc_c3.err | tail -5
python -c "
import json; d=json.load(open('gpurun_out/cc_c3.json')); print(d['e2e'])"
