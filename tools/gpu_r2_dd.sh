#!/bin/bash
mkdir -p gpurun_out
for k in megakernel wavefront; do timeout 600 python tools/tail_fit.py c3_sponza_scale $k 2>&1 | grep -v "^$"; done | tee gpurun_out/dd_tail_fit_c3.log
RT_BLOCK_ORDER=0 timeout 600 python tools/tail_fit.py c3_sponza_scale megakernel 2>&1 | grep -v "^$" | sed 's/^/no block order: /' | tee -a gpurun_out/dd_tail_fit_c3.log
