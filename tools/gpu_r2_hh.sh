#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/hh_order.log
C="RT_ORDER_REGION=64;RT_ORDER_REGION=32;RT_ORDER_REGION=16;RT_ORDER_REGION=8;RT_ORDER_REGION=32,RT_ORDER_PROBES=2;RT_ORDER_REGION=16,RT_ORDER_PROBES=4;RT_ORDER_REGION=8,RT_ORDER_PROBES=4"
for wl in c3_sponza_scale c2_cornell stadium; do
  for r in megakernel wavefront; do
    timeout 600 python tools/tune.py --workload $wl --renderer $r --spp 64 --frames 5 --configs "$C" 2>&1 | grep -E "Mrays|rror" | sed 's/1920x1080 //; s/depth=10 //' >> gpurun_out/hh_order.log
  done
done
cat gpurun_out/hh_order.log
