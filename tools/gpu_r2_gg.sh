#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/gg_order.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for wl in c3_sponza_scale c2_cornell; do
  timeout 600 python tools/tune.py --workload $wl --spp 64 --frames 5 --configs "RT_BLOCK_ORDER=1;RT_BLOCK_ORDER=0;RT_BLOCK_ORDER=1,RT_SAMPLE_PARTS=2;RT_BLOCK_ORDER=0,RT_SAMPLE_PARTS=1" 2>&1 | grep -E "Mrays|rror" | sed 's/1920x1080 //; s/depth=10 //' >> gpurun_out/gg_order.log
done
timeout 600 python tools/tune.py --workload c3_sponza_scale --spp 256 --frames 3 --configs "RT_BLOCK_ORDER=1;RT_BLOCK_ORDER=0" 2>&1 | grep -E "Mrays|rror" | sed 's/1920x1080 //; s/depth=10 //' >> gpurun_out/gg_order.log
timeout 600 python tools/tune.py --workload c2_cornell --spp 256 --frames 3 --configs "RT_BLOCK_ORDER=1;RT_BLOCK_ORDER=0" 2>&1 | grep -E "Mrays|rror" | sed 's/1920x1080 //; s/depth=10 //' >> gpurun_out/gg_order.log
cat gpurun_out/gg_order.log
