#!/bin/bash
# round 2, call EE: sample parts in the megakernel: parity, then A/B (parts 1 / 2 / 3) on C3, C2, C4 and the 8-way tile shards
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ee_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ee_pytest.log; tail -4 gpurun_out/ee_pytest.log
: > gpurun_out/ee_parts.log
for wl in c3_sponza_scale c2_cornell c4_heightfield_10m; do
  s=64; [ $wl = c4_heightfield_10m ] && s=16
  timeout 600 python tools/tune.py --workload $wl --spp $s --frames 5 --configs "RT_SAMPLE_PARTS=1;RT_SAMPLE_PARTS=2;RT_SAMPLE_PARTS=3" 2>&1 | grep -E "Mrays|rror" | sed 's/1920x1080 //; s/3840x2160 //; s/depth=10 //' >> gpurun_out/ee_parts.log
done
timeout 600 python tools/tune.py --workload c3_sponza_scale --spp 256 --frames 3 --configs "RT_SAMPLE_PARTS=1;RT_SAMPLE_PARTS=3" 2>&1 | grep -E "Mrays|rror" | sed 's/1920x1080 //; s/depth=10 //' >> gpurun_out/ee_parts.log
cat gpurun_out/ee_parts.log
for c in 1 3; do echo "== tile shards, parts $c"; RT_SAMPLE_PARTS=$c timeout 600 python tools/tile_probe.py c4_heightfield_10m 8 64 megakernel 2>&1 | tail -4; done | tee gpurun_out/ee_tile.log
