#!/bin/bash
# A/B of the cache-hint variants (tools/build_variant.sh h_*): same process flow as tools/tune.py, one process per (variant, scene, renderer)
mkdir -p gpurun_out; : > gpurun_out/ii_hints.log
run() { # variant workload renderer spp
  RT_LIB_PATH=$PWD/variants/librt_$1.so timeout 300 python tools/tune.py --workload $2 --renderer $3 --spp $4 --frames 5 2>&1 | grep -E "Mrays|rror" | sed "s/^/$1 /; s/1920x1080 //; s/depth=10 //" >> gpurun_out/ii_hints.log
}
for wl in c3_sponza_scale c2_cornell; do
  for r in megakernel wavefront; do
    for v in h_base h_tri1 h_tri2 h_trish1 h_node1 h_all; do run $v $wl $r 64; done
  done
done
for v in h_base h_trish1 h_all; do run $v c4_heightfield_10m megakernel 16; done
cat gpurun_out/ii_hints.log
