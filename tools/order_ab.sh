#!/bin/bash
# A/B of the megakernel's cost-ordered block hand-out (RT_BLOCK_ORDER=0 disables) on the GPU box
for o in 1 0; do
  export RT_BLOCK_ORDER=$o
  echo "== RT_BLOCK_ORDER=$o"
  python tools/tile_probe.py c4_heightfield_10m 8 64 2>&1 | grep -E "unsharded|max "
  for wl in c3_sponza_scale c2_cornell; do python tools/profile_run.py --workload $wl --spp 64 --frames 3 | grep Mrays | tail -1; done
done
