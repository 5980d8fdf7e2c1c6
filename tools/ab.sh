#!/bin/bash
# A/B of library variants on the GPU box: tools/ab.sh <spp> <lib-or-"default"> ...   (RT_LIB_PATH selects the .so)
spp=$1; shift
for lib in "$@"; do
  for wl in c3_sponza_scale c2_cornell c4_heightfield_10m; do
    for rr in megakernel wavefront; do
      s=$spp; [ $wl = c4_heightfield_10m ] && s=$((spp / 4))
      if [ "$lib" = default ]; then unset RT_LIB_PATH; else export RT_LIB_PATH=$PWD/$lib; fi
      echo -n "$lib $wl $rr: "; python tools/profile_run.py --workload $wl --renderer $rr --spp $s --frames 3 | grep Mrays | awk '{print $(NF-3)}' | tr '\n' ' '; echo
    done
  done
done
