#!/bin/bash
# development sweep of the scheduling knobs (run on the GPU box)
for r in ${REFILLS:-4 8 12 16}; do for t in ${TRIDIVS:-100 8 5 3 2}; do
  echo -n "refill=$r tridiv=$t  "; RT_TUNE_REFILL=$r RT_TUNE_TRIDIV=$t python tools/profile_run.py --renderer ${1:-megakernel} --spp ${2:-16} --frames 2 | grep Mrays | tail -1
done; done
