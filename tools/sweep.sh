#!/bin/bash
# development sweep of the refill threshold of the persistent kernels (run on the GPU box)
for r in ${REFILLS:-8 10 12 14 16}; do
  echo -n "refill=$r  "; RT_TUNE_REFILL=$r python tools/profile_run.py --renderer ${1:-megakernel} --spp ${2:-16} --frames 2 | grep Mrays | tail -1
done
