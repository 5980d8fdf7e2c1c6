#!/bin/bash
# development build of librt_b200.so with extra nvcc flags: tools/build_variant.sh <name> [flags...] -> variants/librt_<name>.so
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/sycl-ray-tracer_b200/csrc
out=/tmp/rt_variant_$name
mkdir -p $out $root/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v"
for f in rt_api render bvh_build rt_group; do
  /usr/local/cuda/bin/nvcc $FLAGS "$@" -c $src/$f.cu -o $out/$f.o 2> $out/$f.log &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/variants/librt_$name.so $out/rt_api.o $out/render.o $out/bvh_build.o $out/rt_group.o
grep -A2 "k_megakernel" $out/render.log | grep -E "Used|spill" | tr '\n' ' '; echo " -> variants/librt_$name.so"
