#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=.. ..." : builds the-cooper-mapper_b200/libcoopermap_NAME.so with extra nvcc flags (kernel A/B tests:
# COOPERMAP_LIB=the-cooper-mapper_b200/libcoopermap_NAME.so python bench.py ...)
set -e
cd "$(dirname "$0")/../the-cooper-mapper_b200/csrc"
name=$1; shift
mkdir -p /tmp/cmvar_$name
objs=""
for f in cm_capi cm_match cm_debug cm_scanreg cm_voxel cm_map cm_mapping cm_odometry cm_mapio; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off --expt-relaxed-constexpr -Xptxas -v "$@" -c $f.cu -o /tmp/cmvar_$name/$f.o 2> /tmp/cmvar_$name/$f.log &
  objs="$objs /tmp/cmvar_$name/$f.o"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared --cudart static -o ../libcoopermap_$name.so $objs
grep -A2 "search_kernelILb0" /tmp/cmvar_$name/cm_match.log | grep -E "spill|Used"
