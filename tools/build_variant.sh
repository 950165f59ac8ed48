#!/bin/bash
# Builds a variant of libhm_b200.so with extra -D flags into variants/libhm_b200_<name>.so (selected at run time with
# HM_B200_LIB=<path>): experiment builds and the checked build.   usage: tools/build_variant.sh <name> [-DFLAG ...]
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/hypermapper_b200/csrc
out=$root/variants
mkdir -p $out/obj_$name
for f in hm_core hm_forest hm_acq hm_gp hm_pareto hm_space hm_host; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --fmad=false "$@" \
       -c $src/$f.cu -o $out/obj_$name/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libhm_b200_$name.so $out/obj_$name/*.o -lcudart
rm -rf $out/obj_$name
echo $out/libhm_b200_$name.so
