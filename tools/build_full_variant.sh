#!/bin/bash
# tools/build_full_variant.sh NAME "-DSOME_FLAG=1 ..."  ->  hdiscontinuousgalerkin.jl_b200/variants/lib_NAME.so
# Rebuilds EVERY translation unit with the flags (needed whenever a flag changes a shared header, e.g. hdg_context or the
# constant tables); tools/build_variant.sh only rebuilds the element kernels.  Use with HDG_B200_LIB=<that .so>.
set -e
cd "$(dirname "$0")/../hdiscontinuousgalerkin.jl_b200/csrc"
out=../variants/full_$1
mkdir -p $out
for f in hdg_api hdg_mesh hdg_element hdg_solve hdg_mg hdg_mgx hdg_cg hdg_recover hdg_comm hdg_order; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $2 -c $f.cu -o $out/$f.o 2> $out/$f.log &
done
g++ -O2 -std=c++17 -fPIC -c hdg_tables.cpp -o $out/hdg_tables.o
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_$1.so $out/*.o
rm -rf $out
echo "built hdiscontinuousgalerkin.jl_b200/variants/lib_$1.so"
