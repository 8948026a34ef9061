#!/bin/bash
# tools/build_variant.sh NAME "-DQCB3=1 ..."  ->  hdiscontinuousgalerkin.jl_b200/variants/lib_NAME.so (element kernels rebuilt with the flags)
set -e
cd "$(dirname "$0")/../hdiscontinuousgalerkin.jl_b200/csrc"
mkdir -p ../variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v $2 -c hdg_element.cu -o ../variants/elem_$1.o 2> ../variants/elem_$1.log
grep -A2 "element_quad_kernel" ../variants/elem_$1.log | grep -E "registers|spill" || true
OBJS=$(ls ../build/*.o | grep -v hdg_element.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_$1.so $OBJS ../variants/elem_$1.o
rm -f ../variants/elem_$1.o
