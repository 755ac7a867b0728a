#!/bin/bash
# build_variant.sh NAME -DFLAG[=v] ...  ->  tools/probes/variants/libgt16_NAME.so : libgt16 with gt16_kernels.cu compiled with extra
# defines (experiment knobs such as GT16_WAIT_NS); run with GT16_LIB=tools/probes/variants/libgt16_NAME.so python bench.py ...
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
out=tools/probes/variants
mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -ccbin /usr/bin/g++ -Xcompiler -fPIC,-O2 "$@" -x cu -c 16gt_b200/csrc/gt16_kernels.cu -o $out/k_$name.o
nvcc -shared -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a $out/k_$name.o 16gt_b200/build/gt16_capi.o 16gt_b200/build/gt16_host_io.o \
  -o $out/libgt16_$name.so -cudart static -lpthread -ldl -lrt
rm -f $out/k_$name.o
echo built $out/libgt16_$name.so
