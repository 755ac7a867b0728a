#!/bin/bash
# build_variant.sh NAME -DFLAG[=v] ...  ->  tools/probes/variants/libgt16_NAME.so : libgt16 with gt16_kernels.cu and gt16_capi.cu
# compiled with extra defines (experiment knobs such as GT16_STREAM_ROWS); KSRC=file compiles another kernel source instead
# (e.g. an older revision, with its gt16_kernels.cuh beside it).  Run with GT16_LIB=tools/probes/variants/libgt16_NAME.so python bench.py ...
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
out=tools/probes/variants
mkdir -p $out
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-O2"
nvcc $F "$@" -x cu -c ${KSRC:-16gt_b200/csrc/gt16_kernels.cu} -I 16gt_b200/csrc -o $out/k_$name.o
nvcc $F "$@" -x cu -c ${CSRC:-16gt_b200/csrc/gt16_capi.cu} -I 16gt_b200/csrc -o $out/c_$name.o
nvcc -shared -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a $out/k_$name.o $out/c_$name.o 16gt_b200/build/gt16_host_io.o 16gt_b200/build/gt16_pileup.o 16gt_b200/build/gt16_sam.o \
  -o $out/libgt16_$name.so -cudart static -lpthread -ldl -lrt -lz
rm -f $out/k_$name.o $out/c_$name.o
echo built $out/libgt16_$name.so
