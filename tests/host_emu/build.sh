#!/bin/bash
# builds tests/host_emu/libgt16_emu.so (test-only host compilation of the per-site routines)
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
/usr/bin/g++ -O2 -std=c++17 -fPIC -shared -Wall -Wno-unused-function -ffp-contract=off \
  -I/usr/local/cuda/include "$HERE/gt16_emu.cpp" "$HERE/../../16gt_b200/csrc/gt16_host_io.cpp" \
  -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64 -lcudart -lm -o "$HERE/libgt16_emu.so"
