#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Builds the UNMODIFIED reference implementation of the hot
# path from the sources where they lie under $REF (default /root/reference) into
# oracle/_ref/ (git-ignored; it travels to the GPU box with the snapshot).  Nothing
# from the reference is copied into the repository: the compiler reads the sources in
# place and only objects/binaries are written.
#
# One source defect has to be worked around: likelihood_cache_prefill() is declared
# `int` but falls off its end (likelihood_cache.cpp:36-54); g++ >= 8 at -O1+ turns that
# into unreachable code and the run dies right after the prefill.  That one translation
# unit is therefore compiled from a sed *stream* that inserts `return 1;` before the
# closing brace -- the patched text exists only in the pipe.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
mkdir -p "$OUT/obj" "$OUT/pic"
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF not present (GPU box?) -- using prebuilt files in $OUT" >&2
  exit 0
fi
CXX=${REF_CXX:-/usr/bin/g++}   # not $CXX: the image exports a g++ without libgomp.spec
# Makefile:3 flags + -no-pie (lib/lib.a objects are non-PIC)
CF="-O3 -funroll-loops -fomit-frame-pointer -maccumulate-outgoing-args -static-libgcc -mpopcnt -fopenmp -fpermissive -w -I$REF"

patched_lc() { sed '53a\  return 1;' "$REF/likelihood_cache.cpp"; }

for f in SNP SAMhandler FisherExactTest VariantCaller SnapshotHandler fisher ycsq SNP_Caller SNPFunctions \
         interpreter bam2snapshot CounterReader indexFunction readIndex buildRegionList; do
  [ "$OUT/obj/$f.o" -nt "$REF/$f.cpp" ] || $CXX $CF -c "$REF/$f.cpp" -o "$OUT/obj/$f.o"
done
patched_lc | $CXX $CF -x c++ -c - -o "$OUT/obj/likelihood_cache.o"

O=$OUT/obj
CALLER_OBJS="$O/SNP.o $O/SAMhandler.o $O/FisherExactTest.o $O/VariantCaller.o $O/SnapshotHandler.o $O/fisher.o $O/likelihood_cache.o $O/ycsq.o $O/SNP_Caller.o $O/SNPFunctions.o $O/interpreter.o"
LIBS="$REF/lib/lib.a -lpthread -lm -lz"

$CXX $CF -no-pie "$HERE/ref_drivers/ref_caller_main.cpp" $CALLER_OBJS $LIBS -o "$OUT/ref_caller"
install -m 644 "$REF/snapshotSnpcaller.ini" "$OUT/ref_caller.ini"          # config, not source
# the same driver with FisherExactTest.o replaced by the libgt16 shim: the
# reference's own loop / consumer thread / writer around the GPU operators.  Needs 16gt_b200/lib/libgt16.so.
GT16LIB=$HERE/../16gt_b200/lib
if [ -f "$GT16LIB/libgt16.so" ]; then
  SHIM_OBJS="$O/SNP.o $O/SAMhandler.o $O/VariantCaller.o $O/SnapshotHandler.o $O/fisher.o $O/likelihood_cache.o $O/ycsq.o $O/SNP_Caller.o $O/SNPFunctions.o $O/interpreter.o"
  $CXX $CF -no-pie -I"$HERE/../include" "$HERE/ref_drivers/ref_caller_main.cpp" "$HERE/ref_drivers/gt16_shim.cpp" $SHIM_OBJS $LIBS \
       -L"$GT16LIB" -lgt16 -Wl,-rpath,'$ORIGIN/../../16gt_b200/lib' -Wl,-rpath-link,/usr/local/cuda/lib64 -o "$OUT/ref_caller_gt16"
  install -m 644 "$REF/snapshotSnpcaller.ini" "$OUT/ref_caller_gt16.ini"
fi
$CXX $CF -no-pie "$HERE/ref_drivers/ref_mkindex_main.cpp" $LIBS -o "$OUT/ref_mkindex"
$CXX $CF -no-pie $O/bam2snapshot.o $O/SAMhandler.o $O/SNP.o $O/CounterReader.o $O/SnapshotHandler.o \
     $O/indexFunction.o $O/readIndex.o $LIBS -o "$OUT/ref_bam2snapshot"
install -m 644 "$REF/bam2snapshot.ini" "$OUT/ref_bam2snapshot.ini"
$CXX $CF $O/readIndex.o $O/indexFunction.o $O/buildRegionList.o -o "$OUT/ref_RegionIndexBuilder"

# operator-level shared object for ctypes (no lib.a needed for these TUs)
P=$OUT/pic
for f in FisherExactTest fisher ycsq SNPFunctions; do
  [ "$P/$f.o" -nt "$REF/$f.cpp" ] || $CXX $CF -fPIC -c "$REF/$f.cpp" -o "$P/$f.o"
done
patched_lc | $CXX $CF -fPIC -x c++ -c - -o "$P/likelihood_cache.o"
$CXX $CF -fPIC -shared "$HERE/ref_drivers/ref_ops_shim.cpp" $P/FisherExactTest.o $P/fisher.o $P/ycsq.o \
     $P/SNPFunctions.o $P/likelihood_cache.o -lm -o "$OUT/libref16gt_ops.so"
echo "build_ref: ok -> $OUT"
