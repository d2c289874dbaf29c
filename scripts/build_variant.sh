#!/bin/bash
# Build an A/B variant of the library: scripts/build_variant.sh <name> [-DFOO=1 ...]  -> scratch/variants/<name>.so
# (only fdr_stage.cu and runtime.cu see the extra defines; the other objects come from the normal build)
set -e
NAME=$1; shift
C=python-soxr_b200/csrc; B=$C/build; V=scratch/variants; mkdir -p $V/obj_$NAME
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-O2 -cudart static"
for f in fdr_stage; do
  nvcc $FLAGS "$@" -Xptxas -v -x cu -c $C/$f.cu -o $V/obj_$NAME/$f.o 2>&1 | grep -A2 "Function properties for _ZN2b25k_fdrI" | tail -2
done
nvcc -shared -cudart static -o $V/$NAME.so $B/design.o $B/plan.o $B/runtime.o $B/fft_stage.o $V/obj_$NAME/fdr_stage.o $B/fdr_stage_mono.o $B/vr.o $B/api.o
echo built $V/$NAME.so
