#!/bin/bash
# usage: scripts/ablate.sh name1:"flags" name2:"flags" ...   builds maprasterization.jl_b200/variants/libmaprast_<name>.so
# (mr_fast.cu compiled with -DMR_QUICK <flags>, the other objects from build/)
set -e
cd "$(dirname "$0")/../maprasterization.jl_b200/csrc"
mkdir -p ../variants ../../build
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -DMR_QUICK $flags \
      -c -o ../../build/mr_fast_$name.o mr_fast.cu > ../../build/mr_fast_$name.log 2>&1 && \
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libmaprast_$name.so ../../build/mr_fast_$name.o \
      ../../build/mr_api.o ../../build/mr_generic.o ../../build/mr_synth.o ../../build/mr_ccl.o ../../build/mr_fill.o ../../build/mr_multi.o ../../build/mr_blur.o && echo "built $name" ) &
done
wait
