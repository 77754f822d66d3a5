#!/bin/bash
# Builds tuning variants of the library: scripts/build_variants.sh "<tag>:<nvcc -D flags>" ...
# -> towr_b200/variants/libtowr_b200_<tag>.so (select with TWB_LIB=<path>)
set -e
cd "$(dirname "$0")/../towr_b200/csrc"
make -s >/dev/null
mkdir -p ../variants
for v in "$@"; do
  tag="${v%%:*}"; flags="${v#*:}"
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -v $flags -c kernels.cu -o /tmp/kernels_$tag.o 2> ../variants/ptxas_$tag.log
  g++ -O2 -std=c++17 -fPIC -ffp-contract=off -I/usr/local/cuda/include $flags -c formulation.cc -o /tmp/formulation_$tag.o
  g++ -O2 -std=c++17 -fPIC -ffp-contract=off -I/usr/local/cuda/include $flags -c capi.cc -o /tmp/capi_$tag.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../variants/libtowr_b200_$tag.so spec.o lm_kernel.o /tmp/formulation_$tag.o /tmp/capi_$tag.o /tmp/kernels_$tag.o -lcudart
  echo "$tag: $flags  ->  $(grep -A2 'RomOutILi4\|DynOutILi4\|NodeOut' ../variants/ptxas_$tag.log | grep -o 'Used [0-9]* registers' | tr '\n' ' ')"
done
