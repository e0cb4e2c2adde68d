#!/bin/bash
# build a variant of the library with extra -D flags for attention.cu:  build_variant.sh <name> "<flags>"  -> tools/_ab/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -I include $2 \
     -c imagine360_b200/csrc/attention.cu -o tools/_ab/attention_$1.o
objs=$(ls imagine360_b200/csrc/_obj/*.o | grep -v attention.o)
nvcc -shared -o tools/_ab/lib_$1.so tools/_ab/attention_$1.o $objs -gencode arch=compute_100a,code=sm_100a
echo tools/_ab/lib_$1.so
