#!/bin/bash
# build a variant of the WHOLE library with extra -D flags:  build_variant_all.sh <name> "<flags>"  -> tools/_ab/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_ab/obj_$1
for f in imagine360_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -I include $2 \
       -c $f -o tools/_ab/obj_$1/$(basename $f .cu).o &
done
wait
nvcc -shared -o tools/_ab/lib_$1.so tools/_ab/obj_$1/*.o -gencode arch=compute_100a,code=sm_100a
rm -rf tools/_ab/obj_$1
echo tools/_ab/lib_$1.so
