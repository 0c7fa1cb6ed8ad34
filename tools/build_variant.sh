#!/bin/bash
# Build a tuning variant of the library next to the product build: tools/build_variant.sh NAME "-DIACT_...=.." ["flags for iact_render.cu only"]
# -> variants/libiactrace_b200_NAME.so (select it with IACTRACE_B200_LIB=...; git-ignored, travels with gpurun).
set -e
cd "$(dirname "$0")/.."
mkdir -p variants/obj_$1
for f in iact_core iact_sample iact_render iact_vjp; do
  extra=""; [ "$f" = "iact_render" ] && extra="$3"
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $2 $extra -c iactrace_b200/csrc/$f.cu -o variants/obj_$1/$f.o &
done
wait
nvcc -shared -o variants/libiactrace_b200_$1.so variants/obj_$1/*.o
rm -rf variants/obj_$1
echo variants/libiactrace_b200_$1.so
