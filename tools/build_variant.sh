#!/bin/bash
# usage: tools/build_variant.sh NAME "-DFOO=1 -DBAR=2"   -> tools/_variants/libb200rt_NAME.so (tuning builds, git-ignored)
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_variants/obj_$1
C=homework-18-graphics-raytracer_b200/csrc
for f in rt_kernels rt_wavefront rt_image rt_filter_bench b200rt_api b200rt_group; do
  extra=""; [ $f = b200rt_api ] && extra="-Xcompiler -ffp-contract=off"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I include -I $C -fmad=false $2 -Xcompiler -fPIC $extra -c $C/$f.cu -o tools/_variants/obj_$1/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/_variants/libb200rt_$1.so tools/_variants/obj_$1/*.o homework-18-graphics-raytracer_b200/_lib/obj/host_world.o -ldl -lpthread
echo tools/_variants/libb200rt_$1.so
