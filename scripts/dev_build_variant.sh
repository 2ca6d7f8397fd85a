#!/bin/bash
# Development A/B build: recompile the two fused-stage translation units with extra flags and
# link them with the other objects of the regular build into pystella_b200/_C/<name>.so
#   scripts/dev_build_variant.sh libpsb_polyold -DPSB_POLY_PER_SITE
set -e
name=$1; shift
cd "$(dirname "$0")/.."
obj=build/obj/libpystella_b200
out=build/obj/$name; mkdir -p $out
for tu in fused_stage_f64 fused_stage_f32; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
    -Iinclude -Ipystella_b200/csrc "$@" -c pystella_b200/csrc/$tu.cu -o $out/$tu.o &
done
wait
others=$(ls $obj/*.o | grep -v "fused_stage_f64.o\|fused_stage_f32.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o pystella_b200/_C/$name.so $out/*.o $others \
  -lcudart_static -ldl -lrt -lpthread
echo built pystella_b200/_C/$name.so
