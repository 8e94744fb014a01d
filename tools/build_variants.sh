#!/bin/bash
# build profiling variants of libdbot_b200.so into build/variants: tools/build_variants.sh name:"-DFLAG ..." ...
mkdir -p build/variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  (cd dbot_ros_b200/csrc && /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared $flags dbot_b200.cu -lpthread -ldl -o ../../build/variants/lib_$name.so) &
done
wait
ls build/variants
