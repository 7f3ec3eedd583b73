#!/bin/bash
# Developer tool: builds tuning variants of libtcpt (same sources, different -D knobs) into toy_cpu_pathtracing_b200/lib/variants/
# usage: tools/build_variants.sh name1:"-DFOO=1" name2:"-DBAR=2 -DBAZ=3" ...
set -e
cd "$(dirname "$0")/../toy_cpu_pathtracing_b200/csrc"
mkdir -p ../lib/variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false -Xcompiler -fPIC -Xcompiler -ffp-contract=off"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ( $NV $flags -c tcpt_api.cu -o ../lib/variants/$name.o && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/libtcpt_$name.so ../lib/host_scene.o ../lib/variants/$name.o -Xcompiler -pthread && rm ../lib/variants/$name.o && echo built $name ) &
done
wait
