#!/bin/bash
# tuning builds of libafb200 with other executor geometries (arcanefem_b200/variants/libafb200_<name>.so)
# usage: build_variants.sh name "-Dflags" [name "-Dflags" ...]
set -e
cd "$(dirname "$0")/../arcanefem_b200/csrc"
mkdir -p ../variants
build() {
  name=$1; shift
  rm -rf build_$name; mkdir -p build_$name
  for f in $(ls *.cu | sed s/.cu//); do
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -I../../include -I. --expt-relaxed-constexpr $@ -c $f.cu -o build_$name/$f.o 2>/dev/null &
  done
  wait
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libafb200_$name.so build_$name/*.o -cudart static
  rm -rf build_$name
  echo built $name
}
rm -f ../variants/*.so
while [ $# -gt 0 ]; do build "$1" "$2"; shift; shift; done
