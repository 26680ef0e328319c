#!/bin/bash
# timing experiments: rebuild tiles_exec.cu only with -DAFB_EXP_* knobs, link with the regular objects
# usage: build_exp.sh name "-Dflags" [name "-Dflags" ...]
set -e
cd "$(dirname "$0")/../arcanefem_b200/csrc"
mkdir -p ../variants build_exp
build() {
  name=$1; shift
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -I../../include -I. --expt-relaxed-constexpr $@ -c tiles_exec.cu -o build_exp/tiles_exec_$name.o 2>/dev/null
  objs=$(ls build/*.o | grep -v tiles_exec.o)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libafb200_$name.so $objs build_exp/tiles_exec_$name.o -cudart static
  echo built $name
}
rm -f ../variants/*.so
while [ $# -gt 0 ]; do build "$1" "$2" & shift; shift; done
wait
rm -rf build_exp
