#!/bin/bash
# timing experiments: rebuild tiles_exec.cu only with -DAFB_EXP_* knobs, link with the regular objects
set -e
cd "$(dirname "$0")/../arcanefem_b200/csrc"
mkdir -p ../variants build_exp
build() {
  name=$1; shift
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -I../../include -I. --expt-relaxed-constexpr "$@" -c tiles_exec.cu -o build_exp/tiles_exec_$name.o
  objs=$(ls build/*.o | grep -v tiles_exec.o)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libafb200_$name.so $objs build_exp/tiles_exec_$name.o -cudart static
  echo built $name
}
build a57 -DAFB_EXP_A_PCT=57 &
build a0 -DAFB_EXP_A_PCT=0 &
build b0 -DAFB_EXP_B_PCT=0 &
build b75 -DAFB_EXP_B_PCT=75 &
build o0 -DAFB_EXP_OUT_PCT=0 &
build a0b0 -DAFB_EXP_A_PCT=0 -DAFB_EXP_B_PCT=0 &
build a0b0o0 -DAFB_EXP_A_PCT=0 -DAFB_EXP_B_PCT=0 -DAFB_EXP_OUT_PCT=0 &
build a57b75 -DAFB_EXP_A_PCT=57 -DAFB_EXP_B_PCT=75 &
wait
rm -rf build_exp
