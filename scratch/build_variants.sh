#!/bin/bash
# tuning builds of libafb200 with other executor geometries (arcanefem_b200/variants/libafb200_<name>.so)
set -e
cd "$(dirname "$0")/../arcanefem_b200/csrc"
mkdir -p ../variants
build() {
  name=$1; shift
  rm -rf build_$name; mkdir -p build_$name
  for f in afb_api scan connectivity pattern_rows assemble tiles_plan tiles_exec tiles_pipe pattern_tiled linear mesh_gen; do
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -I../../include -I. --expt-relaxed-constexpr "$@" -c $f.cu -o build_$name/$f.o &
  done
  wait
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libafb200_$name.so build_$name/*.o -cudart static
  rm -rf build_$name
  echo built $name
}
build B -DAFB_TG_THREADS=320 -DAFB_TG_MINB=3 -DAFB_TG_CMAX=704 -DAFB_TG_EMAX=1280 -DAFB_TG_FMAX=256 -DAFB_TG_LMAX=4608 -DAFB_TG_RT3=64 -DAFB_TG_RT2=160
build C -DAFB_TG_THREADS=384 -DAFB_TG_MINB=2 -DAFB_TG_CMAX=1408 -DAFB_TG_EMAX=2304 -DAFB_TG_FMAX=384 -DAFB_TG_LMAX=7680 -DAFB_TG_RT3=125 -DAFB_TG_RT2=288
build D -DAFB_TG_THREADS=256 -DAFB_TG_MINB=4 -DAFB_TG_CMAX=512 -DAFB_TG_EMAX=896 -DAFB_TG_FMAX=192 -DAFB_TG_LMAX=3328 -DAFB_TG_RT3=45 -DAFB_TG_RT2=120
build F -DAFB_TG_THREADS=448 -DAFB_TG_MINB=2 -DAFB_TG_CMAX=1408 -DAFB_TG_EMAX=2304 -DAFB_TG_FMAX=384 -DAFB_TG_LMAX=7680 -DAFB_TG_RT3=125 -DAFB_TG_RT2=288
